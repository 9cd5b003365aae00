python -m pytest tests/test_cuda_properties.py -q -x 2>&1 | tail -15
