python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python gpu14.py 2>&1 | tail -8
