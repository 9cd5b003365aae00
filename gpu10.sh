python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python gpu9.py
