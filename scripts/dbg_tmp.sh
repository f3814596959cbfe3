timeout 600 python -m pytest tests/test_bucketing.py -q -m gpu -k "bucketed_steps" 2>&1 | grep "^E  \|test_bucketing.py:[0-9]" | head -12
