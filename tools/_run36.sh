python tools/e2e_timeline.py 256 2>&1 | tail -22
