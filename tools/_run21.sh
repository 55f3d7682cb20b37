python tools/newton_small_probe.py 2>&1 | tail -20
