python -m pytest tests/test_solvers_gpu.py -x -q -k fused 2>&1 | tail -15
