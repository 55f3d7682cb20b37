python -m pytest tests/test_solvers_gpu.py tests/test_sweep_gpu.py -x -q 2>&1 | grep -v "^$" | tail -40
