python -m pytest tests/test_ddg_gpu.py tests/test_sweep_gpu.py tests/test_solvers_gpu.py -x -q 2>&1 | tail -3
