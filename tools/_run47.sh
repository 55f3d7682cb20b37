python -m pytest tests/test_grid3d_gpu.py -x -q 2>&1 | tail -4
