python tools/newton_small_probe.py 2>&1 | tail -12
( time python -m pytest tests/test_solvers_gpu.py tests/test_cloud_gpu.py tests/test_grid3d_gpu.py -x -q ) 2>&1 | tail -6
python tools/krylov_timing.py 2>&1 | tail -5
