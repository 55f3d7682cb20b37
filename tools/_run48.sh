python -m pytest tests/test_grid3d_gpu.py -x -q 2>&1 | tail -4
python tools/grid3d_profile.py 129 3 | tail -1
python tools/grid3d_profile.py 257 3 | tail -1
PDE_BAND_CLASSES=2 python tools/grid3d_profile.py 257 3 | tail -1
