set -x
python -m pytest tests/test_sweep_gpu.py -x -q 2>&1 | tail -3
python tools/sweep_kernel_timing.py 4 1023 4095 2>&1 | tail -2
PDE_SWEEP_SMEM_KB=72 python tools/sweep_kernel_timing.py 4 4095 2>&1 | tail -1
PDE_SWEEP_SMEM_KB=150 python tools/sweep_kernel_timing.py 4 4095 2>&1 | tail -1
python tools/solve4097.py 4095 --no-cpu-baseline 2>&1 | tail -1
