set -x
python -m pytest tests/test_sweep_gpu.py -x -q 2>&1 | tail -15
python tools/sweep_kernel_timing.py 4 1023 4095 2>&1 | tail -5
PDE_SWEEP_SMEM_KB=72 python tools/sweep_kernel_timing.py 4 4095 2>&1 | tail -3
PDE_SWEEP_SMEM_KB=220 python tools/sweep_kernel_timing.py 4 4095 2>&1 | tail -3
python tools/sweep_timing.py 4 1023 4095 2>&1 | tail -3; cat gpurun_out/sweep_timing.jsonl | tail -2
