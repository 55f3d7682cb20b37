for kb in 40 110; do PDE_SWEEP_SMEM_KB=$kb python tools/sweep_kernel_timing.py 4 4095 2>&1 | tail -1; done
PDE_SWEEP_SMEM_KB=40 python tools/sweep_ablate.py 4095 2>&1 | tail -1
