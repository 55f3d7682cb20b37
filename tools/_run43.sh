for kb in 72 110 140 180 220; do echo "SMEM_KB $kb"; PDE_SWEEP_SMEM_KB=$kb python tools/sweep_relax_exp.py 4095 2>&1 | grep "relax_every 4" | tail -1; done
