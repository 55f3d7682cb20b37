for f in 0 1; do for b in 4 5 6 8; do PDE_JAC_FUSED=$f PDE_ELEM_BLOCKS=$b python tools/certify_counts.py 4095 2>&1 | tail -1; done; done
for f in 0 1; do for b in 5 8; do PDE_JAC_FUSED=$f PDE_ELEM_BLOCKS=$b python tools/certify_counts.py 2047 2>&1 | tail -1; done; done
