for b in 4 5 8; do PDE_ELEM_BLOCKS=$b python tools/certify_counts.py 4095 2>&1 | tail -1; done
python tools/certify_counts.py 4095 2>&1 | tail -1
python tools/certify_counts.py 4095 2>&1 | tail -1
