python tools/e2e_numa.py 2>&1 | tail -12
python tools/e2e_numa.py bind 2>&1 | tail -12
nvidia-smi topo -m 2>/dev/null | head -6
