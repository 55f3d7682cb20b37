python tools/e2e_chunks.py 2>&1 | tail -14
