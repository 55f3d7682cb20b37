python tools/e2e_native_trace.py 256 2>&1 | tail -40
