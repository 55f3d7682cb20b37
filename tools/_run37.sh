python tools/e2e_breakdown.py 2>&1 | grep -E "api|linear|duplex"
