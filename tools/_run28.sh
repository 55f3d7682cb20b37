PDE_JAC_FUSED=0 python tools/fused_check.py 300
PDE_JAC_FUSED=1 python tools/fused_check.py 300
