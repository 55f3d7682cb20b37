PDE_JAC_FUSED=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k3.csv python tools/krylov_timing.py > gpurun_out/ncu_k3.log 2>&1
grep -c . gpurun_out/launches_k3.csv
PDE_JAC_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:k_jac_fused_st -s 3 -c 1 -o gpurun_out/prof_fused -f python tools/krylov_timing.py > gpurun_out/ncu_k3b.log 2>&1
tail -2 gpurun_out/ncu_k3b.log
ncu --set full --clock-control none -k regex:k_elem -s 12 -c 4 -o gpurun_out/prof_elem -f python tools/krylov_timing.py > gpurun_out/ncu_k3c.log 2>&1
tail -1 gpurun_out/ncu_k3c.log
