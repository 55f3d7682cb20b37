python tools/sweep_profile.py 4095 41 > gpurun_out/plain_sweep.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_hull_pass -s 960 -c 6 -o gpurun_out/prof_r2_sweep_g -f python tools/sweep_profile.py 4095 41 > gpurun_out/ncu_sweep2.log 2>&1
tail -3 gpurun_out/ncu_sweep2.log
