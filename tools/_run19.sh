python -m pytest tests/test_grid3d_gpu.py -x -q -k 129 2>&1 | tail -3
python tools/grid3d_profile.py 129 3 && python tools/grid3d_profile.py 257 3
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_g3.csv python tools/grid3d_profile.py 257 1 > gpurun_out/ncu_g3a.log 2>&1
grep -c . gpurun_out/launches_g3.csv
ncu --set full --clock-control none --import-source on -k regex:k_deep3 -c 1 -o gpurun_out/prof_g3 -f python tools/grid3d_profile.py 257 1 > gpurun_out/ncu_g3b.log 2>&1
tail -2 gpurun_out/ncu_g3b.log
