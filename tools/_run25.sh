python tools/sweep_relax_exp.py 4095
python tools/grid3d_profile.py 257 3 | tail -1
PDE_K3_CTAS=1 python tools/grid3d_profile.py 257 3 | tail -1
python bench.py --steps 50 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 > gpurun_out/bench_k2.json 2> gpurun_out/bench_k2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_k2.json').read().strip().splitlines()[-1]); print(d['bicgstab'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_k2.csv python bench.py --steps 5 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 > gpurun_out/ncu_k2.log 2>&1
grep -c . gpurun_out/launches_k2.csv
