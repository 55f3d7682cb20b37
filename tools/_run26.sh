( time python -m pytest tests/test_solvers_gpu.py tests/test_sweep_gpu.py -x -q ) 2>&1 | tail -6
python bench.py --steps 50 --warmup 3 --solve 255 --solve-full 4095 --pucci-full 8191 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err
tail -3 gpurun_out/bench_k3.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_k3.json').read().strip().splitlines()[-1]); print(d['bicgstab']); s=d['solve_4097']; print(s['seconds'], s['converged'], s['sweep']['seconds'], s['certify']); print(d['pucci_8193']); print([(x['lattice'],x['solver'],round(x['seconds'],4),x['iterations'],x.get('krylov_iters')) for x in d['solve']])"
PDE_JAC_FUSED=0 python bench.py --steps 50 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unfused', d['bicgstab'])"
