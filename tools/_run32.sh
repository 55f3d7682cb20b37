( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_r2d.log 2>&1
grep -E "passed|failed" gpurun_out/gputest_r2d.log | tail -2
( time python bench.py ) > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
tail -3 gpurun_out/bench_r2d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
s=d['solve_4097']; print({k:s[k] for k in ('seconds','converged')}, s['sweep']['seconds'], s['sweep']['sweeps'], s['certify']['seconds'], s['certify']['krylov_iters'], s['first_call'])
print(d['bicgstab']); print(d['pucci_8193']); print([ (x['lattice'],x['solver'],round(x['seconds'],4),x['iterations']) for x in d['solve']])
PY
