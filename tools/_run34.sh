( time python bench.py --impl reference ) > gpurun_out/bench_r2d_ref.json 2> gpurun_out/bench_r2d_ref.err
tail -4 gpurun_out/bench_r2d_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2d_ref.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['steps'], d['cpu_baseline']['sample'], d['config'].get('wall_s_total_incl_process_start_and_calibration'))"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
