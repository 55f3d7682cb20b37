( time python -m pytest tests/test_grid3d_gpu.py -x -q ) > gpurun_out/grid3d_test.log 2>&1
tail -30 gpurun_out/grid3d_test.log
( time python bench.py --steps 50 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 ) > gpurun_out/bench_g3.json 2> gpurun_out/bench_g3.err
tail -3 gpurun_out/bench_g3.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_g3.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('grid3d'),indent=1)); print(d['clocks'])
"
