python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -3
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 ) > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
tail -3 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('parity_ok'), d['e2e']['ms_per_step'], d['e2e']['equals_device_resident_result'], d['bicgstab']['ms_per_iteration'], d['slab_4095']['ms_per_step'], d['slab_4095']['parity']['parity_ok'])
PY
