( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 100 --warmup 5 ) > gpurun_out/bench_n4_final.json 2> gpurun_out/bench_n4_final.err
tail -3 gpurun_out/bench_n4_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_ok'), d['e2e']['ms_per_step'], d['e2e']['equals_device_resident_result'], d['bicgstab']['ms_per_iteration'], d['slab_4095']['ms_per_step'], d['slab_4095']['kernel_only_ms'], d['slab_4095']['parity']['parity_ok'])
PY
