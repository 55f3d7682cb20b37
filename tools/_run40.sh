( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 ) > gpurun_out/bench_n8_r2c.json 2> gpurun_out/bench_n8_r2c.err
tail -4 gpurun_out/bench_n8_r2c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8_r2c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('parity_ok'), d['e2e'], d['bicgstab'], d.get('slab_4095'))
PY
