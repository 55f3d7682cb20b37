set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py 2>&1 | tail -12
PDE_HALO=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/check_sharded.py 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err
tail -3 gpurun_out/bench_n2_r2.err
