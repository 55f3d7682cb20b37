python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_n8_r2.json 2> gpurun_out/bench_n8_r2.err
tail -5 gpurun_out/bench_n8_r2.err
