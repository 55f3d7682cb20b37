python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py > gpurun_out/check_sharded2.log 2>&1
tail -25 gpurun_out/check_sharded2.log
