python -m pytest tests/test_solvers_gpu.py tests/test_cloud_gpu.py -x -q 2>&1 | tail -3
( time python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_r2a.err
