( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_r2b.log 2>&1
tail -3 gpurun_out/gputest_r2b.log
( time python bench.py ) > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -3 gpurun_out/bench_r2b.err
( time python bench.py --impl reference ) > gpurun_out/bench_r2b_ref.json 2> gpurun_out/bench_r2b_ref.err
tail -4 gpurun_out/bench_r2b_ref.err
