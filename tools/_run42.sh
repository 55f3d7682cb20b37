export PATH=/usr/local/cuda/bin:$PATH
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_grid3d_gpu.py -x -q -k "structured_3d_kernel_bitexact or band_points" > gpurun_out/sanitizer_3d.log 2>&1; echo "rc3d $?"
tail -5 gpurun_out/sanitizer_3d.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_solvers_gpu.py tests/test_ddg_gpu.py -x -q -k "fused or pinned or bicgstab" > gpurun_out/sanitizer_k.log 2>&1; echo "rck $?"
tail -5 gpurun_out/sanitizer_k.log
