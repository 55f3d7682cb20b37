for rt in 1e-8 1e-10; do
  KRYLOV_RTOL=$rt MAX_ITERS=1500 timeout 240 python tools/solve_scaling.py 4 511 1023 2>&1 | tail -2
done
