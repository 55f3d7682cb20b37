python - <<'PY'
import ctypes, json
from pyellipticfd_b200 import _lib
import torch
torch.cuda.init()
v = ctypes.c_double(); out = {}
for kind, name in ((1,"dadd"),(2,"dmul"),(3,"cmpsel"),(4,"exact_pairs"),(5,"dsetp_independent")):
    _lib.check(_lib.lib().pde_fp64_probe(0, kind, ctypes.byref(v))); out[name+"_per_s"] = v.value
_lib.check(_lib.lib().pde_fp64_peak(0, ctypes.byref(v))); out["dfma_per_s"] = v.value
print(json.dumps(out))
PY
python bench.py --steps 20 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 > gpurun_out/plain_bench_short.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 20 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
