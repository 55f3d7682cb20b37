python -m pytest tests/test_ddg_gpu.py -x -q -k "pinned or torch_inputs" 2>&1 | tail -3
python bench.py --steps 50 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e'])"
python bench.py --steps 50 --warmup 3 --solve 0 --solve-full 0 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 --e2e-steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e'])"
