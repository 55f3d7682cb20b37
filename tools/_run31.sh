for v in "1 0" "1 5" "1 8" "1 6" "0 8" "0 0"; do
set -- $v
export PDE_JAC_FUSED=$1
if [ "$2" = "0" ]; then unset PDE_ELEM_BLOCKS; else export PDE_ELEM_BLOCKS=$2; fi
python bench.py --steps 5 --warmup 3 --solve 255 --solve-full 4095 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solve_4097']
print('FUSED',os.environ.get('PDE_JAC_FUSED'),'BLOCKS',os.environ.get('PDE_ELEM_BLOCKS'),'| bicg ms %.3f'%d['bicgstab']['ms_per_iteration'],'| solve %.3f'%s['seconds'],'certify %.3f its %d'%(s['certify']['seconds'],s['certify']['krylov_iters']),'|',[(x['lattice'],x['solver'],round(x['seconds'],3),x['iterations'],x.get('krylov_iters'),x.get('euler_fallbacks')) for x in d['solve']])"
python -m pytest tests/test_solvers_gpu.py -q 2>&1 | tail -1
done
