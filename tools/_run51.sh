for re in 2 3 4 6; do
export PDE_SWEEP_RELAX_EVERY=$re; python bench.py --steps 5 --warmup 3 --solve 0 --solve-full 4095 --pucci-full 0 --no-cpu-baseline --cloud-h0 0.01 --grid3d 0 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solve_4097']
print('relax_every',os.environ['PDE_SWEEP_RELAX_EVERY'],'solve %.3f'%s['seconds'],'sweep %.3f / %d'%(s['sweep']['seconds'],s['sweep']['sweeps']),'certify %.3f its %d'%(s['certify']['seconds'],s['certify']['krylov_iters']), s['converged'])"
done
