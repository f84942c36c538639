"""Development aid: time GBA / PGO iterations of the CUDA engine on a named synthetic config (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import covins_b200
from covins_b200 import optimization as O, synth_map

name = sys.argv[1] if len(sys.argv) > 1 else "C1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = covins_b200.Context(0)
t0 = time.time(); p = synth_map.make_config(name); print(f"{name}: generated in {time.time()-t0:.1f}s  K={p['K']} L={p['L']} obs={len(p['obs_kf'])}")
for vo in (True, False):
    t0 = time.time(); s = O.BaSolver(ctx, p, visual_only=vo); ctx.sync(); t_setup = time.time() - t0
    l0 = ctx.launch_count(); t0 = time.time(); n = s.iterate(iters); ctx.sync(); dt = time.time() - t0
    r = s.result()
    print(f"  visual_only={vo}: setup {t_setup*1e3:.0f} ms, {n} iterations in {dt*1e3:.1f} ms = {dt/n*1e3:.1f} ms/it ({n/dt:.2f} it/s), "
          f"{(ctx.launch_count()-l0)/n:.0f} launches/it, cost {r['initial_cost']:.4e} -> {r['final_cost']:.4e}, steps {r['steps']}")
    s.close()
