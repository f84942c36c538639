"""Development aid: where the end-to-end GBA call spends its time (host set-up phases + first iterations)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["COVINS_B200_SETUP_TRACE"] = "1"
import covins_b200
from covins_b200 import optimization as O, synth_map
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
ctx = covins_b200.Context(0)
p = synth_map.make_config(name)
for rep in range(2):
    t0 = time.perf_counter(); s = O.BaSolver(ctx, p, visual_only=False); ctx.sync(); t1 = time.perf_counter()
    print(f"rep {rep}: BaSolver() {1e3*(t1-t0):.1f} ms", flush=True)
    for i in range(4):
        t0 = time.perf_counter(); s.iterate(1); ctx.sync(); print(f"  iterate #{i}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
    t0 = time.perf_counter(); r = s.result(); print(f"  result(): {1e3*(time.perf_counter()-t0):.1f} ms"); 
    t0 = time.perf_counter(); s.close(); print(f"  close(): {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
