"""Development aid: one GBA iteration on a synthetic config (profiling target; run under gpurun/ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covins_b200
from covins_b200 import optimization as O, synth_map
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
ctx = covins_b200.Context(0)
p = synth_map.make_config(name)
s = O.BaSolver(ctx, p, visual_only=False)
s.iterate(1); ctx.sync()
print("done")
