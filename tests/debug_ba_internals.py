"""Development aid (run under gpurun): compare the CUDA engine's first-iteration internals with the oracle's."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # checker-side development aid: lives under tests/ because it uses the oracle
import numpy as np, scipy.sparse as sp, math
import covins_b200
from covins_b200 import optimization as O, synth_map
from oracle import ba_oracle as bo

vo = (sys.argv[1] == "1") if len(sys.argv) > 1 else True
p = synth_map.make_config("tiny")
ctx = covins_b200.Context(0)
s = O.BaSolver(ctx, p, visual_only=vo)
pr = bo.Problem(p, visual_only=vo, loop_loss=1.0)
cost, r, J, _ = pr.evaluate(pr.pose, pr.sb, pr.lm)
col_sq = np.asarray(J.multiply(J).sum(0)).reshape(-1)
scale = 1.0 / (1.0 + np.sqrt(col_sq)); scale[~pr.active] = 0
Js = (J @ sp.diags(scale)).tocsr()
nc = pr.ncam
def cmp(name, which, ref):
    c, l = s.debug_vector(which)
    got = np.concatenate([c[:nc], l])
    d = np.abs(got - ref); i = int(np.argmax(d))
    print(f"{name:8s} max abs diff {d.max():.3e} at {i} (got {got[i]:.6e} ref {ref[i]:.6e}) | ref max {np.abs(ref).max():.3e}; cam diff {d[:nc].max():.3e} lm diff {d[nc:].max() if len(d)>nc else 0:.3e}")
cmp("scale", 0, scale)
s.iterate(1)
diag = np.sqrt(np.clip(np.asarray(Js.multiply(Js).sum(0)).reshape(-1), 1e-6, 1e32))
g = Js.T @ r
grad = g / diag
x = bo.solve_normal_equations(Js, r, diag * math.sqrt(1e-8), pr.ncam)
gn = -x * diag
act = pr.active
cmp("colsq", 1, np.where(act, np.asarray(Js.multiply(Js).sum(0)).reshape(-1), 0))
cmp("diag", 2, np.where(act, diag, 1.0))
cmp("g", 3, np.where(act, g, 0) * 1.0)
cmp("grad", 4, np.where(act, grad, 0))
cmp("x", 7, np.where(act, x, 0))
cmp("gn", 5, np.where(act, gn, 0))
print("gn_norm", np.linalg.norm(gn), "radius 1e4")
res = s.result(); print(res["cost"], res["steps"])
ref = bo.solve(pr, 1); print(ref["cost"], ref["steps"])
# ---- second stage: the step and the candidate state
c6, l6 = s.debug_vector(6)
step_gpu = np.concatenate([c6[:nc], l6])
gn_norm = np.linalg.norm(gn)
print("step vs -x: ", np.abs(step_gpu - np.where(act, -x, 0)).max(), " |x| max", np.abs(x).max())
delta = step_gpu * scale
cp, cs, cl = pr.plus(pr.pose, pr.sb, pr.lm, delta)
print("oracle cost at plus(GPU step):", pr.evaluate(cp, cs, cl, with_jac=False)[0])
print("pose diff GPU result vs oracle plus(GPU step):", np.abs(res["pose"] - cp.numpy()).max(), " lm diff:", np.abs(res["lm"] - cl.numpy()).max(),
      " sb diff:", np.abs(res["speedbias"] - cs.numpy()).max())
print("pose diff GPU vs oracle 1-iter result:", np.abs(res["pose"] - ref["pose"].numpy()).max(), "lm:", np.abs(res["lm"] - ref["lm"].numpy()).max())
j = int(np.argmax(np.abs(res["lm"] - ref["lm"].numpy()).max(1))); print("worst lm", j, res["lm"][j], ref["lm"].numpy()[j], p["lm"][j], "in problem", pr.lm_in[j])
