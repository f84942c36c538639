"""BASELINE config 5 (12 agents, 10 008 KF, 1 M landmarks, ~8 M observations) on one GPU: a few trust-region iterations of the
visual-inertial GBA on the packed tile store, and one ORB query keyframe against the 10 M resident descriptors of the map.
Prints one JSON line.  Run under gpurun (map generation ~2 min of host time)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import covins_b200
from covins_b200 import matching as M, optimization as O, synth, synth_map

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = covins_b200.Context(0); dev = torch.device("cuda", 0)
out = {"config": "C5"}
t0 = time.perf_counter(); p = synth_map.make_config("C5"); out["map_generation_s"] = round(time.perf_counter() - t0, 1)
out.update(K=int(p["K"]), L=int(p["L"]), n_obs=int(len(p["obs_kf"])), n_imu=int(len(p["imu_i"])))
torch.cuda.reset_peak_memory_stats()
free0 = torch.cuda.mem_get_info()[0]
t0 = time.perf_counter(); s = O.BaSolver(ctx, p, visual_only=False); ctx.sync(); out["setup_s"] = round(time.perf_counter() - t0, 2)
out["hbm_used_gb"] = round((free0 - torch.cuda.mem_get_info()[0]) / 1e9, 2)
s.iterate(1); ctx.sync(); s.restart(); s.timing(reset=True); ctx.sync()
t0 = time.perf_counter(); n = s.iterate(iters); ctx.sync(); dt = time.perf_counter() - t0
r = s.result(); tm = s.timing(); s.close()
it = max(r["iterations"], 1)
out.update(iterations=int(n), s_per_iteration=round(dt / max(n, 1), 3), initial_cost=r["initial_cost"], final_cost=r["final_cost"], steps=r["steps"],
           phase_ms_per_iteration={k: round(v / it, 2) for k, v in tm.items() if k.endswith("_ms")},
           factor_tflops=round(tm["factor_flops"] / (tm["factor_ms"] * 1e-3) / 1e12, 2) if tm.get("factor_ms") else None,
           factor_tflop_total=round(tm["factor_flops"] / 1e12, 2))
del p
# matching: the whole C5 map resident (10 008 keyframes x 1000 ORB descriptors), one query keyframe
n_kf, nf = 10008, 1000
db = M.DescriptorDatabase(ctx, reserve_rows=n_kf * nf)
t0 = time.perf_counter()
for c in range(0, n_kf, 1112):
    d, _ = synth.orb_keyframes(seed=100 + c, n_kf=min(1112, n_kf - c), n_feat=nf)
    db.append(d.reshape(-1, 32), [nf] * d.shape[0])
    if c == 0: q = torch.from_numpy(np.ascontiguousarray(d[7])).to(dev); hq = np.ascontiguousarray(d[7])
ctx.sync(); out["db_append_s"] = round(time.perf_counter() - t0, 1)
for _ in range(3): db.match_hamming_dev(q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): mt, md, nm = db.match_hamming_dev(q)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter()
for _ in range(5): res = db.match_hamming(hq)
dt = (time.perf_counter() - t0) / 5
out["match"] = {"rows": n_kf * nf, "ms_per_query_kf_device": round(ms, 3), "gpairs_per_s": round(n_kf * nf * nf / ms / 1e6, 1), "accepted": int(nm.sum().item()),
                "ms_per_request_host": round(dt * 1e3, 3), "accepted_host": int(res[0].sum())}
db.close()
print(json.dumps(out))
