"""Multi-GPU parity check (run under torchrun on N GPUs): the landmark-sharded GBA with the all-reduce of the reduced
normal equations must give the same states as the single-GPU solve (differences: floating-point summation order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import covins_b200
from covins_b200 import optimization as O, synth_map

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = covins_b200.Context(local)
ok = True
for name, vo, iters in (("small", False, 6), ("small", True, 6), ("C1", False, 4)):
    p = synth_map.make_config(name)
    s = O.BaSolver(ctx, p, visual_only=vo, rank=rank, world=world, allreduce=O.torch_allreduce())
    n = s.iterate(iters)
    r = s.result()
    s.close()
    # gather the landmark shards
    lm = torch.from_numpy(np.where((r["lm_owner"] == rank)[:, None], r["lm"], 0.0)).cuda()
    dist.all_reduce(lm)
    if rank == 0:
        ref = O.solve(ctx, p, iters, visual_only=vo)
        inc = r["lm_owner"] >= 0
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
        well = inc & (np.abs(ref["lm"]).max(1) < 100)
        e = (rel(r["pose"], ref["pose"]), rel(r["speedbias"], ref["speedbias"]), rel(lm.cpu().numpy()[well], ref["lm"][well]))
        same_steps = r["steps"] == ref["steps"] and r["iterations"] == ref["iterations"]
        good = max(e) < 1e-6 and same_steps and abs(r["final_cost"] - ref["final_cost"]) < 1e-4 * ref["final_cost"]  # ill-posed run-away landmarks make the cost itself chaotic at 1e-6
        ok &= good
        print(f"{name} visual_only={vo} world={world}: iterations {n}, rel err pose/sb/lm {e}, steps equal {same_steps}, "
              f"cost {r['final_cost']:.8e} vs {ref['final_cost']:.8e} -> {'OK' if good else 'MISMATCH'}", flush=True)
# ---- map-wide k-NN, database sharded by keyframe block, one all-gather + merge ----
from covins_b200 import matching as M
dev = torch.device("cuda", local)
for metric, dim, n_kf, nf, nq in (("hamming", 32, 2000, 1000, 1000), ("l2", 128, 4000, 300, 300)):
    g = torch.Generator(device=dev).manual_seed(7)           # same database on every rank, each keeps its slice
    t = torch.randint(0, 256, (n_kf * nf, dim), dtype=torch.uint8, device=dev, generator=g)
    t[nf * 5:nf * 6] = t[:nf]                                 # duplicates → ties across rows
    t[-nf:] = t[:nf]                                          # ... and across shards
    q = t[:nq].clone()
    seg = np.arange(n_kf + 1, dtype=np.int64) * nf
    cuts = M.shard_rows(n_kf * nf, world, seg)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    knn = M.knn_match_hamming if metric == "hamming" else M.knn_match_l2
    for _ in range(2):
        mi, md = M.knn_match_sharded(ctx, q, t[lo:hi].contiguous(), lo, 2, metric)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mi, md = M.knn_match_sharded(ctx, q, t[lo:hi], lo, 2, metric)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ri, rd = knn(ctx, q, t, None, 2)
    good = bool(torch.equal(mi, ri[0]) and torch.equal(md, rd[0]))
    flag = torch.tensor([int(good)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ok &= bool(flag.item())
        print(f"map-wide sharded k-NN {metric}: {nq} queries vs {n_kf*nf} rows over {world} GPUs: {ms.item():.3f} ms per query block "
              f"({nq*n_kf*nf/ms.item()/1e6:.0f} Gpairs/s incl. all-gather + merge) -> {'OK (bit-identical to the single-GPU k-NN)' if flag.item() else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
