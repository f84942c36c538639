"""Multi-GPU check of the distributed factorisation (run under torchrun on N GPUs): the peer path (reduce-scatter by pull +
column-distributed Cholesky over CUDA IPC) must give the single-GPU states (rounding aside) with identical accept/reject
sequences, on small configs and C1/C3; prints per-iteration timing with and without the peer path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import covins_b200
from covins_b200 import optimization as O, synth_map

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
import datetime
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=int(os.environ.get("COVINS_NCCL_TIMEOUT", "90"))))
ctx = covins_b200.Context(local)
ok = True
cases = [("small", False, 6), ("small", True, 6), ("C1", False, 5)]
if len(sys.argv) > 1 and sys.argv[1] == "tiny":
    cases = [("small", False, 3)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cases.append(("C3", False, 4))
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
for name, vo, iters in cases:
    p = synth_map.make_config(name)
    ref = O.solve(ctx, p, iters, visual_only=vo) if rank == 0 else None
    ctx.sync(); dist.barrier()
    for mode in (True, False):
        s = O.BaSolver(ctx, p, visual_only=vo, rank=rank, world=world, allreduce=O.torch_allreduce(), p2p=mode)
        used = s.p2p
        s.iterate(2); ctx.sync(); dist.barrier()     # warm-up (graph capture happens on the 2nd factorisation)
        s.restart(); s.timing(reset=True); ctx.sync(); dist.barrier()
        t0 = time.perf_counter()
        n = s.iterate(iters); ctx.sync()
        dt = time.perf_counter() - t0
        r = s.result(); tm = s.timing(); s.close()
        lm = torch.from_numpy(np.where((r["lm_owner"] == rank)[:, None], r["lm"], 0.0)).cuda()
        dist.all_reduce(lm)
        if rank == 0:
            well = (ref["lm_owner"] >= 0) & (np.abs(ref["lm"]).max(1) < 100)
            e = (rel(r["pose"], ref["pose"]), rel(r["speedbias"], ref["speedbias"]), rel(lm.cpu().numpy()[well], ref["lm"][well]))
            same = r["steps"] == ref["steps"] and r["iterations"] == ref["iterations"]
            good = max(e) < 1e-6 and same
            ok &= good
            it = max(r["iterations"], 1)
            print(f"{name} vo={vo} world={world} p2p={'on' if used else 'off'}: {1e3 * dt / max(n, 1):.2f} ms/it "
                  f"(factor {tm['factor_ms']/it:.2f}, blocks+schur+exchange {tm['build_schur_ms']/it:.2f}, solve {tm['solve_ms']/it:.2f}) "
                  f"rel err pose/sb/lm {e[0]:.1e}/{e[1]:.1e}/{e[2]:.1e} steps equal {same} -> {'OK' if good else 'MISMATCH'}", flush=True)
        dist.barrier()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_P2P_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
