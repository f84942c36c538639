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
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
