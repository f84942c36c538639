"""Development aid: per-role clock64 stamps of the tensor-core matcher's first tiles (CTA 0), for the full kernel and the skeleton."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, covins_b200
from covins_b200 import matching as M, synth
ctx = covins_b200.Context(0); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n_kf, nf, nq = 2000, 1000, 1000
t = torch.randint(0, 256, (n_kf * nf, 32), dtype=torch.uint8, device=dev, generator=g); q = t[:nq].clone()
h_seg = synth.seg_ptr_uniform(n_kf, nf); d_seg = torch.from_numpy(h_seg).to(dev)
os.environ["COVINS_B200_MATCH_KERNEL"] = "tc"
for dbg in (0, 7):
    os.environ["COVINS_B200_TC_DEBUG"] = str(dbg)
    for _ in range(3): M.knn_match_hamming(ctx, q, t, (d_seg, h_seg), 2)
    torch.cuda.synchronize()
    print(f"==== dbg={dbg}", file=sys.stderr, flush=True)
    os.environ["COVINS_B200_TC_DEBUG"] = str(dbg | 16)
    M.knn_match_hamming(ctx, q, t, (d_seg, h_seg), 2)
    torch.cuda.synchronize()
