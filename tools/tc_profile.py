import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, covins_b200
from covins_b200 import matching as M, synth
os.environ["COVINS_B200_MATCH_KERNEL"] = "tc"
ctx = covins_b200.Context(0); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n_kf, nf, nq = 2000, 1000, 1000
t = torch.randint(0, 256, (n_kf * nf, 32), dtype=torch.uint8, device=dev, generator=g); q = t[:nq].clone()
h_seg = synth.seg_ptr_uniform(n_kf, nf); d_seg = torch.from_numpy(h_seg).to(dev)
for _ in range(3):
    M.match_candidates_hamming(ctx, q, t, (d_seg, h_seg), 40.0, 0.8)
torch.cuda.synchronize()
