"""Profiling target (ncu): the C3 matching request on the resident map — 1000-feature query keyframe vs 2000 keyframes x 1000 ORB
descriptors held by the database as packed rows + tensor-core operand tiles (cvb_db_match_hamming_dev → tc_xt_kernel<2>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, covins_b200
from covins_b200 import matching as M, synth
ctx = covins_b200.Context(0); dev = torch.device("cuda", 0)
n_kf, nf = 2000, 1000
desc, _ = synth.orb_keyframes(seed=3, n_kf=n_kf, n_feat=nf)
db = M.DescriptorDatabase(ctx, reserve_rows=n_kf * nf)
db.append(desc.reshape(-1, 32), [nf] * n_kf)
q = torch.from_numpy(np.ascontiguousarray(desc[123])).to(dev)
for _ in range(3):
    out = db.match_hamming_dev(q, 40.0, 0.8)
torch.cuda.synchronize()
print("accepted", int(out[2].sum().item()))
