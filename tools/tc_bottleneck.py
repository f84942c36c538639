"""Development aid: which role limits the tensor-core matcher?  Times the C3 launch with the epilogue's selection and/or
the producers' expansion switched off (results are garbage in those modes)."""
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
for dbg, what in ((0, "full kernel"), (1, "no selection in the epilogue"), (2, "no expansion in the producers"), (3, "neither (MMA + TMEM traffic + barriers)"),
                  (7, "neither, and no TMEM read (MMA + barriers)"), (4, "full but no TMEM read"), (8, "full kernel, packed 16-bit TMEM read"),
                  (11, "neither, packed 16-bit TMEM read")):
    os.environ["COVINS_B200_TC_DEBUG"] = str(dbg)
    for _ in range(3): M.knn_match_hamming(ctx, q, t, (d_seg, h_seg), 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): M.knn_match_hamming(ctx, q, t, (d_seg, h_seg), 2)
    e1.record(); torch.cuda.synchronize()
    print(f"dbg={dbg} {what:45s} {e0.elapsed_time(e1)/10:.3f} ms")

# the resident-tile path: database appended once, device requests (kernel only, no expansion pre-pass)
os.environ["COVINS_B200_TC_DEBUG"] = "0"
db = M.DescriptorDatabase(ctx, reserve_rows=n_kf * nf)
db.append(t.cpu().numpy(), [nf] * n_kf)
for dbg, what in ((0, "resident tiles: full kernel"), (1, "resident tiles: no selection in the epilogue")):
    os.environ["COVINS_B200_TC_DEBUG"] = str(dbg)
    for _ in range(3): db.match_hamming_dev(q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): db.match_hamming_dev(q)
    e1.record(); torch.cuda.synchronize()
    print(f"dbg={dbg} {what:45s} {e0.elapsed_time(e1)/10:.3f} ms")
os.environ["COVINS_B200_TC_DEBUG"] = "0"
db.close()

# parity of the packed-read variant against the plain one on data with true matches
desc, _ = synth.orb_keyframes(5, 64, 1000)
tt = torch.from_numpy(desc.reshape(-1, 32)).to(dev); qq = tt[:1000].clone()
hs = synth.seg_ptr_uniform(64, 1000); ds = torch.from_numpy(hs).to(dev)
os.environ["COVINS_B200_TC_DEBUG"] = "0"; a = M.knn_match_hamming(ctx, qq, tt, (ds, hs), 2)
os.environ["COVINS_B200_TC_DEBUG"] = "8"; b = M.knn_match_hamming(ctx, qq, tt, (ds, hs), 2)
print("pack16 parity:", all(bool((x == y).all()) for x, y in zip(a, b)))
