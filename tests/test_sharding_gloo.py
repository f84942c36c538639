"""CPU, world_size 2 over gloo: the host-side logic of the N>1 paths.
  * matching shards the map by keyframe with no data-path collective — a gather of the per-shard match counts must
    equal the single-process result;
  * map-wide k-NN shards the database rows by keyframe block: local top-k, ONE all-gather of the lists, merge by
    (distance, global trainIdx) must equal the single-process k-NN (ties across shard boundaries included);
  * GBA shards landmark blocks: the sum over ranks of the rank-local Schur complements (all-reduce) must equal the
    reduced camera system of the whole problem, and merge_sharded_landmarks must pick each landmark from its owner.
The CUDA kernels are not involved (no GPU here); the arithmetic comes from the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from covins_b200 import matching as M
from covins_b200 import optimization as O
from covins_b200 import synth, synth_map


def _mapwide_case():
    """query + database with many exact duplicates so that ties straddle the shard boundary"""
    rng = np.random.default_rng(21)
    base = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    t = base[rng.integers(0, 60, 4000)].copy()
    flip = rng.random(t.shape) < 0.02
    t[flip] ^= rng.integers(1, 256, int(flip.sum()), dtype=np.uint8)
    q = base[:50].copy()
    seg = np.concatenate([[0], np.cumsum(rng.integers(1, 400, 30))]).astype(np.int64)
    seg = seg[seg < 4000]; seg = np.concatenate([seg, [4000]])
    return q, t, seg


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import scipy.sparse as sp
    from oracle import ba_oracle as bo
    from oracle import knn as ora
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # ---- matching: shard candidate keyframes ----
    desc, _ = synth.orb_keyframes(seed=4, n_kf=9, n_feat=200, n_lm=400, window=400)
    q, cands = desc[0], desc[1:]
    mine = [i for i in range(len(cands)) if i % world == rank]
    t = cands[mine].reshape(-1, 32); seg = synth.seg_ptr_uniform(len(mine), 200)
    i2, d2 = ora.knn_hamming_batch(q, t, seg, 2)
    _, _, cnt = ora.ratio_filter(i2, d2.astype(np.float32), 40.0, 0.8)
    full = torch.zeros(len(cands), dtype=torch.int64); full[mine] = torch.from_numpy(cnt.astype(np.int64))
    dist.all_reduce(full)
    # ---- map-wide k-NN: database rows sharded by keyframe block, one all-gather, merge ----
    mq, mt, mseg = _mapwide_case()
    cuts = M.shard_rows(len(mt), world, mseg)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    li, ld = ora.knn_hamming(mq, mt[lo:hi], 3)
    gi = [torch.empty(li.shape, dtype=torch.int32) for _ in range(world)]
    gd = [torch.empty(ld.shape, dtype=torch.int32) for _ in range(world)]
    go = [torch.empty(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gi, torch.from_numpy(li)); dist.all_gather(gd, torch.from_numpy(ld))
    dist.all_gather(go, torch.tensor([lo]))
    mw_i, mw_d = ora.merge_shards(torch.stack(gi).numpy(), torch.stack(gd).numpy(), torch.cat(go).numpy(), 3)
    # ---- GBA: landmark-block sharding ----
    p = synth_map.make_config("tiny")
    whole = bo.Problem(p, visual_only=True, loop_loss=1.0)
    owner = O.lm_owner_of_rank(int(whole.lm_in.sum()), world)
    lm_owner = np.full(p["L"], -1); lm_owner[np.flatnonzero(whole.lm_in)] = owner
    obs_lm = np.repeat(np.arange(p["L"]), np.diff(p["lm_obs_ptr"]))
    use = lm_owner[obs_lm] == rank
    part = bo.Problem(p, visual_only=True, loop_loss=1.0, use_obs=use)
    if part.edges is not None:   # factors are dealt round-robin
        keep = np.arange(len(part.edges["i"])) % world == rank
        for k in ("i", "j", "robust"):
            part.edges[k] = part.edges[k][keep]
        for k in ("q", "t", "S"):
            part.edges[k] = part.edges[k][torch.from_numpy(keep)]
    _, r, J, _ = part.evaluate(part.pose, part.sb, part.lm)
    nc = part.ncam
    H = (J.T @ J).tocsr(); g = J.T @ r
    lmk = nc + np.flatnonzero(np.asarray(abs(J[:, nc:]).sum(0)).reshape(-1) > 0)
    Hcc = H[:nc][:, :nc]; W = H[:nc][:, lmk]; Hll = H[lmk][:, lmk] + 1e-9 * sp.eye(len(lmk))
    B = sp.bsr_matrix(Hll, blocksize=(3, 3)); B.sort_indices()
    Hinv = sp.bsr_matrix((np.linalg.inv(B.data), B.indices, B.indptr), shape=Hll.shape).tocsr()
    S = torch.from_numpy((Hcc - W @ Hinv @ W.T).toarray()); gs = torch.from_numpy(g[:nc] - W @ (Hinv @ g[lmk]))
    dist.all_reduce(S); dist.all_reduce(gs)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), counts=full.numpy(), S=S.numpy(), gs=gs.numpy(), lm_owner=lm_owner,
                 mw_i=mw_i, mw_d=mw_d, cuts=cuts)
    dist.destroy_process_group()


def test_world2_gloo_sharding(tmp_path):
    import scipy.sparse as sp
    from oracle import ba_oracle as bo
    from oracle import knn as ora
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npz")
    # matching reference, single process
    desc, _ = synth.orb_keyframes(seed=4, n_kf=9, n_feat=200, n_lm=400, window=400)
    i2, d2 = ora.knn_hamming_batch(desc[0], desc[1:].reshape(-1, 32), synth.seg_ptr_uniform(8, 200), 2)
    _, _, cnt = ora.ratio_filter(i2, d2.astype(np.float32), 40.0, 0.8)
    assert np.array_equal(out["counts"], cnt)
    # map-wide k-NN reference: one k-NN over the whole database; cuts on keyframe boundaries
    mq, mt, mseg = _mapwide_case()
    ri, rd = ora.knn_hamming(mq, mt, 3)
    assert np.array_equal(out["mw_i"], ri) and np.array_equal(out["mw_d"], rd)
    assert set(out["cuts"].tolist()) <= set(mseg.tolist()) and out["cuts"][0] == 0 and out["cuts"][-1] == len(mt)
    # GBA reference: reduced camera system of the whole problem
    p = synth_map.make_config("tiny")
    whole = bo.Problem(p, visual_only=True, loop_loss=1.0)
    _, r, J, _ = whole.evaluate(whole.pose, whole.sb, whole.lm)
    nc = whole.ncam
    H = (J.T @ J).tocsr(); g = J.T @ r
    lmk = np.arange(nc, whole.n)
    Hll = H[lmk][:, lmk] + 1e-9 * sp.eye(len(lmk))
    B = sp.bsr_matrix(Hll, blocksize=(3, 3)); B.sort_indices()
    Hinv = sp.bsr_matrix((np.linalg.inv(B.data), B.indices, B.indptr), shape=Hll.shape).tocsr()
    W = H[:nc][:, lmk]
    S = (H[:nc][:, :nc] - W @ Hinv @ W.T).toarray(); gs = g[:nc] - W @ (Hinv @ g[lmk])
    assert np.allclose(out["S"], S, rtol=1e-9, atol=1e-9 * np.abs(S).max())
    assert np.allclose(out["gs"], gs, rtol=1e-9, atol=1e-9 * np.abs(gs).max())
    # merge rule
    owner = out["lm_owner"]
    res = [dict(pose=np.zeros(1), lm=np.full((p["L"], 3), float(rk)), lm_owner=owner) for rk in range(world)]
    merged = O.merge_sharded_landmarks(res)
    assert np.array_equal(merged["lm"][owner == 1], np.ones(((owner == 1).sum(), 3)))
    assert np.array_equal(merged["lm"][owner <= 0], np.zeros(((owner <= 0).sum(), 3)))
