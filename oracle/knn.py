"""ctypes binding of oracle/knn_oracle.c (CPU restatement of the matching stage) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcovins_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ora_ratio_filter.restype = C.c_int
        _LIB.ora_landmark_match.restype = C.c_int
        _LIB.ora_hamming256.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def knn_hamming_batch(q, t, seg_ptr, k=2, threads=0):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
    nq, nb = q.shape; ns = len(seg_ptr) - 1
    idx = np.empty((ns, nq, k), np.int32); dist = np.empty((ns, nq, k), np.int32)
    lib().ora_knn_hamming_batch(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), _p(seg_ptr, C.c_int32), ns, nb, k,
                                _p(idx, C.c_int32), _p(dist, C.c_int32), threads)
    return idx, dist


def knn_hamming(q, t, k=2, threads=0):
    idx, dist = knn_hamming_batch(q, t, np.array([0, len(t)], np.int32), k, threads)
    return idx[0], dist[0]


def knn_l2_batch(q, t, seg_ptr, k=2, threads=0):
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
    nq, dim = q.shape; ns = len(seg_ptr) - 1
    idx = np.empty((ns, nq, k), np.int32); dist = np.empty((ns, nq, k), np.float32)
    lib().ora_knn_l2_batch(_p(q, C.c_float), nq, _p(t, C.c_float), _p(seg_ptr, C.c_int32), ns, dim, k,
                           _p(idx, C.c_int32), _p(dist, C.c_float), threads)
    return idx, dist


def knn_l2(q, t, k=2, threads=0):
    idx, dist = knn_l2_batch(q, t, np.array([0, len(t)], np.int32), k, threads)
    return idx[0], dist[0]


def ratio_filter(idx2, dist2, thr, ratio):
    """idx2,dist2: [..., n, 2] → (match_train [..., n], match_dist [..., n], count [...])."""
    idx2 = np.ascontiguousarray(idx2, np.int32); dist2 = np.ascontiguousarray(dist2, np.float32)
    lead = idx2.shape[:-2]; n = idx2.shape[-2]
    i2 = idx2.reshape(-1, n, 2); d2 = dist2.reshape(-1, n, 2)
    mt = np.empty((i2.shape[0], n), np.int32); md = np.empty((i2.shape[0], n), np.float32)
    cnt = np.empty(i2.shape[0], np.int32)
    for s in range(i2.shape[0]):
        cnt[s] = lib().ora_ratio_filter(_p(i2[s], C.c_int32), _p(d2[s], C.c_float), n, C.c_float(thr),
                                        C.c_float(ratio), _p(mt[s], C.c_int32), _p(md[s], C.c_float))
    return mt.reshape(*lead, n), md.reshape(*lead, n), cnt.reshape(lead)


def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().ora_hamming256(_p(a, C.c_uint8), _p(b, C.c_uint8))


def landmark_match(A, skipA, B, skipB, thr=50.0, num_best=4, return_lists=False):
    A = np.ascontiguousarray(A, np.uint8); B = np.ascontiguousarray(B, np.uint8)
    nA, nB = len(A), len(B)
    sA = np.ascontiguousarray(skipA, np.uint8) if skipA is not None else None
    sB = np.ascontiguousarray(skipB, np.uint8) if skipB is not None else None
    oA = np.empty(max(nB, 1), np.int32); oB = np.empty(max(nB, 1), np.int32); oD = np.empty(max(nB, 1), np.float32)
    bi = np.empty((nA, num_best), np.int32); bd = np.empty((nA, num_best), np.float32)
    n = lib().ora_landmark_match(_p(A, C.c_uint8), _p(sA, C.c_uint8), nA, _p(B, C.c_uint8), _p(sB, C.c_uint8), nB,
                                 C.c_float(thr), num_best, _p(oA, C.c_int32), _p(oB, C.c_int32), _p(oD, C.c_float),
                                 _p(bi, C.c_int32), _p(bd, C.c_float))
    if return_lists:
        return oA[:n].copy(), oB[:n].copy(), oD[:n].copy(), bi, bd
    return oA[:n].copy(), oB[:n].copy(), oD[:n].copy()


def landmark_match_batch(A, skipA, B, skipB, seg_ptr, thr=50.0, num_best=4, threads=0):
    A = np.ascontiguousarray(A, np.uint8); B = np.ascontiguousarray(B, np.uint8)
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
    sA = np.ascontiguousarray(skipA, np.uint8) if skipA is not None else None
    sB = np.ascontiguousarray(skipB, np.uint8) if skipB is not None else None
    nB = len(B); ns = len(seg_ptr) - 1
    oA = np.full(nB, -1, np.int32); oB = np.full(nB, -1, np.int32); oD = np.zeros(nB, np.float32)
    n = np.zeros(ns, np.int32)
    lib().ora_landmark_match_batch(_p(A, C.c_uint8), _p(sA, C.c_uint8), len(A), _p(B, C.c_uint8), _p(sB, C.c_uint8),
                                   _p(seg_ptr, C.c_int32), ns, C.c_float(thr), num_best, _p(oA, C.c_int32),
                                   _p(oB, C.c_int32), _p(oD, C.c_float), _p(n, C.c_int32), threads)
    return oA, oB, oD, n


def merge_shards(idx_all, dist_all, row_offset, k):
    """CPU restatement of the map-wide sharded k-NN merge: per-shard lists [G, n, k] with shard-local
    trainIdx → the k smallest by (distance, global trainIdx).  Missing entries have idx -1."""
    idx_all = np.asarray(idx_all); dist_all = np.asarray(dist_all)
    G, n, kk = idx_all.shape
    gi = np.where(idx_all >= 0, idx_all.astype(np.int64) + np.asarray(row_offset, np.int64)[:, None, None], -1)
    gi = gi.transpose(1, 0, 2).reshape(n, G * kk); gd = dist_all.transpose(1, 0, 2).reshape(n, G * kk)
    big = np.iinfo(np.int64).max
    key_i = np.where(gi >= 0, gi, big)
    key_d = np.where(gi >= 0, gd.astype(np.float64), np.inf)
    order = np.lexsort((key_i, key_d), axis=1)[:, :k]
    oi = np.take_along_axis(gi, order, 1); od = np.take_along_axis(gd, order, 1)
    empty = np.iinfo(np.int32).max if np.issubdtype(dist_all.dtype, np.integer) else np.finfo(np.float32).max
    od = np.where(oi >= 0, od, empty).astype(dist_all.dtype)
    return oi.astype(np.int32), od


def landmark_descriptor(cand, lm_ptr):
    """Landmark::ComputeDescriptor batched (landmark_be.cpp:49-92) → (best_idx [n_lm] i32, desc [n_lm, 32] u8)."""
    cand = np.ascontiguousarray(cand, np.uint8).reshape(-1, 32); lm_ptr = np.ascontiguousarray(lm_ptr, np.int32)
    n = len(lm_ptr) - 1
    best = np.full(n, -1, np.int32); out = np.zeros((n, 32), np.uint8)
    lib().ora_landmark_descriptor(_p(cand, C.c_uint8), _p(lm_ptr, C.c_int32), n, _p(best, C.c_int32), _p(out, C.c_uint8))
    return best, out
