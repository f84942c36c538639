"""Development aid (CPU only): fill of the reduced camera system's POSE block under different elimination orders.
Builds the keyframe covisibility graph of a synthetic config (two keyframes are adjacent when they observe a common
landmark, share an IMU factor or a loop edge — the block structure of S's pose part), eliminates it symbolically in (a) the
engine's chain order (keyframes agent by agent, in time order) and (b) a greedy minimum-degree order, and reports nnz(L) in
6x6 blocks and the factorisation flops sum_k (6 d_k)^2 * 6  (d_k = number of later neighbours at elimination).
Answers VERDICT r1 item 4: is "dense tiles in chain order" the right shape for the pose part?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from covins_b200 import synth_map

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
p = synth_map.make_config(name)
K = p["K"]
adj = [set() for _ in range(K)]
ptr = p["lm_obs_ptr"]; okf = p["obs_kf"]
for l in range(p["L"]):
    ks = okf[ptr[l]:ptr[l + 1]]
    if len(ks) < 2:
        continue
    for a in ks:
        adj[a].update(ks.tolist())
for i, j in zip(p["imu_i"], p["imu_j"]):
    adj[i].add(j); adj[j].add(i)
for i, j in zip(p["loop_i"], p["loop_j"]):
    adj[i].add(j); adj[j].add(i)
for k in range(K):
    adj[k].discard(k)
nnz0 = sum(len(a) for a in adj) // 2 + K
print(f"{name}: K={K}, pose-block structure of S: {nnz0} blocks of {K*(K+1)//2} ({100*nnz0/(K*(K+1)/2):.1f} % of the lower triangle)")


def eliminate(order):
    g = [set(a) for a in adj]
    pos = np.empty(K, np.int64); pos[order] = np.arange(K)
    nnz = 0; flops = 0.0
    done = np.zeros(K, bool)
    for k in order:
        nb = [v for v in g[k] if not done[v]]
        d = len(nb)
        nnz += d + 1
        flops += (6.0 * (d + 1)) ** 2 * 6.0       # column of 6 dofs: rank-6 update of a (6(d+1))^2 front, 2 flop per MAC /2 (symmetric) ~
        for a in nb:
            g[a].update(nb); g[a].discard(a); g[a].discard(k)
        done[k] = True
    return nnz, flops


def min_degree():
    g = [set(a) for a in adj]
    alive = np.ones(K, bool)
    deg = np.array([len(a) for a in g])
    order = []
    nnz = 0; flops = 0.0
    for _ in range(K):
        cand = np.flatnonzero(alive)
        k = cand[np.argmin(deg[cand])]
        nb = list(g[k])
        d = len(nb)
        nnz += d + 1; flops += (6.0 * (d + 1)) ** 2 * 6.0
        for a in nb:
            g[a].update(nb); g[a].discard(a); g[a].discard(k)
            deg[a] = len(g[a])
        alive[k] = False; g[k] = set()
        order.append(k)
    return np.array(order), nnz, flops


t0 = time.time()
nz_c, fl_c = eliminate(np.arange(K))
print(f"chain order (agent by agent, time order): nnz(L) = {nz_c} blocks ({100*nz_c/(K*(K+1)/2):.1f} % of dense), flops = {fl_c/1e12:.3f} TFLOP   [{time.time()-t0:.0f} s]")
t0 = time.time()
order, nz_m, fl_m = min_degree()
print(f"greedy minimum degree:                    nnz(L) = {nz_m} blocks ({100*nz_m/(K*(K+1)/2):.1f} % of dense), flops = {fl_m/1e12:.3f} TFLOP   [{time.time()-t0:.0f} s]")
print(f"dense pose block: nnz = {K*(K+1)//2}, flops = {(6.0*K)**3/3/1e12:.3f} TFLOP")


def tile_flops(order, T=128):
    """tile-level symbolic elimination of the pose block laid out in `order` (6 columns per keyframe, 128-wide tiles)"""
    pos = np.empty(K, np.int64); pos[order] = np.arange(K)
    nt = (6 * K + T - 1) // T
    mask = np.zeros((nt, nt), bool)
    tiles_of = [sorted({(6 * pos[k]) // T, (6 * pos[k] + 5) // T}) for k in range(K)]
    for k in range(K):
        for a in list(adj[k]) + [k]:
            for ta in tiles_of[k]:
                for tb in tiles_of[a]:
                    mask[max(ta, tb), min(ta, tb)] = True
    pre = int(np.tril(mask).sum())
    gemms = 0
    for c in range(nt):
        rows = np.flatnonzero(mask[c + 1:, c]) + c + 1
        m = len(rows)
        gemms += m + m * (m + 1) // 2
        if m:
            mask[np.ix_(rows, rows)] |= np.tril(np.ones((m, m), bool))
    return pre, int(np.tril(mask).sum()), gemms * 2.0 * T ** 3


for nm, od in (("chain order", np.arange(K)), ("minimum degree", order)):
    pre, post, fl = tile_flops(od)
    print(f"tile level (128), {nm}: tiles before fill {pre}, after fill {post} of {((6*K+127)//128)*((6*K+127)//128+1)//2}, tile-GEMM flops {fl/1e12:.3f} TFLOP")
