#!/bin/bash
# chain stream on one GPU: parity tests + C3 timing with / without it
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_ba.py -x -q -m gpu --timeout 120 -k "c1_matches or c3_matches or tiny or pgo_c2_matches" > $O/r02_pytest7.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest7.log
tail -2 $O/r02_pytest7.log
python - <<'PY'
import os, sys, time
sys.path.insert(0, ".")
import covins_b200
from covins_b200 import optimization as O, synth_map
ctx = covins_b200.Context(0)
p = synth_map.make_config("C3")
for chain, ng in (("1", ""), ("1", "1"), ("0", ""), ("1", "")):
    os.environ["COVINS_B200_CHAIN_STREAM"] = chain
    os.environ.pop("COVINS_B200_NO_GROUP_CHAIN", None)
    if ng: os.environ["COVINS_B200_NO_GROUP_CHAIN"] = "1"
    s = O.BaSolver(ctx, p); s.iterate(2); ctx.sync(); s.restart(); s.timing(reset=True); ctx.sync()
    t0 = time.perf_counter(); n = s.iterate(10); ctx.sync(); dt = time.perf_counter() - t0
    r = s.result(); tm = s.timing(); s.close(); it = max(r["iterations"], 1)
    print(f"C3 chain_stream={chain} no_group_chain={ng or 0}: {1e3*dt/n:.2f} ms/it factor {tm['factor_ms']/it:.2f} solve {tm['solve_ms']/it:.2f} schur {tm['build_schur_ms']/it:.2f} final {r['final_cost']:.6f}", flush=True)
PY
