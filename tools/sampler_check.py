"""Development aid: bench.ClockSampler against the real NVML around ~0.3 s of GPU work."""
import importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
import torch
x = torch.randn(8192, 8192, device="cuda", dtype=torch.float16)
torch.cuda.synchronize()
for label, dur in (("0.3 s region", 0.3), ("10 ms region", 0.01)):
    with b.ClockSampler(0) as c:
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < dur:
            y = x @ x
        torch.cuda.synchronize()
    print(label, c.summary(), "samples", len(c.samples))
