"""Extracts dram bytes per launch of the dominant kernels from committed ncu --set full captures (.ncu-rep under gpurun_out/)
and writes profiles/r02_traffic.json, which bench.py reads for `roofline.traffic` (so the number in the JSON line is tied
to a capture, not pasted).  Usage: python tools/r02_traffic.py"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAPS = {"syrk_kernel": ["gpurun_out/prof_syrk_final.ncu-rep"],   # a BULK launch (the round-2 capture r02_prof_syrk hit a 12 us chain-phase launch)
        "tc_scan_kernel_hamming": ["gpurun_out/r02_prof_tc.ncu-rep", "gpurun_out/prof_tc_final.ncu-rep"],
        "tc_xt_kernel": ["gpurun_out/r02_prof_tc_xt.ncu-rep"]}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for key, cands in CAPS.items():
    rep = next((c for c in cands if os.path.exists(os.path.join(ROOT, c))), None)
    if not rep:
        continue
    txt = subprocess.check_output(["ncu", "-i", os.path.join(ROOT, rep), "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, first = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(first[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
    out[key] = {"dram_bytes": tot, "duration_us": float(first[col["gpu__time_duration.sum"]].replace(",", "")), "kernel": first[col["Kernel Name"]][:120],
                "source": f"ncu --set full capture {os.path.basename(rep)} (dram__bytes_read.sum + dram__bytes_write.sum of the captured launch)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
