#!/bin/bash
# Round-1 evidence capture (run under gpurun on ONE GPU).  Writes everything under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err
timeout 600 python bench.py --impl reference > $O/bench_ref.log 2> $O/bench_ref.err
python tools/potrf_bench.py > $O/potrf_bench.log 2>&1
# launch list of the same bench command (smaller GBA config so the serialised replay stays short)
COVINS_SKIP_CPU_BASELINE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches_final.csv python bench.py --steps 2 --warmup 3 --gba-config C2 > $O/ncu_bench.log 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 1 -c 1 \
  -o $O/prof_tc_final python tools/tc_profile.py > $O/ncu_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:potrf_inv_kernel -s 200 -c 1 \
  -o $O/prof_potrf_final python tools/ba_one_iter.py C3 > $O/ncu_potrf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 148 -c 1 \
  -o $O/prof_syrk_final python tools/ba_one_iter.py C3 > $O/ncu_syrk.log 2>&1
echo done
