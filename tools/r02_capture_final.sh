#!/bin/bash
# Round-2 final evidence (ONE GPU, under gpurun): GPU test-suite, bench line, ncu launch list of the bench command, ncu --set full
# of the two dominant kernels.  Everything under gpurun_out/.
O=gpurun_out; mkdir -p $O
timeout 420 python -m pytest tests -x -q -m gpu --timeout 200 > $O/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_final.log
timeout 420 python bench.py --steps 10 --warmup 3 > $O/r02_bench_final.json 2> $O/r02_bench_final.err
COVINS_SKIP_CPU_BASELINE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_c3.csv \
  python bench.py --steps 2 --warmup 3 > $O/r02_bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_xt_kernel -s 2 -c 1 -o $O/r02_prof_tc_xt -f python tools/tc_profile.py > $O/r02_ncu_tc_xt.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 400 -c 1 -o $O/r02_prof_syrk -f python tools/ba_one_iter.py C3 > $O/r02_ncu_syrk.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chain_gemm_kernel -s 100 -c 2 -o $O/r02_prof_chain -f python tools/ba_one_iter.py C3 > $O/r02_ncu_chain.log 2>&1
gzip -f $O/r02_launches_c3.csv
tail -2 $O/r02_pytest_final.log; tail -c 400 $O/r02_bench_final.json; ls -la $O/*.ncu-rep | tail -4
