#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 120 -c 1 -o $O/r02_prof_syrk -f python tools/ba_one_iter.py C3 > $O/r02_ncu_syrk.log 2>&1
timeout 560 python tools/c5_run.py 2 > $O/r02_c5.json 2> $O/r02_c5.err
tail -2 $O/r02_smoke.log; tail -c 1500 $O/r02_c5.json; tail -3 $O/r02_c5.err
