#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_tc_match.py tests/test_gpu_matching.py -x -q -m gpu --timeout 60 > $O/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest8.log
timeout 200 python tools/c5_match.py > $O/r02_c5_match.log 2>&1
timeout 60 python tools/c5_match.py 2000 >> $O/r02_c5_match.log 2>&1
tail -2 $O/r02_pytest8.log; cat $O/r02_c5_match.log | tail -8
