#!/bin/bash
# Round-end validation on ONE GPU: the driver's own sequence (gpu tests, smoke, both bench arms).
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1
timeout 600 python bench.py --impl reference > $O/bench_ref.log 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err
echo done
