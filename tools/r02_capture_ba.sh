#!/bin/bash
# Round-2: ncu --set full of the HBM-bound BA kernels (K4/K5/K7/K9) at C3 — one GBA iteration (tools/ba_one_iter.py).
set -u
mkdir -p gpurun_out
O=gpurun_out
for k in lin_obs_kernel lm_reduce_kernel obs_Y_kernel kf_visual_kernel schur_kernel lin_imu_kernel backsub_kernel imu_repropagate_kernel jv_obs_kernel plus_kernel lm_fused_kernel lin_fused_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 1 -c 1 \
    -o $O/r02_ba_${k} -f python tools/ba_one_iter.py ${1:-C3} > $O/r02_ncu_${k}.log 2>&1
done
echo done
