#!/bin/bash
# Developer diagnostic: marginal cost of each bulk kernel inside the full schedule.
# Build with the ablation hook first:  ERD_EXTRA_NVCC=-DERD_DEV_ABLATE python erd_b200/build.py
# ids: zero_fill 10, qfl_sweep 11, cls_kd 12, pos_grad 13, box_kd 14, box_fix 15, nms_order 8
for a in 0 0x400 0x800 0x1000 0x4000 0x5000 0x5C00; do
  echo "ERD_ABLATE=$a"
  ERD_ABLATE=$a timeout 100 python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
