#!/bin/bash
# Developer experiment: what bounds the fused teacher head?  Builds an A/B library with -DERD_HEAD_EXP and times the
# kernel with parts switched off (results are garbage in those runs).  Usage (on the GPU box): scripts/exp_head.sh <outdir>
out=${1:-gpurun_out/exp_head}
mkdir -p $out
for m in ${MODES:-0 1 2 3 4 8 5 6}; do
  echo "== ERD_HEAD_EXP=$m (1: no A fill, 2: no B copies, 4: no reg MMAs, 8: no cls MMAs, 64: no MMAs, 128: aligned core matrices)"
  ERD_HEAD_EXP=$m ERD_B200_LIB=$PWD/erd_b200/lib/liberd_b200_exp.so ERD_B200_NO_BUILD=1 timeout 120 python scripts/time_teacher_head.py --iters 10 --kernel-only 2>&1 | tail -1 | tee $out/exp_$m.txt
done
