#!/bin/bash
# compute-sanitizer over the fused teacher head's smallest parity case (kernel + the step behind it); logs -> $1
out=${1:-gpurun_out/sanitizer_head}
mkdir -p $out
filt() { grep -v "^=========     Host Frame\|^=========         in \|libtorch\|libc10\|libcuda\|python3\|^=========     Saved host\|UserWarning\|Consider using\|return dict" ; }
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python -m pytest tests/test_gpu_teacher_head.py -x -q -k "test_fused_teacher_head" 2>&1 | filt > $out/$tool.txt
  echo "== $tool: $(grep -c ' passed' $out/$tool.txt) passed line(s); $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $out/$tool.txt | tail -1)"
done
