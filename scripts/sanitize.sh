#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck over smoke() (one small step of every kernel); logs -> $1
out=${1:-gpurun_out/sanitizer}
mkdir -p $out
filt() { grep -v "^=========     Host Frame\|^=========         in \|libtorch\|libc10\|libcuda\|python3\|^=========     Saved host\|UserWarning\|Consider using\|return dict" ; }
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | filt > $out/$tool.txt
  echo "== $tool: $(grep -c 'smoke ok' $out/$tool.txt) smoke ok, $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $out/$tool.txt | tail -1)"
done
