#!/bin/bash
# compute-sanitizer over smoke() (one small step through every kernel); logs -> $1
#   memcheck, synccheck, racecheck on the production library;
#   racecheck again on a build with -DERD_RACECHECK, in which the student pass's IO warps complete their
#   cp.async copies with cp.async.wait_all instead of cp.async.mbarrier.arrive: racecheck follows cp.async
#   only through wait_group, so on the production build it reports every consumer access to a slot the IO
#   warp filled with cp.async as a hazard although the slot's full barrier orders them.
out=${1:-gpurun_out/sanitizer}
mkdir -p $out build_ab
filt() { grep -v "^=========     Host Frame\|^=========         in \|libtorch\|libc10\|libcuda\|python3\|^=========     Saved host\|UserWarning\|Consider using\|return dict" ; }
run() { tool=$1; tag=$2; shift 2; env "$@" timeout 600 compute-sanitizer --tool $tool --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | filt > $out/$tag.txt
  echo "== $tag: $(grep -c 'smoke ok' $out/$tag.txt) smoke ok; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $out/$tag.txt | tail -1)"; }
run memcheck memcheck
run synccheck synccheck
run racecheck racecheck_production
ERD_B200_LIB=$PWD/build_ab/liberd_racecheck.so ERD_EXTRA_NVCC=-DERD_RACECHECK python erd_b200/build.py > /dev/null
run racecheck racecheck_waitall_build ERD_B200_LIB=$PWD/build_ab/liberd_racecheck.so ERD_B200_NO_BUILD=1
