#!/bin/bash
# One gpurun call: smoke, GPU test-suite, bench, ncu launch list.  Usage: scripts/gpu_round.sh <tag>
tag=${1:-dev}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $out/smoke.txt
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest.txt
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json; tail -5 $out/bench.err
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 60 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_bench.log 2>&1
python scripts/ncu_summary.py $out/launches.csv 2>&1 | tail -30
