mkdir -p gpurun_out/r3f
export ITERS=4 HANG_S=40
timeout 80 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -1
timeout 80 python scripts/stash_stats.py 2>&1 | grep -v Warn | tail -1
REPS=3 timeout 80 python scripts/time_student.py 2>&1 | grep graph_step | tee gpurun_out/r3f/time800.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:teacher_pass -s 3 -c 1 -o gpurun_out/r3f/teacher -f python scripts/time_student.py > gpurun_out/r3f/ncu.log 2>&1
