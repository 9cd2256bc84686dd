mkdir -p gpurun_out/r3i
export ITERS=4 HANG_S=40
timeout 80 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -1
MODE=trained timeout 80 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -1
timeout 120 python scripts/stash_stats.py 2>&1 | grep -v Warn | tee gpurun_out/r3i/stash_stats.txt
REPS=3 timeout 80 python scripts/time_student.py 2>&1 | grep graph_step | tee gpurun_out/r3i/time800.txt
for tf in 8 16; do echo tf=$tf; ERD_TEACHER_FREE_SMS=$tf timeout 80 python scripts/time_student.py 2>&1 | grep graph_step; done
timeout 900 python -m pytest tests -m gpu -x -q --timeout=400 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/r3i/pytest.txt
