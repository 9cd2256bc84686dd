mkdir -p gpurun_out/r2v
export ITERS=4 HANG_S=40
timeout 80 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -1
for sf in 0 24; do for tf in 0 8; do
echo "student_free=$sf teacher_free=$tf"
ERD_STUDENT_FREE_SMS=$sf ERD_TEACHER_FREE_SMS=$tf timeout 80 python scripts/time_student.py 2>&1 | grep graph_step
done; done | tee gpurun_out/r2v/sweep.txt
