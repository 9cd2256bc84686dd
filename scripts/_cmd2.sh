mkdir -p gpurun_out/r3b
for tf in 20 24 28 36; do for sf in 20 24 28; do echo "teacher_free=$tf student_free=$sf"; ERD_TEACHER_FREE_SMS=$tf ERD_STUDENT_FREE_SMS=$sf REPS=2 timeout 80 python scripts/time_student.py 2>&1 | grep graph_step | cut -c1-120; done; done | tee gpurun_out/r3b/sweep.txt
SHIFT=0.5 timeout 200 python scripts/time_predict.py 2>&1 | grep -v Warn | tee gpurun_out/r3b/time_predict_few.txt
