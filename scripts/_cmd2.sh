mkdir -p gpurun_out/r2n
export ERD_B200_NO_BUILD=1 ITERS=8 HANG_S=25
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 60 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -${TAILN:-4} | tee gpurun_out/r2n/$tag.txt; }
A=$PWD/build_ab/libA.so; B=$PWD/build_ab/libB.so
run A1_gauss ERD_B200_LIB=$A
run A2_notma ERD_B200_LIB=$A ERD_STUDENT_TMA=0
run A3_nosparse ERD_B200_LIB=$A ERD_STUDENT_DEV=1
TAILN=14 run A4_blocking ERD_B200_LIB=$A CUDA_LAUNCH_BLOCKING=1
run A5_trained ERD_B200_LIB=$A MODE=trained
run A6_1024 ERD_B200_LIB=$A HW=1024x1024
run B1_gauss ERD_B200_LIB=$B
run B2_trained ERD_B200_LIB=$B MODE=trained
run B3_1024 ERD_B200_LIB=$B HW=1024x1024
ERD_B200_LIB=$B VARIANTS=0:4 timeout 60 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2n/B_time800.txt
ERD_B200_LIB=$B HW=1024x1024 VARIANTS=0:4 timeout 60 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2n/B_time1024.txt
ERD_B200_LIB=$A HW=1024x1024 VARIANTS=0:5 timeout 60 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2n/A_time1024.txt
