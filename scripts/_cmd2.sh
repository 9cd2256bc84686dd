mkdir -p gpurun_out/r2o
export ITERS=6 HANG_S=40
run() { tag=$1; shift; echo "== $tag"; env "$@" timeout 80 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -${TAILN:-3} | tee gpurun_out/r2o/$tag.txt; }
run gauss
run trained MODE=trained
run g1024 HW=1024x1024
timeout 400 python -m pytest tests -m gpu -x -q --timeout=120 2>&1 | tail -12 | tee gpurun_out/r2o/pytest.txt
timeout 80 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2o/time800.txt
HW=1024x1024 timeout 80 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2o/time1024.txt
MODE=trained timeout 80 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2o/time800_trained.txt
