mkdir -p gpurun_out/r2l
timeout 300 python -m pytest tests -m gpu -x -q --timeout=60 2>&1 | tail -12 | tee gpurun_out/r2l/pytest.txt
VARIANTS=0:5,0:4 timeout 100 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2l/variants.txt
HW=1024x1024 VARIANTS=0:5,0:4,1:5 timeout 100 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2l/variants1024.txt
HW=1024x1024 timeout 100 python scripts/trace_student.py 2>&1 | grep -v Warn > gpurun_out/r2l/trace_1024.txt
HW=1024x1024 timeout 100 python scripts/trace_teacher.py 2>&1 | grep -v Warn > gpurun_out/r2l/ttrace_1024.txt
