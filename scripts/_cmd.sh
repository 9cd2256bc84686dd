mkdir -p gpurun_out/r2j
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2j/pytest.txt
VARIANTS=0:5,15:5 timeout 300 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2j/variants.txt
HW=1024x1024 VARIANTS=0:5,15:5 timeout 300 python scripts/time_student.py 2>&1 | grep -v Warn | tee gpurun_out/r2j/variants1024.txt
HW=1024x1024 timeout 300 python scripts/trace_student.py 2>&1 | grep -v Warn > gpurun_out/r2j/trace_1024.txt
