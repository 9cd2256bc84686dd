mkdir -p gpurun_out/r2z
timeout 600 python -m pytest tests/test_gpu_predict.py -x -q --timeout=200 2>&1 | grep -v Warn | tail -25 | tee gpurun_out/r2z/pytest_predict.txt
