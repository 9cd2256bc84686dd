mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests -m gpu -x -q --timeout=400 2>&1 | grep -v Warn | tail -5 | tee gpurun_out/r2y/pytest.txt
timeout 120 python scripts/profile_api.py 2>&1 | grep -v Warn | head -3 | tee gpurun_out/r2y/profile_api.txt
bash scripts/sanitize.sh gpurun_out/r2y/sanitizer
