mkdir -p gpurun_out/r2x
timeout 120 python scripts/profile_api.py 2>&1 | grep -v Warn | cut -c1-150 | tee gpurun_out/r2x/profile_api.txt | head -60
timeout 600 compute-sanitizer --tool racecheck --print-limit 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x/racecheck_full.txt 2>&1
grep "SUMMARY\|smoke ok" gpurun_out/r2x/racecheck_full.txt
