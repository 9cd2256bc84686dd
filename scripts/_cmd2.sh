mkdir -p gpurun_out/r2l
timeout 280 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l/memcheck.txt 2>&1
grep -v "^=========     Host Frame\|^=========         in \|libtorch\|libc10\|python" gpurun_out/r2l/memcheck.txt | head -60
