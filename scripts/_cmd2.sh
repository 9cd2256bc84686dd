mkdir -p gpurun_out/r3o
export ITERS=4 HANG_S=40
timeout 80 python scripts/_dbg.py 2>&1 | grep -v Warn | tail -1
timeout 900 python -m pytest tests -m gpu -x -q --timeout=400 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/r3o/pytest.txt
REPS=2 timeout 80 python scripts/time_student.py 2>&1 | grep graph_step | tee gpurun_out/r3o/time800.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 2 --steps 200 --warmup 5 --no-e2e 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 steps=200', round(d['ms_per_step'],4), 'eager', round(d['launch_mode']['ms_per_step_eager'],4), d['multi_rank_check']['avg_factors_equal_mean_of_ranks_bitwise'])"
