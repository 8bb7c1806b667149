set -x; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compaction or fuzz or mixed_convergence" 2>&1 | tail -5 > gpurun_out/t_compact.log
for e in 0.006 0.008; do for fr in 16384 65536; do
python bench.py --frames $fr --eps $e --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('compact', d['config']['frames_per_gpu'], d['config']['eps'], d['ms_per_step'], d['frame_iters_per_s'], d['value'])" >> gpurun_out/ab_compact.log
DNALDPC_NO_COMPACT=1 python bench.py --frames $fr --eps $e --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('nocompact', d['config']['frames_per_gpu'], d['config']['eps'], d['ms_per_step'], d['frame_iters_per_s'], d['value'])" >> gpurun_out/ab_compact.log
done; done
echo done
