set -x
timeout 600 python -m pytest tests -m gpu -q -x -k "attention or lora or train or unet" > gpurun_out/r2c10_tests.log 2>&1; echo tests=$?; tail -3 gpurun_out/r2c10_tests.log | cut -c1-200
timeout 100 python tools/check_ops2.py attn_perf2 2>&1 | tail -5
CLB_ATTN_POLY_EXP=0 timeout 100 python tools/check_ops2.py attn_perf2 2>&1 | tail -5
timeout 300 python tools/bw_bench.py --iters 5 2>&1 | grep -i "skinny" | cut -c1-110
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2c10_bench.log 2>&1; tail -1 gpurun_out/r2c10_bench.log | cut -c1-300
CLB_ATTN_POLY_EXP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2c10_bench_nopoly.log 2>&1; tail -1 gpurun_out/r2c10_bench_nopoly.log | cut -c1-300
