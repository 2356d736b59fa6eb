set -x
export CLB_LIB=$PWD/controllora_b200/libcontrollora_b200_tl.so
for bn in 64 256; do echo "=== bn=$bn K=1280 EPI=4"; CLB_GEMM_EPI=4 timeout 100 python tools/gemm_timeline.py 4096 1280 1280 $bn 2>&1 | head -120; done > gpurun_out/r2c6_kb_timeline.log 2>&1
echo "=== resident K=320 EPI=4" >> gpurun_out/r2c6_kb_timeline.log; CLB_GEMM_EPI=4 timeout 100 python tools/gemm_timeline.py 32768 320 320 2>&1 | head -100 >> gpurun_out/r2c6_kb_timeline.log
unset CLB_LIB
timeout 600 python -m pytest tests -m gpu -q -x -k "norm or gemm or conv" > gpurun_out/r2c6_tests.log 2>&1; echo tests=$?; tail -3 gpurun_out/r2c6_tests.log
echo "=== GN fused split"; timeout 300 python tools/bw_bench.py --iters 5 2>&1 | grep -i "groupnorm" | cut -c1-110
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2c6_bench.log 2>&1; tail -1 gpurun_out/r2c6_bench.log | cut -c1-300
CLB_GN_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2c6_bench_gnold.log 2>&1; tail -1 gpurun_out/r2c6_bench_gnold.log | cut -c1-300
