set -x
timeout 600 python -m pytest tests -m gpu -q -x -k "gemm or conv or attention" > gpurun_out/r2c3_tests_kernels.log 2>&1; echo tests=$?
tail -8 gpurun_out/r2c3_tests_kernels.log
timeout 200 python tools/check_gemm.py perf > gpurun_out/r2c3_gemm_perf.log 2>&1; grep perf gpurun_out/r2c3_gemm_perf.log
CLB_GEMM_TMA_STORE=0 timeout 200 python tools/check_gemm.py perf > gpurun_out/r2c3_gemm_perf_oldepi.log 2>&1; grep "perf M" gpurun_out/r2c3_gemm_perf_oldepi.log | head -4
timeout 200 python tools/check_ops.py attn_perf > gpurun_out/r2c3_attn.log 2>&1; cat gpurun_out/r2c3_attn.log | tail -6
timeout 100 python tools/check_ops2.py attn_perf2 > gpurun_out/r2c3_attn_bwd.log 2>&1; tail -6 gpurun_out/r2c3_attn_bwd.log
for sh in "32768 320 320" "32768 320 320 lora" "32768 320 320 res"; do
  echo "=== timeline $sh" >> gpurun_out/r2c3_timeline.log
  CLB_LIB=$PWD/controllora_b200/libcontrollora_b200_tl.so timeout 120 python tools/gemm_timeline.py $sh >> gpurun_out/r2c3_timeline.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c3_tests.log 2>&1; echo tests=$?
tail -8 gpurun_out/r2c3_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2c3_bench.log 2>&1; tail -1 gpurun_out/r2c3_bench.log | cut -c1-400
