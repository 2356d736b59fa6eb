set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_tests.log 2>&1; echo tests=$?
tail -5 gpurun_out/r2c1_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench.log 2>&1; tail -1 gpurun_out/r2c1_bench.log | cut -c1-600
for sh in "32768 320 320" "32768 320 320 lora" "32768 320 320 res" "8192 640 2560" "32768 320 2880" "2048 1280 5120"; do
  echo "=== timeline $sh" >> gpurun_out/r2c1_timeline.log
  CLB_LIB=$PWD/controllora_b200/libcontrollora_b200_tl.so timeout 120 python tools/gemm_timeline.py $sh >> gpurun_out/r2c1_timeline.log 2>&1
done
timeout 200 python tools/check_gemm.py perf > gpurun_out/r2c1_gemm_perf.log 2>&1
timeout 300 tools/ncu_capture.sh r2_gemm_k320_lora 'gemm_tc_kernel' 3 python tools/ncu_gemm.py 32768 320 320 lora 6
timeout 300 tools/ncu_capture.sh r2_gemm_2cta_bn160 'gemm_tc_kernel' 3 python tools/ncu_gemm.py 32768 320 2880 nolora 6
timeout 300 tools/ncu_capture.sh r2_gemm_2cta_bn256 'gemm_tc_kernel' 3 python tools/ncu_gemm.py 8192 5120 2560 nolora 6
timeout 300 tools/ncu_capture.sh r2_attn_bwd_dkv 'attn_bwd_dkv' 1 python tools/ncu_attn.py
timeout 300 tools/ncu_capture.sh r2_attn_bwd_dq 'attn_bwd_dq' 1 python tools/ncu_attn.py
timeout 300 tools/ncu_capture.sh r2_gn_reduce 'gn_reduce' 2 python tools/ncu_norm.py
timeout 300 tools/ncu_capture.sh r2_gn_apply 'gn_apply' 2 python tools/ncu_norm.py
ls -la gpurun_out | tail -30
