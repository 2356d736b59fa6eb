timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm or conv or lora or unet or train" > gpurun_out/r2c21_tests.log 2>&1; echo tests=$?; tail -12 gpurun_out/r2c21_tests.log | cut -c1-300
for lib in new old; do
if [ $lib = old ]; then export CLB_LIB=$PWD/controllora_b200/libclb_old.so; else unset CLB_LIB; fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c21_bench_$lib.log; python -c "
import json; d=json.loads(open('gpurun_out/r2c21_bench_$lib.log').read()); print('$lib', d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], [ (r['M'],r['N'],r['K'],r['lora'],round(r['us'],1)) for r in d['aux']['gemm_per_shape']['rows'][:6]])"
done
export CLB_LIB=$PWD/controllora_b200/libcontrollora_b200_tl.so
timeout 120 python tools/gemm_timeline.py 32768 320 320 0 lora > gpurun_out/r2c21_tl_lora.log 2>&1
