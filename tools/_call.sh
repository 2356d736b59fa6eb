timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c24_tests.log 2>&1; echo tests=$?; tail -3 gpurun_out/r2c24_tests.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux 2>/dev/null | tail -1 > gpurun_out/r2c24_bench.log; python -c "
import json; d=json.loads(open('gpurun_out/r2c24_bench.log').read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['gpu_launches'])"
