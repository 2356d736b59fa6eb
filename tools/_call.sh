set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c13_tests.log 2>&1; echo tests=$?; tail -4 gpurun_out/r2c13_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c13_smoke.log 2>&1; tail -2 gpurun_out/r2c13_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c13_bench.log 2>&1; tail -1 gpurun_out/r2c13_bench.log | cut -c1-400
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel --launch-skip 1500 -c 480 --csv --log-file gpurun_out/r2c13_gemm_dram.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-aux > gpurun_out/r2c13_ncu_dram.log 2>&1; tail -1 gpurun_out/r2c13_ncu_dram.log | cut -c1-120
timeout 200 python tools/check_gemm.py perf 2>&1 | grep perf > gpurun_out/r2c13_gemm_perf.log; cat gpurun_out/r2c13_gemm_perf.log
