timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "clip" > gpurun_out/r2c26_tests.log 2>&1; echo tests=$?; grep -E "clip |passed|failed|Error" gpurun_out/r2c26_tests.log | head -12 | cut -c1-300
timeout 300 python tests/check_clip.py full 2>&1 | tail -3 | cut -c1-300
for v in 2 6; do CLB_V2_CTAS_PER_SM=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-aux --no-roofline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v2_per_sm=$v', round(d['ms_per_step'],3))"; done
