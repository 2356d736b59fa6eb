#!/bin/bash
# usage: tools/ncu_capture.sh NAME KERNEL_REGEX SKIP CMD...   (run on the GPU box, one GPU)
# One `ncu --set full` capture of the first launch matching KERNEL_REGEX after SKIP matching launches; writes
# gpurun_out/NAME.ncu-rep plus text exports (details page, selected raw metrics) that can be read without a GPU.
set -u
NAME=$1; REGEX=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -s "$SKIP" -c 1 -f -o "gpurun_out/$NAME" "$@" > "gpurun_out/$NAME.ncu.log" 2>&1
ncu -i "gpurun_out/$NAME.ncu-rep" --page details > "gpurun_out/$NAME.details.txt" 2>&1
ncu -i "gpurun_out/$NAME.ncu-rep" --page raw --csv > "gpurun_out/$NAME.raw.csv" 2>&1
grep -E 'Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum"|dram__bytes_write.sum"|sm__pipe_tensor|sm__inst_executed_pipe_uma|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__dram_throughput.avg.pct|lts__t_sector_hit_rate|sm__throughput.avg.pct|smsp__issue_active.avg.pct' "gpurun_out/$NAME.raw.csv" | head -5 > /dev/null
echo "captured $NAME"
