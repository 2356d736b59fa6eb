"""Print selected metrics of an `ncu --page raw --csv` export (written by tools/ncu_capture.sh): python tools/ncu_raw.py file.raw.csv [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr = i
        break
names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__pipe_tensor|sm__inst_executed_pipe_(uma|tc)|"
                 r"sm__warps_active.avg.pct|launch__registers_per_thread|lts__t_sector_hit_rate.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
                 r"smsp__inst_executed.sum$|sm__throughput.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|sm__cycles_active.avg$|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__cycles_active.avg$|sm__pipe_xu|sm__inst_executed_pipe_xu|lts__t_bytes.sum$|"
                 r"l1tex__m_xbar2l1tex_read_bytes.sum$|sm__sass_inst_executed_op_shared|smsp__pcsamp_warps_issue_stalled")
for n, u, v in zip(names, units, vals):
    if pat.search(n):
        print(f"{n:90s} {v:>18s} {u}")
