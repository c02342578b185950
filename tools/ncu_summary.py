"""Condense `ncu --page raw --csv` output into the per-kernel metrics the design discussion uses (first launch of each
kernel name; --all: the LAST launch of each distinct (kernel name, grid size), i.e. warmed-up launches of every shape).
python tools/ncu_summary.py raw.csv [--all]"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__inst_executed_pipe_xu.sum"]

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, units, data = r, rows[i + 1], rows[i + 2:]
        break
if hdr is None:
    sys.exit("no header row found")
ki = hdr.index("Kernel Name")
rows_sel = []
if "--all" in sys.argv:
    gi = hdr.index("launch__grid_size") if "launch__grid_size" in hdr else None
    last = {}
    for r in data:
        if len(r) == len(hdr):
            last[(r[ki], r[gi] if gi is not None else "")] = r
    rows_sel = list(last.values())
else:
    seen = set()
    for r in data:
        if len(r) != len(hdr) or r[ki] in seen:
            continue
        seen.add(r[ki])
        rows_sel.append(r)
for r in rows_sel:
    print("----", r[ki][:110])
    for m in KEEP:
        if m in hdr:
            j = hdr.index(m)
            print(f"  {m:84s} {r[j]:>18s} {units[j]}")
