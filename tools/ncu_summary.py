"""Summarise an .ncu-rep (read on the CPU box): key metrics + stall breakdown + top stalled SASS lines."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_sector_op_read_hit_rate.pct"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                if "issue_stalled" in h and float(v) < 0.05:
                    continue
            except ValueError:
                pass
            print(f"  {h:95s} {u:14s} {v}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    f = lambda x: float(x) if x.replace(".", "", 1).isdigit() else 0.0
    print("SASS instructions:", len(data), "total samples:", sum(f(r[ix["# Samples"]]) for r in data))
    for r in sorted(data, key=lambda r: -f(r[ix["# Samples"]]))[:int(sys.argv[2])]:
        st = {c: f(r[ix[c]]) for c in hdr if c.startswith("stall_") and "Not" not in c}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"  {r[ix['Address']][-6:]} {f(r[ix['# Samples']]):7.0f} {top}  {r[ix['Source']][:70]}")
