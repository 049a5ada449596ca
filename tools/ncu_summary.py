"""Summarises an .ncu-rep (raw page) into the few numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py report.ncu-rep [--stalls]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]} {units[i]}")
    for i, h in enumerate(hdr):
        if "pipe_tensor" in h and h not in want:
            print(f"{h:75s} {r[i]} {units[i]}")
    if "--stalls" in sys.argv:
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h:
                try:
                    st.append((float(r[i]), h))
                except ValueError:
                    pass
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
