#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics per kernel and the top stall instructions of the source page.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 40]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
        "lts__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct", "launch__registers_per_thread ", "smsp__inst_executed.sum ", "sm__cycles_elapsed.avg ",
        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fma", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "smsp__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts.sum ",
        "smsp__average_warp", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:100])
        for i, h in enumerate(hdr):
            if any(h.startswith(w.strip()) if w.endswith(" ") else w in h for w in WANT):
                print(f"   {h:90s} {units[i]:12s} {r[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for bi, st in enumerate(starts[:1]):
        h = rows[st]
        end = starts[bi + 1] - 1 if bi + 1 < len(starts) else len(rows)
        data = [r for r in rows[st + 1:end] if len(r) == len(h)]
        ci = {n: i for i, n in enumerate(h)}
        tot = sum(int(r[ci["# Samples"]]) for r in data)
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        agg = {n: sum(int(r[ci[n]]) for r in data) for n in stalls}
        print("total samples", tot, "by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
        for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:top]:
            s = int(r[ci["# Samples"]])
            st2 = sorted(((int(r[ci[n]]), n[6:]) for n in stalls), reverse=True)[:2]
            print(f"{s:6d} {100 * s / tot:5.1f}% x{r[ci['Instructions Executed']]:>8s} {r[ci['Source']].strip()[:80]:80s} {st2}")


if __name__ == "__main__":
    main()
