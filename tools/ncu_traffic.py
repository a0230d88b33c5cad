#!/usr/bin/env python
"""Record the measured DRAM traffic of one kernel launch from an `ncu --set full` capture into profiles/traffic.json
(bench.py reports it as roofline.traffic).   python tools/ncu_traffic.py <rep.ncu-rep> <layer> <batch> [launch-index]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, layer, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2 + which]

    def val(name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "%": 1}.get(u, 1)
        return v * scale

    entry = {
        "kernel": r[hdr.index("Kernel Name")],
        "batch": batch,
        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
        "duration_us_under_ncu": val("gpu__time_duration.sum"),
        "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": val("launch__registers_per_thread"),
        "source": os.path.basename(rep),
    }
    path = os.path.join(ROOT, "profiles", "traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[layer] = entry
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
