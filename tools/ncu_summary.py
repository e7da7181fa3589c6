#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt            # one block per captured launch
    python tools/ncu_summary.py --group gpurun_out/x.ncu-rep > profiles/x.summary.txt    # one line per kernel name (sums / means)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
        "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    group = "--group" in sys.argv
    out = subprocess.run(["ncu", "-i", args[0], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: next((i for i, h in enumerate(hdr) if h.endswith(k)), None) for k in KEYS}
    if not group:
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("kernel:", d.get("Kernel Name"), "| grid", d.get("Grid Size"), "| block", d.get("Block Size"))
            for k in KEYS:
                for h, u in zip(hdr, units):
                    if h.endswith(k):
                        print(f"  {k:80s} {d[h]:>16s} {u}")
            print()
        return
    agg = collections.OrderedDict()
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        a = agg.setdefault(r[name_i], dict(n=0, ms=0.0, rd=0.0, wr=0.0, dram_pct=0.0, tensor_pct=0.0, regs=0))
        t = num(r[col["gpu__time_duration.sum"]]) * UNIT.get(units[col["gpu__time_duration.sum"]], 1.0)
        a["n"] += 1
        a["ms"] += t
        a["rd"] += num(r[col["dram__bytes_read.sum"]]) * UNIT.get(units[col["dram__bytes_read.sum"]], 1.0)
        a["wr"] += num(r[col["dram__bytes_write.sum"]]) * UNIT.get(units[col["dram__bytes_write.sum"]], 1.0)
        def get(key):                       # a metric that was not collected (or is n/a for this kernel) counts as 0
            i = col[key]
            v = num(r[i]) if i is not None else 0.0
            return 0.0 if v != v else v
        a["dram_pct"] += t * get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        a["tensor_pct"] += t * get("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
        a["regs"] = max(a["regs"], int(get("launch__registers_per_thread")))
    print(f"# {args[0]}: per kernel name -- launches, total gpu__time_duration (ms, cold-cache replays), DRAM read / written (GB),")
    print("# achieved DRAM GB/s over those launches, time-weighted gpu__dram_throughput % and tensor-pipe active %, registers")
    print(f"# {'n':>4s} {'ms':>9s} {'rd GB':>8s} {'wr GB':>8s} {'GB/s':>8s} {'dram%':>6s} {'tens%':>6s} {'regs':>5s}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        ms = max(a["ms"], 1e-9)
        print(f"  {a['n']:4d} {a['ms']:9.3f} {a['rd'] / 1e9:8.3f} {a['wr'] / 1e9:8.3f} {(a['rd'] + a['wr']) / 1e9 / (ms * 1e-3):8.0f} "
              f"{a['dram_pct'] / ms:6.1f} {a['tensor_pct'] / ms:6.1f} {a['regs']:5d}  {k[:140]}")


if __name__ == "__main__":
    main()
