"""Turn ncu exports into the small JSON summaries committed under profiles/.

    python tools/ncu_digest.py launches <launches.csv> <out.json>
        per-kernel launch count / total time / share from
        `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>`
    python tools/ncu_digest.py full <raw.csv> <out.json> [kernel-substring]
        selected metrics of one kernel from `ncu -i capture.ncu-rep --page raw --csv > raw.csv`
        (capture taken with `ncu --set full --clock-control none --import-source on -k regex:<kernel>`)
"""
import csv
import json
import re
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def short_name(full):
    # "void dpgo::k_rtr_fused<5, 3, 2>(dpgo::FusedParams)" -> "dpgo::k_rtr_fused<5, 3, 2>"
    name = re.sub(r"^void\s+", "", full)
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    kn, mv, mn = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    acc, total = {}, 0.0
    for r in rows:
        if r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        k = short_name(r[kn])
        a = acc.setdefault(k, {"launches": 0, "sum": 0.0})
        a["launches"] += 1
        a["sum"] += v
        total += v
    for a in acc.values():
        a["share"] = a["sum"] / total if total else 0.0
        a["avg"] = a["sum"] / a["launches"]
    json.dump({"unit": "ns", "total": total, "kernels": acc}, open(out, "w"), indent=1)
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1]["sum"])[:12]:
        print(f"{a['share']:7.3%} {a['launches']:5d} x {a['avg'] / 1e3:10.1f} us  {k}")


def full(path, out, needle=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, units, data = rows[0], rows[1], rows[2:]
    kn = head.index("Kernel Name")
    pick = [r for r in data if needle is None or needle in r[kn]]
    if not pick:
        raise SystemExit("no kernel matches %r" % needle)
    r = pick[-1]
    metrics = {}
    for i, h in enumerate(head):
        base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1][:1].isupper() else h
        for want in KEEP:
            if h == want or h.endswith("." + want) or base == want:
                metrics[want] = {"unit": units[i], "value": r[i]}

    def num(key, scale):
        m = metrics.get(key)
        if not m:
            return None
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(m["unit"], 1) if scale else 1
        return float(m["value"].replace(",", "")) * mult

    rd, wr = num("dram__bytes_read.sum", True), num("dram__bytes_write.sum", True)
    res = {"kernel": short_name(r[kn]), "grid": r[head.index("Grid Size")], "block": r[head.index("Block Size")],
           "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": (rd or 0) + (wr or 0) if rd is not None else None, "metrics": metrics}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "metrics"}))


if __name__ == "__main__":
    if len(sys.argv) < 4 or sys.argv[1] not in ("launches", "full"):
        raise SystemExit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
