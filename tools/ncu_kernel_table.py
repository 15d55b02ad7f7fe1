"""Per-kernel table from an ncu metrics CSV of one training step (tools/one_step.py):
    python tools/ncu_kernel_table.py gpurun_out/x.csv > profiles/x.md
Columns: launches, total device time, share of the step, DRAM bytes moved, achieved HBM GB/s (and its fraction of the
measured copy peak in MEASURED_PEAKS.json), tensor-pipe active %, registers.  Times under ncu are serialised and
cold-cache: compare shares and per-kernel ratios, not absolute step time."""
import csv
import json
import re
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches: OrderedDict = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        d = launches.setdefault(r[0], {"name": r[ik]})
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        u = r[iu]
        if r[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)  # -> us
        if r[im].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d[r[im]] = v
    agg: dict = {}
    for d in launches.values():
        name = re.sub(r"\(.*", "", d["name"])
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "bytes": 0.0, "tensor": 0.0, "regs": 0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["us"] += t
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["tensor"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
        a["regs"] = max(a["regs"], int(d.get("launch__registers_per_thread", 0)))
    total = sum(a["us"] for a in agg.values())
    try:
        peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
        src = "measured"
    except (OSError, KeyError, ValueError):
        peak, src = 6650.0, "fallback"
    print(f"# per-kernel ncu metrics, one eager step ({Path(path).name}); HBM peak {peak:.0f} GB/s ({src}); "
          f"{len(launches)} launches, {total / 1e3:.1f} ms summed under ncu\n")
    print("| kernel | launches | ms | share | DRAM MB | HBM GB/s | of peak | tensor pipe % | regs |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = a["bytes"] / a["us"] / 1e3 if a["us"] else 0.0
        print(f"| `{name[:90]}` | {a['n']} | {a['us'] / 1e3:.3f} | {a['us'] / total * 100:.1f}% | {a['bytes'] / 1e6:.1f} | "
              f"{gbs:.0f} | {gbs / peak:.2f} | {a['tensor'] / a['us'] if a['us'] else 0:.1f} | {a['regs']} |")


if __name__ == "__main__":
    main()
