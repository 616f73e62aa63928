#!/usr/bin/env python3
"""Summarise ncu CSV exports into profiles/: kernel shares from a launch list (--metrics gpu__time_duration.sum) and the
per-launch figures of a --set full capture (time, DRAM bytes, pipe utilisation).  Usage:
  tools/ncu_summary.py TAG gpurun_out/TAG_launches_c4.csv gpurun_out/TAG_ncu_full_raw.csv [bench.json]
writes profiles/TAG_summary.md and profiles/lincomb_traffic.json (the file bench.py reads `roofline.traffic` from)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name[:40]


def read_csv(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    return rows[hdr], rows[hdr + 1:]


def launch_shares(path):
    hdr, rows = read_csv(path)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += ms
    return agg


def full_rows(path):
    hdr, rows = read_csv(path)
    units, rows = rows[0], rows[1:]
    out = []
    for r in rows:
        if len(r) < len(hdr):
            continue
        out.append({h: r[i] for i, h in enumerate(hdr)})
    return out, {h: units[i] for i, h in enumerate(hdr)}


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return 0.0


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return num(v) * scale.get(unit, 1)


def main():
    tag, launches, full = sys.argv[1:4]
    bench = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else None
    shares = launch_shares(launches)
    total = sum(v[1] for v in shares.values())
    lines = [f"# {tag}: ncu summary (C4: N_E=2^14, L_E=8, n=1031; 1x B200)", "",
             "## Kernel shares (launch list: `ncu --metrics gpu__time_duration.sum --clock-control none`, "
             "`bench.py --steps 2 --warmup 1`; cold-cache, serialised)", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(shares.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
    if bench:
        kb = bench["kernels_ms_per_step"]
        tb = sum(kb.values())
        lines += ["", "## Same shares from the live CUDA-event timing inside bench.py (warm, per proof)", "",
                  "| kernel | ms per proof | share |", "|---|---|---|"]
        for k, ms in sorted(kb.items(), key=lambda kv: -kv[1]):
            if ms > 0:
                lines.append(f"| {k} | {ms:.3f} | {100 * ms / tb:.1f}% |")
        lines += ["", f"proof: {bench['value']:.2f} ms device-resident, {bench['e2e']['value']:.2f} ms through the C ABI with host buffers; "
                  f"k_crs_lincomb {bench['roofline']['achieved']:.0f} GB/s = {100 * bench['roofline']['frac']:.1f}% of the measured "
                  f"{bench['roofline']['peak']:.0f} GB/s copy peak."]
    rows, units = full_rows(full)
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid")]
    lines += ["", "## `ncu --set full --clock-control none` per launch (bench.py --steps 1 --warmup 1)", "",
              "| kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|" + "---|" * len(cols)]
    lincomb = []
    lincomb_name = "k_crs_lincomb"
    for r in rows:
        name = short(r["Kernel Name"])
        cells = []
        for key, _ in cols:
            v, u = r.get(key, ""), units.get(key, "")
            if key.startswith("dram__bytes"):
                cells.append(f"{to_bytes(v, u) / 1e6:.1f} MB")
            elif key == "gpu__time_duration.sum":
                ms = num(v) / (1e6 if u in ("ns", "nsecond") else 1e3 if u in ("us", "usecond") else 1)
                cells.append(f"{ms:.3f} ms")
            else:
                cells.append(f"{num(v):.1f}" if "." in str(v) else str(v))
        lines.append(f"| {name} | " + " | ".join(cells) + " |")
        if name.startswith("k_crs_lincomb"):
            u = units["gpu__time_duration.sum"]
            lincomb_name = name
            lincomb.append({"grid": r.get("launch__grid_size"),
                            "dram_read_bytes": to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]),
                            "dram_write_bytes": to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"]),
                            "time_ms": num(r["gpu__time_duration.sum"]) / (1e6 if u in ("ns", "nsecond") else 1e3 if u in ("us", "usecond") else 1)})
    open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
    if lincomb and bench:
        per_step = int(round(bench["work_per_step"]["lincomb_launches"]))
        first = lincomb[:per_step]                     # the launches of one proof, in order
        traffic = sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in first)
        out = {"kernel": lincomb_name, "source": f"profiles/{tag}_ncu_full_raw.csv (ncu --set full --clock-control none, "
               "bench.py --steps 1 --warmup 1, C4, 1x B200)", "per_launch": first, "launches_per_step": per_step,
               "dram_bytes_per_step": traffic, "algorithmic_bytes_per_step": bench["roofline"]["algorithmic_bytes_per_step"],
               "ratio": traffic / bench["roofline"]["algorithmic_bytes_per_step"]}
        json.dump(out, open(os.path.join(ROOT, "profiles", "lincomb_traffic.json"), "w"), indent=1)
        print("lincomb DRAM traffic / algorithmic bytes =", out["ratio"])
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
