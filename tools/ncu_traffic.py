#!/usr/bin/env python
"""profiles/traffic.json from the ncu CSV of tools/traffic_run.py: DRAM bytes, L2 bytes, time and
warp instructions per kernel of ONE cast step, and the whole cast's DRAM bytes per ray (what
bench.py scales into roofline.traffic).  usage: ncu_traffic.py <ncu.csv> <scans> [source tag]"""
import collections
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3,
        "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3, "inst": 1, "": 1}


def main():
    path, scans = sys.argv[1], int(sys.argv[2])
    tag = sys.argv[3] if len(sys.argv) > 3 else path
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    h = rows[hdr]
    ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    idi = h.index("ID")
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("b2::", "")
        name = name.split("<")[0]
        key = (r[idi], name)
        d = per.setdefault(key, {})
        try:
            d[r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1)
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for (_, name), d in per.items():
        a = agg.setdefault(name, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "l2_bytes": 0.0, "ms": 0.0,
                                  "warp_inst": 0.0})
        a["launches"] += 1
        a["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
        a["l2_bytes"] += d.get("lts__t_bytes.sum", 0.0)
        a["ms"] += d.get("gpu__time_duration.sum", 0.0)
        a["warp_inst"] += d.get("smsp__inst_executed.sum", 0.0)
    rays = scans * 64 * 2048
    cast = {k: v for k, v in agg.items() if k != "refit_kernel"}
    total = sum(v["dram_read"] + v["dram_write"] for v in cast.values())
    out = {"source": tag, "scans": scans, "rays": rays,
           "cast_dram_bytes_per_ray": total / rays,
           "cast_dram_write_bytes_per_ray": sum(v["dram_write"] for v in cast.values()) / rays,
           "algorithmic_bytes_per_ray": 8,
           "cast_ms_serialised": sum(v["ms"] for v in cast.values()),
           "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                       for k, v in agg.items()}}
    for k, v in out["kernels"].items():
        v["dram_bytes_per_ray"] = round((v["dram_read"] + v["dram_write"]) / rays, 3)
        v["share_of_cast_time"] = round(v["ms"] / max(out["cast_ms_serialised"], 1e-9), 3) if k in cast else None
        v["l2_gbs"] = round(v["l2_bytes"] / max(v["ms"], 1e-9) / 1e6, 1)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
