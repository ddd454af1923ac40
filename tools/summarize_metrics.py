"""Per-kernel table (mean duration, DRAM bytes, GB/s) from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` log:   python tools/summarize_metrics.py file.csv"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[hdr]
    ki, mi, vi, ui = (cols.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    data = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        key = (r[0], re.sub(r"\(.*", "", r[ki]).replace("void ", ""))
        v, u = float(r[vi].replace(",", "")), r[ui]
        if r[mi].startswith("gpu__time"):
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        elif "bytes" in r[mi]:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        data.setdefault(key, {})[r[mi]] = v
    agg = collections.OrderedDict()
    for (_, name), m in data.items():
        agg.setdefault(name, []).append(m)
    print(f"# {path}: ncu --clock-control none, per-launch means (cold caches: ncu flushes between launches)")
    print(f"{'kernel':30s} {'n':>4s} {'us':>9s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s} {'GB/s':>8s}")
    for name, ms in agg.items():
        t = sum(m["gpu__time_duration.sum"] for m in ms) / len(ms)
        rd = sum(m.get("dram__bytes_read.sum", 0) for m in ms) / len(ms)
        wr = sum(m.get("dram__bytes_write.sum", 0) for m in ms) / len(ms)
        print(f"{name[:30]:30s} {len(ms):4d} {t:9.1f} {rd / 1e6:11.1f} {wr / 1e6:11.1f} {(rd + wr) / t / 1e3:8.1f}")


if __name__ == "__main__":
    main()
