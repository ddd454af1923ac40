"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share.

    python tools/summarize_launches.py launches.csv [last_frames | --frames A B]
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[hdr]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        seq.append((name, v))
    return seq


def main():
    path = sys.argv[1]
    last_frames = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "--frames" else 0
    seq = load(path)
    if "--frames" in sys.argv:      # --frames A B: frames A..B (1-based, inclusive) only; a frame starts with its binning
        a, b = (int(v) for v in sys.argv[sys.argv.index("--frames") + 1:][:2])
        first = next((f for f in ("k_bin_sort", "k_bin_bucket") if any(n.startswith(f) for n, _ in seq)), "k_bin_atoms")
        idx = [i for i, (n, _) in enumerate(seq) if n.startswith(first)] + [len(seq)]
        seq = seq[idx[a - 1]:idx[b]]
    if last_frames:      # keep only the launches from the last `last_frames` frames on (a frame starts with its binning)
        first = next((f for f in ("k_bin_sort", "k_bin_bucket") if any(n.startswith(f) for n, _ in seq)), "k_bin_atoms")
        idx = [i for i, (n, _) in enumerate(seq) if n.startswith(first)]
        seq = seq[idx[-last_frames]:]
    agg = collections.OrderedDict()
    for n, v in seq:
        agg.setdefault(n, []).append(v)
    total = sum(v for _, v in seq)
    print(f"# {path}: {len(seq)} launches, {total:.1f} us total device time (cold-cache, serialised)")
    print(f"{'kernel':58s} {'n':>4s} {'mean_us':>10s} {'sum_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:58]:58s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v):10.1f} {100 * sum(v) / total:6.1f}%")


if __name__ == "__main__":
    main()
