"""Per-CUDA-source-line instruction shares from an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py report.ncu-rep [top=40] [kernel name (regex:... allowed)]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    only = ["-k", sys.argv[3]] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else []
    txt = subprocess.run(["ncu", "-i", rep, *only, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
    h = rows[hdr[0]]
    ii = h.index("Instructions Executed")
    isamp = h.index("# Samples") if "# Samples" in h else None
    ia = h.index("Address")
    items, tot, stot = [], 0, 0
    for k, h0 in enumerate(hdr):                       # one table per source file, CUDA lines have no address
        end = hdr[k + 1] if k + 1 < len(hdr) else len(rows)
        fname = ""
        for r in rows[max(0, h0 - 3):h0]:
            if r and r[0] == "File Name":
                fname = r[1].split("/")[-1]
        for r in rows[h0 + 1:end]:
            if len(r) <= ii or r[ia].strip() not in ("", "-"):
                continue
            try:
                n = int(r[ii])
            except ValueError:
                continue
            sm = int(r[isamp]) if isamp is not None and r[isamp].isdigit() else 0
            items.append((n, sm, f"{fname}:{r[0]}", r[1][:105]))
            tot += n
            stot += sm
    print(f"warp instructions {tot}, stall samples {stot}")
    if "--by-samples" in sys.argv:
        items.sort(key=lambda it: it[1], reverse=True)
    else:
        items.sort(reverse=True)
    for n, s, ln, src in items[:top]:
        print(f"{100 * n / tot:5.1f}% inst {100 * s / max(stot, 1):5.1f}% samp  {ln:>16}  {src}")


if __name__ == "__main__":
    main()
