"""Per-source-line warp-stall samples and executed instructions from an ncu report (cuda,sass correlated view).
   python tools/ncu_lines.py gpurun_out/full_k_register.ncu-rep [top_n] [launch_index]"""
import csv
import subprocess
import sys


def main(rep, top=40, which=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # split into per-launch tables
    starts = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
    starts.append(len(rows))
    seg = rows[starts[which]:starts[which + 1]]
    hdr = next(r for r in seg if r and r[0] == "Line No")
    hi = seg.index(hdr)
    ci = {n: k for k, n in enumerate(hdr)}
    samp, inst = ci["# Samples"], ci["Instructions Executed"]
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    lines = []
    for r in seg[hi + 1:]:
        if len(r) < len(hdr) or r[2] != "-":   # keep only the source-line rows (Address == '-')
            continue
        try:
            s, n = int(r[samp]), int(r[inst])
        except ValueError:
            continue
        top_st = sorted(((int(r[ci[x]] or 0), x[6:]) for x in stalls), reverse=True)[:3]
        lines.append((s, n, r[0], r[1], top_st))
    tot_s, tot_i = sum(l[0] for l in lines), sum(l[1] for l in lines)
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for s, n, ln, src, st in sorted(lines, reverse=True)[:top]:
        print(f"{s / max(tot_s, 1) * 100:5.1f}% samp {n / max(tot_i, 1) * 100:5.1f}% inst  L{ln:>4s}  {src.strip()[:100]:100s} {' '.join(f'{b}:{a}' for a, b in st if a)}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
