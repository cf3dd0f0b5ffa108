"""SASS evidence for the kernels of libtbv_b200.so (run in the build container: cuobjdump needs no GPU).

For every kernel: the target architecture, instruction count, and the count of the mnemonics that tell how it moves data and synchronises —
UBLKCP (TMA bulk copy: cp.async.bulk), UTMALDG / UTMASTG (TMA tensor copies), SYNCS (mbarrier arrive / try_wait), UCGABAR (cluster
barrier), LDGSTS (cp.async), ATOMS / RED, DFMA / DADD / DMUL (fp64), HMMA / UTC*MMA (tensor cores: none, by design) — and the first lines on which
the TMA / mbarrier / cluster instructions occur.  Writes profiles/<prefix>_sass_<kernel>.txt for the kernels named on the command line and a
one-table summary for all.   python tools/sass_evidence.py r2 'k1_filter_fused<false, 8, 64>' k_register cells_fused pgo_pcg_cr"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tbv_slam_public_b200", "libtbv_b200.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "LDGSTS", "ATOMS", "ATOMG", "RED", "DFMA", "DADD", "DMUL", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "BAR.SYNC",
         "LDS.128", "LDG.E.128", "STG.E.128", "PREFETCH", "CCTL"]


def kernels():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, arch, body = None, None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = {"arch": arch, "lines": []}
            continue
        m = re.match(r"arch = (\S+)", line.strip())
        if m:
            arch = m.group(1)
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            body[cur]["lines"].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
    return body


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name


def main():
    prefix, wanted = sys.argv[1], sys.argv[2:]
    K = kernels()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    table = []
    for name, k in K.items():
        ops = collections.Counter()
        for ln in k["lines"]:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
            if m:
                ops[m.group(1)] += 1
        cnt = {w: sum(v for o, v in ops.items() if o.startswith(w)) for w in WATCH}
        table.append((demangle(name).split("(")[0], k["arch"], len(k["lines"]), cnt))
        dem = demangle(name)
        short = next((w for w in wanted if w in dem.split("(")[0]), None)
        if short:
            path = os.path.join(ROOT, "profiles", f"{prefix}_sass_{re.sub('[^A-Za-z0-9_]+', '_', short).strip('_')}.txt")
            with open(path, "w") as f:
                f.write(f"# cuobjdump -sass tbv_slam_public_b200/libtbv_b200.so — {demangle(name)}\n# arch {k['arch']}, {len(k['lines'])} SASS instructions\n")
                f.write("# " + ", ".join(f"{w} {c}" for w, c in cnt.items() if c) + "\n")
                for i, ln in enumerate(k["lines"]):
                    if any(t in ln for t in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "MEMBAR", "FENCE")):
                        lo, hi = max(0, i - 2), min(len(k["lines"]), i + 3)
                        f.write("\n".join(k["lines"][lo:hi]) + "\n  ...\n")
    with open(os.path.join(ROOT, "profiles", f"{prefix}_sass_summary.txt"), "w") as f:
        f.write("# every kernel of libtbv_b200.so: arch, SASS instruction count, data-movement / synchronisation / fp64 / tensor mnemonics (tools/sass_evidence.py)\n")
        for nm, arch, n, cnt in sorted(table):
            f.write(f"{nm:60s} {arch} {n:6d}  " + " ".join(f"{w}={c}" for w, c in cnt.items() if c) + "\n")
    print(open(os.path.join(ROOT, "profiles", f"{prefix}_sass_summary.txt")).read())


if __name__ == "__main__":
    main()
