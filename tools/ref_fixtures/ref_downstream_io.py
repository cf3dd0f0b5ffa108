"""Input / output of tools/ref_fixtures/dump_ref_downstream.cpp (numpy only, so that it runs inside the reference's docker image).

  write-input <scans.bin>         the 6 synthetic Oxford-shape scans every downstream fixture is computed on + their ground-truth poses
  convert <dump.txt> <out.npz>    the dumper's text records -> tests/golden/ref_downstream.npz
"""
import os
import sys

import numpy as np

N_SCANS = 6


def stream():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from tbv_slam_public_b200 import synth     # numpy-only module
    return synth.make_stream(N_SCANS)


def write_input(path):
    st = stream()
    with open(path, "wb") as f:
        np.array([len(st.scans), st.scans.shape[1], st.scans.shape[2]], np.int32).tofile(f)
        np.ascontiguousarray(st.scans).tofile(f)
        np.ascontiguousarray(st.gt, np.float64).tofile(f)


def convert(txt, out):
    rec = {}
    for line in open(txt):
        parts = line.split()
        if not parts:
            continue
        tag, a, b, n = parts[0], int(parts[1]), int(parts[2]), int(parts[3])
        vals = np.array(parts[4:4 + n], np.float64)
        rec[f"{tag}_{a}"] = vals
        rec[f"{tag}_{a}_aux"] = np.array([b])
    np.savez_compressed(out, **rec)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "write-input":
        write_input(sys.argv[2])
    elif len(sys.argv) >= 4 and sys.argv[1] == "convert":
        convert(sys.argv[2], sys.argv[3])
    else:
        print(__doc__)
        sys.exit(2)
