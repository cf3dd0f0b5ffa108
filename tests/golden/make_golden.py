"""Generates tests/golden/filters_ref.npz from the REFERENCE'S OWN filter code.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
It builds oracle/_ref/libtbv_ref_filters.so (`make -C oracle ref`: the reference's unmodified radar_filters.cpp / cfar.cpp
compiled where they lie, against the container stand-ins of oracle/ref_shim/) and runs, for every case below, exactly what
radarDriver::Process runs (cfear_radarodometry/src/cfear_radarodometry/radar_driver.cpp:48-60).  Inputs and outputs are stored
together so the fixture is self-contained: the GPU box has neither /root/reference nor a need for it.

Outputs per case: the reference's clouds as float32 x, y, intensity in the reference's order ("filtered point indices" =
this ordered list; x, y are injective in (azimuth, range) so bit-equal clouds mean equal indices).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

OX, MU = 0.0438, 0.0595238


def cases():
    """(name, image, kind, params) — kind 'ks' (k-strongest + peaks) or 'cfar'."""
    from tbv_slam_public_b200 import synth
    out = []
    scan = synth.make_stream(1).scans[0]
    radar_ox = np.ascontiguousarray(scan[100:148])            # 48 azimuths x 3768 bins, radar-like
    radar_mu = np.ascontiguousarray(scan[200:248, :3360])     # MulRan shape
    for k, z in [(40, 60.0), (12, 70.0), (12, 60.0)]:
        out.append((f"radar_ox_k{k}_z{int(z)}", radar_ox, "ks", dict(z_min=z, k=k, min_distance=2.5, range_res=OX)))
    out.append(("radar_mu_k40_z60", radar_mu, "ks", dict(z_min=60.0, k=40, min_distance=2.5, range_res=MU)))
    out.append(("radar_mu_k12_z70", radar_mu, "ks", dict(z_min=70.0, k=12, min_distance=2.5, range_res=MU)))
    for kind in ("uniform", "equal", "zeros", "ramp", "sparse"):
        img = synth.stress_image(kind, n_az=24, n_range=512, seed=4)
        for k, z in [(12, 70.0), (40, 60.0), (5, 0.0), (128, 1.0)]:
            out.append((f"stress_{kind}_k{k}_z{int(z)}", img, "ks", dict(z_min=z, k=k, min_distance=0.3, range_res=OX)))
    rng = np.random.default_rng(77)
    for shape in [(7, 64), (3, 5), (2, 17), (33, 1000), (1, 256), (5, 13)]:
        img = rng.integers(0, 256, size=shape, dtype=np.uint8)
        img[rng.random(shape) < 0.85] = 10
        out.append((f"shape_{shape[0]}x{shape[1]}_z60", img, "ks", dict(z_min=60.0, k=12, min_distance=0.1, range_res=OX)))
        out.append((f"shape_{shape[0]}x{shape[1]}_z5", img, "ks", dict(z_min=5.0, k=12, min_distance=0.1, range_res=OX)))
    # kept bins within 6 of either row end (the NMS window leaves the row; the first / last row leave the buffer)
    edge = np.full((6, 128), 20, np.uint8)
    for a in range(6):
        edge[a, [0, 1, 2, 3, 5, 6, 121, 122, 124, 125, 126, 127]] = [200, 90, 150, 220, 99, 180, 170, 95, 210, 140, 230, 160]
        edge[a, 60 + a] = 250
    out.append(("edge_bins", edge, "ks", dict(z_min=60.0, k=40, min_distance=0.0, range_res=OX)))
    # MulRan min_range_bin float widening: 2.5 / (double)(float)0.0595238 -> ceil = 43
    q = np.zeros((4, 128), np.uint8)
    q[:, 42], q[:, 43], q[:, 44] = 200, 201, 202
    out.append(("min_range_bin_mu", q, "ks", dict(z_min=60.0, k=12, min_distance=2.5, range_res=MU)))
    out.append(("min_range_bin_ox", q, "ks", dict(z_min=60.0, k=12, min_distance=2.5, range_res=OX)))
    # CA-CFAR (cfar.cpp:35-83) over the parameter families of the kstrong_vs_cfar launch files
    for w, g, pfa, z in [(40, 10, 0.01, 20.0), (10, 20, 0.01, 60.0), (100, 5, 0.001, 20.0), (500, 10, 0.1, 20.0), (40, 10, 0.0001, 0.0)]:
        out.append((f"cfar_radar_w{w}_g{g}_p{pfa}_z{int(z)}", radar_ox, "cfar",
                    dict(window_size=w, nb_guard_cells=g, false_alarm_rate=pfa, static_threshold=z, min_distance=2.5, range_res=OX)))
    out.append(("cfar_radar_mu", radar_mu, "cfar", dict(window_size=40, nb_guard_cells=10, false_alarm_rate=0.01, static_threshold=20.0,
                                                       min_distance=2.5, range_res=MU)))
    for kind in ("uniform", "ramp", "sparse", "equal"):
        img = synth.stress_image(kind, n_az=8, n_range=512, seed=5)
        out.append((f"cfar_{kind}", img, "cfar", dict(window_size=20, nb_guard_cells=4, false_alarm_rate=0.05, static_threshold=30.0,
                                                      min_distance=0.5, range_res=OX)))
    return out


def main():
    from oracle import ref_py
    assert ref_py.build(), "oracle/_ref could not be built (is /root/reference present?)"
    arrays, manifest = {}, []
    images = {}
    for name, img, kind, par in cases():
        key = None
        for k2, v in images.items():
            if v.shape == img.shape and np.array_equal(v, img):
                key = k2
        if key is None:
            key = f"img{len(images)}"
            images[key] = img
            arrays[key] = img
        if kind == "ks":
            r = ref_py.kstrongest(img, **par)
            for which in ("filtered", "peaks"):
                for comp, a in zip("xyi", r[which]):
                    arrays[f"{name}.{which}.{comp}"] = a
            n = (len(r["filtered"][0]), len(r["peaks"][0]))
        else:
            x, y, i = ref_py.cacfar(img, max_distance=400.0, **par)
            arrays[f"{name}.x"], arrays[f"{name}.y"], arrays[f"{name}.i"] = x, y, i
            n = (len(x),)
        manifest.append(dict(name=name, image=key, kind=kind, params=par, counts=n))
    arrays["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "filters_ref.npz")
    np.savez_compressed(path, **arrays)
    print(f"{path}: {len(manifest)} cases, {os.path.getsize(path) / 1024:.0f} KiB")
    for m in manifest:
        print(f"  {m['name']:40s} {m['counts']}")


if __name__ == "__main__":
    main()
