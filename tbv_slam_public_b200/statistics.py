"""Per-stage timing under the reference's key names (SURVEY.md §5 "Tracing / profiling").

The reference keeps one global table, `CFEAR_Radarodometry::timing` (cfear_radarodometry/include/cfear_radarodometry/statistics.h:38,
src/cfear_radarodometry/statistics.cpp:10-51): `Document(name, value)` appends a sample, `GetStatistics()` prints, per name,
"<name> avg, <mean>", "<name> dev [σ], <variance>", "<name> count, <n>" with std::to_string's six decimals — the block its tools write to
`pars.txt` / `time_statistics.txt`.  Here the samples are DEVICE times: `Context.profile_begin()` / `profile_end()` return one
(kernel, milliseconds) record per launch, `document_profile` folds them into the reference's stage names.
"""
from __future__ import annotations

# kernel -> the reference's timing key (radar_driver.cpp:87,111; odometrykeyframefuser.cpp:253-256; loopclosure.cpp:647-731; posegraph.cpp:126)
STAGE_OF_KERNEL = {
    "k1_filter_fused": "Filtering", "cfar_rows": "Filtering", "cfar_emit": "Filtering", "k_rotate90ccw": "Filtering",
    "k_compensate": "compensate", "k_compensate_polar": "compensate",
    "cells_fused": "build_normals", "c1_grid": "build_normals", "c2_scan": "build_normals", "c3_scatter": "build_normals", "c4_centroids": "build_normals",
    "c5_cells": "build_normals", "c6_compact": "build_normals",
    "k_register": "register", "k_odom_problems": "register",
    "k_odom_motion": "publish_etc", "k_odom_update": "publish_etc", "k_cellgrid_build": "publish_etc",
    "sc_make": "Descriptor",
    "sc_similarity": "Detect loop", "sc_search": "Detect loop", "sc_distance": "Detect loop",
    "k_pack_constraints": "Register", "k_merge_constraints": "Register", "nccl_all_gather": "Register",
    "k_coral": "Verify loop candidate",
    "pgo_blocks": "Pose grapgh optimization", "pgo_gather": "Pose grapgh optimization", "pgo_cost": "Pose grapgh optimization", "pgo_pcg_cr": "Pose grapgh optimization", "pgo_cr_setup": "Pose grapgh optimization", "pgo_cr_eliminate": "Pose grapgh optimization", "pgo_cr_update": "Pose grapgh optimization", "pgo_scale": "Pose grapgh optimization", "pgo_damping": "Pose grapgh optimization", "pgo_step": "Pose grapgh optimization", "pgo_gradmax": "Pose grapgh optimization",   # sic: the reference's spelling
}


class statistics:
    def __init__(self):
        self.t: dict[str, list[float]] = {}

    def Document(self, name: str, value: float, report: bool = False):
        self.t.setdefault(name, []).append(float(value))
        if report:
            print('Statistics: "%s" = %s' % (name, value))

    def GetStatistics(self) -> str:
        out = []
        for name, v in self.t.items():
            mean = sum(v) / len(v)
            var = sum((x - mean) * (x - mean) for x in v) / len(v)     # the reference prints the variance under "dev [σ]" (:23,46)
            out.append("%s avg, %.6f\n%s dev [σ], %.6f\n%s count, %d\n" % (name, mean, name, var, name, len(v)))
        return "".join(out)


def document_profile(stats: statistics, records, loop_registration: bool = False) -> statistics:
    """records: [(kernel, ms), ...] of ONE step / call -> one sample per stage.  loop_registration: k_register launches belong to
    loopclosure::Register ("Register") instead of the odometry's "register"."""
    acc: dict[str, float] = {}
    for kernel, ms in records:
        stage = STAGE_OF_KERNEL.get(kernel, kernel)
        if loop_registration and stage == "register":
            stage = "Register"
        acc[stage] = acc.get(stage, 0.0) + float(ms)
    for stage, ms in acc.items():
        stats.Document(stage, ms)
    return stats
