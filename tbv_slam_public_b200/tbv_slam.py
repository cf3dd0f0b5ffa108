"""The loop-closure / optimisation driver around the device calls — the callers on the far side of the hot path (SURVEY.md §3.2, §8d
config C4): `ScanContextClosure::SearchAndAddConstraint` (tbv_slam/src/tbv_slam/loopclosure.cpp:553-745), `RegisterLoopCandidate` /
`VerifyLoopCandidate` / `ApplyConstratins` (:261-384, :759-806), `TBVSLAM::ProcessFrame` (tbv_slam/src/tbv_slam/tbv_slam.cpp:32-43) and
`PoseGraph::ForceOptimize` (tbv_slam/src/tbv_slam/posegraph.cpp:112-130), over a `graph_io.SimpleGraph`.

Everything numerical runs on the GPU through `GpuLoopDevice` (Scan-Context descriptors, ring-key search and distances; batched candidate
registration against the device-resident keyframe database; CorAl and CFEAR quality; pose-graph assembly + step solve).  This module is
the reference's bookkeeping between those calls: which clouds are merged into a context, how a candidate becomes a registration guess,
which quality numbers feed which logistic model, which constraints enter the graph.  Per keyframe the <= N_CANDIDATES candidates are
handled as ONE batch per stage (one registration launch, one CorAl launch, one CFEAR launch) instead of the reference's per-candidate
loop; results and their order are the same.

The device is a constructor argument so that the bookkeeping can be exercised without a GPU (tests/test_tbv_slam_cpu.py plugs the CPU
oracle in); the product default fails loudly when there is no GPU (api.Context raises)."""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

from . import graph_io as G
from . import statistics as STAT
from . import verification as V

ODOM_BOUNDS, SC_SIM, COMBINED_COST = "odom-bounds", "sc-sim", "alignment_quality"      # tbv_slam/include/tbv_slam/utils.h:43-47


@dataclass
class LoopClosureParams:
    """loopclosure::Parameters with tbv_slam_offline's defaults (tbv_slam/include/tbv_slam/loopclosure.h:95-140, tbv_slam_offline.cpp:81-117)."""
    N_aggregate: int = 1
    use_peaks: bool = True
    transl_guess: bool = True
    speedup: bool = False
    registration_disabled: bool = False
    verification_disabled: bool = False
    verify_via_odometry: bool = True
    odom_sigma_error: float = 0.05
    model_threshold: float = 0.9
    all_candidates: bool = False
    model_features: tuple = (ODOM_BOUNDS, SC_SIM, COMBINED_COST)
    max_keyframes_per_call: int = 200                  # "if(++count_itrs == 200) break" (:622): the 200th keyframe is NOT processed


def _mat3(xyt):
    x, y, t = (float(v) for v in xyt)
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, -s, x], [s, c, y], [0.0, 0.0, 1.0]])


def _xyt(m):
    return np.array([m[0, 2], m[1, 2], math.atan2(m[1, 0], m[0, 0])])


def transform_cloud(cloud4: np.ndarray, xyt) -> np.ndarray:
    """pcl::transformPointCloud(in, out, Eigen::Affine3d): every coordinate in double, m(r,0) x + m(r,1) y + m(r,2) z + m(r,3) left to right,
    stored as float (PCL 1.10 Transformer<double>::se3); planar pose, so z passes through."""
    out = np.array(cloud4, np.float32, copy=True).reshape(-1, 4)
    m = _mat3(xyt)
    x, y, z = out[:, 0].astype(np.float64), out[:, 1].astype(np.float64), out[:, 2].astype(np.float64)
    out[:, 0] = (((m[0, 0] * x + m[0, 1] * y) + 0.0 * z) + m[0, 2]).astype(np.float32)
    out[:, 1] = (((m[1, 0] * x + m[1, 1] * y) + 0.0 * z) + m[1, 2]).astype(np.float32)
    out[:, 2] = (((0.0 * x + 0.0 * y) + 1.0 * z) + 0.0).astype(np.float32)
    return out


class GpuLoopDevice:
    """The device calls of the loop-closure thread, one method per reference call site."""

    def __init__(self, ctx, max_keyframes: int, cell_capacity: int = 1024, sc_params=None, sharded: bool = False, group=None):
        """sharded: register candidates through parallel.ShardedLoopClosure — every rank of the torch.distributed group holds the whole
        keyframe database and registers the candidates with id_from mod world == rank; the accepted constraints are all-gathered.
        All ranks must then run the same search (same graph, same calls); useful with SearchAndAddConstraintBatched.
        Only ACCEPTED candidates are exchanged (128 bytes each): a rejected candidate's record carries score 0.0 in this mode, where the
        single-GPU path reports the score its failed registration ended with (a log column; no decision reads it)."""
        from . import api
        self.api, self.ctx = api, ctx
        self.rsc = api.RSCManager(ctx, sc_params)
        self.db = api.LoopDB(ctx, max_keyframes, cell_capacity)
        self.sharded = None
        if sharded:
            from . import parallel
            self.sharded = parallel.ShardedLoopClosure(self.db, group)

    def close(self):
        self.db.close()

    # rsc_.makeAndSaveScancontextAndKeysRadarCloud + rsc_.detectLoopClosureID (loopclosure.cpp:585, 649)
    def make_context(self, cloud4, pose_xyt):
        self.rsc.makeAndSaveScancontextAndKeysRadarCloud(cloud4[:, 0], cloud4[:, 1], cloud4[:, 3], pose_xyt)

    def detect(self):
        return self.rsc.detectLoopClosureID()

    # the keyframe's cloud_normal_ becomes resident (tbv_loopdb_add); ids are graph rows
    def add_keyframe(self, cells) -> int:
        return self.db.add([cells])

    # loopclosure::Register for all candidates of a keyframe (:35-97): -> per candidate (ok, t_be xyt, cov (xx, xy, yy, tt), score)
    def register(self, id_from, id_to, T_from, T_to):
        if self.sharded is not None:
            out, summ = self.sharded.register_candidates(id_from, id_to, T_from, T_to), None
        else:
            out, summ = self.db.register_candidates(id_from, id_to, T_from, T_to, want_summaries=True)
        acc = {int(c["candidate"]): c for c in out}
        res = []
        for p in range(len(id_from)):
            if p in acc:
                c = acc[p]
                res.append((True, np.array(c["t_be"]), np.array(c["cov"]), float(c["score"])))
            else:                                                     # Register returned false: Tdiff stays Identity, Cov Identity (:344-352)
                res.append((False, np.zeros(3), np.array([1.0, 0.0, 1.0, 1.0]), float(summ[p].score) if summ is not None else 0.0))
        return res

    # getCorAlQualityMeasure / getCFEARQualityMeasure for all candidates of a keyframe (alignmentinterface.cpp:437-475)
    def coral(self, clouds, src, ref, T_src, T_ref, T_offset=None):
        r = self.api.CorAlRadarQuality(self.ctx, clouds, src, ref, T_src, T_ref, T_offset)
        return np.array([[q.joint, q.sep, q.overlap] for q in r], np.float64).reshape(-1, 3)

    def cfear(self, cellsets, src, ref, T_src, T_ref, T_offset=None):
        return self.ctx.CFEARQualityBatch(cellsets, src, ref, T_src, T_ref, T_offset)

    # CeresLeastSquares(...).Solve() (ceresoptimizer.cpp:13-62)
    def optimize(self, nodes, ids, meas, info, pgo_params, **kw):
        return self.api.pgo_optimize_device(self.ctx, nodes, ids, meas, pgo_params, info=info, **kw)   # whole LM loop on the device (tbv_pgo_optimize)


class ScanLearningInterface:
    """Training of the alignment classifier from odometry (coral_alignment_quality/src/alignment_checker/alignmentinterface.cpp:288-347,
    479-495; driven by cfear_radarodometry's odometry_training_node): every keyframe is compared with the previous one under 13 pose
    perturbations of the PREVIOUS scan — aligned, and +-0.5 / 1 / 2 m in x or y with 0.5 / 2 / 15 degrees — and the CorAl + CFEAR features
    of each become one sample (label 1 only for the unperturbed pair).  The 13 pairs of a keyframe are ONE CorAl launch and ONE CFEAR
    launch (T_offset carries the perturbations)."""
    range_error_, min_dist_btw_scans_ = 0.5, 0.5
    small_th_err, medium_th_err, large_th_err = 0.5 * math.pi / 180.0, 2 * math.pi / 180.0, 15 * math.pi / 180.0

    def __init__(self, device, combined: bool = True, small_errors=True, medium_errors=True, large_errors=True):
        self.dev, self.combined_ = device, combined
        self.cfear_class, self.coral_class, self.combined_class = V.LogisticRegression(), V.LogisticRegression(), V.LogisticRegression()
        self.prev_, self.frame_ = None, 0
        r = self.range_error_
        self.vek_perturbation_ = [(0.0, 0.0, 0.0)]                                   # CreatePerturbations (:479-495)
        for on, k, th in ((small_errors, 1, self.small_th_err), (medium_errors, 2, self.medium_th_err), (large_errors, 4, self.large_th_err)):
            if on:
                self.vek_perturbation_ += [(k * r, 0.0, th), (0.0, k * r, th), (-k * r, 0.0, th), (0.0, -k * r, th)]

    def features(self, current, prev, perturbations):
        """[n, 3] CorAl and [n, 3] CFEAR features of (ref = current, src = prev * perturbation); a scan is (pose xyt, peaks [n,4], cells [m,16])."""
        n = len(perturbations)
        cloud = lambda s: (s[1][:, 0], s[1][:, 1], s[1][:, 3])
        T_src, T_ref = np.tile(np.asarray(prev[0], np.float64), (n, 1)), np.tile(np.asarray(current[0], np.float64), (n, 1))
        off = np.asarray(perturbations, np.float64).reshape(n, 3)
        x_coral = self.dev.coral([cloud(current), cloud(prev)], [1] * n, [0] * n, T_src, T_ref, off)
        x_cfear = self.dev.cfear([current[2], prev[2]], [1] * n, [0] * n, T_src, T_ref, off)
        return np.asarray(x_coral, np.float64).reshape(n, 3), np.asarray(x_cfear, np.float64).reshape(n, 3)

    def AddTrainingData(self, pose_xyt, cloud_peaks, cells) -> int:
        current = (np.asarray(pose_xyt, np.float64), np.asarray(cloud_peaks, np.float32).reshape(-1, 4), np.asarray(cells, np.float64).reshape(-1, 16))
        first = self.frame_ == 0
        self.frame_ += 1
        if first:
            self.prev_ = current
            return 0
        if math.hypot(*(current[0][:2] - self.prev_[0][:2])) < self.min_dist_btw_scans_:
            return 0
        xc, xf = self.features(current, self.prev_, self.vek_perturbation_)
        y = np.array([1.0 if sum(abs(e) for e in v) < 0.0001 else 0.0 for v in self.vek_perturbation_])
        if self.combined_:
            self.combined_class.AddDataPoint(np.concatenate([xc, xf], axis=1), y)
        else:
            self.coral_class.AddDataPoint(xc, y)
            self.cfear_class.AddDataPoint(xf, y)
        self.prev_ = current
        return len(y)

    def FitModels(self):
        if self.combined_:
            self.combined_class.fit()
        else:
            self.coral_class.fit()
            self.cfear_class.fit()

    def PredAlignment(self, current, prev) -> dict:
        """quality map of one pair (:349-368): combined -> {alignment_quality: predict_linear}, else {Coral, CFEAR: predict_proba}."""
        xc, xf = self.features(current, prev, [(0.0, 0.0, 0.0)])
        if self.combined_:
            return {COMBINED_COST: float(self.combined_class.predict_linear(np.concatenate([xc, xf], axis=1))[0])}
        return {"Coral": float(self.coral_class.predict_proba(xc)[0]), "CFEAR": float(self.cfear_class.predict_proba(xf)[0])}

    def SaveCoefficients(self, directory: str):
        if self.combined_:
            self.combined_class.SaveCoefficients(directory + "/trained_alignment_classifier.txt")
        else:
            self.coral_class.SaveCoefficients(directory + "/trained_alignment_classifier_CorAl.txt")
            self.cfear_class.SaveCoefficients(directory + "/trained_alignment_classifier_CFEAR.txt")

    def LoadCoefficients(self, directory: str):
        if self.combined_:
            self.combined_class.LoadCoefficients(directory + "/trained_alignment_classifier.txt")
        else:
            self.coral_class.LoadCoefficients(directory + "/trained_alignment_classifier_CorAl.txt")
            self.cfear_class.LoadCoefficients(directory + "/trained_alignment_classifier_CFEAR.txt")


@dataclass
class CandidateRecord:
    """One evaluated candidate: what graph_->UpdateStatistics receives (loopclosure.cpp:714) — the rows of loop.csv."""
    id_from: int
    id_to: int
    guess_nr: int
    t_be: np.ndarray
    quality: dict
    probability: float
    reg_ok: bool
    applied: bool = False


class ScanContextClosure:
    """loopclosure.cpp:553-745 over a SimpleGraph.  `alignment_classifier`: the combined CorAl + CFEAR logistic model
    (ScanLearningInterface with combined_ = true, coefficients from tbv_slam/model_parameters/trained_alignment_classifier.txt via
    verification.LogisticRegression.LoadCoefficients); `verification_classifier`: optional, else the preset coefficients."""

    def __init__(self, graph: G.SimpleGraph, device, alignment_classifier: V.LogisticRegression, params: LoopClosureParams | None = None,
                 verification_classifier: V.LogisticRegression | None = None, odometry_coupled_closure: bool = True,
                 model_training_file_save: str = ""):
        self.graph, self.dev, self.par = graph, device, params or LoopClosureParams()
        self.model_training_file_save = model_training_file_save          # par_.model_training_file_save: collect (features, is-loop) samples
        self.alignment_classifier, self.verification_classifier = alignment_classifier, verification_classifier
        self.odometry_coupled_closure = odometry_coupled_closure
        self.itr_current = 0                              # row of the next keyframe to process
        self.n_resident = 0                               # rows whose cells are in the device database
        self.statistics: list[CandidateRecord] = []
        self.timing = STAT.statistics()                   # CFEAR_Radarodometry::timing under the reference's keys (loopclosure.cpp:647-731), host wall ms
        self.loop_constraints: dict[tuple[int, int], G.Constraint3d] = {}     # constraints_[loop_appearance], keyed (min, max)

    # ---- helpers ---------------------------------------------------------------------------------------------------------------------
    def _row_of(self):
        return {scan.idx_: r for r, (scan, _) in enumerate(self.graph.graph)}

    def ScansToLocalMap(self, row: int) -> np.ndarray:
        """:553-571 — clouds of the keyframes idx_ - N .. idx_ + N that exist, each moved to the world with its own pose, concatenated, then
        moved into this keyframe's frame.  [n, 4] float32 (x y z intensity)."""
        scan = self.graph.graph[row][0]
        rows = self._row_of()
        parts = []
        for i in range(scan.idx_ - self.par.N_aggregate, scan.idx_ + self.par.N_aggregate + 1):
            if i in rows:
                s = self.graph.graph[rows[i]][0]
                parts.append(transform_cloud(s.cloud_peaks_ if self.par.use_peaks else s.cloud_nopeaks_, G.pose3d_to_xyt(s.T)))
        merged = np.concatenate(parts) if parts else np.zeros((0, 4), np.float32)
        return transform_cloud(merged, _xyt(np.linalg.inv(_mat3(G.pose3d_to_xyt(scan.T)))))

    def _odometry_chain(self, id_to: int, id_from: int):
        """ConstraintsHandler::RelativeMotion(i, i + 1) for i in [to, from): the stored t_be of the odometry constraint between the two nodes,
        whatever its direction (the key is (min, max), cfear_radarodometry/src/cfear_radarodometry/types.cpp:166-171, 208-216); identity when
        missing."""
        table = {}
        for _, cons in self.graph.graph:
            for c in cons:
                if c.type == G.ODOMETRY:
                    table[(min(c.id_begin, c.id_end), max(c.id_begin, c.id_end))] = c
        chain = []
        for i in range(id_to, id_from):
            c = table.get((i, i + 1))
            chain.append(G.pose3d_to_xyt(c.t_be) if c is not None else np.zeros(3))
        return chain

    # ---- the loop --------------------------------------------------------------------------------------------------------------------
    def SearchAndAddConstraint(self) -> bool:
        """Processes up to max_keyframes_per_call - 1 keyframes; returns True while keyframes remain (the reference's return value)."""
        n = len(self.graph.graph)
        count = 0
        while self.itr_current < n:
            count += 1
            if count == self.par.max_keyframes_per_call:
                break
            self._process_keyframe(self.itr_current)
            self.itr_current += 1
        if self.itr_current == n:
            self._maybe_fit_verification_model()
        return self.itr_current != n

    def _maybe_fit_verification_model(self):
        if self.model_training_file_save and self.verification_classifier is not None and self.verification_classifier.DataValid() \
                and not self.verification_classifier.IsFit():
            self.verification_classifier.fit()                        # SaveVerificationTrainingData (:253-259)
            self.verification_classifier.SaveData(self.model_training_file_save)

    # ---- one keyframe = four stages; the batched search runs each stage once for ALL keyframes -----------------------------------------
    def _process_keyframe(self, row: int):
        plan = self._plan_keyframe(row)
        if plan is not None:
            self._register_plans([plan])
            self._verify_plans([plan])
            self._finish_keyframe(plan)

    def SearchAndAddConstraintBatched(self) -> bool:
        """The whole remaining graph at once.  In the offline flow nothing a keyframe does depends on an earlier keyframe's loop result
        (poses change only in ForceOptimize, after the search: tbv_slam_offline.cpp:279-282), so the per-keyframe loop can be cut into its
        stages: the Scan-Context sweep (sequential: the database grows), then ONE registration call for every candidate of every keyframe
        (the call parallel.ShardedLoopClosure shards over GPUs), ONE CorAl and ONE CFEAR call for all of them, then the host bookkeeping
        per keyframe in order.  Same records in the same order, same constraints as SearchAndAddConstraint (tests/test_tbv_slam_cpu.py)."""
        n = len(self.graph.graph)
        plans = []
        while self.itr_current < n:
            plan = self._plan_keyframe(self.itr_current)
            if plan is not None:
                plans.append(plan)
            self.itr_current += 1
        self._register_plans(plans)
        self._verify_plans(plans)
        for plan in plans:
            self._finish_keyframe(plan)
        self._maybe_fit_verification_model()
        return False

    def _plan_keyframe(self, row: int):
        """Stage A (:636-699): context, Scan-Context candidates, registration guesses.  Returns None when the keyframe needs nothing more."""
        scan = self.graph.graph[row][0]
        while self.n_resident <= row:                                 # cells of every keyframe up to this one are on the device
            self.dev.add_keyframe(self.graph.graph[self.n_resident][0].cloud_normal_)
            self.n_resident += 1
        if len(self.graph.graph) == 1:                                # itr_begin == itr_end (:640)
            return None
        pose = G.pose3d_to_xyt(scan.T)
        t1 = time.perf_counter()
        self.dev.make_context(self.ScansToLocalMap(row), pose)        # CreateContext (:573-592): the node's own pose is the odometry pose
        t2 = time.perf_counter()
        candidates = self.dev.detect()
        t3 = time.perf_counter()
        self.timing.Document("Descriptor", 1e3 * (t2 - t1))
        self.timing.Document("Detect loop", 1e3 * (t3 - t2))
        plan = dict(row=row, pose=pose, entries=[], batch=[])         # entries: per guess, a finished record or the index into batch
        if not candidates:
            plan["entries"].append(CandidateRecord(scan.idx_, scan.idx_, -1, np.zeros(3),
                                                   {ODOM_BOUNDS: 1.0, SC_SIM: 1.0 + float(self.odometry_coupled_closure), COMBINED_COST: -20.0},
                                                   0.0, False))
            return plan
        for guess_nr, cand in enumerate(candidates):
            id_to = int(cand["nn_idx"])                               # database index = order of makeAndSave calls = graph row
            scan_to = self.graph.graph[id_to][0]
            if self.odometry_coupled_closure and self.par.speedup and cand["min_dist_odom"] > 0.7:
                plan["entries"].append(CandidateRecord(scan.idx_, scan_to.idx_, guess_nr, np.zeros(3),
                                                       {ODOM_BOUNDS: cand["min_dist_odom"], SC_SIM: cand["min_dist"], COMBINED_COST: -20.0}, 0.0, False))
                continue
            # Tsrcguess = Taug^-1 * Rz(yaw) (:691-696); Tto = Tfrom * guess (:337-338)
            ax, ay = (cand["aug_xy"] if self.par.transl_guess else (0.0, 0.0))
            guess = _mat3((-float(ax), -float(ay), 0.0)) @ _mat3((0.0, 0.0, float(cand["yaw_diff_rad"])))
            plan["entries"].append(len(plan["batch"]))
            plan["batch"].append((guess_nr, cand, id_to, _xyt(_mat3(pose) @ guess)))
        return plan

    def _register_plans(self, plans):
        """Stage B: loopclosure::Register for every candidate of every plan in one device call (:704 -> :320-364 -> :35-97)."""
        todo = [(pl, b) for pl in plans for b in pl["batch"]]
        for pl in plans:
            pl["reg"] = []
        if not todo:
            return
        if self.par.registration_disabled:                            # Tdiff = Tfrom^-1 * T(to) (:349)
            for pl, b in todo:
                Tto = _mat3(G.pose3d_to_xyt(self.graph.graph[b[2]][0].T))
                pl["reg"].append((True, _xyt(np.linalg.inv(_mat3(pl["pose"])) @ Tto), np.array([1.0, 0.0, 1.0, 1.0]), 0.0))
            return
        t0 = time.perf_counter()
        reg = self.dev.register([pl["row"] for pl, _ in todo], [b[2] for _, b in todo], np.array([pl["pose"] for pl, _ in todo]).reshape(-1, 3),
                                np.array([b[3] for _, b in todo]).reshape(-1, 3))
        self.timing.Document("Register", 1e3 * (time.perf_counter() - t0))          # one sample per CALL: all its candidates together
        for (pl, _), r in zip(todo, reg):
            pl["reg"].append(r)

    def _verify_plans(self, plans):
        """Stage C: VerifyByAlignment for every candidate of every plan — one CorAl call, one CFEAR call (:708 -> :365-384 -> :759-775).
        PredAlignment(current = from, prev = to) -> CreateQualityType(ref = current, src = prev) (alignmentinterface.cpp:349-353, 437-475;
        AlignmentQuality.h:263): the candidate (`to`, at Tfrom * t_be) is the MOVING / src scan, the query keyframe the reference one."""
        todo = [(pl, j) for pl in plans for j in range(len(pl["batch"]))]
        for pl in plans:
            pl["X"] = np.zeros((len(pl["batch"]), 6))
        if not todo:
            return
        used = sorted({pl["row"] for pl, _ in todo} | {pl["batch"][j][2] for pl, j in todo})
        slot = {r: i for i, r in enumerate(used)}
        cloud = lambda s: (s.cloud_peaks_[:, 0], s.cloud_peaks_[:, 1], s.cloud_peaks_[:, 3])
        clouds = [cloud(self.graph.graph[r][0]) for r in used]
        cellsets = [self.graph.graph[r][0].cloud_normal_ for r in used]
        src = [slot[pl["batch"][j][2]] for pl, j in todo]
        ref = [slot[pl["row"]] for pl, _ in todo]
        T_from = np.array([pl["pose"] for pl, _ in todo]).reshape(-1, 3)
        T_to_reg = np.array([_xyt(_mat3(pl["pose"]) @ _mat3(pl["reg"][j][1])) for pl, j in todo]).reshape(-1, 3)   # to at Tfrom * t_be
        t0 = time.perf_counter()
        x_coral = np.asarray(self.dev.coral(clouds, src, ref, T_to_reg, T_from), np.float64).reshape(-1, 3)
        x_cfear = np.asarray(self.dev.cfear(cellsets, src, ref, T_to_reg, T_from), np.float64).reshape(-1, 3)
        self.timing.Document("VerifyByAlignment", 1e3 * (time.perf_counter() - t0))
        for (pl, j), xc, xf in zip(todo, x_coral, x_cfear):
            pl["X"][j] = np.concatenate([xc, xf])

    def _finish_keyframe(self, plan):
        """Stage D: quality -> probability -> statistics -> ApplyConstratins, in guess order (:700-724)."""
        scan = self.graph.graph[plan["row"]][0]
        k = len(plan["batch"])
        alignment_quality = self.alignment_classifier.predict_linear(plan["X"]) if k else []   # quality[COMBINED_COST] (PredAlignment :355-360)
        evaluated = []
        for entry in plan["entries"]:
            if isinstance(entry, CandidateRecord):                    # no candidate / skipped by `speedup`
                self.statistics.append(entry)
                continue
            guess_nr, cand, id_to, _ = plan["batch"][entry]
            scan_to = self.graph.graph[id_to][0]
            quality = {ODOM_BOUNDS: 0.0, SC_SIM: float(cand["min_dist"]), COMBINED_COST: float(alignment_quality[entry])}   # CreateAppearanceConstraint
            quality[ODOM_BOUNDS] = V.VerifyByOdometry(self._odometry_chain(scan_to.idx_, scan.idx_), self.par.odom_sigma_error,
                                                      self.par.verify_via_odometry)
            if self.par.verification_disabled:
                prob = 0.0
            else:
                feats = [quality[f] for f in self.par.model_features]
                prob = V.VerificationModel(feats[0], feats[1], feats[2], self.verification_classifier)
            ok, t, cov4, _score = plan["reg"][entry]
            cov6 = np.eye(6)
            if ok and not self.par.registration_disabled:             # reg_cov of the moving scan, xy block rotated (:91-94); singular like the reference's
                cov6 = np.diag([0.0, 0.0, 0.0, 0.0, 0.0, float(cov4[3])])
                cov6[0, 0], cov6[0, 1], cov6[1, 0], cov6[1, 1] = cov4[0], cov4[1], cov4[1], cov4[2]
            con = G.Constraint3d(scan.idx_, scan_to.idx_, G.pose3d_from_xyt(t), G._information(cov6), G.LOOP_APPEARANCE, quality, "")
            rec = CandidateRecord(scan.idx_, scan_to.idx_, guess_nr, np.array(t), dict(quality), float(prob), bool(ok))
            self.statistics.append(rec)
            if self.model_training_file_save:                         # AddVerificationTrainingData (:240-251)
                is_loop, pos_ok = candidate_loop_status(update_statistics(self.graph, rec))
                if pos_ok:
                    if self.verification_classifier is None:
                        self.verification_classifier = V.LogisticRegression()
                    self.verification_classifier.AddDataPoint([[quality[f] for f in self.par.model_features]], [float(is_loop)])
            evaluated.append((prob, con, rec))
        # ApplyConstratins (:261-275)
        t0 = time.perf_counter()
        for i in V.apply_constraints([e[0] for e in evaluated], self.par.model_threshold, self.par.all_candidates):
            _, con, rec = evaluated[i]
            rec.applied = True
            self.loop_constraints[(min(con.id_begin, con.id_end), max(con.id_begin, con.id_end))] = con
        self.timing.Document("Apply contraints", 1e3 * (time.perf_counter() - t0))  # sic


@dataclass
class OptimizeResult:
    summary: object = None
    n_loop_constraints: int = 0
    poses_before: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    poses_after: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))


class TBVSLAM:
    """TBVSLAM::ProcessFrame (tbv_slam.cpp:32-43) + PoseGraph::ForceOptimize (posegraph.cpp:112-130) for the offline flow
    (tbv_slam_offline.cpp:269-285): `while ProcessFrame(optimize=False, loopclosure=True)`, then `ProcessFrame(optimize=True, loopclosure=False)`."""

    def __init__(self, graph: G.SimpleGraph, device, alignment_classifier, loop_params: LoopClosureParams | None = None, pgo_params=None,
                 verification_classifier=None, odometry_coupled_closure: bool = True):
        self.graph, self.dev, self.pgo_params = graph, device, pgo_params
        self.loop = ScanContextClosure(graph, device, alignment_classifier, loop_params, verification_classifier, odometry_coupled_closure)
        self.last_optimization: OptimizeResult | None = None

    def ProcessFrame(self, optimize: bool, loopclosure: bool, batched: bool = False) -> bool:
        more = (self.loop.SearchAndAddConstraintBatched() if batched else self.loop.SearchAndAddConstraint()) if loopclosure else False
        if optimize:
            self.ForceOptimize()
        return more

    def ForceOptimize(self, **solver_options) -> OptimizeResult:
        """Moves the verified loop constraints into the graph (AddConstraintThread drains the queue before the solve, posegraph.cpp:115-116)
        and runs the optimiser in place on the node poses."""
        for con in self.loop.loop_constraints.values():
            if not any(c.type == G.LOOP_APPEARANCE and c.id_begin == con.id_begin and c.id_end == con.id_end
                       for c in self.graph.graph[self.loop._row_of()[con.id_begin]][1]):
                self.graph.AddConstraint(con)
        nodes, ids, meas, info, _ = self.graph.pgo_arrays()
        res = OptimizeResult(n_loop_constraints=int((ids[:, 2] == 1).sum()) if len(ids) else 0, poses_before=self.graph.poses_xyt())
        replace = True if self.pgo_params is None else bool(self.pgo_params.replace_cov_by_identity)
        t0 = time.perf_counter()
        new_nodes, res.summary = self.dev.optimize(nodes, ids, meas, None if replace else info, self.pgo_params, **solver_options)
        self.loop.timing.Document("Pose grapgh optimization", 1e3 * (time.perf_counter() - t0))   # sic (posegraph.cpp:126)
        self.graph.set_poses(new_nodes)
        res.poses_after = self.graph.poses_xyt()
        self.last_optimization = res
        return res

    def Run(self, batched: bool = False) -> OptimizeResult:
        """SLAMEval::RunBasicEvaluation (tbv_slam_offline.cpp:269-285); batched = the whole search as one call per stage."""
        while self.ProcessFrame(False, True, batched):
            pass
        self.ProcessFrame(True, False)
        return self.last_optimization


def _g6(v: float) -> str:
    """operator<< of a double with setprecision(6) and no floatfield: printf %g with 6 significant digits."""
    return "%.6g" % float(v)


def update_statistics(graph: G.SimpleGraph, rec: CandidateRecord) -> dict:
    """PoseGraph::UpdateStatistics (tbv_slam/src/tbv_slam/posegraph.cpp:332-371): ground-truth view of one evaluated candidate — the error of
    the registered transform against Tgt_from^-1 Tgt_to, the ground-truth distance of the candidate, and the closest EARLIER node (more than
    10 keyframes back, with ground truth) as the reference's definition of "a loop exists here"."""
    by_idx = {scan.idx_: (r, scan) for r, (scan, _) in enumerate(graph.graph)}
    _, a = by_idx[rec.id_from]
    _, b = by_idx[rec.id_to]
    Tfrom, Tto = G.pose3d_to_matrix(a.Tgt), G.pose3d_to_matrix(b.Tgt)
    Tguess = np.eye(4)
    m = _mat3(rec.t_be)
    Tguess[:2, :2], Tguess[:2, 3] = m[:2, :2], m[:2, 2]
    Terror = np.linalg.inv(Tguess) @ (np.linalg.inv(Tfrom) @ Tto)
    nearest, cand_dist, Tclosest, close = 100000.0, -1.0, Tfrom, rec.id_from
    if a.has_Tgt_:
        if b.has_Tgt_:
            cand_dist = float(np.linalg.norm(Tfrom[:3, 3] - Tto[:3, 3]))
        for r, (s, _) in enumerate(graph.graph):
            if not s.idx_ < rec.id_from:
                break
            if abs(float(rec.id_from) - float(s.idx_)) > 10 and s.has_Tgt_:
                Ts = G.pose3d_to_matrix(s.Tgt)
                d = float(np.linalg.norm(Tfrom[:3, 3] - Ts[:3, 3]))
                if d < nearest:
                    nearest, Tclosest, close = d, Ts, r
    return dict(Tfrom=Tfrom, Tto=Tto, Tclosest=Tclosest, Tgt_diff=Terror, closest_loop_distance=nearest, candidate_loop_distance=cand_dist,
                id_from=rec.id_from, id_to=rec.id_to, id_close=close, guess_nr=rec.guess_nr, quality=dict(rec.quality))


def candidate_loop_status(row: dict):
    """EvaluationManager::getCandidateLoopStatus (EvaluationManager.cpp:12-27): (is_loop, prediction_pos_ok)."""
    is_loop = row["closest_loop_distance"] < 6
    e = row["Tgt_diff"]
    close = float(np.linalg.norm(e[:3, 3])) < 4 and 180.0 / math.pi * abs(math.atan2(e[1, 0], e[1, 1])) < 2.5
    return is_loop, (not is_loop) or close


def write_loop_csv(path: str, graph: G.SimpleGraph, records, parameter_names: str = "", parameter_values: str = "") -> int:
    """loop/loop.csv as EvaluationManager::writeResultsToCSV writes it (place_recognition_radar/include/place_recognition_radar/
    EvaluationManager.h:57-81, src/place_recognition_radar/EvaluationManager.cpp:29-57), the input of place_recognition_radar/python/
    LoopClosureEval.py: positions with 6 significant digits, the quality map in std::map (sorted-key) order with 6 decimals (Join,
    cfear_radarodometry/include/cfear_radarodometry/utils.h:63-82), the run's parameter names / values appended to every line.
    diff.z is eulerAngles(0,1,2)[2] of the error rotation = its yaw for planar poses.  Returns the number of rows."""
    rows = [update_statistics(graph, r) for r in records]
    if not rows:
        return 0

    def join(q, first):
        keys = sorted(q)
        if len(keys) == 1:
            return keys[0]                                     # Join's size-1 branch prints the key either way
        return ",".join(k if first else "%.6f" % q[k] for k in keys)

    with open(path, "w") as f:
        f.write("from.x,from.y,from.z,to.x,to.y,to.z,close.x,close.y,close.z,diff.x,diff.y,diff.z,closest_loop_distance,candidate_loop_distance,"
                "id_from,id_to,id_close,guess_nr," + join(rows[0]["quality"], True) + "," + parameter_names + "\n")
        for r in rows:
            e = r["Tgt_diff"]
            vals = [_g6(r[k][i, 3]) for k in ("Tfrom", "Tto", "Tclosest") for i in range(3)]
            vals += [_g6(e[0, 3]), _g6(e[1, 3]), _g6(math.atan2(e[1, 0], e[1, 1])), _g6(r["closest_loop_distance"]), _g6(r["candidate_loop_distance"]),
                     str(r["id_from"]), str(r["id_to"]), str(r["id_close"]), str(r["guess_nr"])]
            f.write(",".join(vals) + "," + join(r["quality"], False) + "," + parameter_values + "\n")
    return len(rows)
