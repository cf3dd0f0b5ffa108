"""Loop-candidate verification around the batched kernels: the host-side policy of the reference, fed by
`api.CorAlRadarQuality` (k_coral) and `Context.CFEARQualityBatch` (k_register in evaluation mode).

  LogisticRegression     coral_alignment_quality/src/alignment_checker/alignmentinterface.cpp:224-279 (sklearn bridge: one line
                         "intercept,coef..." per file; predict_linear = coef . x + intercept; predict_proba = sigmoid of it)
  combined_features      ScanLearningInterface::PredAlignment (alignmentinterface.cpp:347-365): X_combined = [X_CorAl, X_CFEAR]
  VerifyByOdometry       tbv_slam/src/tbv_slam/loopclosure.cpp:776-806
  VerificationModel      tbv_slam/src/tbv_slam/loopclosure.cpp:218-238 (preset coefficients when no classifier was trained)
  apply_constraints      tbv_slam/src/tbv_slam/loopclosure.cpp:261-275 (sort by probability, best or all, model_threshold)

Host arithmetic (a handful of dot products per candidate); the expensive part of verification — 20.5 + 2.6 ms per candidate in the
reference — is the two batched kernels.
"""
from __future__ import annotations

import math

import numpy as np


class LogisticRegression:
    def __init__(self, intercept: float | None = None, coef=None):
        self.intercept_ = intercept
        self.coef_ = None if coef is None else np.asarray(coef, float)

    def IsFit(self) -> bool:
        return self.coef_ is not None

    def LoadCoefficients(self, path: str):
        coefs = []
        with open(path) as f:
            for line in f:
                v = [t for t in line.strip().split(",") if t != ""]
                if not v:
                    continue
                self.intercept_ = float(v[0])
                coefs += [float(t) for t in v[1:]]
        self.coef_ = np.asarray(coefs, float)
        return self

    def SaveCoefficients(self, path: str):
        with open(path, "w") as f:
            f.write(",".join(["%g" % self.intercept_] + ["%g" % c for c in self.coef_]) + "\n")

    def predict_linear(self, X) -> np.ndarray:
        X = np.atleast_2d(np.asarray(X, float))
        return X @ self.coef_ + self.intercept_

    def predict_proba(self, X) -> np.ndarray:
        if not self.IsFit():                       # "Model is not fitted yet. Return probability as zero(s)" (:23-28)
            return np.zeros(len(np.atleast_2d(X)))
        return 1.0 / (1.0 + np.exp(-self.predict_linear(X)))


def combined_features(coral_results, cfear_quality) -> np.ndarray:
    """[n, 6] = joint, sep, overlap | score, residuals, mean size — X_combined of PredAlignment."""
    c = np.array([[r.joint, r.sep, r.overlap] for r in coral_results], float).reshape(-1, 3)
    return np.concatenate([c, np.asarray(cfear_quality, float).reshape(-1, 3)], axis=1)


def VerifyByOdometry(rel_motions_xy, odom_sigma_error: float = 0.05, verify_via_odometry: bool = True) -> float:
    """similarity in [0, 1): rel_motions = the chain of relative motions (x, y, yaw) between consecutive keyframes to ... from."""
    if not verify_via_odometry:
        return 1.0
    T = np.eye(3)
    trav = 0.0
    for x, y, th in rel_motions_xy:
        c, s = math.cos(th), math.sin(th)
        trav += math.hypot(x, y)
        T = T @ np.array([[c, -s, x], [s, c, y], [0, 0, 1]])
    est = math.hypot(T[0, 2], T[1, 2])
    error = max(est - 5.0, 0.0)
    rel = error / trav
    return 1.0 - math.exp(-rel * rel / (2 * odom_sigma_error * odom_sigma_error))


PRESET_COEF = (-2.89398535, -9.40230684, 0.23891265)   # odom-bounds, sc-sim, alignment_quality (loopclosure.cpp:224-227)
PRESET_BIAS = 2.67958289


def VerificationModel(odom_bounds: float, sc_sim: float, alignment_quality: float, classifier: LogisticRegression | None = None) -> float:
    x = np.array([odom_bounds, sc_sim, alignment_quality], float)
    z = float(classifier.predict_linear(x)[0]) if (classifier is not None and classifier.IsFit()) else float(np.dot(PRESET_COEF, x) + PRESET_BIAS)
    return 1.0 / (1.0 + math.exp(-z))


def apply_constraints(probabilities, model_threshold: float = 0.9, all_candidates: bool = False):
    """Indices of the candidates whose constraints are added to the graph (ApplyConstratins)."""
    order = sorted(range(len(probabilities)), key=lambda i: -probabilities[i])
    if not order:
        return []
    take = order if all_candidates else order[:1]
    return [i for i in take if probabilities[i] > model_threshold]


def verify_candidates(ctx, clouds, cellsets, src, ref, T_src, T_ref, sc_sim, odom_bounds, alignment_classifier: LogisticRegression,
                      verification_classifier: LogisticRegression | None = None):
    """All candidates of one query at once: CorAl + CFEAR features (two launches), combined alignment score, verification probability."""
    from . import api
    coral = api.CorAlRadarQuality(ctx, clouds, src, ref, T_src, T_ref)
    cfear = ctx.CFEARQualityBatch(cellsets, src, ref, T_src, T_ref)
    X = combined_features(coral, cfear)
    quality = alignment_classifier.predict_linear(X)
    p = [VerificationModel(odom_bounds[i], sc_sim[i], quality[i], verification_classifier) for i in range(len(src))]
    return np.array(p), X, quality
