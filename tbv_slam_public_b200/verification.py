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
        self.X_, self.y_, self.py_clf_ = None, None, None

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

    # ---- training side (PythonClassifierInterface, alignmentinterface.cpp:46-222): sklearn does the fit, as in the reference ----------
    def AddDataPoint(self, X_i, y_i):
        X_i = np.atleast_2d(np.asarray(X_i, float))
        y_i = np.asarray(y_i, float).reshape(-1)
        if X_i.shape[0] != len(y_i) and X_i.shape[1] == len(y_i):       # a column of samples, as Eigen's (n, 1) matrices in the reference's tests
            X_i = X_i.T
        self.X_ = X_i if getattr(self, "X_", None) is None or len(self.X_) == 0 else np.vstack([self.X_, X_i])
        self.y_ = y_i if getattr(self, "y_", None) is None or len(self.y_) == 0 else np.concatenate([self.y_, y_i])

    def DataValid(self) -> bool:
        X, y = getattr(self, "X_", None), getattr(self, "y_", None)
        if X is None or y is None or len(X) != len(y) or len(y) < 1:
            return False
        return bool(np.all(np.isfinite(X)) and np.all(np.isfinite(y)))

    def fit(self):
        """LogisticRegression::fit (alignmentinterface.cpp:196-222): sklearn LogisticRegression(class_weight="balanced", max_iter=1000)."""
        if not self.DataValid():
            raise ValueError("training data invalid (the reference exits here)")
        from sklearn.linear_model import LogisticRegression as SkLR
        self.py_clf_ = SkLR(class_weight="balanced", max_iter=1000).fit(self.X_, self.y_)
        self.coef_ = np.asarray(self.py_clf_.coef_, float)[0].copy()
        self.intercept_ = float(np.asarray(self.py_clf_.intercept_, float).reshape(-1)[0])
        return self

    def predict(self, X) -> np.ndarray:
        if not self.IsFit():
            return np.zeros(len(np.atleast_2d(X)))
        return (self.predict_linear(X) > 0).astype(float)

    def Accuracy(self, y_true=None, y_pred=None) -> float:
        """balanced_accuracy_score; without arguments: on the training data (alignmentinterface.h Accuracy())."""
        if y_true is None:
            y_true, y_pred = self.y_, self.predict(self.X_)
        y_true, y_pred = np.asarray(y_true, float), np.asarray(y_pred, float)
        if len(y_true) != len(y_pred) or len(y_true) == 0:
            return -1.0
        recalls = [float(np.mean(y_pred[y_true == c] == c)) for c in np.unique(y_true)]
        return float(np.mean(recalls))

    def ConfusionMatrix(self, y_true=None, y_pred=None) -> np.ndarray:
        if y_true is None:
            y_true, y_pred = self.y_, self.predict(self.X_)
        y_true, y_pred = np.asarray(y_true, float), np.asarray(y_pred, float)
        if len(y_true) != len(y_pred) or len(y_true) == 0:
            return np.zeros((2, 2))
        m = np.zeros((2, 2))
        for t, p_ in zip(y_true, y_pred):
            m[int(t), int(p_)] += 1
        return m

    def SaveData(self, path: str):
        """One line per sample: y,x0,x1,... (alignmentinterface.cpp:160-180)."""
        with open(path, "w") as f:
            for y, x in zip(self.y_, self.X_):
                f.write(",".join(["%g" % y] + ["%g" % v for v in x]) + "\n")

    def LoadData(self, path: str):
        ys, xs = [], []
        with open(path) as f:
            for line in f:
                v = [t for t in line.strip().split(",") if t != ""]
                if v:
                    ys.append(float(v[0])); xs.append([float(t) for t in v[1:]])
        if ys:
            self.X_, self.y_ = np.asarray(xs, float), np.asarray(ys, float)
        return self

    def predict_linear(self, X) -> np.ndarray:
        X = np.asarray(X, float)
        if X.ndim == 1:                                    # one sample, or (single-feature model) a vector of samples
            X = X.reshape(-1, 1) if len(self.coef_) == 1 else X.reshape(1, -1)
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
