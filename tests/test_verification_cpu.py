"""Host-side verification policy: logistic models, odometry bound, candidate selection — against sklearn (the library
the reference bridges to) and, when /root/reference is present, with the reference's trained coefficient files."""
import math
import os

import numpy as np
import pytest

from tbv_slam_public_b200 import verification as V

REF_MODELS = "/root/reference/tbv_slam/model_parameters"


def test_logistic_regression_equals_sklearn(tmp_path):
    from sklearn.linear_model import LogisticRegression as SK
    rng = np.random.default_rng(0)
    m = V.LogisticRegression(-8.4, rng.normal(0, 3, 6))
    p = str(tmp_path / "c.txt")
    m.SaveCoefficients(p)
    m2 = V.LogisticRegression().LoadCoefficients(p)
    assert np.allclose(m2.coef_, m.coef_, rtol=1e-5) and abs(m2.intercept_ - m.intercept_) < 1e-5
    sk = SK()
    sk.classes_ = np.array([0, 1]); sk.coef_ = m2.coef_[None, :]; sk.intercept_ = np.array([m2.intercept_])
    X = rng.normal(0, 1, (50, 6))
    assert np.allclose(sk.predict_proba(X)[:, 1], m2.predict_proba(X), rtol=1e-12)
    assert np.allclose(sk.decision_function(X), m2.predict_linear(X), rtol=1e-12)
    assert np.array_equal(V.LogisticRegression().predict_proba(X), np.zeros(50))      # unfitted: zeros, like the bridge


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="/root/reference not present (GPU box)")
def test_reference_coefficient_files_load():
    a = V.LogisticRegression().LoadCoefficients(os.path.join(REF_MODELS, "trained_alignment_classifier.txt"))
    v = V.LogisticRegression().LoadCoefficients(os.path.join(REF_MODELS, "trained_loop_classifier.txt"))
    assert len(a.coef_) == 6 and len(v.coef_) == 3 and a.intercept_ < 0 < v.intercept_
    # an aligned-looking candidate (low joint-sep gap, many residuals) scores above a misaligned-looking one
    good = a.predict_linear([[-3.2, -3.1, 0.5, 5.0, 200.0, 320.0]])[0]
    bad = a.predict_linear([[-1.0, -3.1, 0.2, 9.0, 60.0, 320.0]])[0]
    assert good > bad


def test_verification_model_and_selection():
    p = V.VerificationModel(0.0, 0.1, 3.0)
    z = -2.89398535 * 0.0 - 9.40230684 * 0.1 + 0.23891265 * 3.0 + 2.67958289
    assert abs(p - 1 / (1 + math.exp(-z))) < 1e-15
    clf = V.LogisticRegression(4.53196, [-5.06267, -11.9655, 0.268186])
    assert abs(V.VerificationModel(0.2, 0.3, 1.0, clf) - 1 / (1 + math.exp(-(4.53196 - 5.06267 * 0.2 - 11.9655 * 0.3 + 0.268186)))) < 1e-15
    assert V.apply_constraints([0.2, 0.95, 0.97]) == [2]
    assert V.apply_constraints([0.2, 0.95, 0.97], all_candidates=True) == [2, 1]
    assert V.apply_constraints([0.2, 0.5]) == [] and V.apply_constraints([]) == []


def test_verify_by_odometry():
    # a closed square of 40 m sides: estimated distance 0 -> similarity 0; an open 200 m line -> ~1
    sq = [(40.0, 0.0, math.pi / 2)] * 4
    assert V.VerifyByOdometry(sq) < 1e-12
    line = [(2.0, 0.0, 0.0)] * 100
    s = V.VerifyByOdometry(line)
    rel = (200.0 - 5.0) / 200.0
    assert abs(s - (1 - math.exp(-rel * rel / (2 * 0.05 ** 2)))) < 1e-12
    assert V.VerifyByOdometry(line, verify_via_odometry=False) == 1.0


# ---- the reference's own unit tests for the classifier bridge (coral_alignment_quality/test/python_classifier_interface_tests.cpp), restated:
# same fixture (X = 1..6, y = 0 0 0 1 1 1), same assertions.  The decision-tree and ROC-plot cases are not restated: the reference's fit()
# has the decision tree commented out (alignmentinterface.cpp:212-215) and the plot needs matplotlib, which this image lacks.
@pytest.fixture()
def python_classifier():
    clf = V.LogisticRegression()
    clf.AddDataPoint(np.array([1, 2, 3, 4, 5, 6.0]).reshape(6, 1), np.array([0, 0, 0, 1, 1, 1.0]))
    return clf


def test_logisticRegressionPredictTest(python_classifier):
    python_classifier.fit()
    y_pred = python_classifier.predict(np.array([3.0, 4.0]))
    assert y_pred[0] == 0 and y_pred[1] == 1


def test_logisticRegressionPredictProbaTest(python_classifier):
    assert np.array_equal(python_classifier.predict_proba(np.array([[3.0], [4.0]])), [0, 0])     # not fitted yet: zeros (:23-28)
    python_classifier.fit()
    y_prob = python_classifier.predict_proba(np.array([3.0, 4.0]))
    assert y_prob[0] < 0.5 < y_prob[1]
    sk = python_classifier.py_clf_.predict_proba(np.array([[3.0], [4.0]]))[:, 1]                   # numpy.delete(result, 0, 1) in the reference
    assert np.allclose(y_prob, sk, rtol=1e-12)


def test_accuracyTest(python_classifier):
    python_classifier.AddDataPoint(np.array([[2.0], [5.0]]), np.array([1.0, 0.0]))
    python_classifier.fit()
    assert 0.5 < python_classifier.Accuracy() < 1
    from sklearn.metrics import balanced_accuracy_score, confusion_matrix
    y, p = python_classifier.y_, python_classifier.predict(python_classifier.X_)
    assert python_classifier.Accuracy() == pytest.approx(balanced_accuracy_score(y, p))
    assert np.array_equal(python_classifier.ConfusionMatrix(), confusion_matrix(y, p))
    assert python_classifier.Accuracy([1, 0], [1]) == -1 and python_classifier.Accuracy([], []) == -1


def test_saveAndLoadDataTest(python_classifier, tmp_path):
    p = str(tmp_path / "training_data.txt")
    python_classifier.SaveData(p)
    assert open(p).read().splitlines()[:2] == ["0,1", "0,2"]
    loaded = V.LogisticRegression().LoadData(p)
    python_classifier.fit(); loaded.fit()
    assert loaded.predict_proba(np.array([3.0]))[0] == pytest.approx(python_classifier.predict_proba(np.array([3.0]))[0], rel=1e-6)   # EXPECT_FLOAT_EQ
    # coefficients written by one instance drive another without sklearn (LoadCoefficients is what the SLAM run uses)
    c = str(tmp_path / "coef.txt")
    python_classifier.SaveCoefficients(c)
    other = V.LogisticRegression().LoadCoefficients(c)
    # the file holds 6 significant digits, like the reference's `file << coef_(i)` (alignmentinterface.cpp:256-270)
    assert other.predict_proba(np.array([3.0]))[0] == pytest.approx(python_classifier.predict_proba(np.array([3.0]))[0], rel=1e-4)


def test_invalid_training_data_is_refused():
    clf = V.LogisticRegression()
    assert not clf.DataValid()
    with pytest.raises(ValueError):
        clf.fit()
    clf.AddDataPoint(np.array([[1.0], [np.nan]]), np.array([0.0, 1.0]))
    assert not clf.DataValid()
