"""CPU: pin the kriging oracle against reference outputs (tests/golden/kriging.npz).
The reference's own tests hold no numeric kriging goldens,
so the live-reference vectors are the only pin."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

from oracle import kriging as ok


def test_kriging_oracle_matches_reference_golden(golden_kriging):
    g = golden_kriging
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        t = g.out(case, "targets")
        z, sigma, status = ok.krige(
            c, v, t, case["model"], case["effective_range"], case["sill"], case["nugget"],
            case["shape"], case["min_points"], case["max_points"])
        zr, sr = g.out(case, "z"), g.out(case, "sigma")
        assert_array_equal(np.isnan(z), np.isnan(zr), err_msg=str(case))
        assert_array_equal(np.isnan(sigma), np.isnan(sr), err_msg=str(case))
        assert int((status == 1).sum()) == case["no_points_error"]
        # the reference evaluates the model through numba scalar calls, the oracle
        # through numpy ufuncs: last-bit differences amplified by cond(A) ~ 1e3..1e5
        assert_allclose(z, zr, rtol=1e-9, equal_nan=True, err_msg=str(case))
        assert_allclose(sigma, sr, rtol=1e-9, atol=1e-9 * case["sill"], equal_nan=True,
                        err_msg=str(case))


def test_pancake_kriging_known_answer(golden_kriging):
    """BASELINE.md section 3."""
    g = golden_kriging
    case = g.manifest[0]
    assert_allclose(g.out(case, "z")[:4], [216.82668695078203, 213.25019510789258,
                                           208.75398093840414, 204.76546428916612], rtol=1e-12)
    assert_allclose(g.out(case, "sigma")[:4], [821.5990160593558, 607.4358198357556,
                                               457.6864840412478, 250.3293362582594], rtol=1e-12)


def test_jacknife_oracle_matches_reference_golden(golden_cv):
    """util/cross_validation.py through the oracle vs the live reference's deviations."""
    g = golden_cv
    for case in g.manifest:
        c, v = g.inputs(case["dataset"])
        idx = g.out(case, "indices")[:60]
        dev = ok.jacknife_deviations(c, v, idx, case["model"], case["effective_range"],
                                     case["sill"], case["nugget"])
        assert_allclose(dev, g.out(case, "deviations")[:60], rtol=1e-8, atol=1e-9, err_msg=str(case))


def test_oracle_matern_and_estimate_mode_match_reference(golden_kriging_ext):
    """Matern model and mode='estimate' restatements pinned to reference outputs
    (oracle/make_golden.py::make_kriging_ext)."""
    g = golden_kriging_ext
    c, v = g.inputs("rnd1000")
    t = g._z["targets"]
    for case in g.manifest:
        z, s, st = ok.krige(c, v, t, case["model"], case["effective_range"], case["sill"], case["nugget"],
                            case["shape"], 4, case["max_points"], mode=case["mode"],
                            precision=case.get("precision", 100))
        assert (st == 0).all()
        assert_allclose(z, g.out(case, "z"), rtol=1e-9)
        assert_allclose(s, g.out(case, "sigma"), rtol=1e-9, atol=1e-9 * case["sill"])
    # the Matern semivariance itself (reference models.matern on a grid of h and smoothness)
    hs = g._z["matern/h"]
    for k, sm in enumerate(g._z["matern/s"]):
        assert_allclose(ok.matern(hs, 200.0, 3.0, sm, 0.1), g._z["matern/%d" % k], rtol=1e-12)
