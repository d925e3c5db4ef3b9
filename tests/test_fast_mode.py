"""Tolerance mode (DIS_OPT_ARITH = 1): the search and refinement kernels compiled with FMA contraction.

It is NOT the parity claim (the exact engine is, bit for bit); what is asserted here is the tolerance BASELINE.json's
north_star states -- mean |dflow| <= 1e-3 px and max <= 1e-2 px outside a border margin of patchsz * 2^lv_l -- against
the oracle on the 1080p configs with 16 Gauss-Newton iterations (C3, a C5 pair), what it does at 4K (C4a: the mean is
met, the max is not), and that the mode is refused for the 128-iteration operating points, where rounding differences
grow to 0.25 px."""
import numpy as np
import pytest

import flowonthego_b200 as F
from flowonthego_b200 import api
from oracle import port
from tests.test_full_size import check, load_case

pytestmark = pytest.mark.gpu

TOL_MEAN, TOL_MAX = 1e-3, 1e-2


def within_tolerance(got, ref, patchsz, lv_l):
    m = patchsz * (1 << lv_l)
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))[m:-m, m:-m]
    return float(d.mean()), float(d.max())


@pytest.mark.parametrize("key", ["c3", "c5_37"])
def test_fast_mode_within_tolerance_of_the_oracle(key):
    a, b, p, dig = load_case(key)
    ref = port.run_u8(a, b, p)  # the oracle, full resolution
    with F.Engine(F.Params.from_dict(p), a.shape[1], a.shape[0]) as e:
        exact = e.run_u8(a, b).copy()
        check(e.level_flow(a.shape[1], a.shape[0]), dig, key)          # exact mode: bit-identical, as ever
        e.set_option(api.OPT_ARITH, 1)
        fast = e.run_u8(a, b).copy()
        e.set_option(api.OPT_ARITH, 0)
        again = e.run_u8(a, b)
    assert np.array_equal(exact.view(np.uint32), again.view(np.uint32))  # switching back restores the exact engine
    assert np.array_equal(exact.view(np.uint32), ref.view(np.uint32))
    mean, mx = within_tolerance(fast, ref, p["patchsz"], p["lv_l"])
    assert mean <= TOL_MEAN and mx <= TOL_MAX, (mean, mx)
    assert not np.array_equal(fast.view(np.uint32), ref.view(np.uint32))  # it really is another arithmetic


def test_fast_mode_4k_characterised():
    """C4a (3840x2160, lv 7->0): the exact engine reproduces the reference digest.  The fast engine does NOT meet the
    max tolerance here -- measured: mean 1.7e-4 px, but 0.5 % of the pixels differ by more than 1e-2 px (max 0.17 px),
    in the low-texture regions where the Gauss-Newton steps are ill-conditioned.  The test pins that characterisation
    (mean within tolerance, outliers below 1 %); DESIGN.md states that the tolerance claim of the mode ends at 1080p."""
    a, b, p, dig = load_case("c4a")
    with F.Engine(F.Params.from_dict(p), a.shape[1], a.shape[0]) as e:
        exact = e.run_u8(a, b).copy()
        check(e.level_flow(a.shape[1], a.shape[0]), dig, "c4a")
        e.set_option(api.OPT_ARITH, 1)
        fast = e.run_u8(a, b)
    m = p["patchsz"] << p["lv_l"]
    d = np.abs(fast.astype(np.float64) - exact.astype(np.float64))[m:-m, m:-m]
    assert d.mean() <= TOL_MEAN, d.mean()
    assert (d > TOL_MAX).mean() < 0.01 and d.max() < 1.0, ((d > TOL_MAX).mean(), d.max())


def test_fast_mode_refused_for_long_iterations():
    a, b, p, _ = load_case("c2")  # 128 Gauss-Newton iterations
    with F.Engine(F.Params.from_dict(p), a.shape[1], a.shape[0]) as e:
        e.set_option(api.OPT_ARITH, 1)
        with pytest.raises(F.DisError):
            e.run_u8(a, b)
        e.set_option(api.OPT_ARITH, 0)
        e.run_u8(a, b)
