"""CPU tests: the oracle (oracle/dis_oracle.c, a plain-C restatement of the reference hot path) is
pinned against the reference's only known-answer vector and against fixtures produced by the
reference's own code compiled verbatim (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import port, ref_driver

PARAMS2 = dict(lv_f=5, lv_l=3, maxiter=12, miniter=12, mindprate=0.05, mindrrate=0.95, minimgerr=0.0, patchsz=8,
               poverl=0.4, usefbcon=0, patnorm=1, costfct=0, usetvref=1, tv_alpha=10.0, tv_gamma=10.0,
               tv_delta=5.0, tv_innerit=1, tv_solverit=3, tv_sor=1.6, verbosity=0)


def bits_differ(x, y):
    return int((np.ascontiguousarray(x, np.float32).view(np.uint32) !=
                np.ascontiguousarray(y, np.float32).view(np.uint32)).sum())


def test_golden_flo_bit_exact(alley_pair, golden_dir):
    """kroeger/flows/alley_0001.flo == run_OF_INT frame_0001 frame_0002 at operating point 2."""
    a, b = alley_pair
    golden = np.load(os.path.join(golden_dir, "alley_0001_flo.npz"))["flow"]
    assert ref_driver.preset_params(a.shape[1], 2) == PARAMS2
    flow = port.run_u8(a, b, PARAMS2)
    assert flow.shape == golden.shape == (436, 1024, 2)
    assert bits_differ(flow, golden) == 0


def test_reference_fixtures_bit_exact(alley_pair, golden_dir):
    """Raw engine output of the verbatim-compiled reference on crops / other parameter sets."""
    a, b = alley_pair
    z = np.load(os.path.join(golden_dir, "ref_cases.npz"))
    names = sorted(k[:-5] for k in z.files if k.endswith("_flow"))
    assert len(names) >= 10
    for name in names:
        y0, y1, x0, x1 = z[name + "_crop"]
        p = ref_driver.parse_params(list(z[name + "_params"]))
        _, lvl = port.run_u8(a[y0:y1, x0:x1], b[y0:y1, x0:x1], p, want_level=True)
        assert lvl.shape == z[name + "_flow"].shape, name
        assert bits_differ(lvl, z[name + "_flow"]) == 0, name


def test_reference_fixtures_rgb_bit_exact(golden_dir):
    """Colour build (SELECTCHANNEL=3): raw engine output of the verbatim-compiled RGB reference."""
    z = np.load(os.path.join(golden_dir, "ref_cases_rgb.npz"))
    A, B = z["img_a"], z["img_b"]
    names = sorted(k[:-5] for k in z.files if k.endswith("_flow"))
    assert len(names) >= 8
    for name in names:
        p = ref_driver.parse_params(list(z[name + "_params"]))
        _, lvl = port.run_u8(A, B, p, want_level=True)
        assert lvl.shape == z[name + "_flow"].shape, name
        assert bits_differ(lvl, z[name + "_flow"]) == 0, name


def test_pyramid_matches_opencv(alley_pair):
    """P1 restatement vs the cv2 call sequence of ConstructImgPyramide (kroeger/run_dense.cpp:130-178)."""
    a, _ = alley_pair
    for crop, lv_f, ps in (((0, 436, 0, 1024), 5, 8), ((3, 275, 10, 490), 4, 12), ((0, 101, 0, 203), 2, 6)):
        img = a[crop[0]:crop[1], crop[2]:crop[3]]
        cvp = ref_driver.build_pyramids_cv2(img, dict(lv_f=lv_f, patchsz=ps))
        orp = port.build_pyramid(img, lv_f, ps)
        for k in range(3):
            for l in range(lv_f + 1):
                assert cvp[k][l].shape == orp[k][l].shape
                assert bits_differ(cvp[k][l], orp[k][l]) == 0, (k, l)


def test_finish_matches_opencv():
    """O1 restatement vs cv2.resize(INTER_LINEAR) + crop (kroeger/run_dense.cpp:407-414)."""
    rng = np.random.default_rng(5)
    fl = rng.standard_normal((28, 64, 2)).astype(np.float32) * 3
    p = dict(lv_f=5, lv_l=3)
    ref = ref_driver.finish_flow_cv2(fl, p, 500, 218)
    got = port.finish(fl, 3, *ref_driver.pad_geometry(500, 218, 5)[2:], 500, 218)
    assert ref.shape == got.shape
    assert np.abs(ref - got).max() <= 2e-6  # cv2's own arithmetic; tolerance, not bit-exact


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(port.__file__), "_ref", "libdis_ref.so")),
                    reason="oracle/_ref not built (needs /root/reference at build time)")
def test_verbatim_reference_build_matches_golden_and_port(alley_pair, golden_dir):
    a, b = alley_pair
    golden = np.load(os.path.join(golden_dir, "alley_0001_flo.npz"))["flow"]
    flow = ref_driver.run_dense_ref(a, b, PARAMS2)
    assert bits_differ(flow, golden) == 0
    p = dict(PARAMS2, lv_f=4, lv_l=1, patchsz=12, poverl=0.75, maxiter=16, miniter=16)
    A, B = a[100:356, 200:584], b[100:356, 200:584]
    r = ref_driver.run_dense_ref(A, B, p, full_res=False)
    _, o = port.run_u8(A, B, p, want_level=True)
    assert bits_differ(r, o) == 0
