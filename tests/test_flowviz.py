"""Evaluation tools behind the path: Middlebury colour coding and endpoint error.
CPU part: the numpy restatement (oracle/flowcolor.py) against the reference's known-answer pair
alley_0001.flo -> alley_0001.png and against the verbatim build of colorcode.cpp.  GPU part: the CUDA
kernels through the C-ABI against both."""
import os
import subprocess

import cv2
import numpy as np
import pytest

import flowonthego_b200 as F
from oracle import flowcolor as fc


def cases():
    rng = np.random.default_rng(11)
    f = (rng.standard_normal((97, 131, 2)) * 4).astype(np.float32)
    f[3, 4] = 1e10            # unknown flow (flowIO.h UNKNOWN_FLOW)
    f[5, 6, 1] = np.nan
    f[7, 7] = 0
    f[9, 9] = (-2.5, 0.0)      # on the atan2 branch cut
    yield "random", f, -1.0
    yield "out_of_range", f, 3.0   # maxmotion below the largest motion: the x0.75 branch
    yield "zero", np.zeros((8, 9, 2), np.float32), -1.0
    unk = np.full((4, 5, 2), 1e10, np.float32)
    yield "all_unknown", unk, -1.0


@pytest.fixture(scope="module")
def golden(golden_dir):
    flow = np.load(os.path.join(golden_dir, "alley_0001_flo.npz"))["flow"]
    png = cv2.imread(os.path.join(golden_dir, "alley_0001_color.png"), cv2.IMREAD_COLOR)
    return flow, png


def test_color_oracle_matches_reference_png(golden):
    flow, png = golden
    col, maxrad = fc.motion_to_color(flow)
    assert np.array_equal(col, png)
    assert abs(maxrad - 7.1662478) < 1e-5


def test_color_oracle_matches_verbatim_colorcode(golden):
    if not fc.ref_available():
        pytest.skip("oracle/_ref/libcolor_ref.so not built (no reference tree)")
    flow, png = golden
    assert np.array_equal(fc.motion_to_color_ref(flow)[0], png)
    # computeColor calls the host libm's atan2f (C++ overload), which glibc 2.39 does not round correctly;
    # the oracle uses the correctly rounded float arctangent.  The two may therefore disagree by one grey
    # level on a few pixels per million (measured here: 23 of 12e6) -- never more.
    rng = np.random.default_rng(3)
    big = (rng.standard_normal((600, 700, 2)) * 5).astype(np.float32)
    for name, f, mm in list(cases()) + [("big", big, -1.0), ("big_clip", big, 2.0)]:
        a, ra = fc.motion_to_color(f, mm)
        b, rb = fc.motion_to_color_ref(f, mm)
        d = np.abs(a.astype(int) - b.astype(int))
        assert d.max() <= 1 and (d > 0).any(-1).mean() <= 1e-4, name
        assert ra == rb, name


@pytest.mark.gpu
def test_gpu_color_matches_golden_and_oracle(golden):
    flow, png = golden
    col, st = F.flow_to_color(flow, want_stats=True)
    assert np.array_equal(col, png)
    assert st["maxrad"] == np.float32(fc.motion_to_color(flow)[1])
    assert st["minu"] == flow[..., 0].min() and st["maxv"] == flow[..., 1].max()
    for name, f, mm in cases():
        assert np.array_equal(F.flow_to_color(f, mm), fc.motion_to_color(f, mm)[0]), name


@pytest.mark.gpu
def test_gpu_epe(golden):
    flow, _ = golden
    rng = np.random.default_rng(5)
    other = flow + (rng.standard_normal(flow.shape) * 0.3).astype(np.float32)
    other[10, 10] = 1e10
    for margin in (0, 24):
        mean, mx, cnt = F.flow_epe(flow, other, margin)
        rm, rx, rc = fc.epe(flow, other, margin)
        assert cnt == rc
        assert abs(mean - rm) <= 1e-12 * max(1.0, rm) and mx == pytest.approx(rx, rel=1e-6)
    assert F.flow_epe(flow, flow) == (0.0, 0.0, flow.shape[0] * flow.shape[1])
    with pytest.raises(F.DisError):
        F.flow_epe(flow[:20, :20], flow[:20, :20], margin=10)


@pytest.mark.gpu
def test_color_flow_and_epe_cli(tmp_path, golden):
    """color_flow [-quiet] in.flo out.png [maxmotion] (flow_code/C/color_flow.cpp) and flow_epe."""
    flow, png = golden
    bindir = os.path.dirname(F.api.__file__)
    flo, out = str(tmp_path / "a.flo"), str(tmp_path / "a.png")
    F.write_flo(flo, flow)
    r = subprocess.run([os.path.join(bindir, "color_flow"), flo, out], capture_output=True, text=True, check=True)
    u, v = flow[..., 0], flow[..., 1]
    assert r.stdout == "max motion: %.4f  motion range: u = %.3f .. %.3f;  v = %.3f .. %.3f\n" % (
        fc.motion_to_color(flow)[1], u.min(), u.max(), v.min(), v.max())
    assert "normalizing by 7.16625" in r.stderr
    assert np.array_equal(cv2.imread(out, cv2.IMREAD_COLOR), png)      # decoded by OpenCV
    assert np.array_equal(F.read_image_bgr(out), png)                  # and by the native reader
    out2 = str(tmp_path / "b.png")
    subprocess.run([os.path.join(bindir, "color_flow"), "-quiet", flo, out2, "3.5"], check=True)
    assert np.array_equal(cv2.imread(out2, cv2.IMREAD_COLOR), fc.motion_to_color(flow, 3.5)[0])
    assert subprocess.run([os.path.join(bindir, "color_flow"), flo]).returncode != 0
    flo2 = str(tmp_path / "b.flo")
    F.write_flo(flo2, flow + np.float32(0.5))
    r = subprocess.run([os.path.join(bindir, "flow_epe"), flo, flo2, "8"], capture_output=True, text=True, check=True)
    assert r.stdout.startswith("EPE mean 0.70710678")
