"""GPU parity tests (B200): libdis_b200.so through its C-ABI vs the oracle on identical inputs.

Tolerance (BASELINE.json north_star): mean |dflow| <= 1e-3 px, max <= 1e-2 px outside a border
margin of patchsz*2^lv_l.  The kernels are built to reproduce the reference's float order, so the
tests additionally require bit-identical results wherever that has been achieved (everything
except nothing, so far): a weaker result would mean a reduction order or FMA slipped in."""
import os

import numpy as np
import pytest

import flowonthego_b200 as F
from flowonthego_b200 import api
from oracle import port, ref_driver
from tests.synth import synth_pair, synth_pair_bgr

pytestmark = pytest.mark.gpu

TOL_MEAN, TOL_MAX = 1e-3, 1e-2


def bits_differ(x, y):
    return int((np.ascontiguousarray(x, np.float32).view(np.uint32) !=
                np.ascontiguousarray(y, np.float32).view(np.uint32)).sum())


def assert_flow_parity(got, ref, patchsz, lv_l, exact=True):
    assert got.shape == ref.shape
    m = patchsz * (1 << lv_l)
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    inner = d[m:-m, m:-m] if d.shape[0] > 2 * m and d.shape[1] > 2 * m else d
    assert inner.mean() <= TOL_MEAN and inner.max() <= TOL_MAX, (inner.mean(), inner.max())
    if exact:
        assert bits_differ(got, ref) == 0


def params(preset, w, **kw):
    return F.Params.preset(preset, w, verbosity=0).copy(**kw)


def test_golden_flo_through_c_abi(alley_pair, golden_dir):
    """C1/C2 input, operating point 2: the reference's own golden file, bit for bit."""
    a, b = alley_pair
    golden = np.load(os.path.join(golden_dir, "alley_0001_flo.npz"))["flow"]
    flow = F.run_dense(a, b, None, 2)
    assert_flow_parity(flow, golden, 8, 3)


def test_c1_no_variational(alley_pair):
    a, b = alley_pair
    p = F.Params.from_argv("5 3 12 12 0.05 0.95 0 8 0.40 0 1 0 0 10 10 5 1 3 1.6 0".split())
    with F.Engine(p, 1024, 436) as e:
        flow = e.run_u8(a, b)
    assert_flow_parity(flow, port.run_u8(a, b, p.to_dict()), 8, 3)


def test_c2_preset4_crop(alley_pair):
    """C2 (128 Gauss-Newton iterations, lv_l=0) on a crop the scalar oracle finishes in seconds."""
    a, b = alley_pair
    A, B = a[60:260, 100:420], b[60:260, 100:420]
    p = params(4, 1024, lv_f=3, lv_l=0)
    with F.Engine(p, 320, 200) as e:
        flow = e.run_u8(A, B)
    assert_flow_parity(flow, port.run_u8(A, B, p.to_dict()), 12, 0)


def test_reference_fixtures(alley_pair, golden_dir):
    """Outputs of the verbatim-compiled reference engine (tests/golden/ref_cases.npz)."""
    a, b = alley_pair
    z = np.load(os.path.join(golden_dir, "ref_cases.npz"))
    names = sorted(k[:-5] for k in z.files if k.endswith("_flow"))
    ran = 0
    for name in names:
        y0, y1, x0, x1 = z[name + "_crop"]
        pd = ref_driver.parse_params(list(z[name + "_params"]))
        A, B = a[y0:y1, x0:x1], b[y0:y1, x0:x1]
        with F.Engine(F.Params.from_dict(pd), A.shape[1], A.shape[0]) as e:
            e.run_u8(A, B)
            lvl = e.level_flow(A.shape[1], A.shape[0])
        assert lvl.shape == z[name + "_flow"].shape, name
        assert bits_differ(lvl, z[name + "_flow"]) == 0, name
        ran += 1
    assert ran >= 13


def test_stage_taps_bit_exact():
    """Per-stage comparison: pyramid (P1), patch search (D1-D7), densify (A1), refinement (V1-V8)."""
    a, b, _ = synth_pair(250, 190, seed=3)  # odd size: exercises the divisibility padding
    p = params(3, 256, lv_f=3, lv_l=0)
    wp, hp, left, top = F.padded_size(250, 190, 3)
    pa, pb = port.build_pyramid(a, 3, 12), port.build_pyramid(b, 3, 12)
    fo, pf, dn = port.run_engine(pa, pb, wp, hp, p.to_dict(), taps=True)
    with F.Engine(p, 250, 190) as e:
        e.enable_taps(True)
        full = e.run_u8(a, b)
        for l in range(4):
            assert bits_differ(e.tap(api.TAP_IMG_A, l), pa[0][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_IMG_A_DX, l), pa[1][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_IMG_A_DY, l), pa[2][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_IMG_B, l), pb[0][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_PATCH_FLOW, l), pf[l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_FLOW_DENSE, l), dn[l].ravel()) == 0
        assert bits_differ(e.level_flow(250, 190), fo) == 0
    assert bits_differ(full, port.finish(fo, 0, left, top, 250, 190)) == 0


def test_product_pyramid_path_bit_exact():
    """Untapped runs build level lv_l straight from the u8 frames (k_block_mean) and skip the finer levels: the
    images and gradients the engine sees must still be those of ConstructImgPyramide, for every lv_l, odd sizes
    (divisibility padding -> clamped blocks) and colour."""
    for (w, h, lv_f, lv_l, ch) in ((250, 190, 3, 2, 1), (322, 198, 4, 1, 1), (256, 192, 3, 3, 1), (250, 190, 3, 2, 3),
                                   (1000, 600, 5, 4, 1)):
        a, b, _ = (synth_pair if ch == 1 else synth_pair_bgr)(w, h, seed=w)
        p = params(2, 1024, lv_f=lv_f, lv_l=lv_l)
        pa, pb = port.build_pyramid(a, lv_f, 8), port.build_pyramid(b, lv_f, 8)
        with F.Engine(p, w, h, channels=ch) as e:
            e.enable_taps(2)
            e.run_u8(a, b)
            for l in range(lv_l, lv_f + 1):
                assert bits_differ(e.tap(api.TAP_IMG_A, l), pa[0][l].ravel()) == 0, (w, l)
                assert bits_differ(e.tap(api.TAP_IMG_A_DX, l), pa[1][l].ravel()) == 0, (w, l)
                assert bits_differ(e.tap(api.TAP_IMG_A_DY, l), pa[2][l].ravel()) == 0, (w, l)
                assert bits_differ(e.tap(api.TAP_IMG_B, l), pb[0][l].ravel()) == 0, (w, l)


def test_engine_boundary_run_pyramids(alley_pair):
    """dis_run_pyramids == OFC::OFClass ctor (kroeger/oflow.h:84-111): caller-built (OpenCV) pyramids in,
    level-lv_l flow out; also through the Python OFClass mirror with the reference's argument list."""
    a, b = alley_pair
    pd = ref_driver.preset_params(1024, 2)
    pa, pb = ref_driver.build_pyramids_cv2(a, pd), ref_driver.build_pyramids_cv2(b, pd)
    ref = port.run_engine(pa, pb, 1024, 448, pd)
    with F.Engine(F.Params.from_dict(pd), 1024, 448) as e:
        got = e.run_pyramids(pa, pb, 1024, 448)
    assert bits_differ(got, ref) == 0
    out = np.zeros_like(ref)
    F.OFClass(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2], 8, out, None, 1024, 448, 5, 3, 12, 12, 0.05, 0.95, 0.0, 8,
              0.4, False, 0, 1, 1, True, 10.0, 10.0, 5.0, 1, 3, 1.6, 0)
    assert bits_differ(out, ref) == 0
    # initflow (kroeger/oflow.cpp:217-220): resolution of level lv_f+1
    init = (np.random.default_rng(1).standard_normal((7, 16, 2)) * 0.5).astype(np.float32)
    ref_i = port.run_engine(pa, pb, 1024, 448, pd, initflow=init)
    with F.Engine(F.Params.from_dict(pd), 1024, 448) as e:
        got_i = e.run_pyramids(pa, pb, 1024, 448, initflow=init)
    assert bits_differ(got_i, ref_i) == 0 and bits_differ(ref_i, ref) != 0


def test_cost_functions_and_patch_sizes():
    a, b, _ = synth_pair(224, 160, seed=11)
    for kw in (dict(costfct=1, maxiter=20, miniter=20), dict(costfct=2, maxiter=20, miniter=20), dict(patnorm=0),
               dict(patchsz=4, poverl=0.5), dict(patchsz=6, poverl=0.5), dict(patchsz=10, poverl=0.6),
               dict(patchsz=14, poverl=0.7), dict(patchsz=16, poverl=0.75),
               dict(miniter=2, maxiter=30, mindprate=0.2, mindrrate=0.9, minimgerr=1.0),
               dict(tv_innerit=2, tv_solverit=5, tv_sor=1.9), dict(tv_gamma=0.0), dict(tv_delta=0.0)):
        p = params(2, 224, lv_f=2, lv_l=0, **kw)
        with F.Engine(p, 224, 160) as e:
            flow = e.run_u8(a, b)
        ref = port.run_u8(a, b, p.to_dict())
        assert bits_differ(flow, ref) == 0, kw


def test_forward_backward_merging():
    """usefbcon=1 (kroeger/patchgrid.cpp:278-375, oflow.cpp:162-170,193-197,269-294)."""
    for (w, h, kw) in ((224, 160, dict(lv_f=2, lv_l=0)), (250, 190, dict(lv_f=3, lv_l=1, patchsz=12, poverl=0.75)),
                       (320, 200, dict(lv_f=3, lv_l=2, usetvref=0))):
        a, b, _ = synth_pair(w, h, seed=w + 1, shift=(6.5, -4.25), rot_deg=1.5)
        p = params(2, w, usefbcon=1, **kw)
        with F.Engine(p, w, h) as e:
            flow = e.run_u8(a, b)
        assert bits_differ(flow, port.run_u8(a, b, p.to_dict())) == 0, (w, h, kw)


def test_ragged_and_small_inputs():
    """Edge shapes: sizes that need padding on both axes, 30-wide coarsest level (the reference's
    stride != width case), minimal sizes, large displacements that push patches out of bounds."""
    for (w, h, lf, ll, ps) in ((501, 301, 3, 1, 8), (480, 272, 4, 3, 8), (97, 61, 1, 0, 8), (64, 40, 2, 0, 4)):
        a, b, _ = synth_pair(w, h, seed=w, shift=(9.5, -7.25), rot_deg=2.0)
        p = params(2, w, lv_f=lf, lv_l=ll, patchsz=ps)
        with F.Engine(p, w, h) as e:
            flow = e.run_u8(a, b)
        assert bits_differ(flow, port.run_u8(a, b, p.to_dict())) == 0, (w, h)


def test_finish_block_kernel_edges():
    """lv_l = 2 with padding offsets that are multiples of 4 takes the 4x4-block upsampling kernel: sizes whose last
    block column / row is partial, and sizes small enough that every block touches a border clamp."""
    for (w, h, lv_f) in ((323, 203, 2), (327, 207, 3), (320, 200, 2), (19, 15, 2), (324, 8, 2)):
        wp, hp, left, top = F.padded_size(w, h, lv_f)
        assert left % 4 == 0 and top % 4 == 0
        a, b, _ = synth_pair(w, h, seed=w + h)
        p = params(2, 1024, lv_f=lv_f, lv_l=2, patchsz=4 if min(w, h) < 32 else 8)
        try:
            eng = F.Engine(p, w, h)
        except F.DisError as ex:
            assert "too small" in str(ex)
            continue
        with eng as e:
            assert bits_differ(e.run_u8(a, b), port.run_u8(a, b, p.to_dict())) == 0, (w, h, lv_f)


def test_c3_full_size_1080p():
    """C3: 1920x1080 synthetic affine pair, preset 3 + variational; oracle comparison and EPE vs ground truth."""
    a, b, gt = synth_pair(1920, 1080, seed=1)
    p = params(3, 1920)
    assert (p.lv_f, p.lv_l) == (6, 2)
    with F.Engine(p, 1920, 1080) as e:
        flow = e.run_u8(a, b)
        again = e.run_u8(a, b)
    assert bits_differ(flow, again) == 0  # deterministic (no atomics anywhere)
    assert_flow_parity(flow, port.run_u8(a, b, p.to_dict()), 12, 2)
    epe = np.sqrt(((flow - gt) ** 2).sum(-1))[48:-48, 48:-48]
    assert epe.mean() < 0.5, epe.mean()


def test_c4_full_size_4k_properties():
    """C4a: 3840x2160, lv 7->0, 16 iterations, variational.  The scalar oracle needs about a minute at
    this size, so check size-independent properties: determinism, EPE vs the known affine flow, and
    agreement with the oracle on the coarse-to-fine prefix (levels 7..4 via taps, cheap for the oracle)."""
    a, b, gt = synth_pair(3840, 2160, seed=2)
    p = F.Params.from_argv("7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split())
    with F.Engine(p, 3840, 2160) as e:
        flow = e.run_u8(a, b)
        again = e.run_u8(a, b)
        assert bits_differ(flow, again) == 0
        e.enable_taps(True)
        e.run_u8(a, b)
        coarse = e.tap(api.TAP_FLOW_REFINED, 4).reshape(2176 >> 4, 3840 >> 4, 2)
    epe = np.sqrt(((flow - gt) ** 2).sum(-1))[64:-64, 64:-64]
    assert epe.mean() < 0.25, epe.mean()
    # same pair, engine stopped at level 4 (lv_l=4): identical prefix of the coarse-to-fine recursion
    p4 = p.copy(lv_l=4)
    _, lvl = port.run_u8(a, b, p4.to_dict(), want_level=True)
    assert bits_differ(coarse, lvl) == 0


def test_engine_reuse_across_sizes_and_params():
    """One handle, changing image sizes and parameter sets (re-planning drops the recorded CUDA graph)."""
    p = params(2, 320, lv_f=3, lv_l=1)
    with F.Engine(p, 320, 240) as e:
        for (w, h) in ((320, 240), (200, 136), (320, 240), (352, 260)):  # the last one exceeds the create-time size
            a, b, _ = synth_pair(w, h, seed=w + h)
            assert bits_differ(e.run_u8(a, b), port.run_u8(a, b, p.to_dict())) == 0, (w, h)
        p2 = p.copy(patchsz=12, poverl=0.75, lv_l=0, usetvref=0)
        e.set_params(p2)
        a, b, _ = synth_pair(320, 240, seed=9)
        assert bits_differ(e.run_u8(a, b), port.run_u8(a, b, p2.to_dict())) == 0
    with pytest.raises(F.DisError):
        F.Engine(params(2, 64, lv_f=6, lv_l=3), 64, 48)  # coarsest level would be 1x1


def test_execution_options_do_not_change_results():
    """DIS_OPT_SOR_GROUP (8 | 16), DIS_OPT_SOR_SMALL and DIS_OPT_USE_GRAPH only change how the work is scheduled."""
    a, b, _ = synth_pair(500, 300, seed=21)
    p = params(3, 512, lv_f=3, lv_l=0)
    ref = port.run_u8(a, b, p.to_dict())
    with F.Engine(p, 500, 300) as e:
        for grp in (8, 16):
            for graph in (1, 0):
                e.set_option(api.OPT_SOR_GROUP, grp)
                e.set_option(api.OPT_USE_GRAPH, graph)
                assert bits_differ(e.run_u8(a, b), ref) == 0, (grp, graph)
                assert bits_differ(e.run_u8(a, b), ref) == 0, (grp, graph)  # replay
        with pytest.raises(F.DisError):
            e.set_option(api.OPT_SOR_GROUP, 12)
        # OPT_SOR_SMALL: which levels take the one-CTA SOR (never, the default, every level it can hold: 320 rows -> 288 max)
        e.set_option(api.OPT_SOR_GROUP, 0)
        for small in (0, 2, 9, -1):
            e.set_option(api.OPT_SOR_SMALL, small)
            assert bits_differ(e.run_u8(a, b), ref) == 0, small
        with pytest.raises(F.DisError):
            e.set_option(api.OPT_SOR_SMALL, 10)
        # OPT_LEVEL_OUTPUT: the OFClass-style output; the caller's resize + crop (here: the oracle's) gives the same flow
        e.set_option(api.OPT_LEVEL_OUTPUT, 1)
        lvl = e.run_u8(a, b)
        wp, hp, left, top = F.padded_size(500, 300, 3)
        assert lvl.shape == (hp, wp, 2)
        assert bits_differ(port.finish(lvl, 0, left, top, 500, 300), ref) == 0
    p1 = params(2, 1024, lv_f=3, lv_l=1)
    with F.Engine(p1, 500, 300) as e:
        full = e.run_u8(a, b)
        e.set_option(api.OPT_LEVEL_OUTPUT, 1)
        lvl = e.run_u8(a, b)
        assert lvl.shape == (hp >> 1, wp >> 1, 2)
        assert bits_differ(lvl, e.level_flow(500, 300)) == 0
        assert bits_differ(port.finish(lvl, 1, left, top, 500, 300), full) == 0


def test_many_sor_sweeps_and_inner_iterations():
    """tv_solverit beyond 32 (the sweep index lives in the record tags) and several inner iterations per level."""
    a, b, _ = synth_pair(200, 140, seed=8)
    for kw in (dict(tv_solverit=40, tv_innerit=1), dict(tv_solverit=2, tv_innerit=4, tv_sor=1.2)):
        p = params(2, 1024, lv_f=2, lv_l=0, **kw)
        with F.Engine(p, 200, 140) as e:
            assert bits_differ(e.run_u8(a, b), port.run_u8(a, b, p.to_dict())) == 0, kw
    with pytest.raises(F.DisError):
        F.Engine(params(2, 1024, tv_solverit=300), 200, 140)


def test_sor_small_kernel_shapes():
    """k_sor_small (one CTA per pair, two columns per step): 1 ... 5 sweeps (fewer than three sweeps add fetch-only
    warps), odd widths (the last step of a row holds one pixel), levels narrower than a step pair, 1 ... 5 row blocks,
    several inner iterations (sweep 0 reads the previous launch's records), colour."""
    for (w, h, ch, kw) in ((61, 45, 1, dict(tv_solverit=1, tv_innerit=2)), (130, 97, 1, dict(tv_solverit=4)),
                           (35, 33, 1, dict(tv_solverit=3, tv_innerit=3, tv_sor=1.9)), (258, 150, 1, dict(tv_solverit=3)),
                           (22, 140, 1, dict(tv_solverit=2)), (99, 70, 3, dict(tv_solverit=5, tv_innerit=2))):
        a, b, _ = (synth_pair if ch == 1 else synth_pair_bgr)(w, h, seed=w + h)
        p = params(2, 1024, lv_f=1, lv_l=0, maxiter=6, miniter=6, **kw)
        ref = port.run_u8(a, b, p.to_dict())
        with F.Engine(p, w, h, channels=ch) as e:
            assert bits_differ(e.run_u8(a, b), ref) == 0, (w, h, kw)
            e.set_option(api.OPT_SOR_SMALL, 0)  # the wavefront pipeline on the same levels
            assert bits_differ(e.run_u8(a, b), ref) == 0, (w, h, kw, "wavefront")


@pytest.mark.parametrize("channels,usefbcon,batch", [(1, 0, 4), (1, 1, 3), (3, 0, 2)])
def test_batched_handle(channels, usefbcon, batch):
    """dis_create_batch: every launch serves `batch` pairs; each pair equals a separate run, also with fewer pairs
    than slots and through the single-pair entry points."""
    import torch
    w, h = 322, 198
    p = params(2, 1024, lv_f=3, lv_l=1, usefbcon=usefbcon)
    mk = synth_pair if channels == 1 else synth_pair_bgr
    pairs = [mk(w, h, seed=40 + k)[:2] for k in range(batch)]
    refs = [port.run_u8(x[0], x[1], p.to_dict()) for x in pairs]
    da = [torch.from_numpy(x[0]).cuda() for x in pairs]
    db = [torch.from_numpy(x[1]).cuda() for x in pairs]
    out = torch.zeros((batch, h, w, 2), dtype=torch.float32, device="cuda")
    with F.Engine(p, w, h, channels=channels, batch=batch) as e:
        for rep in range(2):  # capture, then replay
            out.zero_()
            e.submit_u8_device_batch([x.data_ptr() for x in da], [x.data_ptr() for x in db], w, h, w * channels,
                                     [out[k].data_ptr() for k in range(batch)])
            e.wait()
            for k in range(batch):
                assert bits_differ(out[k].cpu().numpy(), refs[k]) == 0, (rep, k)
        out.zero_()
        e.submit_u8_device_batch([da[-1].data_ptr()], [db[-1].data_ptr()], w, h, w * channels, [out[0].data_ptr()])
        e.wait()
        assert bits_differ(out[0].cpu().numpy(), refs[-1]) == 0
        assert float(out[1:].abs().sum()) == 0.0  # idle slots write to their own scratch
        assert bits_differ(e.run_u8(pairs[0][0], pairs[0][1]), refs[0]) == 0  # host-buffer entry point, slot 0
    with pytest.raises(F.DisError):
        F.Engine(p, w, h, batch=9)


def test_async_and_device_entry_points():
    import torch
    a, b, _ = synth_pair(320, 240, seed=4)
    p = params(2, 320, lv_f=3, lv_l=1)
    ref = port.run_u8(a, b, p.to_dict())
    with F.Engine(p, 320, 240) as e1, F.Engine(p, 320, 240) as e2:
        o1, o2 = F.pinned_empty((240, 320, 2), np.float32), F.pinned_empty((240, 320, 2), np.float32)
        e1.submit_u8(a, b, o1)
        e2.submit_u8(a, b, o2)
        e1.wait(), e2.wait()
        assert bits_differ(o1, ref) == 0 and bits_differ(o2, ref) == 0
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        dflow = torch.empty((240, 320, 2), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        e1.submit_u8_device(da.data_ptr(), db.data_ptr(), 320, 240, 320, dflow.data_ptr())
        e1.wait()
        assert bits_differ(dflow.cpu().numpy(), ref) == 0
        t = e1.timings()
        assert t["launches"] > 10


def test_run_dense_cli_reproduces_golden(tmp_path, golden_dir):
    """The compiled CLI with the reference's argv grammar (kroeger/README.md:48-88) on the fixture frames."""
    import subprocess
    exe = os.path.join(os.path.dirname(F.api.__file__), "run_dense")
    a, b = os.path.join(golden_dir, "alley_0001_gray.png"), os.path.join(golden_dir, "alley_0002_gray.png")
    golden = np.load(os.path.join(golden_dir, "alley_0001_flo.npz"))["flow"]
    out1, out2 = str(tmp_path / "v1.flo"), str(tmp_path / "v3.flo")
    subprocess.check_call([exe, a, b, out1])  # variant 1: operating point 2
    argv = "5 3 12 12 0.05 0.95 0 8 0.40 0 1 0 1 10 10 5 1 3 1.6 0".split()
    subprocess.check_call([exe, a, b, out2] + argv)  # variant 3: the same point, explicit
    for p in (out1, out2):
        assert os.path.getsize(p) == 3571724  # = the reference's golden file size
        assert bits_differ(F.read_flo(p), golden) == 0
    log = subprocess.run([exe, a, b, out1, "2"], capture_output=True, text=True).stdout
    assert "TIME (O.Flow Run-Time   ) (ms):" in log and "TIME (Sc: 3, #p:   448" in log


# ---- colour mode (the reference's run_OF_RGB build, SELECTCHANNEL=3) --------------------------------------

def test_rgb_reference_fixtures(golden_dir):
    """Outputs of the verbatim-compiled RGB reference engine (tests/golden/ref_cases_rgb.npz)."""
    z = np.load(os.path.join(golden_dir, "ref_cases_rgb.npz"))
    A, B = z["img_a"], z["img_b"]
    names = sorted(k[:-5] for k in z.files if k.endswith("_flow"))
    assert len(names) >= 8
    for name in names:
        pd = ref_driver.parse_params(list(z[name + "_params"]))
        with F.Engine(F.Params.from_dict(pd), A.shape[1], A.shape[0], channels=3) as e:
            e.run_u8(A, B)
            lvl = e.level_flow(A.shape[1], A.shape[0])
        assert lvl.shape == z[name + "_flow"].shape, name
        assert bits_differ(lvl, z[name + "_flow"]) == 0, name


def test_rgb_stage_taps_bit_exact():
    a, b, _ = synth_pair_bgr(250, 190, seed=5)
    p = params(3, 256, lv_f=3, lv_l=0)
    wp, hp, left, top = F.padded_size(250, 190, 3)
    pa, pb = port.build_pyramid(a, 3, 12), port.build_pyramid(b, 3, 12)
    fo, pf, dn = port.run_engine(pa, pb, wp, hp, p.to_dict(), taps=True, noc=3)
    with F.Engine(p, 250, 190, channels=3) as e:
        e.enable_taps(True)
        full = e.run_u8(a, b)
        for l in range(4):
            assert bits_differ(e.tap(api.TAP_IMG_A, l), pa[0][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_IMG_A_DX, l), pa[1][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_IMG_A_DY, l), pa[2][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_IMG_B, l), pb[0][l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_PATCH_FLOW, l), pf[l].ravel()) == 0
            assert bits_differ(e.tap(api.TAP_FLOW_DENSE, l), dn[l].ravel()) == 0
        assert bits_differ(e.level_flow(250, 190), fo) == 0
    assert bits_differ(full, port.finish(fo, 0, left, top, 250, 190)) == 0


def test_rgb_configs_and_engine_boundary():
    """Patch sizes / cost functions / fb merging in colour, and the OFClass constructor with noc=3."""
    a, b, gt = synth_pair_bgr(320, 200, seed=7)
    for kw in (dict(patchsz=4, poverl=0.5), dict(patchsz=8, costfct=1), dict(patchsz=12, poverl=0.75, costfct=2),
               dict(patchsz=14, poverl=0.5, patnorm=0), dict(patchsz=16, poverl=0.5, usefbcon=1)):
        p = params(2, 1024, lv_f=3, lv_l=1, **kw)
        with F.Engine(p, 320, 200, channels=3) as e:
            flow = e.run_u8(a, b)
        assert_flow_parity(flow, port.run_u8(a, b, p.to_dict()), p.patchsz, 1)
    m = 16
    assert np.abs(flow - gt)[m:-m, m:-m].mean() < 0.25
    p = params(2, 1024, lv_f=3, lv_l=1)
    pa, pb = port.build_pyramid(a, 3, 8), port.build_pyramid(b, 3, 8)
    out = np.zeros((100, 160, 2), np.float32)
    F.OFClass(*pa, *pb, 8, out, None, 320, 200, 3, 1, p.maxiter, p.miniter, p.mindprate, p.mindrrate, p.minimgerr,
              8, p.poverl, False, 0, 3, 1, True, p.tv_alpha, p.tv_gamma, p.tv_delta, p.tv_innerit, p.tv_solverit,
              p.tv_sor, 0)
    assert bits_differ(out, port.run_engine(pa, pb, 320, 200, p.to_dict(), noc=3)) == 0
    with pytest.raises(ValueError):
        with F.Engine(p, 320, 200, channels=3) as e:
            e.run_u8(a[..., 0], b[..., 0])  # grey input on a colour engine


def test_rgb_cli(tmp_path, golden_dir):
    """run_dense_rgb (= the reference's run_OF_RGB binary): native BGR decode of a PPM + colour engine."""
    import subprocess
    exe = os.path.join(os.path.dirname(F.api.__file__), "run_dense_rgb")
    z = np.load(os.path.join(golden_dir, "ref_cases_rgb.npz"))
    A, B = z["img_a"], z["img_b"]
    names = []
    for name, img in (("a.ppm", A), ("b.ppm", B)):
        names.append(str(tmp_path / name))
        with open(names[-1], "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
            f.write(np.ascontiguousarray(img[..., ::-1]).tobytes())  # PPM stores RGB
        assert np.array_equal(F.read_image_bgr(names[-1]), img)
    argv = "3 1 12 12 0.05 0.95 0 8 0.40 0 1 0 1 10 10 5 1 3 1.6 0".split()
    out = str(tmp_path / "rgb.flo")
    subprocess.check_call([exe, names[0], names[1], out] + argv)
    pd = ref_driver.parse_params(argv)
    assert bits_differ(F.read_flo(out), port.run_u8(A, B, pd)) == 0
    assert bits_differ(F.run_dense(names[0], names[1], None, *argv, channels=3), port.run_u8(A, B, pd)) == 0


def test_rgb_full_size_1080p():
    """Colour at the benchmark size: 1920x1080 BGR pair, preset 3 + variational, against the oracle."""
    a, b, gt = synth_pair_bgr(1920, 1080, seed=4)
    p = params(3, 1920)
    with F.Engine(p, 1920, 1080, channels=3) as e:
        flow = e.run_u8(a, b)
        assert bits_differ(flow, e.run_u8(a, b)) == 0
    assert_flow_parity(flow, port.run_u8(a, b, p.to_dict()), 12, 2)
    epe = np.sqrt(((flow - gt) ** 2).sum(-1))[48:-48, 48:-48]
    assert epe.mean() < 0.5, epe.mean()


def test_randomized_configurations():
    """Seeded sweep over sizes (incl. non-multiples of 2^lv_f), patch sizes, overlaps, iteration settings, cost
    functions, forward-backward merging, refinement settings and channel counts: every result bit-identical."""
    rng = np.random.default_rng(20261017)
    ran = 0
    for case in range(28):
        ch = 3 if case % 4 == 3 else 1
        w, h = int(rng.integers(70, 420)), int(rng.integers(60, 300))
        patchsz = int(rng.choice([4, 6, 8, 10, 12, 14, 16]))
        lv_f = int(rng.integers(0, 4))
        while lv_f > 0 and (min(w, h) >> lv_f) < max(8, patchsz // 2 + 2):  # keep the coarsest level meaningful
            lv_f -= 1
        lv_l = int(rng.integers(0, lv_f + 1))
        maxit = int(rng.integers(1, 20))
        kw = dict(lv_f=lv_f, lv_l=lv_l, patchsz=patchsz, poverl=float(rng.choice([0.0, 0.3, 0.5, 0.75, 0.9])),
                  maxiter=maxit, miniter=int(rng.integers(0, maxit + 1)), mindprate=float(rng.choice([0.05, 0.2, 0.5])),
                  mindrrate=float(rng.choice([0.95, 0.8])), minimgerr=float(rng.choice([0.0, 0.5, 2.0])),
                  usefbcon=int(rng.integers(0, 2)), patnorm=int(rng.integers(0, 2)), costfct=int(rng.integers(0, 3)),
                  usetvref=int(rng.integers(0, 4) > 0), tv_alpha=float(rng.choice([10.0, 3.0, 30.0])),
                  tv_gamma=float(rng.choice([10.0, 0.0, 5.0])), tv_delta=float(rng.choice([5.0, 0.0, 1.0])),
                  tv_innerit=int(rng.integers(0, 3)), tv_solverit=int(rng.integers(1, 6)),
                  tv_sor=float(rng.choice([1.6, 1.0, 1.9])))
        p = params(2, 1024, **kw)
        a, b, _ = (synth_pair if ch == 1 else synth_pair_bgr)(w, h, seed=100 + case, shift=(float(rng.uniform(-6, 6)),
                                                                                              float(rng.uniform(-4, 4))))
        try:
            eng = F.Engine(p, w, h, channels=ch)
        except F.DisError as ex:  # e.g. coarsest level too small for this patch size: the engine says so
            assert "too small" in str(ex), (case, kw, str(ex))
            continue
        with eng as e:
            got = e.run_u8(a, b)
        ref = port.run_u8(a, b, p.to_dict())
        assert bits_differ(got, ref) == 0, (case, ch, w, h, kw)
        ran += 1
    assert ran >= 20


def test_pageable_buffers_and_row_pitch():
    """dis_submit_u8 / dis_run_u8 with ordinary (non-pinned) host memory and a row pitch larger than the width."""
    w, h, pitch = 333, 201, 352
    a, b, _ = synth_pair(w, h, seed=21)
    big_a, big_b = np.full((h, pitch), 77, np.uint8), np.full((h, pitch), 99, np.uint8)
    big_a[:, :w], big_b[:, :w] = a, b
    va, vb = big_a[:, :w], big_b[:, :w]          # views: strides (352, 1)
    assert va.strides == (pitch, 1) and not va.flags["C_CONTIGUOUS"]
    p = params(2, 1024, lv_f=3, lv_l=1)
    ref = port.run_u8(a, b, p.to_dict())
    out = np.empty((h, w, 2), np.float32)        # pageable result buffer
    with F.Engine(p, w, h) as e:
        assert bits_differ(e.run_u8(va, vb, out), ref) == 0
        e.submit_u8(va, vb, out)                 # asynchronous form, same buffers
        assert bits_differ(e.wait(), ref) == 0
        with pytest.raises(F.DisError):          # pitch smaller than a row
            api._check(F.lib().dis_run_u8(e._h, va.ctypes.data, vb.ctypes.data, w, h, w - 1, api._as_fp(out)), e._h)


def test_two_processes_share_one_gpu(tmp_path):
    """Two host processes, each with its own handle and context on the same device, run concurrently."""
    import subprocess
    import sys
    script = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import flowonthego_b200 as F\n"
        "from oracle import port\n"
        "from tests.synth import synth_pair\n"
        "seed = int(sys.argv[1])\n"
        "a, b, _ = synth_pair(320, 200, seed=seed)\n"
        "p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=3, lv_l=1)\n"
        "ref = port.run_u8(a, b, p.to_dict())\n"
        "with F.Engine(p, 320, 200) as e:\n"
        "    for _ in range(20):\n"
        "        got = e.run_u8(a, b)\n"
        "        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))\n"
        "print('ok', seed)\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = [subprocess.Popen([sys.executable, "-c", script, str(s)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for s in (31, 32)]
    for s, pr in zip((31, 32), procs):
        out, err = pr.communicate(timeout=300)
        assert pr.returncode == 0 and out.strip() == "ok %d" % s, err[-2000:]


def test_many_concurrent_handles_stay_bit_exact():
    """Stress for the SOR hand-off (16-byte {du, dv, tag} records published without fences): many
    handles run concurrently, so that wavefront warps of different pairs interleave on every SM; every result must still
    be bit-identical to the oracle.  A torn record would show up as a differing (or non-deterministic) flow."""
    import torch
    w, h, n = 334, 210, 24
    p = params(2, 1024, lv_f=3, lv_l=0, tv_solverit=5, tv_innerit=2)
    pairs = [synth_pair(w, h, seed=70 + k)[:2] for k in range(4)]
    refs = [port.run_u8(a, b, p.to_dict()) for a, b in pairs]
    da = [torch.from_numpy(x[0]).cuda() for x in pairs]
    db = [torch.from_numpy(x[1]).cuda() for x in pairs]
    engines = [F.Engine(p, w, h) for _ in range(n)]
    out = torch.zeros((n, h, w, 2), dtype=torch.float32, device="cuda")
    try:
        for rep in range(6):
            for i, e in enumerate(engines):
                k = (i + rep) % 4
                e.submit_u8_device(da[k].data_ptr(), db[k].data_ptr(), w, h, w, out[i].data_ptr())
            for i, e in enumerate(engines):
                e.wait()
            got = out.cpu().numpy()
            for i in range(n):
                assert bits_differ(got[i], refs[(i + rep) % 4]) == 0, (rep, i)
    finally:
        for e in engines:
            e.close()
