"""Full-size parity of every BASELINE.json config against the verbatim-compiled reference engine.

tests/golden/full_digests.npz (written by tests/golden/make_golden_full.py where /root/reference exists) holds,
for C2, C3, four pairs of the C5 stream, C4a and C4b: sha256 of the reference's raw level-lv_l flow (the
OFC::OFClass output), every 16th pixel of it, and sha256 of both input frames.  Inputs: the committed grey
first frames (alley, road_HD, yosemite_4k) and second frames derived with tests/synth (cv2.warpAffine on u8,
fixed-point) -- their hashes are checked first so that a different OpenCV build fails loudly, not as "parity".

CPU part: the C restatement (oracle/dis_oracle.c) reproduces the digests (C2, C3, one C5 pair).
GPU part: libdis_b200.so reproduces all of them bit for bit through the C-ABI.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import port, ref_driver
from tests import synth

HERE = os.path.dirname(os.path.abspath(__file__))
DIG = os.path.join(HERE, "golden", "full_digests.npz")
C5_PAIRS = (0, 1, 37, 63)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_case(key):
    """-> (a_u8, b_u8, params dict, digest dict); input hashes verified."""
    z = np.load(DIG)
    cfg = key.split("_")[0]
    if cfg == "c2":
        a, b = synth.load_gray("alley_0001_gray.png"), synth.load_gray("alley_0002_gray.png")
    elif cfg == "c3":
        a = synth.load_gray("road_HD_gray.png")
        b = synth.warp(a, synth.affine(a.shape[1], a.shape[0]))
    elif cfg in ("c4a", "c4b"):
        a = synth.load_gray("yosemite_4k_gray.png")
        b = synth.warp(a, synth.affine(a.shape[1], a.shape[0]))
    else:
        k = int(key.split("_")[1])
        base = synth.load_gray("road_HD_gray.png")
        a, b = synth.c5_frame(base, k), synth.c5_frame(base, k + 1)
    want = [str(x) for x in z[key + "_in_sha"]]
    assert [sha(a), sha(b)] == want, "input frames of %s differ from the ones the reference fixture was made from " \
                                     "(different OpenCV warpAffine / PNG decode?)" % key
    p = ref_driver.parse_params(str(z[key + "_argv"]).split())
    return a, b, p, dict(sha=str(z[key + "_sha"]), sub=z[key + "_sub"], shape=tuple(z[key + "_shape"]))


def check(level_flow, dig, key):
    assert tuple(level_flow.shape) == dig["shape"], key
    sub = np.ascontiguousarray(level_flow[::16, ::16])
    nd = int((sub.view(np.uint32) != dig["sub"].view(np.uint32)).sum())
    assert nd == 0, "%s: %d of %d sampled values differ from the reference (max |d| %.3g)" % (
        key, nd, sub.size, float(np.abs(sub - dig["sub"]).max()))
    assert sha(level_flow) == dig["sha"], key + ": sampled pixels agree but the sha256 of the whole field does not"


def test_digest_file_complete():
    z = np.load(DIG)
    for key in ["c2", "c3", "c4a", "c4b"] + ["c5_%d" % k for k in C5_PAIRS]:
        for s in ("_sha", "_sub", "_shape", "_in_sha", "_argv"):
            assert key + s in z.files, key + s


@pytest.mark.parametrize("key", ["c3", "c5_37", "c2"])
def test_oracle_restatement_full_size(key):
    """oracle/dis_oracle.c == verbatim reference build at the real size of the config."""
    a, b, p, dig = load_case(key)
    _, lvl = port.run_u8(a, b, p, want_level=True)
    check(lvl, dig, key)


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["c2", "c3", "c5_0", "c5_1", "c5_37", "c5_63", "c4a", "c4b"])
def test_gpu_full_size(key):
    """Every BASELINE config at full size through the C-ABI: bit-identical to the reference's OFClass output."""
    import flowonthego_b200 as F
    a, b, p, dig = load_case(key)
    with F.Engine(F.Params.from_dict(p), a.shape[1], a.shape[0]) as e:
        full = e.run_u8(a, b)
        lvl = e.level_flow(a.shape[1], a.shape[0])
        check(lvl, dig, key)
        assert full.shape == a.shape + (2,) and np.isfinite(full).all()
        if key == "c3":  # the same pair again through the graph replay, and EPE vs the known affine motion
            e.run_u8(a, b)
            check(e.level_flow(a.shape[1], a.shape[0]), dig, key + " (replay)")
            gt = synth.gt_flow(a.shape[1], a.shape[0], synth.affine(a.shape[1], a.shape[0]))
            m = 48
            epe = np.sqrt(((full - gt)[m:-m, m:-m] ** 2).sum(-1)).mean()
            assert epe < 1.0, epe  # the reference scores 0.63 px on this pair


@pytest.mark.gpu
def test_gpu_c5_stream_batched_and_video():
    """The bench's two arms on real C5 pairs: batched handles (device-resident) and the video front end (level
    output) both reproduce the reference digests."""
    import torch
    import flowonthego_b200 as F
    base = synth.load_gray("road_HD_gray.png")
    z = np.load(DIG)
    p = F.Params.from_argv(str(z["c5_0_argv"]).split())
    frames = [synth.c5_frame(base, k) for k in range(3)]
    digs = [dict(sha=str(z["c5_%d_sha" % k]), sub=z["c5_%d_sub" % k], shape=tuple(z["c5_%d_shape" % k])) for k in (0, 1)]
    with F.FlowStream(p, 1920, 1080, depth=2) as s:
        flows = list(s.flows(frames))
    for k in (0, 1):
        check(flows[k], digs[k], "video c5_%d" % k)
    d = [torch.from_numpy(f).cuda() for f in frames]
    out = torch.empty((2, 1080, 1920, 2), dtype=torch.float32, device="cuda")
    with F.Engine(p, 1920, 1080, batch=2) as e:
        stage = torch.zeros((2,) + e.level_flow_shape(), dtype=torch.float32, device="cuda")
        e.set_level_export([stage[0].data_ptr(), stage[1].data_ptr()])  # what bench.py hands to the NCCL gather
        for rep in range(2):  # capture, then graph replay
            stage.zero_()
            torch.cuda.synchronize()
            e.submit_u8_device_batch([d[0].data_ptr(), d[1].data_ptr()], [d[1].data_ptr(), d[2].data_ptr()], 1920, 1080, 1920,
                                     [out[0].data_ptr(), out[1].data_ptr()])
            e.wait()
            for k in (0, 1):
                check(stage[k].cpu().numpy(), digs[k], "batched c5_%d" % k)
        e.set_level_export([])
        stage.zero_()
        torch.cuda.synchronize()
        e.submit_u8_device_batch([d[0].data_ptr()], [d[1].data_ptr()], 1920, 1080, 1920, [out[0].data_ptr()])
        e.wait()
        assert float(stage.abs().max()) == 0.0
        e.copy_level_flow_device(0, stage[1].data_ptr())
        e.wait()
        check(stage[1].cpu().numpy(), digs[0], "copy_level_flow_device")
