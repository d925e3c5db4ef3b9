"""Video-stream front end (dis_video_*, FlowStream, run_dense_stream): every flow of the pipelined sequence
must be bit-identical to a separate run on that pair."""
import os
import subprocess

import numpy as np
import pytest

import flowonthego_b200 as F
from oracle import port
from tests.synth import affine, texture, warp

pytestmark = pytest.mark.gpu


def sequence(w, h, n, seed=2, channels=1):
    """n frames of a texture under a growing affine motion."""
    if channels == 1:
        base = texture(w, h, seed)
        return [warp(base, affine(w, h, rot_deg=0.1 * k, scale=1 + 0.001 * k, shift=(0.8 * k, -0.5 * k)))
                for k in range(n)]
    chans = [sequence(w, h, n, seed + 10 * c, 1) for c in range(3)]
    return [np.ascontiguousarray(np.stack([chans[c][k] for c in range(3)], -1)) for k in range(n)]


def bits_differ(x, y):
    return int((np.ascontiguousarray(x, np.float32).view(np.uint32) !=
                np.ascontiguousarray(y, np.float32).view(np.uint32)).sum())


@pytest.mark.parametrize("channels,depth,reuse", [(1, 3, None), (1, 1, None), (3, 2, None), (1, 3, False), (1, 12, True)])
def test_stream_equals_pairwise(channels, depth, reuse):
    """Every flow of the pipelined sequence equals a separate run on that pair -- with and without pyramid reuse
    between consecutive pairs (default: on for 2 <= depth <= 8)."""
    w, h, n = 322, 198, 8 if depth < 12 else 30
    frames = sequence(w, h, n, channels=channels)
    p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=3, lv_l=1)
    with F.FlowStream(p, w, h, depth=depth, channels=channels, output="full", reuse=reuse) as s:
        assert s.reuse == ((2 <= depth <= 8) if reuse is None else reuse)
        flows = list(s.flows(frames))
        assert s.pending == 0
    with F.FlowStream(p, w, h, depth=depth, channels=channels) as s:  # default: the engine's level-lv_l output
        lflows = list(s.flows(frames))
        assert s.flow_shape == (200 // 2, 328 // 2, 2)  # padded to a multiple of 2^lv_f = 8, level lv_l = 1
    assert len(flows) == n - 1 and len(lflows) == n - 1
    with F.Engine(p, w, h, channels=channels) as e:
        for k in range(n - 1):
            assert bits_differ(flows[k], e.run_u8(frames[k], frames[k + 1])) == 0, k
            assert bits_differ(lflows[k], e.level_flow(w, h)) == 0, k
    # and the oracle on one pair, so that this file stands on its own (pair 2 has a reused first-frame pyramid)
    assert bits_differ(flows[2], port.run_u8(frames[2], frames[3], p.to_dict())) == 0
    with pytest.raises(F.DisError):
        F.FlowStream(p, w, h, depth=1, reuse=True)


@pytest.mark.parametrize("channels,depth,nb,output", [(1, 8, 4, "level"), (1, 6, 3, "full"), (3, 4, 2, "level"), (1, 8, 8, "level")])
def test_batched_stream_equals_pairwise(channels, depth, nb, output):
    """dis_video_create_batched: the pairs of nb pushes go out as one launch chain; every flow still equals a separate
    run on its pair, including the partial batches a pop forces out and the tail of the sequence."""
    w, h, n = 322, 198, 14
    frames = sequence(w, h, n, channels=channels)
    p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=3, lv_l=1)
    with F.FlowStream(p, w, h, depth=depth, channels=channels, output=output, pairs_per_launch=nb) as s:
        assert not s.reuse
        flows = list(s.flows(frames))
        # a second pass on the same object, popping early: partial batches
        s2 = []
        for k, fr in enumerate(frames[:6]):
            s.push(fr)
            if k in (2, 3):
                s2.append(s.pop().copy())
        while s.pending:
            s2.append(s.pop().copy())
    assert len(flows) == n - 1
    with F.Engine(p, w, h, channels=channels) as e:
        for k in range(n - 1):
            full = e.run_u8(frames[k], frames[k + 1])
            ref = full if output == "full" else e.level_flow(w, h)
            assert bits_differ(flows[k], ref) == 0, k
        # second pass: its first pair is (last frame of the first pass, frames[0]), then frames[0..5]
        full = e.run_u8(frames[-1], frames[0])
        ref = [full if output == "full" else e.level_flow(w, h)]
        for k in range(5):
            full = e.run_u8(frames[k], frames[k + 1])
            ref.append(full.copy() if output == "full" else e.level_flow(w, h))
        assert len(s2) == len(ref)
        for k in range(len(ref)):
            assert bits_differ(s2[k], ref[k]) == 0, ("second pass", k)
    with pytest.raises(F.DisError):
        F.FlowStream(p, w, h, depth=6, pairs_per_launch=4)   # must divide depth
    with pytest.raises(F.DisError):
        F.FlowStream(p, w, h, depth=8, pairs_per_launch=4, reuse=True)


def test_stream_protocol_errors():
    p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=2, lv_l=1)
    fr = sequence(128, 96, 4)
    with F.FlowStream(p, 128, 96, depth=2, output="full") as s:
        with pytest.raises(F.DisError):
            s.pop()                      # nothing in flight
        s.push(fr[0])
        assert s.pending == 0
        s.push(fr[1])
        s.push(fr[2])
        assert s.pending == 2
        with pytest.raises(F.DisError):
            s.push(fr[3])                # depth reached
        a = s.pop().copy()
        s.push(fr[3])
        b, c = s.pop().copy(), s.pop().copy()
    with F.Engine(p, 128, 96) as e:
        for got, k in ((a, 0), (b, 1), (c, 2)):
            assert bits_differ(got, e.run_u8(fr[k], fr[k + 1])) == 0


def test_run_dense_stream_cli(tmp_path):
    w, h, n = 320, 200, 5
    frames = sequence(w, h, n, seed=4)
    names = []
    for k, f in enumerate(frames):
        names.append(str(tmp_path / ("f%02d.pgm" % k)))
        with open(names[-1], "wb") as fh:
            fh.write(b"P5\n%d %d\n255\n" % (w, h))
            fh.write(f.tobytes())
    argv = "3 1 12 12 0.05 0.95 0 8 0.40 0 1 0 1 10 10 5 1 3 1.6 1".split()
    exe = os.path.join(os.path.dirname(F.api.__file__), "run_dense_stream")
    r = subprocess.run([exe, "-depth", "2", "-params"] + argv + [str(tmp_path / "out_")] + names,
                       capture_output=True, text=True, check=True)
    assert "TIME (4 pairs, decode + flow + save)" in r.stdout
    one = os.path.join(os.path.dirname(F.api.__file__), "run_dense")
    for k in range(n - 1):
        got = F.read_flo(str(tmp_path / ("out_%04d.flo" % (k + 1))))
        single = str(tmp_path / "single.flo")
        subprocess.check_call([one, names[k], names[k + 1], single] + argv[:-1] + ["0"])
        assert bits_differ(got, F.read_flo(single)) == 0, k
    assert subprocess.run([exe, "out_", names[0]], capture_output=True).returncode == 2
