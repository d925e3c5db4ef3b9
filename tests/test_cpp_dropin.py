"""The header-only C++ drop-in (include/oflow.h): a program written against the reference's OFC::OFClass constructor
compiles, links against libdis_b200.so and -- on a GPU -- produces the oracle's flow."""
import os
import subprocess

import numpy as np
import pytest

import flowonthego_b200 as F
from oracle import port
from tests.synth import synth_pair, synth_pair_bgr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.dirname(F.api.__file__)


def build(tmp_path):
    exe = str(tmp_path / "ofclass_demo")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "ofclass_demo.cpp"), "-o", exe, "-L", LIBDIR,
                           "-ldis_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def write_case(path, a, b, lv_f, lv_l, patchsz, usefbcon, usetvref):
    noc = 1 if a.ndim == 2 else 3
    h, w = a.shape[:2]
    wp, hp, _, _ = port.padded_size(w, h, lv_f)
    A = np.full((hp, wp) + a.shape[2:], 0, np.uint8)  # the engine boundary takes padded sizes: pad by replication
    pad = lambda x: np.pad(x, ((0, hp - h), (0, wp - w)) + ((0, 0),) * (x.ndim - 2), mode="edge")
    a, b = pad(a), pad(b)
    pa, pb = port.build_pyramid(a, lv_f, patchsz), port.build_pyramid(b, lv_f, patchsz)
    with open(path, "wb") as f:
        f.write(np.array([noc, wp, hp, lv_f, lv_l, patchsz, usefbcon, usetvref], np.int32).tobytes())
        for lst in (*pa, *pb):
            for l in range(lv_f + 1):
                f.write(np.ascontiguousarray(lst[l], np.float32).tobytes())
    return pa, pb, wp, hp, noc


def test_cpp_dropin_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = build(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu test")
    except ImportError:
        pass
    a, b, _ = synth_pair(64, 48, seed=1)
    write_case(str(tmp_path / "p.bin"), a, b, 1, 0, 8, 0, 0)
    r = subprocess.run([exe, str(tmp_path / "p.bin"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "OFClass failed" in r.stderr  # no CPU fallback: the error surfaces


@pytest.mark.gpu
@pytest.mark.parametrize("colour,usefbcon", [(False, 0), (False, 1), (True, 0)])
def test_cpp_dropin_matches_oracle(tmp_path, colour, usefbcon):
    exe = build(tmp_path)
    a, b, _ = (synth_pair_bgr if colour else synth_pair)(256, 192, seed=6)
    pa, pb, wp, hp, noc = write_case(str(tmp_path / "p.bin"), a, b, 3, 1, 8, usefbcon, 1)
    subprocess.check_call([exe, str(tmp_path / "p.bin"), str(tmp_path / "o.bin")])
    got = np.fromfile(str(tmp_path / "o.bin"), np.float32).reshape(hp >> 1, wp >> 1, 2)
    p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=3, lv_l=1, usefbcon=usefbcon).to_dict()
    ref = port.run_engine(pa, pb, wp, hp, p, noc=noc)
    assert int((got.view(np.uint32) != ref.view(np.uint32)).sum()) == 0
