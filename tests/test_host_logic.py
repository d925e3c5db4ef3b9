"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/dis_c.h
declares, parameter handling mirrors kroeger/run_dense.cpp, .flo I/O round-trips, and the product
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import flowonthego_b200 as F
from flowonthego_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dis_c.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(dis_[a-z0-9_]+)\s*\(", hdr))
    assert {"dis_create", "dis_run_pyramids", "dis_run_u8", "dis_submit_u8_device", "dis_write_flo"} <= names
    L = F.lib()
    for n in sorted(names):
        assert hasattr(L, n), "libdis_b200.so does not export " + n
    assert b"sm_100a" in L.dis_version()


def test_presets_follow_reference_cli():
    # kroeger/run_dense.cpp:239-267 on a 1024-wide image
    expect = {1: (8, 0.3, 5, 3, 16, 0), 2: (8, 0.4, 5, 3, 12, 1), 3: (12, 0.75, 5, 1, 16, 1), 4: (12, 0.75, 5, 0, 128, 1)}
    for k, (ps, ov, lf, ll, it, tv) in expect.items():
        p = F.Params.preset(k, 1024)
        assert (p.patchsz, p.lv_f, p.lv_l, p.maxiter, p.miniter, p.usetvref) == (ps, lf, ll, it, it, tv)
        assert abs(p.poverl - ov) < 1e-7
    assert F.Params.preset(7, 1024).to_dict() == F.Params.preset(2, 1024).to_dict()  # `default:` label
    assert F.lib().dis_auto_first_scale(1920, 5, 12) == 6 and F.lib().dis_auto_first_scale(3840, 5, 12) == 7
    assert F.lib().dis_auto_first_scale(10, 5, 12) == 0


def test_params_from_argv_order():
    argv = "5 3 12 12 0.05 0.95 0 8 0.40 0 1 0 1 10 10 5 1 3 1.6 2".split()  # kroeger/README.md:62
    p = F.Params.from_argv(argv)
    d = p.to_dict()
    assert [d[k] for k in ("lv_f", "lv_l", "maxiter", "miniter", "patchsz", "usefbcon", "patnorm", "costfct",
                           "usetvref", "tv_innerit", "tv_solverit", "verbosity")] == [5, 3, 12, 12, 8, 0, 1, 0, 1, 1, 3, 2]
    assert np.float32(d["tv_sor"]) == np.float32(1.6) and np.float32(d["poverl"]) == np.float32(0.4)
    with pytest.raises(F.DisError):
        F.Params.from_argv(argv[:19])


def test_validate_rejects_out_of_scope():
    L = F.lib()
    buf = ctypes.create_string_buffer(200)
    ok = F.Params.preset(2, 1024)
    assert L.dis_params_validate(ctypes.byref(ok), buf, 200) == 0
    for kw in (dict(patchsz=7), dict(patchsz=18), dict(lv_l=6), dict(costfct=10), dict(poverl=1.0), dict(tv_solverit=0), dict(tv_solverit=257)):
        bad = ok.copy(**kw)
        assert L.dis_params_validate(ctypes.byref(bad), buf, 200) != 0, kw
        assert len(buf.value) > 0


def test_padded_size_matches_reference():
    # kroeger/run_dense.cpp:298-311: 1024x436 at lv_f=5 -> 1024x448, pad split floor/ceil
    assert F.padded_size(1024, 436, 5) == (1024, 448, 0, 6)
    assert F.padded_size(1920, 1080, 6) == (1920, 1088, 0, 4)
    assert F.padded_size(3840, 2160, 7) == (3840, 2176, 0, 8)
    assert F.padded_size(501, 301, 3) == (504, 304, 1, 1)


def test_flo_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    fl = rng.standard_normal((7, 13, 2)).astype(np.float32)
    path = str(tmp_path / "x.flo")
    F.write_flo(path, fl)
    raw = open(path, "rb").read()
    assert raw[:4] == b"PIEH" and len(raw) == 12 + 8 * 7 * 13  # flow_code/C/flowIO.cpp:5-25
    assert np.frombuffer(raw[4:12], np.int32).tolist() == [13, 7]
    assert np.array_equal(F.read_flo(path), fl)
    open(path, "wb").write(b"XXXX" + raw[4:])
    with pytest.raises(F.DisError):
        F.read_flo(path)


def test_native_image_decode_matches_opencv(tmp_path, alley_pair):
    """PNG (grey + RGB) / PGM / PPM decode == cv2.imread(IMREAD_GRAYSCALE) (kroeger/run_dense.cpp:208-209)."""
    import cv2
    a, _ = alley_pair
    rgb = np.stack([a, np.roll(a, 7, 1), 255 - a], -1)[:97, :131]
    for name, img in (("g.png", a), ("c.png", rgb), ("g.pgm", a[:50, :70]), ("c.ppm", rgb)):
        p = str(tmp_path / name)
        assert cv2.imwrite(p, img)
        assert np.array_equal(F.read_image_gray(p), cv2.imread(p, cv2.IMREAD_GRAYSCALE)), name
    with pytest.raises(F.DisError):
        F.read_image_gray(str(tmp_path / "missing.png"))


def test_no_cpu_fallback():
    """Without a CUDA device the engine must refuse to exist (and say why)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(F.DisError) as ei:
        F.Engine(F.Params.preset(2, 256), 256, 192)
    assert ei.value.code == 3 and "no CPU path" in str(ei.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under flowonthego_b200/ or bench.py's product arm may use it."""
    pkg = os.path.join(ROOT, "flowonthego_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), os.path.join(dp, f)


def test_cli_binaries_usage_and_loud_failure(tmp_path, golden_dir):
    """The compiled CLIs exist next to the library, print their usage on bad arguments and -- without a GPU --
    fail with an error instead of computing anything on the CPU."""
    import subprocess
    bindir = os.path.dirname(F.api.__file__)
    for exe in ("run_dense", "run_dense_rgb", "run_dense_stream", "color_flow", "flow_epe"):
        r = subprocess.run([os.path.join(bindir, exe)], capture_output=True, text=True)
        assert r.returncode != 0 and "usage" in (r.stderr + r.stdout).lower(), exe
    try:
        import torch
        if torch.cuda.is_available():
            return
    except ImportError:
        pass
    a = os.path.join(golden_dir, "alley_0001_gray.png")
    b = os.path.join(golden_dir, "alley_0002_gray.png")
    r = subprocess.run([os.path.join(bindir, "run_dense"), a, b, str(tmp_path / "o.flo")], capture_output=True, text=True)
    assert r.returncode == 1 and "run_dense:" in r.stderr and not (tmp_path / "o.flo").exists()


def test_image_reader_rejects_malformed_files(tmp_path):
    """Hostile headers must come back as DIS_ERR_IO, not as an abort or a multi-gigabyte allocation."""
    import struct
    import zlib
    import flowonthego_b200 as F

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    sig = b"\x89PNG\r\n\x1a\n"
    cases = {
        "short_ihdr.png": sig + chunk(b"IHDR", b"\0\0\0\x10\0\0\0\x10\x08") + chunk(b"IEND", b""),
        "huge.png": sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 0x7fffffff, 0x7fffffff, 8, 0, 0, 0, 0)) + chunk(b"IEND", b""),
        "big_area.png": sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 90000, 90000, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b""),
        "huge.pgm": b"P5\n99999999999999999999 5\n255\n" + b"\0" * 16,
        "trunc.pgm": b"P5\n64 64\n255\n" + b"\0" * 100,
    }
    for name, data in cases.items():
        path = tmp_path / name
        path.write_bytes(data)
        with pytest.raises(F.DisError) as ei:
            F.read_image_gray(str(path))
        assert ei.value.code == 4, name  # DIS_ERR_IO


def test_native_jpeg_decode_matches_opencv(tmp_path):
    """Baseline JPEG -> grey: bit-identical to cv2.imread(.., IMREAD_GRAYSCALE) (libjpeg's luminance plane, islow IDCT)
    for 4:2:0 / 4:4:4 / grey files, odd sizes, restart intervals; progressive files are refused.  The reference's own
    fixtures images/road_HD.jpg and images/yosemite_4k.jpg are checked where the reference tree exists."""
    import cv2
    import flowonthego_b200 as F
    from tests.synth import texture
    rng = np.random.default_rng(5)
    cases = []
    for k, (w, h) in enumerate(((64, 48), (123, 77), (17, 9), (640, 360))):
        col = np.stack([texture(w, h, 20 + 3 * k + c) for c in range(3)], -1)
        col = np.clip(col.astype(int) + rng.integers(-20, 20, col.shape), 0, 255).astype(np.uint8)
        for q in (35, 90):
            cases.append(("c%d_q%d.jpg" % (k, q), col, [cv2.IMWRITE_JPEG_QUALITY, q]))
        cases.append(("g%d.jpg" % k, col[..., 0].copy(), [cv2.IMWRITE_JPEG_QUALITY, 80]))
        cases.append(("r%d.jpg" % k, col, [cv2.IMWRITE_JPEG_QUALITY, 75, cv2.IMWRITE_JPEG_RST_INTERVAL, 3]))
        if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
            cases.append(("s444_%d.jpg" % k, col, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]))
            cases.append(("s422_%d.jpg" % k, col, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]))
    for name, img, flags in cases:
        path = str(tmp_path / name)
        assert cv2.imwrite(path, img, flags)
        want = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        got = F.read_image_gray(path)
        assert got.shape == want.shape and np.array_equal(got, want), name
    prog = str(tmp_path / "prog.jpg")
    cv2.imwrite(prog, cases[0][1], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(F.DisError):
        F.read_image_gray(prog)
    with pytest.raises(F.DisError):  # colour output from JPEG is not provided
        F.read_image_bgr(str(tmp_path / cases[0][0]))
    ref = os.environ.get("DIS_REFERENCE", "/root/reference")
    for n in ("road_HD", "yosemite_4k"):
        path = os.path.join(ref, "images", n + ".jpg")
        if os.path.exists(path):
            assert np.array_equal(F.read_image_gray(path), cv2.imread(path, cv2.IMREAD_GRAYSCALE)), n


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs without a GPU and without the product library, and prints the one JSON line the
    driver reads: the metric / unit / config of the B200 arm, `impl`, a `cpu_baseline` that describes this very run and
    an `e2e` object with zero copy bytes (the round-1 arm timed process start-up instead of the engine)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frame_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["config_id"] == "c5" and d["config"]["resolution"] == [1920, 1080]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0
    # steady state: the pool's rate agrees with cores / single-thread time (not with process start-up)
    assert 0.3 < cb["agreement"] < 1.5, cb
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "libdis_b200" not in r.stderr  # never loads the product
