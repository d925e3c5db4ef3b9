"""Generates the committed fixtures in tests/golden/ from the reference tree (run in the build
container, where /root/reference exists; the GPU box has no reference tree and only reads the
fixtures).

  alley_0001_gray.png / alley_0002_gray.png
      images/alley_1/frame_0001.png, frame_0002.png decoded ONCE with
      cv2.imread(IMREAD_GRAYSCALE) -- the reference's own decode (kroeger/run_dense.cpp:208-209)
      -- and stored losslessly, so oracle and GPU path see identical u8 arrays.
  alley_0001_flo.npz
      kroeger/flows/alley_0001.flo, the reference's only known-answer vector
      (= run_OF_INT frame_0001.png frame_0002.png out.flo, operating point 2), bit-for-bit.
  ref_cases.npz
      outputs of the verbatim-compiled reference engine (oracle/_ref/libdis_ref.so) on small
      crops for parameter sets the golden file does not cover (preset 3/4 style, L1/Huber cost,
      forward-backward merging, odd patch sizes, 30-wide level).  Raw level-lv_l flow.
  alley_0001_color.png
      kroeger/flows/alley_0001.png, the reference's colour coding of alley_0001.flo (flow_code/C/color_flow),
      pixel for pixel (re-encoded) -- the known-answer vector of the colour-coding tool.
  ref_cases_rgb.npz
      the same for the reference's colour build (SELECTCHANNEL=3, oracle/_ref/libdis_ref_rgb.so): one BGR
      crop of the two frames as cv2.imread(IMREAD_COLOR) decodes them (kroeger/run_dense.cpp:203-206),
      stored in the file, and the raw level-lv_l flow for a few parameter sets.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402

REF = os.environ.get("DIS_REFERENCE", "/root/reference")

# (name, crop (y0,y1,x0,x1), parameter overrides on top of operating point 2)
CASES = [
    ("p3like", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, patchsz=12, poverl=0.75, maxiter=16, miniter=16)),
    ("p4like", (60, 260, 100, 420), dict(lv_f=3, lv_l=0, patchsz=12, poverl=0.75, maxiter=128, miniter=128)),
    ("notv", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, usetvref=0)),
    ("fbcon", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, usefbcon=1)),
    ("l1", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, costfct=1, maxiter=24, miniter=24)),
    ("huber", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, costfct=2, maxiter=24, miniter=24)),
    ("nonorm", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, patnorm=0)),
    ("p6", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, patchsz=6, poverl=0.5)),
    ("p10", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, patchsz=10, poverl=0.5)),
    ("p16", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, patchsz=16, poverl=0.5)),
    ("early", (40, 232, 100, 420), dict(lv_f=3, lv_l=1, miniter=2, maxiter=30, mindprate=0.2, mindrrate=0.9, minimgerr=1.0)),
    ("w30", (0, 272, 0, 480), dict(lv_f=4, lv_l=3, patchsz=8, poverl=0.4)),  # level 4 is 30x17
    ("tvheavy", (0, 200, 0, 320), dict(lv_f=2, lv_l=0, tv_innerit=2, tv_solverit=5, tv_sor=1.9)),
]


RGB_CROP = (40, 232, 100, 420)
RGB_CASES = [
    ("op2", dict(lv_f=3, lv_l=1)),
    ("p3like", dict(lv_f=3, lv_l=1, patchsz=12, poverl=0.75, maxiter=16, miniter=16)),
    ("notv", dict(lv_f=3, lv_l=1, usetvref=0)),
    ("fbcon", dict(lv_f=3, lv_l=1, usefbcon=1)),
    ("huber", dict(lv_f=3, lv_l=1, costfct=2, maxiter=24, miniter=24)),
    ("l1_nonorm", dict(lv_f=3, lv_l=1, costfct=1, patnorm=0)),
    ("p6", dict(lv_f=3, lv_l=1, patchsz=6, poverl=0.5)),
    ("p10_lvl0", dict(lv_f=3, lv_l=0, patchsz=10, poverl=0.5)),
    ("early", dict(lv_f=3, lv_l=1, miniter=2, maxiter=30, mindprate=0.2, mindrrate=0.9, minimgerr=1.0)),
]


def main_rgb():
    a = cv2.imread(os.path.join(REF, "images/alley_1/frame_0001.png"), cv2.IMREAD_COLOR)
    b = cv2.imread(os.path.join(REF, "images/alley_1/frame_0002.png"), cv2.IMREAD_COLOR)
    y0, y1, x0, x1 = RGB_CROP
    A, B = np.ascontiguousarray(a[y0:y1, x0:x1]), np.ascontiguousarray(b[y0:y1, x0:x1])
    out = dict(img_a=A, img_b=B)
    base = rd.preset_params(a.shape[1], 2)
    for name, kw in RGB_CASES:
        p = dict(base)
        p.update(kw)
        fl = rd.run_dense_ref(A, B, p, full_res=False, rgb=True)
        out[name + "_flow"] = fl
        out[name + "_params"] = np.array([p[k] for k in rd.PARAM_NAMES], np.float64)
        print("rgb", name, fl.shape, float(np.abs(fl).max()))
    np.savez_compressed(os.path.join(HERE, "ref_cases_rgb.npz"), **out)


def main():
    a = cv2.imread(os.path.join(REF, "images/alley_1/frame_0001.png"), cv2.IMREAD_GRAYSCALE)
    b = cv2.imread(os.path.join(REF, "images/alley_1/frame_0002.png"), cv2.IMREAD_GRAYSCALE)
    cv2.imwrite(os.path.join(HERE, "alley_0001_gray.png"), a, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    cv2.imwrite(os.path.join(HERE, "alley_0002_gray.png"), b, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    g = rd.read_flo(os.path.join(REF, "kroeger/flows/alley_0001.flo"))
    np.savez_compressed(os.path.join(HERE, "alley_0001_flo.npz"), flow=g)
    col = cv2.imread(os.path.join(REF, "kroeger/flows/alley_0001.png"), cv2.IMREAD_COLOR)
    cv2.imwrite(os.path.join(HERE, "alley_0001_color.png"), col, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    out = {}
    base = rd.preset_params(a.shape[1], 2)
    for name, (y0, y1, x0, x1), kw in CASES:
        p = dict(base)
        p.update(kw)
        fl = rd.run_dense_ref(a[y0:y1, x0:x1], b[y0:y1, x0:x1], p, full_res=False)
        out[name + "_flow"] = fl
        out[name + "_crop"] = np.array([y0, y1, x0, x1], np.int32)
        out[name + "_params"] = np.array([p[k] for k in rd.PARAM_NAMES], np.float64)
        print(name, fl.shape, float(np.abs(fl).max()))
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **out)


if __name__ == "__main__":
    if "--rgb-only" not in sys.argv:
        main()
    main_rgb()
