/* TEST INFRASTRUCTURE ONLY -- this file is the parity checker, never the product.
 *
 * Plain-C (scalar, single-thread) restatement of the reference's hot path
 * (zhaorz/FlowOnTheGo, CPU tree kroeger/): pyramid + gradients, per-patch inverse search,
 * densification, variational refinement, final upsampling.  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
 * may load the resulting liboracle; the product (libdis_b200.so) never does.
 *
 * Pinning: tests/test_oracle_*.py check this restatement
 *   (1) against the reference's only known-answer vector kroeger/flows/alley_0001.flo
 *       (bit-exact, through the committed fixture tests/golden/), and
 *   (2) against oracle/_ref/libdis_ref.so = the reference's own sources compiled verbatim
 *       (bit-exact on every config tried, incl. stride != width levels, L1/Huber cost,
 *       forward-backward merging).
 *
 * Arithmetic rules: fp32 round-to-nearest, no FMA contraction (the
 * reference is built -O3 -msse4, kroeger/CMakeLists.txt:4-5; this file is built with
 * -ffp-contract=off), C evaluation order of the reference expressions, IEEE sqrt/div,
 * Eigen-3.3/SSE reduction order for per-patch sums.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dis_c.h"

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Eigen 3.3 linear-vectorised reduction, SSE packet of 4, two accumulators
 * (Eigen/src/Core/Redux.h; Eigen is an un-vendored dependency, kroeger/CMakeLists.txt:12).
 * call sites: kroeger/patch.cpp:74-78, 178-184, 278, 331, 401.
 * ---------------------------------------------------------------------------------------- */
static float eig_sum3(const float* a, const float* b, int n, int mode) {
  /* mode 0: sum a[i]; 1: sum a[i]*b[i] (product rounded first); 2: sum |a[i]| */
#define EL(i) (mode == 0 ? a[i] : (mode == 1 ? (float)(a[i] * b[i]) : fabsf(a[i])))
  const int n8 = (n / 8) * 8, n4 = (n / 4) * 4;
  float res;
  if (n4) {
    float p0[4], p1[4];
    int i, k;
    for (k = 0; k < 4; ++k) p0[k] = EL(k);
    if (n4 > 4) {
      for (k = 0; k < 4; ++k) p1[k] = EL(4 + k);
      for (i = 8; i < n8; i += 8) {
        for (k = 0; k < 4; ++k) p0[k] = p0[k] + EL(i + k);
        for (k = 0; k < 4; ++k) p1[k] = p1[k] + EL(i + 4 + k);
      }
      for (k = 0; k < 4; ++k) p0[k] = p0[k] + p1[k];
      if (n4 > n8)
        for (k = 0; k < 4; ++k) p0[k] = p0[k] + EL(n8 + k);
    }
    res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
    for (i = n4; i < n; ++i) res = res + EL(i);
  } else {
    int i;
    res = EL(0);
    for (i = 1; i < n; ++i) res = res + EL(i);
  }
#undef EL
  return res;
}
static float eig_sum(const float* a, int n) { return eig_sum3(a, 0, n, 0); }
static float eig_dot(const float* a, const float* b, int n) { return eig_sum3(a, b, n, 1); }
static float eig_abssum(const float* a, int n) { return eig_sum3(a, 0, n, 2); }

/* Eigen LLT of a 2x2 + solve (Eigen/src/Cholesky/LLT.h llt_inplace::unblocked, triangular
 * solves by division); call site kroeger/patch.cpp:184. */
static void llt2_solve(float h00, float h10, float h11, float b0, float b1, float* x0o, float* x1o) {
  float L00 = h00, L10 = h10, L11 = h11;
  float x = L00;
  if (!(x <= 0.0f)) {
    L00 = x = sqrtf(x);
    L10 = L10 / x;
    x = L11 - L10 * L10;
    if (!(x <= 0.0f)) L11 = sqrtf(x);
  }
  {
    float y0 = b0 / L00;
    float y1 = b1 - L10 * y0;
    float x1, x0;
    y1 = y1 / L11;
    x1 = y1 / L11;
    x0 = y0 - L10 * x1;
    x0 = x0 / L00;
    *x0o = x0;
    *x1o = x1;
  }
}

/* ------------------------------------------------------------------------------------------
 * P1: image pyramid + gradients + padding.  kroeger/run_dense.cpp:298-311 (divisibility pad),
 * :326-327 (u8 -> f32), :130-178 ConstructImgPyramide (cv::resize 0.5 INTER_LINEAR, cv::Sobel
 * ksize=1 BORDER_DEFAULT, copyMakeBorder REPLICATE / CONSTANT 0).  OpenCV is un-vendored
 * (kroeger/CMakeLists.txt:11, version unpinned); restated from its published semantics:
 *   resize x0.5 INTER_LINEAR == 2x2 box mean ((a+b)+(c+d))*0.25 (OpenCV's own code path; the IPP
 *   path computes a+(b-a)/2 lerps -- identical bits for u8-derived input up to level 8);
 *   Sobel ksize=1: I(x+1)-I(x-1) with reflect-101 (=> 0 on the border column/row).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_padded_size(int w, int h, int lv_f, int* wp, int* hp, int* left, int* top) {
  const int sc = 1 << lv_f;
  int padw = 0, padh = 0;
  if (w % sc) padw = sc - w % sc;
  if (h % sc) padh = sc - h % sc;
  *wp = w + padw;
  *hp = h + padh;
  *left = (int)floorf((float)padw / 2.0f);
  *top = (int)floorf((float)padh / 2.0f);
  return 0;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Builds levels 0..lv_f.  I/Ix/Iy: arrays of lv_f+1 pointers, each malloc'ed here as
 * (w_l+2*pad) x (h_l+2*pad).  Ix/Iy may be NULL (no gradients). */
/* noc = 1: grey (SELECTCHANNEL=1); noc = 3: interleaved BGR (SELECTCHANNEL=3, run_dense.cpp:203-206).  Every
 * OpenCV call of ConstructImgPyramide works per channel, so the colour case is the grey one on 3 planes. */
ORACLE_API int oracle_build_pyramid_c(const uint8_t* img, int w, int h, int pitch, int noc, int lv_f, int pad,
                                      float** I, float** Ix, float** Iy) {
  int wp, hp, left, top, l, x, y, ch;
  float* prev = NULL;
  oracle_padded_size(w, h, lv_f, &wp, &hp, &left, &top);
  for (l = 0; l <= lv_f; ++l) {
    const int wl = wp >> l, hl = hp >> l;
    float* cur = (float*)malloc(sizeof(float) * wl * hl * noc);
    if (l == 0) {
      for (y = 0; y < hl; ++y)
        for (x = 0; x < wl; ++x)
          for (ch = 0; ch < noc; ++ch)
            cur[(y * wl + x) * noc + ch] =
                (float)img[clampi(y - top, 0, h - 1) * pitch + clampi(x - left, 0, w - 1) * noc + ch];
    } else {
      const int wq = wl * 2;
      for (y = 0; y < hl; ++y)
        for (x = 0; x < wl; ++x)
          for (ch = 0; ch < noc; ++ch) {
            const float a = prev[((2 * y) * wq + 2 * x) * noc + ch], b = prev[((2 * y) * wq + 2 * x + 1) * noc + ch];
            const float c = prev[((2 * y + 1) * wq + 2 * x) * noc + ch], d = prev[((2 * y + 1) * wq + 2 * x + 1) * noc + ch];
            cur[(y * wl + x) * noc + ch] = ((a + b) + (c + d)) * 0.25f;
          }
    }
    {
      const int tw = wl + 2 * pad, th = hl + 2 * pad;
      float* Ip = (float*)malloc(sizeof(float) * tw * th * noc);
      float* Ixp = Ix ? (float*)calloc((size_t)tw * th * noc, sizeof(float)) : NULL;
      float* Iyp = Iy ? (float*)calloc((size_t)tw * th * noc, sizeof(float)) : NULL;
      for (y = 0; y < th; ++y)
        for (x = 0; x < tw; ++x)
          for (ch = 0; ch < noc; ++ch)
            Ip[(y * tw + x) * noc + ch] = cur[(clampi(y - pad, 0, hl - 1) * wl + clampi(x - pad, 0, wl - 1)) * noc + ch];
      if (Ixp)
        for (y = 0; y < hl; ++y)
          for (x = 0; x < wl; ++x) {
            /* reflect-101: index -1 -> 1, wl -> wl-2 */
            const int xm = x == 0 ? (wl > 1 ? 1 : 0) : x - 1, xq = x == wl - 1 ? (wl > 1 ? wl - 2 : 0) : x + 1;
            const int ym = y == 0 ? (hl > 1 ? 1 : 0) : y - 1, yq = y == hl - 1 ? (hl > 1 ? hl - 2 : 0) : y + 1;
            for (ch = 0; ch < noc; ++ch) {
              Ixp[((y + pad) * tw + x + pad) * noc + ch] = cur[(y * wl + xq) * noc + ch] - cur[(y * wl + xm) * noc + ch];
              Iyp[((y + pad) * tw + x + pad) * noc + ch] = cur[(yq * wl + x) * noc + ch] - cur[(ym * wl + x) * noc + ch];
            }
          }
      I[l] = Ip;
      if (Ix) Ix[l] = Ixp;
      if (Iy) Iy[l] = Iyp;
    }
    free(prev);
    prev = cur;
  }
  free(prev);
  return 0;
}

ORACLE_API int oracle_build_pyramid(const uint8_t* img, int w, int h, int pitch, int lv_f, int pad,
                                    float** I, float** Ix, float** Iy) {
  return oracle_build_pyramid_c(img, w, h, pitch, 1, lv_f, pad, I, Ix, Iy);
}

ORACLE_API void oracle_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Engine parameters derived in OFClass::OFClass (kroeger/oflow.cpp:75-108) and the per-level
 * camparam (oflow.cpp:138-160, oflow.h:16-29).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int p, novals, steps, max_iter, min_iter, patnorm, costfct, noc;
  float outlierthresh, dp_thresh, dr_thresh, res_thresh;
} opt_t;

typedef struct {
  int w, h, pad, tmp_w, lv;
  float lb, ubw, ubh;
  int nopw, noph, offw, offh, nop;
} lvl_t;

static void make_opt(const dis_params* q, int noc, opt_t* o) {
  o->noc = noc;
  o->p = q->patchsz;
  o->outlierthresh = (float)o->p / 2;
  o->max_iter = q->maxiter;
  o->min_iter = q->miniter;
  o->dp_thresh = q->mindprate * q->mindprate; /* oflow.cpp:88 */
  o->dr_thresh = q->mindrrate;
  o->res_thresh = q->minimgerr;
  {
    int s = (int)floor(o->p * (1 - q->poverl)); /* oflow.cpp:91: int*float -> float, floor(double) */
    o->steps = s > 1 ? s : 1;
  }
  o->novals = noc * o->p * o->p; /* oflow.cpp:92 */
  o->patnorm = q->patnorm;
  o->costfct = q->costfct;
}

static void make_lvl(const dis_params* q, const opt_t* o, int width, int height, int pad, int sl, lvl_t* c) {
  const float sc_fct = (float)pow(2, -sl);
  c->h = (int)(height * sc_fct); /* oflow.cpp:143-144 */
  c->w = (int)(width * sc_fct);
  c->pad = pad;
  c->lb = -(float)o->p / 2;
  c->ubw = (float)(c->w + o->p / 2 - 2);
  c->ubh = (float)(c->h + o->p / 2 - 2);
  c->tmp_w = c->w + 2 * pad;
  c->lv = sl;
  /* grid geometry, kroeger/patchgrid.cpp:42-49 */
  c->nopw = (int)ceil((float)c->w / (float)o->steps);
  c->noph = (int)ceil((float)c->h / (float)o->steps);
  c->offw = (int)floor((c->w - (c->nopw - 1) * o->steps) / 2); /* integer division first */
  c->offh = (int)floor((c->h - (c->noph - 1) * o->steps) / 2);
  c->nop = c->nopw * c->noph;
  (void)q;
}

/* ------------------------------------------------------------------------------------------
 * D1-D7: one patch.  kroeger/patch.cpp.
 * ---------------------------------------------------------------------------------------- */
/* getPatchStaticBil, patch.cpp:335-402 */
static void patch_bilinear(const float* img, const lvl_t* c, const opt_t* o, float mx, float my, float* out) {
  int posx = (int)ceilf(mx + .00001f), posy = (int)ceilf(my + .00001f);
  const int flx = (int)floorf(mx), fly = (int)floorf(my);
  const float rx = mx - (float)flx, ry = my - (float)fly;
  const float w0 = rx * ry, w1 = (1 - rx) * ry, w2 = rx * (1 - ry), w3 = (1 - rx) * (1 - ry);
  const int lb = -o->p / 2, ub = o->p / 2 - 1;
  int x, y, k = 0;
  posx += c->pad;
  posy += c->pad;
  for (y = posy + lb; y <= posy + ub; ++y)
    for (x = posx + lb; x <= posx + ub; ++x) {
      int ch;
      for (ch = 0; ch < o->noc; ++ch, ++k) { /* RGB: 3 interleaved channels, same weights (patch.cpp:392-396) */
        const int n = o->noc;
        const float a = img[(y * c->tmp_w + x) * n + ch], b = img[(y * c->tmp_w + x - 1) * n + ch];
        const float cc = img[((y - 1) * c->tmp_w + x) * n + ch], d = img[((y - 1) * c->tmp_w + x - 1) * n + ch];
        out[k] = w0 * a + w1 * b + w2 * cc + w3 * d;
      }
    }
  if (o->patnorm > 0) {
    const float m = eig_sum(out, o->novals) / o->novals;
    for (k = 0; k < o->novals; ++k) out[k] = out[k] - m;
  }
}

/* LossComputeErrorImage, patch.cpp:223-262 (constants oflow.h:62-63, oflow.cpp:107-109) */
static void patch_loss(const opt_t* o, float* pdiff, float* pweight, const float* tmpl) {
  int k;
  const float bsq = 5.0f * 5.0f, bsq2 = bsq * 2.0f;
  for (k = 0; k < o->novals; ++k) {
    float d = pdiff[k] - tmpl[k];
    if (o->costfct == 1)
      d = copysignf(sqrtf(fabsf(d)), d);
    else if (o->costfct == 2)
      d = copysignf(sqrtf((sqrtf(1.0f + (d * d) / bsq) - 1.0f) * bsq2), d);
    pdiff[k] = d;
    pweight[k] = fabsf(d);
  }
}

typedef struct {
  float p_in[2], p_iter[2], delta_p[2], pt_iter[2], pt_st[2];
  float dp_sq, dp_sq_init, mares, mares_old;
  int cnt, converged;
} pstate_t;

/* OptimizeComputeErrImg, patch.cpp:264-284 */
static void patch_errimg(const float* imb, const lvl_t* c, const opt_t* o, const float* tmpl,
                         pstate_t* s, float* pdiff, float* pweight) {
  patch_bilinear(imb, c, o, s->pt_iter[0], s->pt_iter[1], pdiff);
  patch_loss(o, pdiff, pweight, tmpl);
  s->dp_sq = s->delta_p[0] * s->delta_p[0] + s->delta_p[1] * s->delta_p[1];
  if (s->cnt == 1) s->dp_sq_init = s->dp_sq;
  s->mares_old = s->mares;
  s->mares = eig_abssum(pweight, o->novals) / (o->novals);
  if (!((s->cnt < o->max_iter) & (s->mares > o->res_thresh) &
        ((s->cnt < o->min_iter) | (s->dp_sq / s->dp_sq_init >= o->dp_thresh)) &
        ((s->cnt < o->min_iter) | (s->mares / s->mares_old <= o->dr_thresh))))
    s->converged = 1;
}

static int patch_oob(const lvl_t* c, const float* pt) {
  return pt[0] < c->lb || pt[1] < c->lb || pt[0] > c->ubw || pt[1] > c->ubh;
}

/* InitializePatch (patch.cpp:57-69, getPatchStaticNNGrad :287-332, ComputeHessian :71-88) followed
 * by OptimizeIter(p_init, true) (patch.cpp:159-212, OptimizeStart :120-156).
 * out: p_iter[2]; pweight[novals] (zero when the start position is out of bounds: the reference
 * leaves it uninitialised there, see oracle/standin/Eigen/Core). */
static void patch_run(const float* ima, const float* imax, const float* imay, const float* imb,
                      const lvl_t* c, const opt_t* o, int cx, int cy, const float* p_init,
                      float* p_out, float* pweight, float* scratch) {
  const int n = o->novals, lb = -o->p / 2, ub = o->p / 2 - 1;
  float *tmpl = scratch, *gx = scratch + n, *gy = scratch + 2 * n, *pdiff = scratch + 3 * n;
  float H00, H01, H11;
  pstate_t s;
  int i, j, k = 0;
  const int px = cx + c->pad, py = cy + c->pad;
  for (j = lb; j <= ub; ++j)
    for (i = lb; i <= ub; ++i) {
      int ch;
      for (ch = 0; ch < o->noc; ++ch, ++k) { /* patch.cpp:316-325 */
        const int idx = ((px + i) + (py + j) * c->tmp_w) * o->noc + ch;
        tmpl[k] = ima[idx];
        gx[k] = imax[idx];
        gy[k] = imay[idx];
      }
    }
  if (o->patnorm > 0) {
    const float m = eig_sum(tmpl, n) / n;
    for (k = 0; k < n; ++k) tmpl[k] = tmpl[k] - m;
  }
  H00 = eig_dot(gx, gx, n);
  H01 = eig_dot(gx, gy, n);
  H11 = eig_dot(gy, gy, n);
  if (H00 * H11 - H01 * H01 == 0) {
    H00 = (float)(H00 + 1e-10); /* float += double literal */
    H11 = (float)(H11 + 1e-10);
  }
  memset(pweight, 0, sizeof(float) * n);
  memset(&s, 0, sizeof(s));
  s.p_in[0] = s.p_iter[0] = p_init[0];
  s.p_in[1] = s.p_iter[1] = p_init[1];
  s.pt_iter[0] = (float)cx + s.p_iter[0];
  s.pt_iter[1] = (float)cy + s.p_iter[1];
  s.pt_st[0] = s.pt_iter[0];
  s.pt_st[1] = s.pt_iter[1];
  if (patch_oob(c, s.pt_iter)) {
    s.converged = 1;
  } else {
    s.cnt = 0;
    s.dp_sq = 1e-10f;
    s.dp_sq_init = 1e-10f;
    s.mares = 1e5f;
    s.mares_old = 1e20f;
    s.converged = 0;
    patch_errimg(imb, c, o, tmpl, &s, pdiff, pweight);
  }
  while (!s.converged) {
    float dx, dy;
    s.cnt++;
    s.delta_p[0] = eig_dot(gx, pdiff, n);
    s.delta_p[1] = eig_dot(gy, pdiff, n);
    llt2_solve(H00, H01, H11, s.delta_p[0], s.delta_p[1], &s.delta_p[0], &s.delta_p[1]);
    s.p_iter[0] = s.p_iter[0] - s.delta_p[0];
    s.p_iter[1] = s.p_iter[1] - s.delta_p[1];
    s.pt_iter[0] = (float)cx + s.p_iter[0];
    s.pt_iter[1] = (float)cy + s.p_iter[1];
    dx = s.pt_st[0] - s.pt_iter[0];
    dy = s.pt_st[1] - s.pt_iter[1];
    if (sqrtf(dx * dx + dy * dy) > o->outlierthresh || patch_oob(c, s.pt_iter)) {
      s.p_iter[0] = s.p_in[0];
      s.p_iter[1] = s.p_in[1];
      s.pt_iter[0] = (float)cx + s.p_iter[0];
      s.pt_iter[1] = (float)cy + s.p_iter[1];
      s.converged = 1;
    }
    patch_errimg(imb, c, o, tmpl, &s, pdiff, pweight);
  }
  p_out[0] = s.p_iter[0];
  p_out[1] = s.p_iter[1];
}

/* ------------------------------------------------------------------------------------------
 * Grid at one level: InitializeGrid + SetTargetImage + InitializeFromCoarserOF + Optimize
 * (kroeger/patchgrid.cpp:98-141, 195-211).  pflow: nop*2, pweight: nop*novals.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_grid_search(const float* ima, const float* imax, const float* imay,
                                   const float* imb, const lvl_t* c, const opt_t* o,
                                   const float* flow_coarse, float* pflow, float* pweight) {
  float* scratch = (float*)malloc(sizeof(float) * 4 * o->novals);
  int gx, gy;
  for (gx = 0; gx < c->nopw; ++gx)
    for (gy = 0; gy < c->noph; ++gy) {
      const int ip = gx * c->noph + gy;
      const int cx = gx * o->steps + c->offw, cy = gy * o->steps + c->offh;
      float pinit[2] = {0.0f, 0.0f};
      if (flow_coarse) { /* patchgrid.cpp:200-206 */
        const int x = (int)floor((float)cx / 2), y = (int)floor((float)cy / 2);
        const int i = y * (c->w / 2) + x;
        pinit[0] = flow_coarse[2 * i] * 2;
        pinit[1] = flow_coarse[2 * i + 1] * 2;
      }
      patch_run(ima, imax, imay, imb, c, o, cx, cy, pinit, pflow + 2 * ip,
                pweight + (size_t)ip * o->novals, scratch);
    }
  free(scratch);
}

/* pixel weight, kroeger/patchgrid.cpp:253-260 (and :330-337) */
static float patch_absw(const opt_t* o, const float* pw) {
  if (o->noc == 1) return 1.0f / (2.0f < pw[0] ? pw[0] : 2.0f);
  {
    float a = (2.0f < pw[0] ? pw[0] : 2.0f);
    a += (2.0f < pw[1] ? pw[1] : 2.0f);
    a += (2.0f < pw[2] ? pw[2] : 2.0f);
    return 1.0f / a;
  }
}

/* A1: AggregateFlowDense, kroeger/patchgrid.cpp:213-397.
 * pt_bw/pflow_bw/pweight_bw (complementary grid, forward-backward merge :278-375) may be NULL;
 * ptpos_bw holds the backward patches' final positions pt_iter. */
ORACLE_API void oracle_densify(const lvl_t* c, const opt_t* o, const float* pflow, const float* pweight,
                               const float* pflow_bw, const float* pweight_bw, float* flowout) {
  const int w = c->w, h = c->h, lb = -o->p / 2, ub = o->p / 2 - 1;
  float* we = (float*)calloc((size_t)w * h, sizeof(float));
  int gx, gy, x, y;
  memset(flowout, 0, sizeof(float) * 2 * w * h);
  for (gx = 0; gx < c->nopw; ++gx)
    for (gy = 0; gy < c->noph; ++gy) {
      const int ip = gx * c->noph + gy;
      const float* pw = pweight + (size_t)ip * o->novals;
      const float cx = (float)(gx * o->steps + c->offw), cy = (float)(gy * o->steps + c->offh);
      /* RGB quirk of the reference (patchgrid.cpp:241-260): the weight pointer advances by 3 only for pixels
       * inside the image and by 1 for skipped ones, so after a skipped pixel the channels are misaligned.
       * Reproduced as is. */
      for (y = lb; y <= ub; ++y)
        for (x = lb; x <= ub; ++x, ++pw) {
          const int yt = (int)(y + cy), xt = (int)(x + cx);
          if (xt >= 0 && yt >= 0 && xt < w && yt < h) {
            const int i = yt * w + xt;
            const float absw = patch_absw(o, pw); /* std::max(minerrval,*pweight) */
            pw += o->noc - 1;
            const float f0 = pflow[2 * ip] * absw, f1 = pflow[2 * ip + 1] * absw;
            we[i] += absw;
            flowout[2 * i] += f0;
            flowout[2 * i + 1] += f1;
          }
        }
    }
  if (pflow_bw) {
    for (gx = 0; gx < c->nopw; ++gx)
      for (gy = 0; gy < c->noph; ++gy) {
        const int ip = gx * c->noph + gy;
        const float* pw = pweight_bw + (size_t)ip * o->novals;
        const float cx = (float)(gx * o->steps + c->offw), cy = (float)(gy * o->steps + c->offh);
        /* GetPointPos(): pt_iter = pt_ref + p_iter */
        const float rx = cx + pflow_bw[2 * ip], ry = cy + pflow_bw[2 * ip + 1];
        /* patchgrid.cpp:304-307: double arithmetic on the ceil argument */
        const int p0 = (int)ceil(rx + .00001), p1 = (int)ceil(ry + .00001);
        const int p2 = (int)floor(rx), p3 = (int)floor(ry);
        const float r0 = rx - p2, r1 = ry - p3;
        const float wb0 = r0 * r1, wb1 = (1 - r0) * r1, wb2 = r0 * (1 - r1), wb3 = (1 - r0) * (1 - r1);
        for (y = lb; y <= ub; ++y)
          for (x = lb; x <= ub; ++x, ++pw) { /* same pointer quirk as above (patchgrid.cpp:321-337) */
            const int yt = y + p1, xt = x + p0;
            if (xt >= 1 && yt >= 1 && xt < (w - 1) && yt < (h - 1)) {
              const float absw = patch_absw(o, pw);
              pw += o->noc - 1;
              const float f0 = pflow_bw[2 * ip] * absw, f1 = pflow_bw[2 * ip + 1] * absw;
              const int cc = xt + yt * w, fc = (xt - 1) + yt * w, cf = xt + (yt - 1) * w, ff = (xt - 1) + (yt - 1) * w;
              we[cc] += wb0 * absw;
              we[fc] += wb1 * absw;
              we[cf] += wb2 * absw;
              we[ff] += wb3 * absw;
              flowout[2 * cc] -= wb0 * f0;
              flowout[2 * cc + 1] -= wb0 * f1;
              flowout[2 * fc] -= wb1 * f0;
              flowout[2 * fc + 1] -= wb1 * f1;
              flowout[2 * cf] -= wb2 * f0;
              flowout[2 * cf + 1] -= wb2 * f1;
              flowout[2 * ff] -= wb3 * f0;
              flowout[2 * ff + 1] -= wb3 * f1;
            }
          }
      }
  }
  for (y = 0; y < h; ++y)
    for (x = 0; x < w; ++x) {
      const int i = y * w + x;
      if (we[i] > 0) {
        flowout[2 * i] /= we[i];
        flowout[2 * i + 1] /= we[i];
      }
    }
  free(we);
}

/* ------------------------------------------------------------------------------------------
 * V1-V8: variational refinement of one level, grey images.
 * kroeger/refine_variational.cpp:25-116 (ctor), :153-241 (RefLevelOF);
 * kroeger/FDF1.0.1/opticalflow_aux.c, image.c, solver.c.
 * Planar images here use stride == width: the reference's stride padding columns (width%4 != 0)
 * never feed a valid pixel (checked against _ref on 30-wide levels).
 * ---------------------------------------------------------------------------------------- */
/* convolve_horiz_fast_5 (image.c:466-502) / convolve_horiz_fast_3 (:436-464): replicate borders */
static void conv_h(float* dst, const float* src, int w, int h, const float* cf, int order) {
  int x, y, k;
  for (y = 0; y < h; ++y)
    for (x = 0; x < w; ++x) {
      float acc = cf[0] * src[y * w + clampi(x - order, 0, w - 1)];
      for (k = 1; k <= 2 * order; ++k) acc = acc + cf[k] * src[y * w + clampi(x - order + k, 0, w - 1)];
      dst[y * w + x] = acc;
    }
}
/* convolve_vert_fast_5 (image.c:401-434) / convolve_vert_fast_3 (:376-399): border rows use
 * pre-summed coefficients, not replicated samples */
static void conv_v(float* dst, const float* src, int w, int h, const float* cf, int order) {
  int x, y;
  for (y = 0; y < h; ++y)
    for (x = 0; x < w; ++x) {
      const float* s = src + x;
#define S(r) s[(r) * w]
      float v;
      if (order == 2) {
        if (y == 0)
          v = (cf[0] + cf[1] + cf[2]) * S(0) + cf[3] * S(1) + cf[4] * S(2);
        else if (y == 1)
          v = (cf[0] + cf[1]) * S(0) + cf[2] * S(1) + cf[3] * S(2) + cf[4] * S(3);
        else if (y == h - 2)
          v = cf[0] * S(y - 2) + cf[1] * S(y - 1) + cf[2] * S(y) + (cf[3] + cf[4]) * S(y + 1);
        else if (y == h - 1)
          v = cf[0] * S(y - 2) + cf[1] * S(y - 1) + (cf[2] + cf[3] + cf[4]) * S(y);
        else
          v = cf[0] * S(y - 2) + cf[1] * S(y - 1) + cf[2] * S(y) + cf[3] * S(y + 1) + cf[4] * S(y + 2);
      } else {
        if (y == 0)
          v = (cf[0] + cf[1]) * S(0) + cf[2] * S(1);
        else if (y == h - 1)
          v = cf[0] * S(y - 1) + (cf[1] + cf[2]) * S(y);
        else
          v = cf[0] * S(y - 1) + cf[1] * S(y) + cf[2] * S(y + 1);
      }
#undef S
      dst[y * w + x] = v;
    }
}

/* convolve_extract_coeffs odd branch (image.c:339-342) */
static void deriv_coeffs(float* cf5, float* cf3) {
  const float h5[3] = {0.0f, -8.0f / 12.0f, 1.0f / 12.0f}; /* refine_variational.cpp:45 */
  const float h3[2] = {0.0f, -0.5f};                       /* :47 */
  int i;
  for (i = 0; i <= 2; ++i) {
    cf5[2 - i] = +h5[i];
    cf5[2 + i] = -h5[i];
  }
  for (i = 0; i <= 1; ++i) {
    cf3[1 - i] = +h3[i];
    cf3[1 + i] = -h3[i];
  }
}

#define MINMAX_TA(a, b) ((((a) > 0 ? (a) : 0)) < ((b)-1) ? ((a) > 0 ? (a) : 0) : ((b)-1))

ORACLE_API void oracle_varref(const float* ima_pad, const float* imb_pad, const lvl_t* c,
                              const dis_params* q, int noc, float* flow) {
  const int w = c->w, h = c->h, n = w * h;
  const int n_inner = q->tv_innerit * (c->lv + 1); /* refine_variational.cpp:36 */
  const float qa = 0.25f * q->tv_alpha, hg = q->tv_gamma * 0.5f / 3.0f, hd = q->tv_delta * 0.5f / 3.0f;
  const float omega = q->tv_sor;
  const float dnorm = 0.1f * 0.1f, eps = 0.001f * 0.001f; /* opticalflow_aux.c:10-14 */
  float cf5[5], cf3[3];
  float* buf = (float*)calloc((size_t)n * (20 + 12 * noc), sizeof(float));
  float *wx = buf, *wy = buf + n, *du = buf + 2 * n, *dv = buf + 3 * n;
  float *mask = buf + 4 * n, *sh = buf + 5 * n, *sv = buf + 6 * n, *uu = buf + 7 * n, *vv = buf + 8 * n;
  float *a11 = buf + 9 * n, *a12 = buf + 10 * n, *a22 = buf + 11 * n, *b1 = buf + 12 * n, *b2 = buf + 13 * n;
  float *ux = buf + 14 * n, *uy = buf + 15 * n, *vx = buf + 16 * n, *vy = buf + 17 * n, *sm = buf + 18 * n;
  /* per-channel planes (color_image_t: c1, c2, c3), 12 each */
  float *im1[3], *im2[3], *wim[3], *avg[3], *Ix[3], *Iy[3], *Iz[3], *Ixx[3], *Ixy[3], *Iyy[3], *Ixz[3], *Iyz[3];
  int i, j, k, it, iter, ch;
  for (ch = 0; ch < noc; ++ch) {
    float* pch = buf + (size_t)(20 + 12 * ch) * n;
    im1[ch] = pch; im2[ch] = pch + n; wim[ch] = pch + 2 * n; avg[ch] = pch + 3 * n; Ix[ch] = pch + 4 * n;
    Iy[ch] = pch + 5 * n; Iz[ch] = pch + 6 * n; Ixx[ch] = pch + 7 * n; Ixy[ch] = pch + 8 * n; Iyy[ch] = pch + 9 * n;
    Ixz[ch] = pch + 10 * n; Iyz[ch] = pch + 11 * n;
  }
  deriv_coeffs(cf5, cf3);
  /* refine_variational.cpp:61-82, copyimage :120-149 (de-interleaves the colour channels) */
  for (j = 0; j < h; ++j)
    for (i = 0; i < w; ++i) {
      wx[j * w + i] = flow[2 * (j * w + i)];
      wy[j * w + i] = flow[2 * (j * w + i) + 1];
      for (ch = 0; ch < noc; ++ch) {
        im1[ch][j * w + i] = ima_pad[((j + c->pad) * c->tmp_w + i + c->pad) * noc + ch];
        im2[ch][j * w + i] = imb_pad[((j + c->pad) * c->tmp_w + i + c->pad) * noc + ch];
      }
    }
  /* image_warp, opticalflow_aux.c:18-60 */
  for (j = 0; j < h; ++j)
    for (i = 0; i < w; ++i) {
      const int o = j * w + i;
      const float xx = i + wx[o], yy = j + wy[o];
      const int x = (int)floor(xx), y = (int)floor(yy);
      const float dx = xx - x, dy = yy - y;
      const int x1 = MINMAX_TA(x, w), x2 = MINMAX_TA(x + 1, w), y1 = MINMAX_TA(y, h), y2 = MINMAX_TA(y + 1, h);
      mask[o] = (xx >= 0 && xx <= w - 1 && yy >= 0 && yy <= h - 1);
      for (ch = 0; ch < noc; ++ch) {
        const float* s2 = im2[ch];
        wim[ch][o] = s2[y1 * w + x1] * (1.0f - dx) * (1.0f - dy) + s2[y1 * w + x2] * dx * (1.0f - dy) +
                     s2[y2 * w + x1] * (1.0f - dx) * dy + s2[y2 * w + x2] * dx * dy;
      }
    }
  /* get_derivatives, opticalflow_aux.c:65-116 */
  for (ch = 0; ch < noc; ++ch) {
    for (k = 0; k < n; ++k) {
      avg[ch][k] = 0.5f * (wim[ch][k] + im1[ch][k]);
      Iz[ch][k] = wim[ch][k] - im1[ch][k];
    }
    conv_h(Ix[ch], avg[ch], w, h, cf5, 2);
    conv_v(Iy[ch], avg[ch], w, h, cf5, 2);
    conv_h(Ixx[ch], Ix[ch], w, h, cf5, 2);
    conv_v(Ixy[ch], Ix[ch], w, h, cf5, 2);
    conv_v(Iyy[ch], Iy[ch], w, h, cf5, 2);
    conv_h(Ixz[ch], Iz[ch], w, h, cf5, 2);
    conv_v(Iyz[ch], Iz[ch], w, h, cf5, 2);
  }
  /* refine_variational.cpp:184-189: du = dv = 0 (calloc), uu = wx, vv = wy */
  memcpy(uu, wx, sizeof(float) * n);
  memcpy(vv, wy, sizeof(float) * n);
  for (it = 0; it < n_inner; ++it) {
    /* compute_smoothness, opticalflow_aux.c:123-165 */
    conv_h(ux, uu, w, h, cf3, 1);
    conv_h(vx, vv, w, h, cf3, 1);
    conv_v(uy, uu, w, h, cf3, 1);
    conv_v(vy, vv, w, h, cf3, 1);
    for (k = 0; k < n; ++k)
      sm[k] = qa / sqrtf(ux[k] * ux[k] + uy[k] * uy[k] + vx[k] * vx[k] + vy[k] * vy[k] + eps);
    for (j = 0; j < h; ++j)
      for (i = 0; i < w; ++i) {
        sh[j * w + i] = (i < w - 1) ? sm[j * w + i] + sm[j * w + i + 1] : 0.0f;
        sv[j * w + i] = (j < h - 1) ? sm[j * w + i] + sm[(j + 1) * w + i] : 0.0f;
      }
    /* compute_data, opticalflow_aux.c:310-438 */
    for (k = 0; k < n; ++k) {
      float A11 = 0.0f, A12 = 0.0f, A22 = 0.0f, B1 = 0.0f, B2 = 0.0f;
      if (noc == 1) { /* single channel branch */
        float tmp, tmp2, n1, n2;
        const float *ix = Ix[0], *iy = Iy[0], *iz = Iz[0], *ixx = Ixx[0], *ixy = Ixy[0], *iyy = Iyy[0], *ixz = Ixz[0], *iyz = Iyz[0];
        if (hd) {
          tmp = iz[k] + ix[k] * du[k] + iy[k] * dv[k];
          n1 = ix[k] * ix[k] + iy[k] * iy[k] + dnorm;
          tmp = mask[k] * hd / sqrtf(3 * tmp * tmp / n1 + eps);
          tmp /= n1;
          A11 += tmp * ix[k] * ix[k];
          A12 += tmp * ix[k] * iy[k];
          A22 += tmp * iy[k] * iy[k];
          B1 -= tmp * iz[k] * ix[k];
          B2 -= tmp * iz[k] * iy[k];
        }
        n1 = ixx[k] * ixx[k] + ixy[k] * ixy[k] + dnorm;
        n2 = iyy[k] * iyy[k] + ixy[k] * ixy[k] + dnorm;
        tmp = ixz[k] + ixx[k] * du[k] + ixy[k] * dv[k];
        tmp2 = iyz[k] + ixy[k] * du[k] + iyy[k] * dv[k];
        tmp = mask[k] * hg / sqrtf(3 * tmp * tmp / n1 + 3 * tmp2 * tmp2 / n2 + eps);
        tmp2 = tmp / n2;
        tmp /= n1;
        A11 += tmp * ixx[k] * ixx[k] + tmp2 * ixy[k] * ixy[k];
        A12 += tmp * ixx[k] * ixy[k] + tmp2 * ixy[k] * iyy[k];
        A22 += tmp2 * iyy[k] * iyy[k] + tmp * ixy[k] * ixy[k];
        B1 -= tmp * ixx[k] * ixz[k] + tmp2 * ixy[k] * iyz[k];
        B2 -= tmp2 * iyy[k] * iyz[k] + tmp * ixy[k] * ixz[k];
        A11 *= 3; /* :420-425, single channel only */
        A12 *= 3;
        A22 *= 3;
        B1 *= 3;
        B2 *= 3;
      } else { /* RGB branch: one robust weight over the three channels, no final x3 */
        float t[6], nn[6], psi;
        if (hd) {
          for (ch = 0; ch < 3; ++ch) {
            t[ch] = Iz[ch][k] + Ix[ch][k] * du[k] + Iy[ch][k] * dv[k];
            nn[ch] = Ix[ch][k] * Ix[ch][k] + Iy[ch][k] * Iy[ch][k] + dnorm;
          }
          psi = mask[k] * hd / sqrtf(t[0] * t[0] / nn[0] + t[1] * t[1] / nn[1] + t[2] * t[2] / nn[2] + eps);
          for (ch = 0; ch < 3; ++ch) {
            const float tc = psi / nn[ch]; /* tmp3 = tmp/n3; tmp2 = tmp/n2; tmp /= n1 */
            A11 += tc * Ix[ch][k] * Ix[ch][k];
            A12 += tc * Ix[ch][k] * Iy[ch][k];
            A22 += tc * Iy[ch][k] * Iy[ch][k];
            B1 -= tc * Iz[ch][k] * Ix[ch][k];
            B2 -= tc * Iz[ch][k] * Iy[ch][k];
          }
        }
        for (ch = 0; ch < 3; ++ch) {
          nn[2 * ch] = Ixx[ch][k] * Ixx[ch][k] + Ixy[ch][k] * Ixy[ch][k] + dnorm;
          nn[2 * ch + 1] = Iyy[ch][k] * Iyy[ch][k] + Ixy[ch][k] * Ixy[ch][k] + dnorm;
          t[2 * ch] = Ixz[ch][k] + Ixx[ch][k] * du[k] + Ixy[ch][k] * dv[k];
          t[2 * ch + 1] = Iyz[ch][k] + Ixy[ch][k] * du[k] + Iyy[ch][k] * dv[k];
        }
        psi = mask[k] * hg / sqrtf(t[0] * t[0] / nn[0] + t[1] * t[1] / nn[1] + t[2] * t[2] / nn[2] + t[3] * t[3] / nn[3] +
                                   t[4] * t[4] / nn[4] + t[5] * t[5] / nn[5] + eps);
        for (ch = 0; ch < 3; ++ch) {
          const float ta = psi / nn[2 * ch], tb = psi / nn[2 * ch + 1];
          A11 += ta * Ixx[ch][k] * Ixx[ch][k] + tb * Ixy[ch][k] * Ixy[ch][k];
          A12 += ta * Ixx[ch][k] * Ixy[ch][k] + tb * Ixy[ch][k] * Iyy[ch][k];
          A22 += tb * Iyy[ch][k] * Iyy[ch][k] + ta * Ixy[ch][k] * Ixy[ch][k];
          B1 -= ta * Ixx[ch][k] * Ixz[ch][k] + tb * Ixy[ch][k] * Iyz[ch][k];
          B2 -= tb * Iyy[ch][k] * Iyz[ch][k] + ta * Ixy[ch][k] * Ixz[ch][k];
        }
      }
      a11[k] = A11;
      a12[k] = A12;
      a22[k] = A22;
      b1[k] = B1;
      b2[k] = B2;
    }
    /* sub_laplacian x2, opticalflow_aux.c:172-199 (horizontal pass, then vertical pass) */
    for (k = 0; k < 2; ++k) {
      float* dst = k ? b2 : b1;
      const float* src = k ? wy : wx;
      for (j = 0; j < h; ++j)
        for (i = 0; i < w - 1; ++i) {
          const float t = sh[j * w + i] * (src[j * w + i + 1] - src[j * w + i]);
          dst[j * w + i] += t;
          dst[j * w + i + 1] -= t;
        }
      for (j = 0; j < h - 1; ++j)
        for (i = 0; i < w; ++i) {
          const float t = sv[j * w + i] * (src[(j + 1) * w + i] - src[j * w + i]);
          dst[j * w + i] += t;
          dst[(j + 1) * w + i] -= t;
        }
    }
    /* sor_coupled, solver.c:77-421 (w>=2, h>=2, iterations>=1; else the _slow_but_readable
     * variant :20-72 is used by the reference) */
    if (w < 2 || h < 2 || q->tv_solverit < 1) {
      for (iter = 0; iter < q->tv_solverit; ++iter)
        for (j = 0; j < h; ++j)
          for (i = 0; i < w; ++i) {
            float sigma_u = 0.0f, sigma_v = 0.0f, sum_dpsis = 0.0f, A11, A22, A12, B1, B2;
            const int o = j * w + i;
            if (j > 0) {
              sigma_u -= sv[o - w] * du[o - w];
              sigma_v -= sv[o - w] * dv[o - w];
              sum_dpsis += sv[o - w];
            }
            if (i > 0) {
              sigma_u -= sh[o - 1] * du[o - 1];
              sigma_v -= sh[o - 1] * dv[o - 1];
              sum_dpsis += sh[o - 1];
            }
            if (j < h - 1) {
              sigma_u -= sv[o] * du[o + w];
              sigma_v -= sv[o] * dv[o + w];
              sum_dpsis += sv[o];
            }
            if (i < w - 1) {
              sigma_u -= sh[o] * du[o + 1];
              sigma_v -= sh[o] * dv[o + 1];
              sum_dpsis += sh[o];
            }
            A11 = a11[o] + sum_dpsis;
            A12 = a12[o];
            A22 = a22[o] + sum_dpsis;
            B1 = b1[o] - sigma_u;
            B2 = b2[o] - sigma_v;
            du[o] = (1.0f - omega) * du[o] + omega / A11 * (B1 - A12 * dv[o]);
            dv[o] = (1.0f - omega) * dv[o] + omega / A22 * (B2 - A12 * du[o]);
          }
    } else {
      for (iter = 0; iter < q->tv_solverit; ++iter)
        for (j = 0; j < h; ++j)
          for (i = 0; i < w; ++i) {
            const int o = j * w + i;
            const float hl = i > 0 ? sh[o - 1] : 0.0f; /* f1[0] = 0 */
            const float hr = sh[o];                    /* zero in the last column */
            const float dur = i < w - 1 ? du[o + 1] : 0.0f, dvr = i < w - 1 ? dv[o + 1] : 0.0f;
            float s1, s2, B1, B2;
            if (iter == 0) { /* invert the 2x2 block in place, solver.c:115-120 / 173-178 / 231-236 */
              float dpsis, A11, A22, det;
              if (j == 0)
                dpsis = hl + hr + sv[o];
              else if (j == h - 1)
                dpsis = hl + hr + sv[o - w];
              else
                dpsis = hl + hr + sv[o - w] + sv[o];
              A11 = a22[o] + dpsis;
              A22 = a11[o] + dpsis;
              det = A11 * A22 - a12[o] * a12[o];
              a11[o] = A11 / det;
              a22[o] = A22 / det;
              a12[o] = a12[o] / (-det);
            }
            if (j == 0) {
              s1 = hr * dur + sv[o] * du[o + w] + b1[o];
              s2 = hr * dvr + sv[o] * dv[o + w] + b2[o];
            } else if (j == h - 1) {
              s1 = hr * dur + sv[o - w] * du[o - w] + b1[o];
              s2 = hr * dvr + sv[o - w] * dv[o - w] + b2[o];
            } else {
              s1 = hr * dur + sv[o - w] * du[o - w] + sv[o] * du[o + w] + b1[o];
              s2 = hr * dvr + sv[o - w] * dv[o - w] + sv[o] * dv[o + w] + b2[o];
            }
            if (i == 0) {
              B1 = s1;
              B2 = s2;
            } else {
              B1 = hl * du[o - 1] + s1;
              B2 = hl * dv[o - 1] + s2;
            }
            du[o] += omega * (a11[o] * B1 + a12[o] * B2 - du[o]);
            dv[o] += omega * (a12[o] * B1 + a22[o] * B2 - dv[o]);
          }
    }
    /* refine_variational.cpp:208-216 */
    for (k = 0; k < n; ++k) {
      uu[k] = wx[k] + du[k];
      vv[k] = wy[k] + dv[k];
    }
  }
  /* refine_variational.cpp:219-221, 92-99 */
  for (k = 0; k < n; ++k) {
    flow[2 * k] = uu[k];
    flow[2 * k + 1] = vv[k];
  }
  free(buf);
}

/* ------------------------------------------------------------------------------------------
 * E1: OFClass::OFClass, kroeger/oflow.cpp:32-363 (grey, optical flow).
 * Pyramids as in the reference boundary (oflow.h:84-111).  taps (optional, may be NULL): per level
 * index sl, pointers receiving malloc'ed copies of patch flow, dense flow (pre-refinement).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API int oracle_engine(const float* const* im_ao, const float* const* im_ao_dx,
                             const float* const* im_ao_dy, const float* const* im_bo,
                             const float* const* im_bo_dx, const float* const* im_bo_dy, int imgpadding,
                             float* outflow, const float* initflow, int width, int height,
                             const dis_params* q, float** tap_pflow, float** tap_dense, int noc) {
  opt_t o;
  const int noscales = q->lv_f - q->lv_l + 1;
  float** flow_fw = (float**)calloc(noscales, sizeof(float*));
  float** flow_bw = (float**)calloc(noscales, sizeof(float*));
  int sl;
  make_opt(q, noc, &o);
  for (sl = q->lv_f; sl >= q->lv_l; --sl) {
    const int ii = sl - q->lv_l;
    lvl_t c;
    float *pflow, *pweight, *pflow_bw = NULL, *pweight_bw = NULL, *out;
    make_lvl(q, &o, width, height, imgpadding, sl, &c);
    flow_fw[ii] = (float*)malloc(sizeof(float) * 2 * c.w * c.h);
    pflow = (float*)malloc(sizeof(float) * 2 * c.nop);
    pweight = (float*)malloc(sizeof(float) * (size_t)c.nop * o.novals);
    {
      const float* coarse = NULL;
      if (sl < q->lv_f)
        coarse = flow_fw[ii + 1];
      else if (initflow)
        coarse = initflow;
      oracle_grid_search(im_ao[sl], im_ao_dx[sl], im_ao_dy[sl], im_bo[sl], &c, &o, coarse, pflow, pweight);
    }
    if (q->usefbcon) {
      const float* coarse = sl < q->lv_f ? flow_bw[ii + 1] : NULL; /* initflow only seeds fw, oflow.cpp:217-220 */
      flow_bw[ii] = (float*)malloc(sizeof(float) * 2 * c.w * c.h);
      pflow_bw = (float*)malloc(sizeof(float) * 2 * c.nop);
      pweight_bw = (float*)malloc(sizeof(float) * (size_t)c.nop * o.novals);
      oracle_grid_search(im_bo[sl], im_bo_dx[sl], im_bo_dy[sl], im_ao[sl], &c, &o, coarse, pflow_bw, pweight_bw);
    }
    out = (sl == q->lv_l) ? outflow : flow_fw[ii];
    oracle_densify(&c, &o, pflow, pweight, pflow_bw, pweight_bw, out);
    if (q->usefbcon && sl > q->lv_l) oracle_densify(&c, &o, pflow_bw, pweight_bw, pflow, pweight, flow_bw[ii]);
    if (tap_pflow) {
      tap_pflow[sl] = (float*)malloc(sizeof(float) * 2 * c.nop);
      memcpy(tap_pflow[sl], pflow, sizeof(float) * 2 * c.nop);
    }
    if (tap_dense) {
      tap_dense[sl] = (float*)malloc(sizeof(float) * 2 * c.w * c.h);
      memcpy(tap_dense[sl], out, sizeof(float) * 2 * c.w * c.h);
    }
    if (q->usetvref) {
      oracle_varref(im_ao[sl], im_bo[sl], &c, q, noc, out);
      if (q->usefbcon && sl > q->lv_l) oracle_varref(im_bo[sl], im_ao[sl], &c, q, noc, flow_bw[ii]);
    }
    free(pflow);
    free(pweight);
    free(pflow_bw);
    free(pweight_bw);
  }
  for (sl = 0; sl < noscales; ++sl) {
    free(flow_fw[sl]);
    free(flow_bw[sl]);
  }
  free(flow_fw);
  free(flow_bw);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * O1: kroeger/run_dense.cpp:407-414: flow *= 2^lv_l; cv::resize(x 2^lv_l, INTER_LINEAR); crop.
 * OpenCV's bilinear: src = (dst+0.5)/s - 0.5 in float, floor, clamp to the edge; separable
 * lerp, horizontal then vertical (OpenCV arithmetic; cv2 agrees to ~1e-6, not bit-exact).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_finish(const float* flow_l, int wl, int hl, int lv_l, int left, int top,
                              int w_org, int h_org, float* out) {
  const int sc = 1 << lv_l;
  int x, y;
  for (y = 0; y < h_org; ++y)
    for (x = 0; x < w_org; ++x) {
      const int X = x + left, Y = y + top;
      float u, v;
      if (lv_l == 0) {
        u = flow_l[2 * (Y * wl + X)];
        v = flow_l[2 * (Y * wl + X) + 1];
      } else {
        float fx = (float)((X + 0.5) * (1.0 / sc) - 0.5), fy = (float)((Y + 0.5) * (1.0 / sc) - 0.5);
        int sx = (int)floorf(fx), sy = (int)floorf(fy), sx1, sy1;
        fx -= sx;
        fy -= sy;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= wl - 1) { fx = 0; sx = wl - 1; }
        if (sy < 0) { fy = 0; sy = 0; }
        if (sy >= hl - 1) { fy = 0; sy = hl - 1; }
        sx1 = sx + 1 < wl ? sx + 1 : wl - 1;
        sy1 = sy + 1 < hl ? sy + 1 : hl - 1;
        {
          const float* r0 = flow_l + 2 * (sy * wl);
          const float* r1 = flow_l + 2 * (sy1 * wl);
          const float s = (float)sc;
          const float a0 = (r0[2 * sx] * s) * (1.f - fx) + (r0[2 * sx1] * s) * fx;
          const float a1 = (r1[2 * sx] * s) * (1.f - fx) + (r1[2 * sx1] * s) * fx;
          const float b0 = (r0[2 * sx + 1] * s) * (1.f - fx) + (r0[2 * sx1 + 1] * s) * fx;
          const float b1 = (r1[2 * sx + 1] * s) * (1.f - fx) + (r1[2 * sx1 + 1] * s) * fx;
          u = a0 * (1.f - fy) + a1 * fy;
          v = b0 * (1.f - fy) + b1 * fy;
        }
      }
      out[2 * (y * w_org + x)] = u;
      out[2 * (y * w_org + x) + 1] = v;
    }
}

/* Whole run_dense data path on decoded grey images (kroeger/run_dense.cpp:298-414).
 * flow_out: w*h*2 (full resolution); level_out (optional): raw engine output at level lv_l. */
ORACLE_API int oracle_run_u8c(const uint8_t* a, const uint8_t* b, int w, int h, int pitch, int noc,
                              const dis_params* q, float* flow_out, float* level_out) {
  const int nl = q->lv_f + 1;
  float** P[6];
  int wp, hp, left, top, k, l;
  float* lvl;
  oracle_padded_size(w, h, q->lv_f, &wp, &hp, &left, &top);
  for (k = 0; k < 6; ++k) P[k] = (float**)calloc(nl, sizeof(float*));
  oracle_build_pyramid_c(a, w, h, pitch, noc, q->lv_f, q->patchsz, P[0], P[1], P[2]);
  oracle_build_pyramid_c(b, w, h, pitch, noc, q->lv_f, q->patchsz, P[3], P[4], P[5]);
  lvl = (float*)malloc(sizeof(float) * 2 * (wp >> q->lv_l) * (hp >> q->lv_l));
  oracle_engine((const float* const*)P[0], (const float* const*)P[1], (const float* const*)P[2],
                (const float* const*)P[3], (const float* const*)P[4], (const float* const*)P[5],
                q->patchsz, lvl, NULL, wp, hp, q, NULL, NULL, noc);
  if (level_out) memcpy(level_out, lvl, sizeof(float) * 2 * (wp >> q->lv_l) * (hp >> q->lv_l));
  if (flow_out) oracle_finish(lvl, wp >> q->lv_l, hp >> q->lv_l, q->lv_l, left, top, w, h, flow_out);
  free(lvl);
  for (k = 0; k < 6; ++k) {
    for (l = 0; l < nl; ++l) free(P[k][l]);
    free(P[k]);
  }
  return 0;
}

ORACLE_API int oracle_run_u8(const uint8_t* a, const uint8_t* b, int w, int h, int pitch,
                             const dis_params* q, float* flow_out, float* level_out) {
  return oracle_run_u8c(a, b, w, h, pitch, 1, q, flow_out, level_out);
}
