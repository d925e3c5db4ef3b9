// varref.cu -- stage 4: variational refinement of one pyramid level (grey images).
//
// Replaces VarRefClass::VarRefClass / RefLevelOF (kroeger/refine_variational.cpp:25-116, 153-241)
// and the FDF1.0.1 C kernels underneath:
//   image_warp            kroeger/FDF1.0.1/opticalflow_aux.c:18-60      -> k_warp
//   get_derivatives       opticalflow_aux.c:65-116, image.c:401-434,466-502 -> k_deriv1, k_deriv2
//   compute_smoothness    opticalflow_aux.c:123-165, image.c:376-399,436-464 \
//   compute_data          opticalflow_aux.c:310-438 (1-channel branch)        > k_assemble
//   sub_laplacian (x2)    opticalflow_aux.c:172-199                          /
//   sor_coupled           kroeger/FDF1.0.1/solver.c:77-421              -> k_assemble (2x2 block
//                         inverse of the first sweep) + k_sor_wavefront (all sweeps)
//   uu = wx+du, write-back  refine_variational.cpp:208-221, 92-99       -> k_assemble / k_update
//
// Layout: the level flow, (du,dv), (b1,b2) and (smooth_horiz,smooth_vert) are float2 images of
// w x h (pitch = w); the derivative stack and the inverted 2x2 blocks are planar w x h.  The
// reference's `stride` padding columns (width % 4 != 0) never feed a valid pixel, so pitch = w.
// The padded pyramid images are read in place (no copyimage pass).
//
// SOR keeps the reference's *lexicographic* Gauss-Seidel dependency order exactly (pixel (i,j)
// sees new (i-1,j), new (i,j-1), old (i+1,j), old (i,j+1)) by sweeping a skewed wavefront: one
// warp owns 32 consecutive rows, lane l works on column s-l at step s, so the upper neighbour
// is the previous step's result of lane l-1 (one shuffle) and the left neighbour is the lane's
// own previous result.  Row blocks and successive sweeps chase each other through per-item
// progress counters in global memory (items are handed out through a ticket in dependency
// order, so a running warp only ever waits for warps that have already started).
#include "common.cuh"

namespace dis {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr float kDnorm = 0.1f * 0.1f;    // datanorm, opticalflow_aux.c:10
constexpr float kEps = 0.001f * 0.001f;  // epsilon_color/grad/smooth, :11-14

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- image_warp + first half of get_derivatives ------------------------------------------------
__global__ void __launch_bounds__(256) k_warp(int w, int h, int pad, int pitch, const float* __restrict__ I0,
                                              const float* __restrict__ I1, const float2* __restrict__ flow,
                                              float* __restrict__ avg, float* __restrict__ Iz,
                                              float* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= w || j >= h) return;
  const int o = j * w + i;
  const float2 f = flow[o];
  const float xx = (float)i + f.x, yy = (float)j + f.y;
  const int x = (int)floorf(xx), y = (int)floorf(yy);
  const float dx = xx - (float)x, dy = yy - (float)y;
  const float m = (xx >= 0 && xx <= (float)(w - 1) && yy >= 0 && yy <= (float)(h - 1)) ? 1.0f : 0.0f;
  const int x1 = clampi(x, 0, w - 1), x2 = clampi(x + 1, 0, w - 1);
  const int y1 = clampi(y, 0, h - 1), y2 = clampi(y + 1, 0, h - 1);
  const float* s = I1 + (size_t)pad * pitch + pad;
  const float s11 = __ldg(s + (size_t)y1 * pitch + x1), s12 = __ldg(s + (size_t)y1 * pitch + x2);
  const float s21 = __ldg(s + (size_t)y2 * pitch + x1), s22 = __ldg(s + (size_t)y2 * pitch + x2);
  const float wv = s11 * (1.0f - dx) * (1.0f - dy) + s12 * dx * (1.0f - dy) + s21 * (1.0f - dx) * dy +
                   s22 * dx * dy;
  const float i0 = __ldg(I0 + (size_t)(j + pad) * pitch + i + pad);
  avg[o] = 0.5f * (wv + i0);
  Iz[o] = wv - i0;
  mask[o] = m;
}

// 5-tap derivative filter (refine_variational.cpp:45 + convolve_extract_coeffs image.c:339-342):
// coeffs = {1/12, -8/12, -0, 8/12, -1/12}
struct Cf5 {
  float c0, c1, c2, c3, c4;
};
__device__ __forceinline__ Cf5 cf5() {
  Cf5 c;
  c.c0 = 1.0f / 12.0f;
  c.c1 = -8.0f / 12.0f;
  c.c2 = -0.0f;
  c.c3 = -(-8.0f / 12.0f);
  c.c4 = -(1.0f / 12.0f);
  return c;
}
// convolve_horiz_fast_5 (image.c:466-502): replicated border samples
__device__ __forceinline__ float conv_h5(const float* __restrict__ s, int w, int i, int rowoff) {
  const Cf5 c = cf5();
  const float* r = s + rowoff;
  return c.c0 * r[max(i - 2, 0)] + c.c1 * r[max(i - 1, 0)] + c.c2 * r[i] + c.c3 * r[min(i + 1, w - 1)] +
         c.c4 * r[min(i + 2, w - 1)];
}
// convolve_vert_fast_5 (image.c:401-434): border rows use pre-summed coefficients
__device__ __forceinline__ float conv_v5(const float* __restrict__ s, int w, int h, int i, int j) {
  const Cf5 c = cf5();
  const float* p = s + i;
#define S(r) p[(size_t)(r)*w]
  if (j == 0) return (c.c0 + c.c1 + c.c2) * S(0) + c.c3 * S(1) + c.c4 * S(2);
  if (j == 1) return (c.c0 + c.c1) * S(0) + c.c2 * S(1) + c.c3 * S(2) + c.c4 * S(3);
  if (j == h - 2) return c.c0 * S(j - 2) + c.c1 * S(j - 1) + c.c2 * S(j) + (c.c3 + c.c4) * S(j + 1);
  if (j == h - 1) return c.c0 * S(j - 2) + c.c1 * S(j - 1) + (c.c2 + c.c3 + c.c4) * S(j);
  return c.c0 * S(j - 2) + c.c1 * S(j - 1) + c.c2 * S(j) + c.c3 * S(j + 1) + c.c4 * S(j + 2);
#undef S
}

__global__ void __launch_bounds__(256) k_deriv1(int w, int h, const float* __restrict__ avg,
                                                const float* __restrict__ Iz, float* __restrict__ Ix,
                                                float* __restrict__ Iy, float* __restrict__ Ixz,
                                                float* __restrict__ Iyz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= w || j >= h) return;
  const int o = j * w + i;
  Ix[o] = conv_h5(avg, w, i, j * w);
  Iy[o] = conv_v5(avg, w, h, i, j);
  Ixz[o] = conv_h5(Iz, w, i, j * w);
  Iyz[o] = conv_v5(Iz, w, h, i, j);
}

__global__ void __launch_bounds__(256) k_deriv2(int w, int h, const float* __restrict__ Ix,
                                                const float* __restrict__ Iy, float* __restrict__ Ixx,
                                                float* __restrict__ Ixy, float* __restrict__ Iyy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= w || j >= h) return;
  const int o = j * w + i;
  Ixx[o] = conv_h5(Ix, w, i, j * w);
  Ixy[o] = conv_v5(Ix, w, h, i, j);
  Iyy[o] = conv_v5(Iy, w, h, i, j);
}

// ---- one inner fixed-point iteration: smoothness + data term + Laplacian RHS + block inverse ----
struct AssembleArgs {
  int w, h;
  float qa, hg, hd;
  int first;  // first inner iteration: uu = wx (memcpy), du = dv = 0
  const float2* flow;  // wx, wy
  const float2* duv;
  const float *mask, *Ix, *Iy, *Iz, *Ixx, *Ixy, *Iyy, *Ixz, *Iyz;
  float *a11, *a12, *a22;
  float2 *b, *hv;
};

__device__ __forceinline__ float2 uu_at(const AssembleArgs& a, int i, int j) {
  const int o = j * a.w + i;
  const float2 f = a.flow[o];
  if (a.first) return f;
  const float2 d = a.duv[o];
  return make_float2(f.x + d.x, f.y + d.y);  // refine_variational.cpp:212-213
}

// smoothness weight s(i,j), opticalflow_aux.c:128-137 with the 3-tap filters of image.c:376-399,436-464
__device__ __forceinline__ float smooth_at(const AssembleArgs& a, int i, int j) {
  const float c0 = -0.5f, c1 = -0.0f, c2 = 0.5f;  // deriv_flow, refine_variational.cpp:47
  const int w = a.w, h = a.h;
  const float2 m = uu_at(a, i, j);
  const float2 l = uu_at(a, max(i - 1, 0), j), r = uu_at(a, min(i + 1, w - 1), j);
  const float ux = c0 * l.x + c1 * m.x + c2 * r.x;
  const float vx = c0 * l.y + c1 * m.y + c2 * r.y;
  float uy, vy;
  if (j == 0) {
    const float2 d = uu_at(a, i, 1);
    uy = (c0 + c1) * m.x + c2 * d.x;
    vy = (c0 + c1) * m.y + c2 * d.y;
  } else if (j == h - 1) {
    const float2 u = uu_at(a, i, j - 1);
    uy = c0 * u.x + (c1 + c2) * m.x;
    vy = c0 * u.y + (c1 + c2) * m.y;
  } else {
    const float2 u = uu_at(a, i, j - 1), d = uu_at(a, i, j + 1);
    uy = c0 * u.x + c1 * m.x + c2 * d.x;
    vy = c0 * u.y + c1 * m.y + c2 * d.y;
  }
  return a.qa / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + kEps);
}

__global__ void __launch_bounds__(256) k_assemble(const AssembleArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int w = a.w, h = a.h;
  if (i >= w || j >= h) return;
  const int o = j * w + i;
  // compute_smoothness: horiz(i,j) = s(i,j)+s(i+1,j) (0 for i >= w-1), vert likewise
  const float sc = smooth_at(a, i, j);
  const float hr = (i < w - 1) ? sc + smooth_at(a, i + 1, j) : 0.0f;
  const float vb = (j < h - 1) ? sc + smooth_at(a, i, j + 1) : 0.0f;
  const float hl = (i > 0) ? smooth_at(a, i - 1, j) + sc : 0.0f;
  const float vt = (j > 0) ? smooth_at(a, i, j - 1) + sc : 0.0f;

  // compute_data, 1-channel branch
  const float2 d = a.first ? make_float2(0.0f, 0.0f) : a.duv[o];
  const float du = d.x, dv = d.y;
  const float mk = a.mask[o];
  const float ix = a.Ix[o], iy = a.Iy[o], iz = a.Iz[o];
  const float ixx = a.Ixx[o], ixy = a.Ixy[o], iyy = a.Iyy[o], ixz = a.Ixz[o], iyz = a.Iyz[o];
  float A11 = 0.0f, A12 = 0.0f, A22 = 0.0f, B1 = 0.0f, B2 = 0.0f;
  float tmp, tmp2, n1, n2;
  if (a.hd != 0.0f) {
    tmp = iz + ix * du + iy * dv;
    n1 = ix * ix + iy * iy + kDnorm;
    tmp = mk * a.hd / sqrtf(3 * tmp * tmp / n1 + kEps);
    tmp /= n1;
    A11 += tmp * ix * ix;
    A12 += tmp * ix * iy;
    A22 += tmp * iy * iy;
    B1 -= tmp * iz * ix;
    B2 -= tmp * iz * iy;
  }
  n1 = ixx * ixx + ixy * ixy + kDnorm;
  n2 = iyy * iyy + ixy * ixy + kDnorm;
  tmp = ixz + ixx * du + ixy * dv;
  tmp2 = iyz + ixy * du + iyy * dv;
  tmp = mk * a.hg / sqrtf(3 * tmp * tmp / n1 + 3 * tmp2 * tmp2 / n2 + kEps);
  tmp2 = tmp / n2;
  tmp /= n1;
  A11 += tmp * ixx * ixx + tmp2 * ixy * ixy;
  A12 += tmp * ixx * ixy + tmp2 * ixy * iyy;
  A22 += tmp2 * iyy * iyy + tmp * ixy * ixy;
  B1 -= tmp * ixx * ixz + tmp2 * ixy * iyz;
  B2 -= tmp2 * iyy * iyz + tmp * ixy * ixz;
  A11 *= 3;
  A12 *= 3;
  A22 *= 3;
  B1 *= 3;
  B2 *= 3;

  // sub_laplacian(b1, wx) and (b2, wy): horizontal pass then vertical pass, in source order
  const float2 fc = a.flow[o];
  if (i > 0) {
    const float2 fl = a.flow[o - 1];
    B1 -= hl * (fc.x - fl.x);
    B2 -= hl * (fc.y - fl.y);
  }
  if (i < w - 1) {
    const float2 fr = a.flow[o + 1];
    B1 += hr * (fr.x - fc.x);
    B2 += hr * (fr.y - fc.y);
  }
  if (j > 0) {
    const float2 fu = a.flow[o - w];
    B1 -= vt * (fc.x - fu.x);
    B2 -= vt * (fc.y - fu.y);
  }
  if (j < h - 1) {
    const float2 fd = a.flow[o + w];
    B1 += vb * (fd.x - fc.x);
    B2 += vb * (fd.y - fc.y);
  }

  // 2x2 block inverse of sor_coupled's first sweep (solver.c:115-120, 173-178, 231-236)
  float dpsis;
  if (j == 0)
    dpsis = hl + hr + vb;
  else if (j == h - 1)
    dpsis = hl + hr + vt;
  else
    dpsis = hl + hr + vt + vb;
  const float iA11 = A22 + dpsis, iA22 = A11 + dpsis;
  const float det = iA11 * iA22 - A12 * A12;
  a.a11[o] = iA11 / det;
  a.a22[o] = iA22 / det;
  a.a12[o] = A12 / (-det);
  a.b[o] = make_float2(B1, B2);
  a.hv[o] = make_float2(hr, vb);
}

// ---- sor_coupled: exact lexicographic sweeps as a skewed wavefront ------------------------------
struct SorArgs {
  int w, h, T, K;  // T sweeps, K = ceil(h/32) row blocks
  float omega;
  const float *a11, *a12, *a22;
  const float2 *b, *hv;
  float2* duv;
  int* prog;  // [T*K] completed steps per item, then [1] ticket
};

constexpr int kSorPublish = 8;  // publish progress every this many steps

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__global__ void __launch_bounds__(32) k_sor_wavefront(const SorArgs a) {
  const int lane = threadIdx.x;
  const int w = a.w, h = a.h, T = a.T, K = a.K;
  int tk = 0;
  if (lane == 0) tk = atomicAdd(a.prog + T * K, 1);
  tk = __shfl_sync(FULL, tk, 0);
  // decode ticket -> (t,k): items ordered by key = 2t + k, so that (t,k-1), (t-1,k), (t-1,k+1)
  // all hold smaller tickets
  int t = -1, k = -1;
  {
    int cnt = 0;
    for (int key = 0; key <= 2 * (T - 1) + K - 1 && t < 0; ++key)
      for (int tt = 0; tt < T; ++tt) {
        const int kk = key - 2 * tt;
        if (kk < 0 || kk >= K) continue;
        if (cnt == tk) {
          t = tt;
          k = kk;
          break;
        }
        ++cnt;
      }
  }
  if (t < 0) return;
  const int j = k * 32 + lane;
  const bool rowok = j < h;
  const int total = w + 31;
  const int* p_up = (k > 0) ? a.prog + t * K + k - 1 : nullptr;
  const int* p_prev = (t > 0) ? a.prog + (t - 1) * K + k : nullptr;
  const int* p_below = (t > 0 && k < K - 1) ? a.prog + (t - 1) * K + k + 1 : nullptr;
  int seen_up = 0, seen_prev = 0, seen_below = 0;
  float2 res_prev = make_float2(0.0f, 0.0f);    // my result of the previous step (new (i-1,j))
  float2 right_prev = make_float2(0.0f, 0.0f);  // old (i,j), loaded as "right" one step earlier
  float hl = 0.0f;                              // horiz(i-1,j); 0 in the first column (f1[0] = 0)
  float v_prev = 0.0f;                          // vert(i-1..): my vert weight of the previous step
  const float omega = a.omega;

  for (int s = 0; s < total; ++s) {
    // ---- wait for the producers of this step
    {
      bool polled = false;
      if (p_up) {
        const int need = min(s + 32, total);
        if (seen_up < need) {
          do seen_up = ld_volatile(p_up); while (seen_up < need);
          polled = true;
        }
      }
      if (p_prev) {
        const int need = min(s + 2, total);
        if (seen_prev < need) {
          do seen_prev = ld_volatile(p_prev); while (seen_prev < need);
          polled = true;
        }
      }
      if (p_below) {
        const int need = min(max(s - 30, 0), total);
        if (seen_below < need) {
          do seen_below = ld_volatile(p_below); while (seen_below < need);
          polled = true;
        }
      }
      if (polled) __threadfence();
    }
    const int i = s - lane;
    const bool act = rowok && i >= 0 && i < w;
    float2 up = make_float2(__shfl_up_sync(FULL, res_prev.x, 1), __shfl_up_sync(FULL, res_prev.y, 1));
    float vt = __shfl_up_sync(FULL, v_prev, 1);
    if (act) {
      const int o = j * w + i;
      if (lane == 0 && j > 0) {
        up = __ldcg(a.duv + o - w);
        vt = __ldg(a.hv + o - w).y;
      }
      const float2 hvv = __ldg(a.hv + o);
      const float2 bb = __ldg(a.b + o);
      const float A11 = __ldg(a.a11 + o), A12 = __ldg(a.a12 + o), A22 = __ldg(a.a22 + o);
      const float2 right = (i < w - 1) ? __ldcg(a.duv + o + 1) : make_float2(0.0f, 0.0f);
      const float2 self = (i == 0) ? __ldcg(a.duv + o) : right_prev;
      float s1, s2;
      if (j == 0) {
        const float2 below = __ldcg(a.duv + o + w);
        s1 = hvv.x * right.x + hvv.y * below.x + bb.x;
        s2 = hvv.x * right.y + hvv.y * below.y + bb.y;
      } else if (j == h - 1) {
        s1 = hvv.x * right.x + vt * up.x + bb.x;
        s2 = hvv.x * right.y + vt * up.y + bb.y;
      } else {
        const float2 below = __ldcg(a.duv + o + w);
        s1 = hvv.x * right.x + vt * up.x + hvv.y * below.x + bb.x;
        s2 = hvv.x * right.y + vt * up.y + hvv.y * below.y + bb.y;
      }
      float B1, B2;
      if (i == 0) {
        B1 = s1;
        B2 = s2;
      } else {
        B1 = hl * res_prev.x + s1;
        B2 = hl * res_prev.y + s2;
      }
      float2 nv;
      nv.x = self.x + omega * (A11 * B1 + A12 * B2 - self.x);
      nv.y = self.y + omega * (A12 * B1 + A22 * B2 - self.y);
      __stcg(a.duv + o, nv);
      res_prev = nv;
      right_prev = right;
      hl = hvv.x;
      v_prev = hvv.y;
    }
    // ---- publish progress
    if (((s + 1) % kSorPublish) == 0 || s + 1 == total) {
      __threadfence();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile int*>(a.prog + t * K + k) = s + 1;
    }
  }
}

// final flow = wx + du (refine_variational.cpp:212-221)
__global__ void __launch_bounds__(256) k_update(int n, float2* __restrict__ flow, const float2* __restrict__ duv) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const float2 f = flow[o], d = duv[o];
  flow[o] = make_float2(f.x + d.x, f.y + d.y);
}

}  // namespace

size_t varref_progress_ints(int h, int n_solver) { return (size_t)n_solver * ((h + 31) / 32) + 1; }

// Returns the number of kernels launched, or -1 for an unsupported level shape.
int launch_varref(const LevelGeom& g, const VarParams& v, const float* I0, const float* I1, float2* flow,
                  const VarRefBuffers& b, cudaStream_t st, Prof* prof) {
  const int w = g.w, h = g.h, n = w * h;
  if (w < 2 || h < 4 || v.n_solver < 1) return -1;  // reference would take its slow path / read out of range
  int launches = 0;
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  // algorithmic bytes per SURVEY.md section 8(d): warp+mask 28 B/px, derivative stack 40 B/px
  {
    ProfScope ps(prof, "k_warp", g.lv, 28.0 * n);
    k_warp<<<grid, block, 0, st>>>(w, h, g.pad, g.pitch, I0, I1, flow, b.avg, b.Iz, b.mask);
  }
  {
    ProfScope ps(prof, "k_deriv1", g.lv, 24.0 * n);
    k_deriv1<<<grid, block, 0, st>>>(w, h, b.avg, b.Iz, b.Ix, b.Iy, b.Ixz, b.Iyz);
  }
  {
    ProfScope ps(prof, "k_deriv2", g.lv, 16.0 * n);
    k_deriv2<<<grid, block, 0, st>>>(w, h, b.Ix, b.Iy, b.Ixx, b.Ixy, b.Iyy);
  }
  launches += 3;
  if (v.n_inner <= 0) return launches;
  const int K = (h + 31) / 32, T = v.n_solver;
  for (int it = 0; it < v.n_inner; ++it) {
    AssembleArgs aa{w, h, v.qa, v.hg, v.hd, it == 0 ? 1 : 0, flow, b.duv, b.mask, b.Ix, b.Iy, b.Iz,
                    b.Ixx, b.Ixy, b.Iyy, b.Ixz, b.Iyz, b.a11, b.a12, b.a22, b.b, b.hv};
    {
      // smoothness 16 + data term 64 + sub_laplacian 32 + flow update 24 B/px
      ProfScope ps(prof, "k_assemble", g.lv, 136.0 * n);
      k_assemble<<<grid, block, 0, st>>>(aa);
    }
    if (it == 0) cudaMemsetAsync(b.duv, 0, sizeof(float2) * n, st);
    cudaMemsetAsync(b.progress, 0, sizeof(int) * (T * K + 1), st);
    SorArgs sa{w, h, T, K, v.omega, b.a11, b.a12, b.a22, b.b, b.hv, b.duv, b.progress};
    {
      // each sweep reads 9 arrays and writes 2: 44 B/px
      ProfScope ps(prof, "k_sor_wavefront", g.lv, 44.0 * T * n);
      k_sor_wavefront<<<T * K, 32, 0, st>>>(sa);
    }
    launches += 2;
  }
  {
    ProfScope ps(prof, "k_update", g.lv, 8.0 * n);
    k_update<<<(n + 255) / 256, 256, 0, st>>>(n, flow, b.duv);
  }
  return launches + 1;
}

}  // namespace dis
