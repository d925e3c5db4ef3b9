// varref.cu -- stage 4: variational refinement of one pyramid level (grey images).
//
// Replaces VarRefClass::VarRefClass / RefLevelOF (kroeger/refine_variational.cpp:25-116, 153-241)
// and the FDF1.0.1 C kernels underneath:
//   image_warp            kroeger/FDF1.0.1/opticalflow_aux.c:18-60      -> k_front (stage A)
//   get_derivatives       opticalflow_aux.c:65-116, image.c:401-434,466-502 -> k_front (stages B, C)
//   compute_smoothness    opticalflow_aux.c:123-165, image.c:376-399,436-464 -> k_assemble
//   compute_data          opticalflow_aux.c:310-438 (1-channel branch)        -> k_assemble
//   sub_laplacian (x2)    opticalflow_aux.c:172-199                          -> k_assemble
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
// warp owns 32 consecutive rows ("row block"), lane l works on column s - SK*l at step s, so the
// upper neighbour is lane l-1's result of SK steps ago (one shuffle, off the critical path for
// SK = 2) and the left neighbour is the lane's own previous result.
//
// Wavefront-major layout: everything the sweep touches is stored per row block as [step][lane]
// (element (row 32k+l, column i) lives at step s = i + SK*l), so each step of a warp is one
// fully coalesced 512-byte access per array -- {a11,a12,a22,horiz} and {b1,b2,vert,-} as two
// float4 streams, (du,dv) as float4 pairs of two consecutive steps -- prefetched D steps ahead
// into a small shared-memory ring with cp.async.cg (L1 is bypassed: other SMs write these lines).
// Row blocks hand their last row to the block below through a flag-in-data boundary buffer
// (16-byte {du,dv,tag} stores, no fences); successive sweeps chase each other through coarse
// progress counters.  Items (sweep t, row block k) are handed out through a ticket in
// dependency order, so a running warp only ever waits for warps that have already started.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

// Compiled twice like patch_search.cu: varref_fast.o (-fmad=true -DDIS_ARITH_FAST) serves DIS_OPT_ARITH = 1.
#ifdef DIS_ARITH_FAST
#define launch_varref launch_varref_fast
#define varref_init_device varref_init_device_fast
#define varref_sizes varref_sizes_fast
#endif

namespace dis {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr float kDnorm = 0.1f * 0.1f;    // datanorm, opticalflow_aux.c:10
constexpr float kEps = 0.001f * 0.001f;  // epsilon_color/grad/smooth, :11-14

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

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
// ---- fused, shared-memory-tiled front end: image_warp + both passes of get_derivatives in one kernel -----------------
// Tile of FTX x FTY pixels per block.  Stage A warps I1 and forms avg = 0.5 (I1w + I0), Iz = I1w - I0 on the tile plus
// a halo of 4 in shared memory; stage B takes the 5-tap derivatives Ix, Iy of avg on the tile plus a halo of 2 (shared)
// and Ixz, Iyz of Iz on the tile (global); stage C the second derivatives Ixx, Ixy = d/dy Ix, Iyy on the tile.  avg
// never leaves the SM.  noc channels: the padded pyramid images are interleaved (copyimage de-interleaves them in the
// reference, refine_variational.cpp:120-149); the stack holds noc planes of n = w*h floats per array, the mask one.
// Every value is produced by the reference's expression from the same operands -- the horizontal filter replicates border samples by clamping the index, the vertical one
// uses the reference's pre-summed border coefficients on the image's first / last two rows -- so the planes are
// bit-identical; samples outside the image are never read.
constexpr int FTX = 32, FTY = 16;
struct FrontArgs {
  int w, h, pad, pitch, noc;
  const float *I0, *I1;
  const float2* flow;
  float* stack;  // derivative stack + mask (VarRefBuffers): array k at stack + k * astride
  unsigned astride;
  size_t bstride;
};

__global__ void __launch_bounds__(256) k_front(const FrontArgs a) {
  pdl_wait();
  enum { S_IZ = 1, S_IX = 2, S_IY = 3, S_IXX = 4, S_IXY = 5, S_IYY = 6, S_IXZ = 7, S_IYZ = 8, S_MASK = 9 };
  __shared__ float avg_s[FTY + 8][FTX + 8], iz_s[FTY + 8][FTX + 8];
  __shared__ float ix_s[FTY + 4][FTX + 4], iy_s[FTY + 4][FTX + 4];
  const size_t boff = (size_t)blockIdx.z * a.bstride;
  const float* __restrict__ I0 = bshift_nn(a.I0, boff);
  const float* __restrict__ I1 = bshift_nn(a.I1, boff);
  const float2* __restrict__ flow = bshift_nn(a.flow, boff);
  float* __restrict__ S = bshift_nn(a.stack, boff);
  const unsigned AS = a.astride;
  const int w = a.w, h = a.h, pad = a.pad, pitch = a.pitch, noc = a.noc, n = w * h;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int i0 = blockIdx.x * FTX, j0 = blockIdx.y * FTY;
  const Cf5 c = cf5();
  const float* s1 = I1 + (size_t)pad * pitch + pad * noc;
  for (int ch = 0; ch < noc; ++ch) {
    // ---- stage A: warp (opticalflow_aux.c:18-60) and avg / Iz (:70-71) on the tile + halo 4
    for (int q = tid; q < (FTY + 8) * (FTX + 8); q += 256) {
      const int ly = q / (FTX + 8), lx = q - ly * (FTX + 8);
      const int i = i0 + lx - 4, j = j0 + ly - 4;
      if (i < 0 || i >= w || j < 0 || j >= h) continue;
      const int o = j * w + i;
      const float2 f = flow[o];
      const float xx = (float)i + f.x, yy = (float)j + f.y;
      const int x = (int)floorf(xx), y = (int)floorf(yy);
      const float dx = xx - (float)x, dy = yy - (float)y;
      const int x1 = clampi(x, 0, w - 1), x2 = clampi(x + 1, 0, w - 1);
      const int y1 = clampi(y, 0, h - 1), y2 = clampi(y + 1, 0, h - 1);
      const float s11 = __ldg(s1 + (size_t)y1 * pitch + x1 * noc + ch), s12 = __ldg(s1 + (size_t)y1 * pitch + x2 * noc + ch);
      const float s21 = __ldg(s1 + (size_t)y2 * pitch + x1 * noc + ch), s22 = __ldg(s1 + (size_t)y2 * pitch + x2 * noc + ch);
      const float wv = s11 * (1.0f - dx) * (1.0f - dy) + s12 * dx * (1.0f - dy) + s21 * (1.0f - dx) * dy +
                       s22 * dx * dy;
      const float v0 = __ldg(I0 + (size_t)(j + pad) * pitch + (i + pad) * noc + ch);
      avg_s[ly][lx] = 0.5f * (wv + v0);
      const float iz = wv - v0;
      iz_s[ly][lx] = iz;
      if (lx >= 4 && lx < FTX + 4 && ly >= 4 && ly < FTY + 4) {  // the tile itself
        S[S_IZ * AS + (unsigned)(ch * n + o)] = iz;
        if (ch == 0) S[S_MASK * AS + (unsigned)o] = (xx >= 0 && xx <= (float)(w - 1) && yy >= 0 && yy <= (float)(h - 1)) ? 1.0f : 0.0f;
      }
    }
    __syncthreads();
    // 5-tap filters on a shared tile whose element (ly, lx) is image pixel (j0 + ly - H, i0 + lx - H)
    auto hconv = [&](const float* row, int i, int H) {  // convolve_horiz_fast_5 (image.c:466-502)
      const int b = i0 - H;
      return c.c0 * row[max(i - 2, 0) - b] + c.c1 * row[max(i - 1, 0) - b] + c.c2 * row[i - b] + c.c3 * row[min(i + 1, w - 1) - b] +
             c.c4 * row[min(i + 2, w - 1) - b];
    };
    // ---- stage B: Ix, Iy of avg on the tile + halo 2; Ixz, Iyz of Iz on the tile
    for (int q = tid; q < (FTY + 4) * (FTX + 4); q += 256) {
      const int ly = q / (FTX + 4), lx = q - ly * (FTX + 4);
      const int i = i0 + lx - 2, j = j0 + ly - 2;
      if (i < 0 || i >= w || j < 0 || j >= h) continue;
      const int ay = ly + 2, ax = lx + 2;  // the same pixel in the halo-4 tiles
#define V5(T_, X_)                                                                                                      \
  (j == 0       ? (c.c0 + c.c1 + c.c2) * T_[ay][X_] + c.c3 * T_[ay + 1][X_] + c.c4 * T_[ay + 2][X_]                                 \
   : j == 1     ? (c.c0 + c.c1) * T_[ay - 1][X_] + c.c2 * T_[ay][X_] + c.c3 * T_[ay + 1][X_] + c.c4 * T_[ay + 2][X_]                  \
   : j == h - 2 ? c.c0 * T_[ay - 2][X_] + c.c1 * T_[ay - 1][X_] + c.c2 * T_[ay][X_] + (c.c3 + c.c4) * T_[ay + 1][X_]                  \
   : j == h - 1 ? c.c0 * T_[ay - 2][X_] + c.c1 * T_[ay - 1][X_] + (c.c2 + c.c3 + c.c4) * T_[ay][X_]                                 \
                : c.c0 * T_[ay - 2][X_] + c.c1 * T_[ay - 1][X_] + c.c2 * T_[ay][X_] + c.c3 * T_[ay + 1][X_] + c.c4 * T_[ay + 2][X_])
      const float ix = hconv(avg_s[ay], i, 4), iy = V5(avg_s, ax);
      ix_s[ly][lx] = ix;
      iy_s[ly][lx] = iy;
      if (lx >= 2 && lx < FTX + 2 && ly >= 2 && ly < FTY + 2) {
        const unsigned o = (unsigned)(ch * n + j * w + i);
        S[S_IX * AS + o] = ix;
        S[S_IY * AS + o] = iy;
        S[S_IXZ * AS + o] = hconv(iz_s[ay], i, 4);
        S[S_IYZ * AS + o] = V5(iz_s, ax);
      }
#undef V5
    }
    __syncthreads();
    // ---- stage C: second derivatives on the tile
    for (int q = tid; q < FTY * FTX; q += 256) {
      const int ty = q / FTX, tx = q - ty * FTX;
      const int i = i0 + tx, j = j0 + ty;
      if (i >= w || j >= h) continue;
      const int ay = ty + 2, ax = tx + 2;
#define V5(T_, X_)                                                                                                      \
  (j == 0       ? (c.c0 + c.c1 + c.c2) * T_[ay][X_] + c.c3 * T_[ay + 1][X_] + c.c4 * T_[ay + 2][X_]                                 \
   : j == 1     ? (c.c0 + c.c1) * T_[ay - 1][X_] + c.c2 * T_[ay][X_] + c.c3 * T_[ay + 1][X_] + c.c4 * T_[ay + 2][X_]                  \
   : j == h - 2 ? c.c0 * T_[ay - 2][X_] + c.c1 * T_[ay - 1][X_] + c.c2 * T_[ay][X_] + (c.c3 + c.c4) * T_[ay + 1][X_]                  \
   : j == h - 1 ? c.c0 * T_[ay - 2][X_] + c.c1 * T_[ay - 1][X_] + (c.c2 + c.c3 + c.c4) * T_[ay][X_]                                 \
                : c.c0 * T_[ay - 2][X_] + c.c1 * T_[ay - 1][X_] + c.c2 * T_[ay][X_] + c.c3 * T_[ay + 1][X_] + c.c4 * T_[ay + 2][X_])
      const unsigned o = (unsigned)(ch * n + j * w + i);
      S[S_IXX * AS + o] = hconv(ix_s[ay], i, 2);
      S[S_IXY * AS + o] = V5(ix_s, ax);
      S[S_IYY * AS + o] = V5(iy_s, ax);
#undef V5
    }
    __syncthreads();
  }
}

// ---- one inner fixed-point iteration: smoothness + data term + Laplacian RHS + block inverse ----
constexpr int SK = 1;  // wavefront skew (columns per row) of k_sor_wavefront; levels that take k_sor_small use 2

// wavefront-major addressing of one level
struct Skew {
  int w, h, K, sk, nsteps, nsp;  // sk: columns per row; nsp: steps per row block as allocated (whole TMA chunks + prefetch overrun)
  __host__ __device__ Skew(int w_, int h_, int sk_ = SK)
      : w(w_), h(h_), K((h_ + 31) / 32), sk(sk_), nsteps(w_ + sk_ * 31), nsp((w_ + sk_ * 31 + 15) / 16 * 16 + 32) {}
  __host__ __device__ size_t at(int i, int j) const {  // float4 index of pixel (i,j)
    const int k = j >> 5, l = j & 31;
    return ((size_t)k * nsp + i + sk * l) * 32 + l;
  }
};

struct AssembleArgs {
  int w, h;
  float qa, hg, hd;
  int first;  // first inner iteration: uu = wx (memcpy), du = dv = 0
  int noc;    // 1: single-channel data term; 3: RGB data term (planes of w*h floats)
  const float2* flow;  // wx, wy
  const float4* du4;   // (du,dv) records, wavefront-major
  const float* stack;  // derivative stack + mask: array k at stack + k * astride (VarRefBuffers)
  unsigned astride;
  float4 *coefA, *coefB;  // {a11,a12,a22,horiz}, {b1,b2,vert,0}, wavefront-major
  int* prog;              // SOR flags: [0] epoch, [1] ticket, [2..2+n_prog) per-item progress counters
  int n_prog;
  int skew;               // columns per row of the wavefront-major layout (Skew::sk)
  size_t bstride;         // batched handles: blockIdx.z = pair
};

__device__ __forceinline__ float2 uu_at(const AssembleArgs& a, int i, int j, float2* d_out) {
  const float2 f = a.flow[j * a.w + i];
  if (a.first) {
    *d_out = make_float2(0.0f, 0.0f);
    return f;
  }
  const float4 d = a.du4[Skew(a.w, a.h, a.skew).at(i, j)];
  *d_out = make_float2(d.x, d.y);
  return make_float2(f.x + d.x, f.y + d.y);  // refine_variational.cpp:212-213
}

// Tile of 32 x 8 pixels per block.  Phase 1 stages uu = wx + du with a halo of 2 (clamped to the image,
// which is exactly the replicate border of the reference's horizontal 3-tap filter); phase 2 computes the
// smoothness weight s once per pixel of the tile + halo 1; phase 3 does the per-pixel assembly; phase 4
// writes the wavefront-major coefficient streams through shared memory so that the 8 lanes of one SOR
// step that live in this tile are written as one contiguous 128-byte run.
constexpr int ATX = 32, ATY = 8;

// 8 blocks per SM (32 registers, a few spills): the kernel waits on global loads, so resident warps beat registers
// (+3 % pairs/s against the compiler's own choice of 48 registers, A/B in one session)
__global__ void __launch_bounds__(ATX* ATY, 8) k_assemble(const AssembleArgs a_in) {
  pdl_wait();
  AssembleArgs a = a_in;
  {
    const size_t boff = (size_t)blockIdx.z * a.bstride;
    a.flow = bshift_nn(a.flow, boff); a.du4 = bshift_nn(a.du4, boff); a.stack = bshift_nn(a.stack, boff);
    a.coefA = bshift_nn(a.coefA, boff); a.coefB = bshift_nn(a.coefB, boff); a.prog = bshift_nn(a.prog, boff);
  }
  // array k of the derivative stack (written by k_front of this level): 32-bit indices off one base
  enum { S_IZ = 1, S_IX = 2, S_IY = 3, S_IXX = 4, S_IXY = 5, S_IYY = 6, S_IXZ = 7, S_IYZ = 8, S_MASK = 9 };
  const float* __restrict__ S = a.stack;
  const unsigned AS = a.astride;
#define SP(k, idx) S[(unsigned)(k) * AS + (unsigned)(idx)]
  __shared__ float2 uu_s[ATY + 4][ATX + 4];
  __shared__ float2 du_s[ATY][ATX];
  __shared__ float s_s[ATY + 2][ATX + 2];
  __shared__ float4 oa_s[ATY][ATX + 1], ob_s[ATY][ATX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * ATX + tx;
  const int i0 = blockIdx.x * ATX, j0 = blockIdx.y * ATY;
  const int i = i0 + tx, j = j0 + ty;
  const int w = a.w, h = a.h;
  if (blockIdx.x == 0 && blockIdx.y == 0) {  // reset the SOR flags of the sweep that follows
    for (int q = tid; q <= a.n_prog; q += ATX * ATY) a.prog[1 + q] = 0;  // ticket + counters
    if (tid == 0) a.prog[0] += 1;  // epoch: makes the boundary tags of every launch unique
  }
  // ---- phase 1: uu (and du for the tile itself)
  for (int q = tid; q < (ATY + 4) * (ATX + 4); q += ATX * ATY) {
    const int ly = q / (ATX + 4), lx = q - ly * (ATX + 4);
    const int gi = min(max(i0 + lx - 2, 0), w - 1), gj = min(max(j0 + ly - 2, 0), h - 1);
    float2 d;
    uu_s[ly][lx] = uu_at(a, gi, gj, &d);
    if (lx >= 2 && lx < ATX + 2 && ly >= 2 && ly < ATY + 2) du_s[ly - 2][lx - 2] = d;
  }
  __syncthreads();
  // ---- phase 2: smoothness weight s (opticalflow_aux.c:128-137; 3-tap filters image.c:376-399, 436-464)
  for (int q = tid; q < (ATY + 2) * (ATX + 2); q += ATX * ATY) {
    const int ly = q / (ATX + 2), lx = q - ly * (ATX + 2);
    const int gj = j0 + ly - 1;
    const float c0 = -0.5f, c1 = -0.0f, c2 = 0.5f;  // deriv_flow, refine_variational.cpp:47
    const float2 m = uu_s[ly + 1][lx + 1], l = uu_s[ly + 1][lx], r = uu_s[ly + 1][lx + 2];
    const float2 u = uu_s[ly][lx + 1], dn = uu_s[ly + 2][lx + 1];
    const float ux = c0 * l.x + c1 * m.x + c2 * r.x;
    const float vx = c0 * l.y + c1 * m.y + c2 * r.y;
    float uy, vy;
    if (gj <= 0) {
      uy = (c0 + c1) * m.x + c2 * dn.x;
      vy = (c0 + c1) * m.y + c2 * dn.y;
    } else if (gj >= h - 1) {
      uy = c0 * u.x + (c1 + c2) * m.x;
      vy = c0 * u.y + (c1 + c2) * m.y;
    } else {
      uy = c0 * u.x + c1 * m.x + c2 * dn.x;
      vy = c0 * u.y + c1 * m.y + c2 * dn.y;
    }
    s_s[ly][lx] = a.qa / sqrtf(ux * ux + uy * uy + vx * vx + vy * vy + kEps);
  }
  __syncthreads();
  // ---- phase 3: per-pixel assembly
  const bool inside = i < w && j < h;
  if (inside) {
    const int o = j * w + i;
    // compute_smoothness: horiz(i,j) = s(i,j)+s(i+1,j) (0 for i >= w-1), vert likewise
    const float sc = s_s[ty + 1][tx + 1];
    const float hr = (i < w - 1) ? sc + s_s[ty + 1][tx + 2] : 0.0f;
    const float vb = (j < h - 1) ? sc + s_s[ty + 2][tx + 1] : 0.0f;
    const float hl = (i > 0) ? s_s[ty + 1][tx] + sc : 0.0f;
    const float vt = (j > 0) ? s_s[ty][tx + 1] + sc : 0.0f;

    // compute_data (opticalflow_aux.c:310-438)
    const float du = du_s[ty][tx].x, dv = du_s[ty][tx].y;
    const float mk = SP(S_MASK, o);
    float A11 = 0.0f, A12 = 0.0f, A22 = 0.0f, B1 = 0.0f, B2 = 0.0f;
    if (a.noc == 1) {  // 1-channel branch
      const float ix = SP(S_IX, o), iy = SP(S_IY, o), iz = SP(S_IZ, o);
      const float ixx = SP(S_IXX, o), ixy = SP(S_IXY, o), iyy = SP(S_IYY, o), ixz = SP(S_IXZ, o), iyz = SP(S_IYZ, o);
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
      A11 *= 3;  // single channel only (:420-425)
      A12 *= 3;
      A22 *= 3;
      B1 *= 3;
      B2 *= 3;
    } else {  // RGB branch: one robust weight over the three channels, no final x3
      const unsigned n = (unsigned)(w * h);
      float t[6], nn[6], psi;
      if (a.hd != 0.0f) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float ix = SP(S_IX, ch * n + o), iy = SP(S_IY, ch * n + o), iz = SP(S_IZ, ch * n + o);
          t[ch] = iz + ix * du + iy * dv;
          nn[ch] = ix * ix + iy * iy + kDnorm;
        }
        psi = mk * a.hd / sqrtf(t[0] * t[0] / nn[0] + t[1] * t[1] / nn[1] + t[2] * t[2] / nn[2] + kEps);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float ix = SP(S_IX, ch * n + o), iy = SP(S_IY, ch * n + o), iz = SP(S_IZ, ch * n + o);
          const float tc = psi / nn[ch];
          A11 += tc * ix * ix;
          A12 += tc * ix * iy;
          A22 += tc * iy * iy;
          B1 -= tc * iz * ix;
          B2 -= tc * iz * iy;
        }
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float ixx = SP(S_IXX, ch * n + o), ixy = SP(S_IXY, ch * n + o), iyy = SP(S_IYY, ch * n + o);
        nn[2 * ch] = ixx * ixx + ixy * ixy + kDnorm;
        nn[2 * ch + 1] = iyy * iyy + ixy * ixy + kDnorm;
        t[2 * ch] = SP(S_IXZ, ch * n + o) + ixx * du + ixy * dv;
        t[2 * ch + 1] = SP(S_IYZ, ch * n + o) + ixy * du + iyy * dv;
      }
      psi = mk * a.hg / sqrtf(t[0] * t[0] / nn[0] + t[1] * t[1] / nn[1] + t[2] * t[2] / nn[2] + t[3] * t[3] / nn[3] +
                              t[4] * t[4] / nn[4] + t[5] * t[5] / nn[5] + kEps);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float ixx = SP(S_IXX, ch * n + o), ixy = SP(S_IXY, ch * n + o), iyy = SP(S_IYY, ch * n + o);
        const float ixz = SP(S_IXZ, ch * n + o), iyz = SP(S_IYZ, ch * n + o);
        const float ta = psi / nn[2 * ch], tb = psi / nn[2 * ch + 1];
        A11 += ta * ixx * ixx + tb * ixy * ixy;
        A12 += ta * ixx * ixy + tb * ixy * iyy;
        A22 += tb * iyy * iyy + ta * ixy * ixy;
        B1 -= ta * ixx * ixz + tb * ixy * iyz;
        B2 -= tb * iyy * iyz + ta * ixy * ixz;
      }
    }

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
    oa_s[ty][tx] = make_float4(iA11 / det, A12 / (-det), iA22 / det, hr);
    ob_s[ty][tx] = make_float4(B1, B2, vb, 0.0f);
  }
  __syncthreads();
  // ---- phase 4: wavefront-major stores.  Pixels of this tile with equal tx + skew * ty share one SOR step; their
  // 8 lanes (rows j0..j0+7 of one 32-row block) are contiguous in the [step][lane] layout.
  const Skew sk(w, h, a.skew);
  for (int q = tid; q < (ATX + a.skew * (ATY - 1)) * ATY; q += ATX * ATY) {
    const int sd = q / ATY, ry = q - sd * ATY, rx = sd - a.skew * ry;
    if (rx < 0 || rx >= ATX) continue;
    const int gi = i0 + rx, gj = j0 + ry;
    if (gi >= w || gj >= h) continue;
    const size_t oc = sk.at(gi, gj);
    a.coefA[oc] = oa_s[ry][rx];
    a.coefB[oc] = ob_s[ry][rx];
  }
#undef SP
}

// ---- sor_coupled: exact lexicographic sweeps as a skewed wavefront ------------------------------
struct SorArgs {
  int w, h, T, K;  // T sweeps, K = ceil(h/32) row blocks
  float omega;
  const float4 *coefA, *coefB;  // wavefront-major [K][nsp][32]
  float4* du4;                  // wavefront-major records {du, dv, tag, -}: [K+1][nsp][32]
  int* prog;                    // [0] epoch, [1] ticket, [2 + t*K + k] pacing hint: completed steps of item (t,k)
  size_t bstride;               // batched handles: blockIdx.y = pair
};

// kG = steps per group: prefetch / validation / pacing-hint granularity, = steps per TMA chunk of the coefficient
// streams (kCH); the record ring holds the group being consumed and the prefetched one (kRD = 2 kG).
// Two instantiations: kG = 16 has the lowest latency for a lone pair (fewest per-group overheads; 49.5 KB of shared
// memory per warp), kG = 8 is ~10 % slower alone but holds 24.9 KB, which is worth +5.6 % pairs/s when many pairs
// share the GPU.
constexpr int kNS = 2;  // TMA stages in flight
template <int kG>
constexpr size_t sor_smem() {
  return (size_t)kNS * kG * 32 * 16 * 2 + (size_t)(2 * kG) * 32 * 16 + 3 * (2 * kG) * 16 + kNS * 8;
}

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 ld_volatile4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// TMA (bulk async copy) + mbarrier
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int kG, bool kPersistent>
__global__ void __launch_bounds__(32, kG == 8 ? 16 : 8) k_sor_wavefront(const SorArgs a_in) {
  pdl_wait();
  constexpr int kCH = kG, kRD = 2 * kG;
  SorArgs a = a_in;
  {
    const size_t boff = (size_t)blockIdx.y * a.bstride;
    a.coefA = bshift(a.coefA, boff); a.coefB = bshift(a.coefB, boff); a.du4 = bshift(a.du4, boff);
    a.prog = bshift(a.prog, boff);
  }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const sA = smem_raw;                                          // [kNS*kCH][32] float4
  unsigned char* const sB = smem_raw + (size_t)kNS * kCH * 512;                // [kNS*kCH][32] float4
  unsigned char* const sD = smem_raw + (size_t)kNS * kCH * 1024;               // [kRD][32] float4, slot x: old[x+2]
  unsigned char* const sUp = sD + (size_t)kRD * 512;                           // [kRD] float4
  unsigned char* const sDn = sUp + kRD * 16;                                   // [kRD] float4
  unsigned char* const sVt = sDn + kRD * 16;                                   // [kRD] float4
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sVt + kRD * 16);  // [kNS]

  const int lane = threadIdx.x;
  const int w = a.w, h = a.h, T = a.T, K = a.K;
  const Skew sk(w, h);
  const int nsteps = sk.nsteps, nsp = sk.nsp;
  // Persistent warp: the launch has only as many CTAs as items are busy at a time (launch_varref); a warp that has
  // finished an item takes the next ticket.  Tickets are handed out in dependency order and a warp takes a new one
  // only after finishing, so every item a running warp waits for is finished or running: no deadlock for any CTA
  // count.  (With one CTA per item, two thirds of the resident warps only waited for their turn, holding their 25 KB.)
  if (lane == 0) {
    for (int q = 0; q < kNS; ++q) mbar_init(&bars[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncwarp();
  unsigned chunks_done = 0;  // chunks consumed by this warp so far: stage and phase parity of the coefficient ring
  for (;;) {
  int tk = 0;
  if (lane == 0) tk = atomicAdd(a.prog + 1, 1);
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
  const bool has_up = k > 0, has_dn = (k < K - 1);
  const bool no_up = (j == 0), no_dn = (j >= h - 1);
  const bool edge_blk = (k == 0) || (k == K - 1);
  const int epoch = ld_volatile(a.prog);
  // epoch >= 1, T <= 256 (dis_params_validate); unsigned arithmetic: the epoch may wrap, tags only have to differ
  // from those of the launches that last wrote the same records
  const int tag_cur = (int)(((unsigned)epoch << 8) | (unsigned)t), tag_prev = (int)(((unsigned)epoch << 8) | (unsigned)(t - 1));
  const bool chk_old = (t > 0);                           // my old records carry the previous sweep's tag
  const int* p_prev = (t > 0) ? a.prog + 2 + (t - 1) * K + k : nullptr;                // same block, previous sweep
  const int* p_upb = has_up ? a.prog + 2 + t * K + k - 1 : nullptr;                     // block above, same sweep
  const int* p_dnb = (t > 0 && has_dn) ? a.prog + 2 + (t - 1) * K + k + 1 : nullptr;    // block below, previous sweep
  int seen_prev = 0, seen_upb = 0, seen_dnb = 0;
  const float4* gA = a.coefA + (size_t)k * nsp * 32;
  const float4* gB = a.coefB + (size_t)k * nsp * 32;
  float4* gD = a.du4 + (size_t)k * nsp * 32 + lane;                             // my records
  const float4* gUp = a.du4 + ((size_t)(has_up ? k - 1 : 0) * nsp) * 32 + 31;  // block above, lane 31
  const float4* gDn = a.du4 + ((size_t)(k + 1) * nsp) * 32;                    // block below, lane 0
  const float4* gVt = a.coefB + ((size_t)(has_up ? k - 1 : 0) * nsp) * 32 + 31;
  const float omega = a.omega;
  const int nchunks = (nsteps + kCH - 1) / kCH;

  // ---- coefficient streams: TMA bulk copies of kCH steps (8 KB per array) into a kNS-stage ring; chunk c of this
  // item is the warp's chunk number chunks_done + c
  const unsigned c_base = chunks_done;
  auto tma_chunk = [&](int c) {  // lane 0 only
    const int st = (int)((c_base + (unsigned)c) % kNS);
    mbar_expect_tx(&bars[st], 2 * kCH * 512);
    tma_bulk_g2s(sA + (size_t)st * kCH * 512, gA + (size_t)c * kCH * 32, kCH * 512, &bars[st]);
    tma_bulk_g2s(sB + (size_t)st * kCH * 512, gB + (size_t)c * kCH * 32, kCH * 512, &bars[st]);
  };
  if (lane == 0)
    for (int c = 0; c < kNS && c < nchunks; ++c) tma_chunk(c);

  // ---- record streams in groups of 8 steps: paced by the producers' progress hints, validated by tag
  auto pace = [&](const int* p, int& seen, int need) {
    need = min(need, nsteps);
    if (p && seen < need) {
      do seen = ld_volatile(p); while (seen < need);
    }
  };
  const unsigned sD_u = smem_u32(sD) + lane * 16, sUp_u = smem_u32(sUp), sDn_u = smem_u32(sDn), sVt_u = smem_u32(sVt);
  // loads that steps x0..x0+7 will consume.  Addresses past a row block's end fall into the next block's
  // (allocated) storage and are never used, so nothing here is predicated on the column range.
  auto tick8 = [&](int x0) {
    pace(p_prev, seen_prev, x0 + kG - 1 + 2 + 1);
    pace(p_upb, seen_upb, min(x0 + kG - 1, w - 1) + 31 * SK + 1);
    pace(p_dnb, seen_dnb, min(x0 + kG - 1 - 31 * SK, w - 1) + 1);
    const unsigned slot = (unsigned)(x0 & (kRD - 1));
    const float4* src = gD + (size_t)(x0 + 2) * 32;
#pragma unroll
    for (int q = 0; q < kG; ++q)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sD_u + ((slot + q) << 9)), "l"(src + q * 32) : "memory");
    if (lane < kG && has_up) {  // lane q: column x0+q of the row above = block k-1, lane 31, step x0+q + 31*SK
      const size_t o = (size_t)(x0 + lane + 31 * SK) * 32;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sUp_u + ((slot + lane) << 4)), "l"(gUp + o) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sVt_u + ((slot + lane) << 4)), "l"(gVt + o) : "memory");
    }
    if (lane >= 32 - kG && has_dn) {  // lane 24+q: column x0+q - 31*SK of the row below = block k+1, lane 0
      const long long o = (long long)(x0 + (lane - (32 - kG)) - 31 * SK) * 32;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sDn_u + ((slot + lane - (32 - kG)) << 4)), "l"(gDn + o) : "memory");
    }
    cp_async_commit();
  };
  pace(p_prev, seen_prev, 2);
  float4 rec0 = ld_volatile4(gD), rec1 = ld_volatile4(gD + 32);
  if (chk_old) {
    while (__float_as_int(rec0.z) != tag_prev) rec0 = ld_volatile4(gD);
    while (__float_as_int(rec1.z) != tag_prev) rec1 = ld_volatile4(gD + 32);
  }
  tick8(0);

  float2 self_old = make_float2(rec0.x, rec0.y);  // old[s]   (lane 0's first "self" is old[0])
  float2 old1 = make_float2(rec1.x, rec1.y);      // old[s+1]
  float2 res1 = make_float2(0.f, 0.f), res2 = make_float2(0.f, 0.f);  // my results 1 and 2 steps ago
  float v1 = 0.f, v2 = 0.f;                                           // my vert weights 1 and 2 steps ago
  float hl = 0.f;                                                     // horiz(i-1,j)
  int i = -SK * lane;                                                 // my column at step s
  float4* gOut = gD;
  volatile int* my_hint = a.prog + 2 + t * K + k;
  const float tagf = __int_as_float(tag_cur);

  for (int c = 0; c < nchunks; ++c) {
    const int st = (int)((c_base + (unsigned)c) % kNS);
    mbar_wait(&bars[st], ((c_base + (unsigned)c) / kNS) & 1u);
    const unsigned char* cA_p = sA + (size_t)st * kCH * 512 + lane * 16;
    const unsigned char* cB_p = sB + (size_t)st * kCH * 512 + lane * 16;
#pragma unroll
    for (int g = 0; g < kCH / kG; ++g) {
      const int s0 = c * kCH + g * kG;
      tick8(s0 + kG);
      cp_async_wait<1>();  // everything issued for steps < s0+8 has landed
      // ---- stage this group's records in registers and validate their tags (a mismatch means a
      // prefetch overtook its producer: rare, handled by polling the source)
      const unsigned slot = (unsigned)(s0 & (kRD - 1));
      float2 o2[kG];
      {
        float4 r[kG];
        int bad = 0;
#pragma unroll
        for (int q = 0; q < kG; ++q) {
          r[q] = *reinterpret_cast<const float4*>(sD + ((slot + q) << 9) + lane * 16);
          bad |= __float_as_int(r[q].z) ^ tag_prev;
        }
        if (chk_old && bad) {
#pragma unroll
          for (int q = 0; q < kG; ++q)
            if (s0 + q + 2 < nsteps)
              while (__float_as_int(r[q].z) != tag_prev) r[q] = ld_volatile4(gD + (size_t)(s0 + q + 2) * 32);
        }
#pragma unroll
        for (int q = 0; q < kG; ++q) o2[q] = make_float2(r[q].x, r[q].y);
      }
      if (has_up && lane < kG) {  // lane q checks the record lane 0 will use at step s0+q and adds the vert weight
        float4 r = *reinterpret_cast<const float4*>(sUp + ((slot + lane) << 4));
        if (s0 + lane < w && __float_as_int(r.z) != tag_cur)
          do r = ld_volatile4(gUp + (size_t)(s0 + lane + 31 * SK) * 32); while (__float_as_int(r.z) != tag_cur);
        r.w = reinterpret_cast<const float4*>(sVt + ((slot + lane) << 4))->z;
        *reinterpret_cast<float4*>(sUp + ((slot + lane) << 4)) = r;
      }
      if (chk_old && has_dn && lane >= 32 - kG) {  // lane 24+q checks the record lane 31 will use at step s0+q
        const int iq = s0 + (lane - (32 - kG)) - 31 * SK;
        float4 r = *reinterpret_cast<const float4*>(sDn + ((slot + lane - (32 - kG)) << 4));
        if (iq >= 0 && iq < w && __float_as_int(r.z) != tag_prev) {
          do r = ld_volatile4(gDn + (size_t)iq * 32); while (__float_as_int(r.z) != tag_prev);
          *reinterpret_cast<float4*>(sDn + ((slot + lane - (32 - kG)) << 4)) = r;
        }
      }
      __syncwarp();
      // ---- 8 steps.  EDGE: this block holds the first or last image row (those rows drop a term);
      // INTERIOR: every lane is strictly inside its row for the whole group (no row start / end).
      auto steps8 = [&](auto edge_c, auto interior_c) {
        constexpr bool EDGE = decltype(edge_c)::value, INTERIOR = decltype(interior_c)::value;
#pragma unroll
        for (int q = 0; q < kG; ++q) {
          const int so = g * kG + q;
          const float4 cA = *reinterpret_cast<const float4*>(cA_p + so * 512);
          const float4 cB = *reinterpret_cast<const float4*>(cB_p + so * 512);
          const float2 old2 = o2[q];
          // Neighbour rows by rotation: lane l takes "below" from lane l+1 and "up"/vt from lane l-1 (mod 32).
          // The wrap-around sources are the hand-off records of the neighbouring row blocks: lane 0 offers
          // the old value of the first row of the block below (read by lane 31), lane 31 offers the new
          // value + vert weight of the last row of the block above (read by lane 0).  Their own values are
          // not needed by anyone through these shuffles, so no branch and no extra shuffle is required.
          float2 bsrc = (SK == 2) ? old2 : old1, usrc = (SK == 2) ? res2 : res1;
          float vsrc = (SK == 2) ? v2 : v1;
          if (has_dn) {
            const float2 r = *reinterpret_cast<const float2*>(sDn + ((slot + q) << 4));
            bsrc.x = (lane == 0) ? r.x : bsrc.x;
            bsrc.y = (lane == 0) ? r.y : bsrc.y;
          }
          if (has_up) {
            const float4 r = *reinterpret_cast<const float4*>(sUp + ((slot + q) << 4));
            usrc.x = (lane == 31) ? r.x : usrc.x;
            usrc.y = (lane == 31) ? r.y : usrc.y;
            vsrc = (lane == 31) ? r.w : vsrc;
          }
          const float2 below = make_float2(__shfl_sync(FULL, bsrc.x, (lane + 1) & 31), __shfl_sync(FULL, bsrc.y, (lane + 1) & 31));
          const float2 up = make_float2(__shfl_sync(FULL, usrc.x, (lane + 31) & 31), __shfl_sync(FULL, usrc.y, (lane + 31) & 31));
          const float vt = __shfl_sync(FULL, vsrc, (lane + 31) & 31);
          // ---- the reference's update (solver.c:122-131 / 180-190 / 237-247), both components.
          // old1 is exactly 0 past the last column (inactive steps store zeros), like the reference's f2/f3.
          float px = cA.w * old1.x, py = cA.w * old1.y;
          const float ux = px + vt * up.x, uy = py + vt * up.y;
          if (EDGE) {
            px = no_up ? px : ux;
            py = no_up ? py : uy;
          } else {
            px = ux;
            py = uy;
          }
          const float qx = px + cB.z * below.x, qy = py + cB.z * below.y;
          if (EDGE) {
            px = no_dn ? px : qx;
            py = no_dn ? py : qy;
          } else {
            px = qx;
            py = qy;
          }
          const float s1 = px + cB.x, s2 = py + cB.y;
          const float l1 = hl * res1.x + s1, l2 = hl * res1.y + s2;
          const float B1 = (!INTERIOR && i == 0) ? s1 : l1, B2 = (!INTERIOR && i == 0) ? s2 : l2;
          float2 nv;
          nv.x = self_old.x + omega * (cA.x * B1 + cA.y * B2 - self_old.x);
          nv.y = self_old.y + omega * (cA.y * B1 + cA.z * B2 - self_old.y);
          if (!INTERIOR) {
            const bool act = (unsigned)i < (unsigned)w;
            nv.x = act ? nv.x : 0.0f;
            nv.y = act ? nv.y : 0.0f;
            ++i;
          }
          __stcg(gOut + so * 32, make_float4(nv.x, nv.y, tagf, 0.f));
          if (SK == 2) {
            res2 = res1;
            v2 = v1;
          }
          res1 = nv;
          v1 = cB.z;
          hl = cA.w;
          self_old = old1;
          old1 = old2;
        }
        if (INTERIOR) i += kG;
      };
      const bool interior = (s0 > SK * 31) && (s0 + kG <= w);
      if (edge_blk) {
        if (interior)
          steps8(std::true_type{}, std::true_type{});
        else
          steps8(std::true_type{}, std::false_type{});
      } else {
        if (interior)
          steps8(std::false_type{}, std::true_type{});
        else
          steps8(std::false_type{}, std::false_type{});
      }
      // ---- pacing hint for the consumers of this item (no fence: records are validated by tag)
      if (lane == 0) *my_hint = min(s0 + kG, nsteps);
    }
    gOut += kCH * 32;
    // ---- chunk consumed: refill its stage
    __syncwarp();
    if (lane == 0 && c + kNS < nchunks) tma_chunk(c + kNS);
  }
  cp_async_wait<0>();
  if (!kPersistent) return;  // one CTA per item (DIS_OPT_SOR_GROUP = 16 asked for explicitly: the latency setting)
  chunks_done += (unsigned)nchunks;
  __syncwarp();
  }  // next ticket
}

// ---- the SOR of a small level: ONE CTA per pair, one warp per (sweep, row block), one barrier per step ------------
// Thread (s, j) owns image row j in sweep s (32 rows per warp = the row blocks of the wavefront-major layout).  At
// step t it updates the C pixels (i0 ... i0 + C - 1, j), i0 = C (t - j - 2 s): the lexicographic Gauss-Seidel order
// of sor_coupled only asks that (i-1, j) and (i, j-1) of the same sweep and (i+1, j), (i, j+1) of the previous sweep
// are final; with rows C columns apart and sweeps two steps apart all of them were produced at step t - 1 or by the
// thread itself.  The whole launch is w / C + h - 1 + 2 (T - 1) steps -- the wavefront kernel's pipeline of (sweep, row
// block) items needs a lag of 2 kG + 2 steps per sweep and ~54 per row block, which is most of its
// time on a coarse level.  Every thread publishes its pixels of the step in a double-buffered shared array
// res[parity][s][j][C]; neighbours read it after the step's barrier.  Levels that take this kernel use the
// wavefront-major layout with a skew of C columns per row (Skew::sk), so that the C columns of a step are C whole
// wavefront steps of the layout: the coefficient streams (and, for sweep 0, the du records of the previous launch)
// arrive in shared rings of wavefront steps filled by cp.async D steps ahead, 512 contiguous bytes per warp, array and
// wavefront step, the three arrays shared out over the sweeps' warps; only the last sweep writes records.  A warp
// whose rows are all outside the image columns at a step only publishes zeros.  Arithmetic per pixel: that of
// k_sor_wavefront, same order.  Measured (profiles/README.md): a step costs ~250 cycles + ~110 per pixel, so C = 2.
constexpr size_t kSorSmallMaxSmem = 200 * 1024;
constexpr int kSmallC = 2;  // columns per step (template parameter C of k_sor_small; 1, 2 and 4 work)

__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a) : "memory");
  return v;
}
template <int C>
struct SorSmallDist {
  static constexpr int value = C == 4 ? 3 : C == 2 ? 4 : 8;  // prefetch distance in steps
};
template <int C>
inline size_t sor_small_smem(int rows, int T) {
  constexpr int D = SorSmallDist<C>::value;
  return (size_t)rows * C * ((D + 2 * T) * 32 + D * 8 + 2 * T * 8);
}

template <int C>
__global__ void __launch_bounds__(512) k_sor_small(const SorArgs a_in, const int first) {
  pdl_wait();
  constexpr int D = SorSmallDist<C>::value;
  SorArgs a = a_in;
  {
    const size_t boff = (size_t)blockIdx.x * a.bstride;
    a.coefA = bshift_nn(a.coefA, boff); a.coefB = bshift_nn(a.coefB, boff); a.du4 = bshift_nn(a.du4, boff);
  }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = a.w, h = a.h, T = a.T, K = a.K;
  const int rows = K * 32;
  const int s = (int)(threadIdx.x >> 5) / K, k = (int)(threadIdx.x >> 5) - s * K, l = threadIdx.x & 31, j = k * 32 + l;
  // shared memory (byte offsets).  One slot = one wavefront step of the layout = one column of every row:
  // rA [C (D + 2 T)][rows] float4 {inverse block a11 a12 a22, horiz} | rB likewise {b1, b2, vert, -} |
  // rD [C D][rows] float2 du, dv before this launch | rR [2][T][rows][C] float2 results of the previous step
  unsigned sb = smem_u32(smem_raw);
  asm volatile("" : "+r"(sb));  // keep it in a register (the compiler would re-derive it from %cluster_ctaid every step)
  const unsigned strideA = (unsigned)rows * 16, ringA = (unsigned)(C * (D + 2 * T)) * strideA;
  const unsigned strideD = (unsigned)rows * 8, ringD = (unsigned)(C * D) * strideD;
  const unsigned offB = ringA, offD = 2 * ringA, offR = offD + ringD, resHalf = (unsigned)(T * rows) * (C * 8);
  const Skew sk(w, h, C);
  const int nsteps = sk.nsteps;
  const size_t blk = ((size_t)k * sk.nsp) * 32 + l;
  const float omega = a.omega;
  {
    float2* const z = reinterpret_cast<float2*>(smem_raw + offD);  // du = dv = 0 (first iteration) and no results yet
    for (int q = threadIdx.x; q < C * D * rows + 2 * T * rows * C; q += blockDim.x) z[q] = make_float2(0.f, 0.f);
  }
  __syncthreads();
  // Fetch duty: per step the C wavefront steps u .. u + C - 1 of the layout (columns u - C j ... of row j) = steps
  // u - 32 C k of a row block, 512 contiguous bytes per array, row block and wavefront step.  Warp (s, k) fetches
  // array s of row block k: 0 coefA, 1 coefB, 2 the du records (not in the first iteration); with fewer than three
  // sweeps the block has fetch-only warps s = T ... 2.
  const int duty = (s > 2 || (s == 2 && first)) ? -1 : s;
  const unsigned char* fsrc = reinterpret_cast<const unsigned char*>((duty == 0 ? a.coefA : duty == 1 ? a.coefB : a.du4) + blk) -
                              (ptrdiff_t)(32 * C * k) * 512;
  const unsigned fstride = duty == 2 ? strideD : strideA, fring = duty == 2 ? ringD : ringA;
  const unsigned fdst = sb + (duty == 1 ? offB : duty == 2 ? offD : 0u) + j * (duty == 2 ? 8 : 16);
  int fsp = -32 * C * k;
  unsigned fo = 0;
  auto fetch = [&]() {
    if (duty == 2) {
#pragma unroll
      for (int q = 0; q < C; ++q)
        if ((unsigned)(fsp + q) < (unsigned)nsteps)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(fdst + fo + q * fstride), "l"(fsrc + q * 512) : "memory");
    } else if (duty >= 0) {
#pragma unroll
      for (int q = 0; q < C; ++q)
        if ((unsigned)(fsp + q) < (unsigned)nsteps)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(fdst + fo + q * fstride), "l"(fsrc + q * 512) : "memory");
    }
    cp_async_commit();
    fsrc += 512 * C;
    fsp += C;
    fo += C * fstride;
    fo = (fo == fring) ? 0u : fo;
  };
  for (int u = 0; u < D; ++u) fetch();
  cp_async_wait<D - 2>();  // steps 0 and 1 have landed
  __syncthreads();

  const bool no_up = (j == 0), no_dn = (j >= h - 1), row_ok = j < h, last = (s == T - 1);
  const int jp = max(j - 1, 0), jn = min(j + 1, rows - 1);
  // per-thread addresses; the ring slots of step t - 2 s advance by C strideA per step, the result buffers alternate
  const unsigned aSelf = sb + j * 16, aVt = sb + offB + jp * 16 + 8;
  const unsigned aUp = sb + offR + (unsigned)(s * rows + jp) * (C * 8), aOut = sb + offR + (unsigned)(s * rows + j) * (C * 8);
  const unsigned aOld = s == 0 ? sb + offD + j * 8 : sb + offR + (unsigned)((s - 1) * rows + j) * (C * 8);
  const unsigned aBel = s == 0 ? sb + offD + jn * 8 : sb + offR + (unsigned)((s - 1) * rows + jn) * (C * 8);
  const int RS = D + 2 * T;                                          // steps the coefficient ring holds
  unsigned o = (unsigned)(((-2 * s) % RS + RS) % RS) * C * strideA;      // slots of step t - 2 s: columns i0 ... i0 + C - 1
  unsigned om = (unsigned)(((-2 * s - 1) % RS + RS) % RS) * C * strideA;  // the step before: the row above at these columns
  unsigned od = (D > 1 ? 1u : 0u) * C * strideD;                     // sweep 0: slots of step t + 1 in the du ring
  unsigned pc = 0;                                                   // byte offset of this step's result buffer
  float2 res1 = make_float2(0.f, 0.f);  // my last result: (i0 - 1, j)
  float hl = 0.f;                       // horiz(i0 - 1, j)
  int i0 = -C * j - 2 * C * s;
  // the previous sweep at my columns i0 ... i0 + C - 1: what I read one step ago as the columns ahead ...
  float2 Cc[C];
#pragma unroll
  for (int q = 0; q < C; ++q) Cc[q] = make_float2(0.f, 0.f);
  if (s == 0) {  // ... except for the first pixels of row 0, active at step 0 (ring slots 0 ... C - 1)
#pragma unroll
    for (int q = 0; q < C; ++q) {
      Cc[q] = lds64(aOld + q * strideD);
      if (i0 + q >= w) Cc[q] = make_float2(0.f, 0.f);
    }
  }
  __syncthreads();  // the loop's first fetch refills the du ring slots of step 0 just read
  float4* gOut = a.du4 + blk + (ptrdiff_t)(i0 + C * l) * 32;  // record of pixel (i0, j): step i0 + C l of the row block
  const int nst = (w + C - 1) / C + h - 1 + 2 * (T - 1);
  // first and last step at which a lane of this warp is inside its row (one step early: the carried columns)
  const int t_lo = 32 * k + 2 * s - 1, t_hi = 32 * k + 31 + 2 * s + (w + C - 1) / C;
  for (int t = 0; t < nst; ++t) {
    fetch();
    const unsigned pp = resHalf - pc;  // the previous step's results
    unsigned on = o + C * strideA;
    on = (on == ringA) ? 0u : on;
    float2 nv[C];
    if (s < T && t >= t_lo && t < t_hi) {
      float4 cA[C], cB[C];
      float vt[C];
      float2 U[C], Bl[C], P[C];  // (i, j-1) of this sweep; (i, j+1) and (i + C, j) of the previous sweep
#pragma unroll
      for (int q = 0; q < C; ++q) {
        cA[q] = lds128(aSelf + o + q * strideA);
        cB[q] = lds128(aSelf + offB + o + q * strideA);
        vt[q] = lds32(aVt + om + q * strideA);  // vert(i, j-1)
      }
      if (C == 1) U[0] = lds64(aUp + pp);
#pragma unroll
      for (int q = 0; q + 1 < C; q += 2) {
        const float4 u = lds128(aUp + pp + q * 8);
        U[q] = make_float2(u.x, u.y);
        U[q + 1] = make_float2(u.z, u.w);
      }
      if (s == 0) {
#pragma unroll
        for (int q = 0; q < C; ++q) {
          P[q] = lds64(aOld + od + q * strideD);
          Bl[q] = lds64(aBel + od + q * strideD);
          if (i0 + C + q >= w) P[q] = make_float2(0.f, 0.f);  // the reference's zero past the last column
        }
      } else {
        if (C == 1) {
          P[0] = lds64(aOld + pp);
          Bl[0] = lds64(aBel + pp);
        }
#pragma unroll
        for (int q = 0; q + 1 < C; q += 2) {
          const float4 p = lds128(aOld + pp + q * 8), b = lds128(aBel + pp + q * 8);
          P[q] = make_float2(p.x, p.y);
          P[q + 1] = make_float2(p.z, p.w);
          Bl[q] = make_float2(b.x, b.y);
          Bl[q + 1] = make_float2(b.z, b.w);
        }
      }
      // ---- the reference's update (solver.c:122-131 / 180-190 / 237-247), both components.  First everything that
      // does not depend on the left neighbour, for all C pixels ...
      float s1[C], s2[C];
#pragma unroll
      for (int q = 0; q < C; ++q) {
        const float2 old1 = (q < C - 1) ? Cc[q + 1 < C ? q + 1 : 0] : P[0];
        float px = cA[q].w * old1.x, py = cA[q].w * old1.y;
        const float ux = px + vt[q] * U[q].x, uy = py + vt[q] * U[q].y;
        px = no_up ? px : ux;
        py = no_up ? py : uy;
        const float qx = px + cB[q].z * Bl[q].x, qy = py + cB[q].z * Bl[q].y;
        px = no_dn ? px : qx;
        py = no_dn ? py : qy;
        s1[q] = px + cB[q].x;
        s2[q] = py + cB[q].y;
      }
      // ... then the chain along the row
      float2 left = res1;
      float hq = hl;
#pragma unroll
      for (int q = 0; q < C; ++q) {
        const float l1 = hq * left.x + s1[q], l2 = hq * left.y + s2[q];
        const float B1 = (q == 0 && i0 == 0) ? s1[q] : l1, B2 = (q == 0 && i0 == 0) ? s2[q] : l2;
        float2 v;
        v.x = Cc[q].x + omega * (cA[q].x * B1 + cA[q].y * B2 - Cc[q].x);
        v.y = Cc[q].y + omega * (cA[q].y * B1 + cA[q].z * B2 - Cc[q].y);
        const bool act = row_ok && (unsigned)(i0 + q) < (unsigned)w;
        v.x = act ? v.x : 0.0f;
        v.y = act ? v.y : 0.0f;
        nv[q] = v;
        left = v;
        hq = cA[q].w;
        if (last && act) __stcg(gOut + q * 32, make_float4(v.x, v.y, 0.f, 0.f));
      }
      res1 = left;
      hl = hq;
#pragma unroll
      for (int q = 0; q < C; ++q) Cc[q] = P[q];
    } else {
#pragma unroll
      for (int q = 0; q < C; ++q) nv[q] = make_float2(0.f, 0.f);  // outside the rows: the zeros the next sweep expects
    }
    if (C == 1 && s < T) asm volatile("st.shared.v2.f32 [%0], {%1,%2};\n" ::"r"(aOut + pc), "f"(nv[0].x), "f"(nv[0].y) : "memory");
#pragma unroll
    for (int q = 0; q + 1 < C && s < T; q += 2)
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(aOut + pc + q * 8), "f"(nv[q].x), "f"(nv[q].y), "f"(nv[q + 1 < C ? q + 1 : q].x), "f"(nv[q + 1 < C ? q + 1 : q].y) : "memory");
    om = o;
    o = on;
    od += C * strideD;
    od = (od == ringD) ? 0u : od;
    pc = pp;
    gOut += 32 * C;
    i0 += C;
    cp_async_wait<D - 2>();  // steps <= t + 2 have landed (made visible to the other warps by the barrier)
    __syncthreads();
  }
  cp_async_wait<0>();
}

// final flow = wx + du (refine_variational.cpp:212-221)
__global__ void __launch_bounds__(256) k_update(int w, int h, int skew, float2* __restrict__ flow,
                                                const float4* __restrict__ du4, size_t bstride) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= w || j >= h) return;
  flow = bshift_nn(flow, (size_t)blockIdx.z * bstride);
  du4 = bshift_nn(du4, (size_t)blockIdx.z * bstride);
  const int o = j * w + i;
  const float2 f = flow[o];
  const float4 d = du4[Skew(w, h, skew).at(i, j)];
  flow[o] = make_float2(f.x + d.x, f.y + d.y);
}

}  // namespace

// per-device opt-in to > 48 KB dynamic shared memory (called from dis_create on the handle's device)
void varref_init_device() {
  cudaFuncSetAttribute(k_sor_wavefront<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sor_smem<16>());
  cudaFuncSetAttribute(k_sor_wavefront<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sor_smem<16>());
  cudaFuncSetAttribute(k_sor_wavefront<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sor_smem<8>());
  cudaFuncSetAttribute(k_sor_small<kSmallC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSorSmallMaxSmem);
}

void varref_sizes(int w, int h, int n_solver, size_t* n_coef4, size_t* n_du4, size_t* n_prog) {
  const Skew sk(w, h, kSmallC > SK ? kSmallC : SK);  // the larger of the two layouts
  *n_coef4 = (size_t)sk.K * sk.nsp * 32;
  *n_du4 = (size_t)(sk.K + 1) * sk.nsp * 32;
  *n_prog = (size_t)n_solver * sk.K + 2;
}

// Returns the number of kernels launched, or -1 for an unsupported level shape.
int launch_varref(const LevelGeom& g, const VarParams& v, const float* I0, const float* I1, float2* flow,
                  const VarRefBuffers& b, cudaStream_t st, Prof* prof) {
  const int w = g.w, h = g.h, n = w * h;
  if (w < 2 || h < 4 || v.n_solver < 1 || v.n_solver > 256) return -1;  // reference would take its slow path / read out of range
  int launches = 0;
  const int nb = g.nb;  // pairs per launch (batched handles): rides on grid.z (grid.y for the SOR)
  const size_t bs = g.bstride;
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, nb);
  // algorithmic bytes per SURVEY.md section 8(d): warp+mask 28 B/px, derivative stack 40 B/px
  {
    ProfScope ps(prof, "k_front", g.lv, 68.0 * n);
    FrontArgs fa{w, h, g.pad, g.pitch, g.noc, I0, I1, flow, b.stack, b.astride, bs};
    launch_pdl(k_front, dim3((w + FTX - 1) / FTX, (h + FTY - 1) / FTY, nb), dim3(32, 8), 0, st, fa);
  }
  launches += 1;

  if (v.n_inner <= 0) return launches;
  const int K = (h + 31) / 32, T = v.n_solver;
  // small level: one CTA per pair, one warp per (sweep, row block), kSmallC columns per step and a layout skewed by
  // as many columns per row (measured: 2 beats 1 and 4 on every level of a 1080p pair, profiles/README.md)
  const int small = (K <= v.sor_small && K * std::max(T, 3) <= 16 && sor_small_smem<kSmallC>(K * 32, T) <= kSorSmallMaxSmem) ? kSmallC : 0;
  const Skew sk(w, h, small ? small : SK);
  size_t n_coef4, n_du4, n_prog;
  varref_sizes(w, h, T, &n_coef4, &n_du4, &n_prog);
  // du = dv = 0 (image_erase, refine_variational.cpp:184-185), for every pair of the batch
  cudaMemset2DAsync(b.du4, nb > 1 ? bs : sizeof(float4) * n_du4, 0, sizeof(float4) * n_du4, nb, st);
  for (int it = 0; it < v.n_inner; ++it) {
    AssembleArgs aa{w, h, v.qa, v.hg, v.hd, it == 0 ? 1 : 0, g.noc, flow, b.du4,
                    b.stack, b.astride, b.coefA, b.coefB, b.progress, T * K, sk.sk, bs};
    {
      // smoothness 16 + data term 64 + sub_laplacian 32 + flow update 24 B/px
      ProfScope ps(prof, "k_assemble", g.lv, 136.0 * n);
      launch_pdl(k_assemble, grid, block, 0, st, aa);
    }
    SorArgs sa{w, h, T, K, v.omega, b.coefA, b.coefB, b.du4, b.progress, bs};
    if (small) {
      ProfScope ps(prof, "k_sor_small", g.lv, 44.0 * T * n);
      launch_pdl(k_sor_small<kSmallC>, dim3(nb), dim3(K * std::max(T, 3) * 32), sor_small_smem<kSmallC>(K * 32, T), st, sa,
                 it == 0 ? 1 : 0);
    } else {
      // each sweep reads 9 arrays and writes 2: 44 B/px
      ProfScope ps(prof, "k_sor_wavefront", g.lv, 44.0 * T * n);
      // CTAs = items busy at a time: total work (T K items of nsteps steps) over the critical path (section 4.4),
      // plus slack; never more than one per item
      const int lag = 80, path = sk.nsteps + (K - 1 + 2 * (T - 1)) * lag;
      const int ctas = v.sor_full ? T * K : std::min(T * K, (int)(((long long)T * K * sk.nsteps + path - 1) / path) + 3);
      if (v.sor_full)
        launch_pdl(k_sor_wavefront<16, false>, dim3(ctas, nb), dim3(32), sor_smem<16>(), st, sa);
      else if (v.sor_group == 16)
        launch_pdl(k_sor_wavefront<16, true>, dim3(ctas, nb), dim3(32), sor_smem<16>(), st, sa);
      else
        launch_pdl(k_sor_wavefront<8, true>, dim3(ctas, nb), dim3(32), sor_smem<8>(), st, sa);
    }
    launches += 2;
  }
  {
    ProfScope ps(prof, "k_update", g.lv, 8.0 * n);
    launch_pdl(k_update, grid, block, 0, st, w, h, sk.sk, flow, (const float4*)b.du4, bs);
  }
  return launches + 1;
}

}  // namespace dis
