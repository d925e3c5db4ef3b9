// patch_search.cu -- stage 2: per-patch inverse-compositional Gauss-Newton search.
//
// Replaces, for one pyramid level, PatGridClass::InitializeGrid / SetTargetImage /
// InitializeFromCoarserOF / Optimize (kroeger/patchgrid.cpp:98-141, 195-211) and everything in
// PatClass underneath (kroeger/patch.cpp): template + gradient extraction (:287-332), Hessian
// (:71-88), OptimizeStart/OptimizeIter (:120-212), bilinear patch sampling (:335-402), the
// L2/L1/pseudo-Huber error image (:223-262) and the termination test (:264-284).
//
// Mapping: 8 lanes ("octet") per patch, 4 patches per warp.  Lane c of the octet owns exactly
// the elements e = 8i + c of the row-major p x p patch -- i.e. one of the 8 accumulator chains
// of Eigen 3.3's SSE reduction (two 4-wide packets, Eigen/src/Core/Redux.h) -- so the per-patch
// sums are formed in the reference's float order: each lane adds its chain sequentially in
// registers, then the 8 partial sums are combined with three xor-shuffles as
// (p0+p1) [+ tail packet] -> (r0+r2)+(r1+r3).  Template, both gradient patches and the current
// residual live in registers; the only memory traffic in the iteration loop is the 4 bilinear
// taps per element of the target image.  No FMA (-fmad=false), IEEE div/sqrt.
#include "common.cuh"

// This file is compiled twice (csrc/Makefile): with -fmad=false for the exact engine (the arithmetic contract of
// common.cuh) and, as patch_search_fast.o with -fmad=true -DDIS_ARITH_FAST, for the opt-in tolerance mode
// DIS_OPT_ARITH = 1: same code, same reduction order, but ptxas may contract a*b+c into FMA.
#ifdef DIS_ARITH_FAST
#define launch_patch_search launch_patch_search_fast
#define patch_search_init_device patch_search_init_device_fast
#endif

namespace dis {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// Combine the 8 chain sums of an octet in Eigen order.  `extra` is the element of the trailing
// 4-wide packet (only when n % 8 == 4; held by lanes c < 4).
template <bool HAS_EXTRA>
__device__ __forceinline__ float octet_reduce(float chain, float extra) {
  float r = chain + __shfl_xor_sync(FULL, chain, 4);  // p0[k] + p1[k]
  if (HAS_EXTRA) r = r + extra;                       // p0 += packet(alignedEnd2)
  const float t = r + __shfl_xor_sync(FULL, r, 2);    // r0+r2 | r1+r3
  // every lane ends with the same bits: lanes c and c+4 hold equal r (a+b == b+a), the tail element is
  // duplicated on lanes 4..7, and (r0+r2)+(r1+r3) is symmetric under the remaining swaps
  return t + __shfl_xor_sync(FULL, t, 1);             // (r0+r2)+(r1+r3)
}

__host__ __device__ constexpr int gcd_ce(int a, int b) { return b == 0 ? a : gcd_ce(b, a % b); }

// Shared-memory window of the target image per patch: every position the search may sample lies
// within outlierthresh = p/2 of the start position (else the patch is reset to its start), so a
// (2p+4)^2 window anchored at floor(start) - p - 1 covers all bilinear taps of all iterations.
#ifndef DIS_PS_THREADS
#define DIS_PS_THREADS 128
#endif
constexpr int kPsThreads = DIS_PS_THREADS;  // threads per CTA: 8 per patch

template <int P>
struct Win {
  // width: covers 2p+3 columns, and W - P is a multiple of 8 so that the two row segments an octet
  // may touch in one load never share a bank (see the interleaving below)
  static constexpr int W = P + 8 * ((P + 3 + 7) / 8);
  static constexpr int SIZE = W * W;
};

//
// NC = 3 (the reference's SELECTCHANNEL=3 build): a patch row is 3P interleaved floats and the same
// chain rule applies to the 3P^2 values (patch.cpp:392-396); the target taps are then read straight
// from global memory (L1/L2) instead of a staged window, and for p > 8 the template and gradients are
// re-read per iteration instead of being held in registers.
template <int P, bool L2, int NC>
__global__ void __launch_bounds__(kPsThreads, (P == 12 && NC == 1) ? 512 / kPsThreads : 1) k_patch_search(const PatchSearchArgs a) {
  pdl_wait();
  // batched handles: blockIdx.y = pair, all buffers of that pair sit a.g.bstride bytes further (common.cuh)
  const size_t boff = (size_t)blockIdx.y * a.g.bstride;
  const float* __restrict__ pI0 = bshift_nn(a.I0, boff);
  const float* __restrict__ pI0x = bshift_nn(a.I0x, boff);
  const float* __restrict__ pI0y = bshift_nn(a.I0y, boff);
  const float* __restrict__ pI1 = bshift_nn(a.I1, boff);
  const float2* __restrict__ pcoarse = bshift(a.flow_coarse, boff);
  float2* __restrict__ ppflow = bshift_nn(a.pflow, boff);
  float* __restrict__ ppweight = bshift_nn(a.pweight, boff);
  constexpr int N = P * P * NC;
  constexpr int PW = P * NC;            // floats per patch row
  constexpr int NI = N / 8;             // chain length
  constexpr bool EX = (N % 8) == 4;     // trailing packet present
  constexpr int NE = NI + (EX ? 1 : 0);
  constexpr int LB = -P / 2;
  constexpr int M = PW / gcd_ce(8, PW); // period of the (row, col) pattern of elements 8i + c
  constexpr int ROWS = 8 * M / PW;      // rows advanced per period
  constexpr int WIN = Win<P>::W;
  constexpr bool SMEM = NC == 1;        // target window staged in shared memory
  constexpr bool REG = NE <= 32;        // template + gradients held in registers
  extern __shared__ float smem_win[];

  const int c = threadIdx.x & 7;
  const int oct = threadIdx.x >> 3;
  int ip = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool live = ip < a.g.nop;
  if (!live) ip = a.g.nop - 1;  // dead octets shadow the last patch (shuffles stay warp-uniform)

  const int pitch = a.g.pitch, pad = a.g.pad;
  const int gx = ip / a.g.noph, gy = ip - gx * a.g.noph;
  const int cx = gx * a.o.steps + a.g.offw, cy = gy * a.o.steps + a.g.offh;

  // element (8i + c) of the row-major patch = (row, col); the pattern repeats every M chain steps
  int prow[M], pcol[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int e = 8 * m + c;
    prow[m] = e / PW;
    pcol[m] = e % PW;
  }
  const int te = 8 * NI + (c & 3);  // tail element (EX only)
  const int trow = te / PW, tcol = te % PW;

  // ---- InitializePatch: template and gradients at the integer patch centre (patch.cpp:287-332)
  constexpr int NR = REG ? NE : 1;
  float T[NR], GX[NR], GY[NR];
  const size_t base = (size_t)(cy + pad + LB) * pitch + (cx + pad + LB) * NC;
  // offset of chain element i from the patch origin (padded image coordinates)
  auto eoff = [&](int i) -> int {
    return (i < NI) ? (prow[i % M] + (i / M) * ROWS) * pitch + pcol[i % M] : trow * pitch + tcol;
  };
  float tmean = 0.0f;  // x - 0.0f == x exactly: no branch needed where it is subtracted
  if (REG) {
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const size_t o = base + eoff(i);
      T[i] = __ldg(pI0 + o);
      GX[i] = __ldg(pI0x + o);
      GY[i] = __ldg(pI0y + o);
    }
  }
  if (a.o.patnorm > 0) {
    float ch = 0.0f, ex = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const float t = REG ? T[i] : __ldg(pI0 + base + eoff(i));
      if (i == 0) ch = t; else if (i < NI) ch = ch + t; else ex = t;
    }
    tmean = octet_reduce<EX>(ch, ex) / (float)N;
    if (REG) {
#pragma unroll
      for (int i = 0; i < NE; ++i) T[i] = T[i] - tmean;
    }
  }
  // ---- ComputeHessian (patch.cpp:71-88)
  float H00, H01, H11;
  {
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, e0 = 0.0f, e1 = 0.0f, e2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const float gx_ = REG ? GX[i] : __ldg(pI0x + base + eoff(i));
      const float gy_ = REG ? GY[i] : __ldg(pI0y + base + eoff(i));
      if (i == 0) {
        c0 = gx_ * gx_; c1 = gx_ * gy_; c2 = gy_ * gy_;
      } else if (i < NI) {
        c0 = c0 + gx_ * gx_; c1 = c1 + gx_ * gy_; c2 = c2 + gy_ * gy_;
      } else {
        e0 = gx_ * gx_; e1 = gx_ * gy_; e2 = gy_ * gy_;
      }
    }
    H00 = octet_reduce<EX>(c0, e0);
    H01 = octet_reduce<EX>(c1, e1);
    H11 = octet_reduce<EX>(c2, e2);
    if (H00 * H11 - H01 * H01 == 0.0f) {
      H00 = (float)((double)H00 + 1e-10);
      H11 = (float)((double)H11 + 1e-10);
    }
  }
  // LLT factor of H (Eigen LLT unblocked, early exit on a non-positive pivot); H is constant per
  // patch so the factor is hoisted out of the iteration loop (same values every iteration).
  float L00 = H00, L10 = H01, L11 = H11;
  {
    float x = L00;
    if (!(x <= 0.0f)) {
      L00 = x = sqrtf(x);
      L10 = L10 / x;
      x = L11 - L10 * L10;
      if (!(x <= 0.0f)) L11 = sqrtf(x);
    }
  }

  // ---- InitializeFromCoarserOF (patchgrid.cpp:195-211)
  float pinx = 0.0f, piny = 0.0f;
  if (pcoarse != nullptr) {
    const int x = (int)floorf((float)cx / 2), y = (int)floorf((float)cy / 2);
    const float2 f = __ldg(pcoarse + (size_t)y * (a.g.w / 2) + x);
    pinx = f.x * 2;
    piny = f.y * 2;
  }

  // ---- OptimizeStart (patch.cpp:120-156)
  float px = pinx, py = piny;
  float ptx = (float)cx + px, pty = (float)cy + py;
  const float stx = ptx, sty = pty;
  float dpx = 0.0f, dpy = 0.0f;
  float dp_sq_init = 1e-10f, mares = 1e5f;
  int cnt = 0;
  const bool oob_start = ptx < a.g.lb || pty < a.g.lb || ptx > a.g.ubw || pty > a.g.ubh;
  bool conv = oob_start;
  if (oob_start) {  // sample somewhere legal; the result is discarded
    ptx = (float)cx;
    pty = (float)cy;
  }

  // ---- stage the window of the target image (padded coordinates, clamped to the padded array)
  // The four windows of a warp are interleaved word by word (word w of octet o at 4w + o): the 8 lanes
  // of an octet read consecutive words -> banks o, o+4, .., o+28, and the four octets occupy the four
  // residue classes mod 4, so a warp-wide tap load is conflict-free whatever the patches' positions.
  float* win = smem_win + (oct >> 2) * (4 * Win<P>::SIZE) + (oct & 3);
  const int wx0 = (int)floorf(ptx) - P - 1 + pad, wy0 = (int)floorf(pty) - P - 1 + pad;  // window origin
  if (SMEM) {
    // lane c stages columns c, c+8, c+16, ... of every window row (column clamps hoisted out of the row loop)
    constexpr int NCOL = (WIN + 7) / 8;
    const int tw1 = a.g.tw - 1, th1 = a.g.th - 1;
    int xs[NCOL];
#pragma unroll
    for (int u = 0; u < NCOL; ++u) xs[u] = min(max(wx0 + c + 8 * u, 0), tw1);
    float* wrow = win + c * 4;
#pragma unroll 4
    for (int wy = 0; wy < WIN; ++wy) {
      const float* src = pI1 + (size_t)min(max(wy0 + wy, 0), th1) * pitch;
#pragma unroll
      for (int u = 0; u < NCOL; ++u)
        if (c + 8 * u < WIN) wrow[32 * u] = __ldg(src + xs[u]);
      wrow += WIN * 4;
    }
  }
  __syncwarp();

  float R[NE];   // |residual| per element (pweight)
  float bx = 0.0f, by = 0.0f;
  bool first = true;
  while (true) {
    if (!first) {
      if (!__any_sync(FULL, !conv)) break;
      if (!conv) {
        // ---- OptimizeIter body (patch.cpp:172-208)
        cnt++;
        // delta_p = Hes.llt().solve(delta_p)
        float y0 = bx / L00;
        float y1 = by - L10 * y0;
        y1 = y1 / L11;
        const float x1 = y1 / L11;
        float x0 = y0 - L10 * x1;
        x0 = x0 / L00;
        dpx = x0;
        dpy = x1;
        px = px - dpx;
        py = py - dpy;
        ptx = (float)cx + px;
        pty = (float)cy + py;
        const float ox = stx - ptx, oy = sty - pty;
        if (ox * ox + oy * oy > a.o.outlier_sq || ptx < a.g.lb || pty < a.g.lb ||  // sqrtf(.) > outlierthresh
            ptx > a.g.ubw || pty > a.g.ubh) {
          px = pinx;
          py = piny;
          ptx = (float)cx + px;
          pty = (float)cy + py;
          conv = true;  // error image is still recomputed below (patch.cpp:210)
        }
      }
    }
    // ---- OptimizeComputeErrImg (patch.cpp:264-284); converged octets recompute identical values
    {
      // getPatchStaticBil (patch.cpp:335-402)
      const int posx = (int)ceilf(ptx + .00001f), posy = (int)ceilf(pty + .00001f);
      const float rx = ptx - (float)(int)floorf(ptx), ry = pty - (float)(int)floorf(pty);
      const float w0 = rx * ry, w1 = (1 - rx) * ry, w2 = rx * (1 - ry), w3 = (1 - rx) * (1 - ry);
      // window-relative position of tap `a` of element (0,0); clamped so that no read can leave the window
      const int ax = min(max(posx + pad + LB - wx0, 1), WIN - P), ay = min(max(posy + pad + LB - wy0, 1), WIN - P);
      const float* wb = win + (ay * WIN + ax) * 4;
      const float* bases[M];
#pragma unroll
      for (int m = 0; m < M; ++m) bases[m] = wb + (prow[m] * WIN + pcol[m]) * 4;
      const float* tbase = wb + (trow * WIN + tcol) * 4;
      // NC > 1: taps straight from the padded target image (the position is inside [lb, ub], so the
      // p x p footprint plus its left/upper neighbours lies inside the padding)
      const float* gb = pI1 + (size_t)(posy + pad + LB) * pitch + (posx + pad + LB) * NC;
      float ch = 0.0f;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        float va, vb, vc, vd;
        if (SMEM) {
          const float* q = (i < NI) ? bases[i % M] + (i / M) * ROWS * WIN * 4 : tbase;
          va = q[0], vb = q[-4], vc = q[-4 * WIN], vd = q[-4 * WIN - 4];
        } else {
          const float* q = gb + eoff(i);
          va = __ldg(q), vb = __ldg(q - NC), vc = __ldg(q - pitch), vd = __ldg(q - pitch - NC);
        }
        R[i] = w0 * va + w1 * vb + w2 * vc + w3 * vd;
        if (i == 0)
          ch = R[0];
        else if (i < NI)
          ch = ch + R[i];
      }
      float m = 0.0f;  // x - 0.0f == x exactly, so the no-patnorm case needs no branch below
      if (a.o.patnorm > 0) m = octet_reduce<EX>(ch, EX ? R[NE - 1] : 0.0f) / (float)N;
      // LossComputeErrorImage (patch.cpp:223-262) fused with the projections of the next
      // iteration (patch.cpp:178-179) and the L1 norm of the weights (:278)
      float cgx = 0.0f, cgy = 0.0f, cab = 0.0f;
      float egx = 0.0f, egy = 0.0f, eab = 0.0f;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        float d = R[i] - m;
        d = d - (REG ? T[i] : __ldg(pI0 + base + eoff(i)) - tmean);
        if (!L2) {
          if (a.o.costfct == 1)
            d = copysignf(sqrtf(fabsf(d)), d);
          else
            d = copysignf(sqrtf((sqrtf(1.0f + (d * d) / 25.0f) - 1.0f) * 50.0f), d);
        }
        const float ad = fabsf(d);
        R[i] = d;  // |d| is taken when the weights are written out
        const float tx = (REG ? GX[i] : __ldg(pI0x + base + eoff(i))) * d;
        const float ty = (REG ? GY[i] : __ldg(pI0y + base + eoff(i))) * d;
        if (i == 0) {
          cgx = tx; cgy = ty; cab = ad;
        } else if (i < NI) {
          cgx = cgx + tx; cgy = cgy + ty; cab = cab + ad;
        } else {
          egx = tx; egy = ty; eab = ad;
        }
      }
      const float nbx = octet_reduce<EX>(cgx, egx);
      const float nby = octet_reduce<EX>(cgy, egy);
      const float asum = octet_reduce<EX>(cab, eab);
      if (!conv || (first && !oob_start)) {
        bx = nbx;
        by = nby;
      }
      if (first ? !oob_start : true) {
        // state update + termination test; harmless for octets that are already converged
        // because nothing below is read again once conv is set.
        const float dp_sq = dpx * dpx + dpy * dpy;
        if (cnt == 1) dp_sq_init = dp_sq;
        if (!conv) {
          const float mares_old = mares;
          mares = asum / (float)N;
          bool go = (cnt < a.o.max_iter) & (mares > a.o.res_thresh);
          if (cnt >= a.o.min_iter)  // the two rate tests (and their divisions) only once min_iter is reached
            go = go & (dp_sq / dp_sq_init >= a.o.dp_thresh) & (mares / mares_old <= a.o.dr_thresh);
          if (!go) conv = true;
        }
      }
    }
    first = false;
  }

  // ---- results: p_iter and the weight patch (read by AggregateFlowDense)
  if (live) {
    if (c == 0) ppflow[ip] = make_float2(px, py);
    float* pw = ppweight + (size_t)ip * N;
#pragma unroll
    for (int i = 0; i < NI; ++i) pw[8 * i + c] = oob_start ? 0.0f : fabsf(R[i]);
    if (EX && c < 4) pw[8 * NI + c] = oob_start ? 0.0f : fabsf(R[NE - 1]);
  }
}

template <int P>
int launch_p(const PatchSearchArgs& a, cudaStream_t st) {
  const int threads = kPsThreads;
  const int blocks = (a.g.nop * 8 + threads - 1) / threads;
  const size_t smem = (size_t)(threads / 8) * Win<P>::SIZE * sizeof(float);
  if (a.o.noc == 3) {
    if (a.o.costfct == 0)
      launch_pdl(k_patch_search<P, true, 3>, dim3(blocks, a.g.nb), dim3(threads), 0, st, a);
    else
      launch_pdl(k_patch_search<P, false, 3>, dim3(blocks, a.g.nb), dim3(threads), 0, st, a);
  } else if (a.o.costfct == 0)
    launch_pdl(k_patch_search<P, true, 1>, dim3(blocks, a.g.nb), dim3(threads), smem, st, a);
  else
    launch_pdl(k_patch_search<P, false, 1>, dim3(blocks, a.g.nb), dim3(threads), smem, st, a);
  return 0;
}

template <int P>
void init_p() {
  const int smem = (int)((kPsThreads / 8) * Win<P>::SIZE * sizeof(float));
  cudaFuncSetAttribute(k_patch_search<P, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_patch_search<P, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

}  // namespace

// per-device opt-in to > 48 KB dynamic shared memory (called from dis_create on the handle's device)
void patch_search_init_device() {
  init_p<4>();
  init_p<6>();
  init_p<8>();
  init_p<10>();
  init_p<12>();
  init_p<14>();
  init_p<16>();
}

int launch_patch_search(const PatchSearchArgs& a, cudaStream_t st) {
  switch (a.o.p) {
    case 4: return launch_p<4>(a, st);
    case 6: return launch_p<6>(a, st);
    case 8: return launch_p<8>(a, st);
    case 10: return launch_p<10>(a, st);
    case 12: return launch_p<12>(a, st);
    case 14: return launch_p<14>(a, st);
    case 16: return launch_p<16>(a, st);
    default: return 1;
  }
}

}  // namespace dis
