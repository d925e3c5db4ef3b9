// densify.cu -- stage 3: densification (weighted average of patch displacements per pixel).
//
// Replaces PatGridClass::AggregateFlowDense (kroeger/patchgrid.cpp:213-397).  The reference
// splats patch after patch into two accumulators; here every pixel *gathers* from the patches
// covering it, visiting them in ascending patch index ip = gx*noph + gy, which is exactly the
// order in which the sequential splat adds to that pixel -- so the float sums are identical
// and no atomics are needed.  Per covering patch: a = 1/max(2,|r|) (std::max(minerrval,w),
// NaN -> 2), we += a, flow += p*a; finally flow /= we where we > 0.
//
// Forward-backward merging (usefbcon, patchgrid.cpp:278-375): after the forward splat the reference
// walks the complementary grid's patches in index order and splats -flow bilinearly (weights of the
// patch's final sub-pixel position) at the *displaced* footprint.  In gather form a pixel visits, again
// in ascending patch index, every backward patch whose displaced footprint can reach it -- the search
// radius comes from the largest displacement of the level (k_bw_anchors) -- and adds the up to four
// terms cc, fc, cf, ff in the order the reference's y/x loop produces them.
#include <type_traits>

#include "common.cuh"

namespace dis {
namespace {

__device__ __forceinline__ float absw_of(float w) { return 1.0f / (2.0f < w ? w : 2.0f); }

// Pixel weight of patch element (lx,ly) (kroeger/patchgrid.cpp:253-260, 330-337).  Grey: 1/max(2,w).  RGB:
// 1/(max(2,w0)+max(2,w1)+max(2,w2)) -- read through the reference's own pointer walk, which advances by 3 for
// pixels that pass the bounds test and by 1 for skipped ones (so the channels of border patches are misaligned;
// reproduced as is): offset = (#pixels before) + 2*(#accepted pixels before), the accepted pixels forming the
// rectangle [xlo,xhi] x [ylo,yhi] in patch coordinates.
__device__ __forceinline__ float patch_absw(const float* pw, int P, int noc, int lx, int ly, int xlo, int xhi, int ylo) {
  if (noc == 1) return absw_of(__ldg(pw + ly * P + lx));
  const int off = (ly * P + lx) + 2 * ((ly - ylo) * (xhi - xlo + 1) + (lx - xlo));
  const float w0 = __ldg(pw + off), w1 = __ldg(pw + off + 1), w2 = __ldg(pw + off + 2);
  float a = (2.0f < w0 ? w0 : 2.0f);
  a += (2.0f < w1 ? w1 : 2.0f);
  a += (2.0f < w2 ? w2 : 2.0f);
  return 1.0f / a;
}

// Per-patch integer anchor ceil(pt_iter + 1e-5) (double arithmetic, patchgrid.cpp:304-305), bilinear
// weights (:310-315) and the level-wide maximum anchor displacement.
__global__ void __launch_bounds__(256) k_bw_anchors(const LevelGeom g, const OptParams o, const float2* __restrict__ pflow0,
                                                    int2* __restrict__ anchor0, float4* __restrict__ wbil0, int* __restrict__ maxdisp0) {
  pdl_wait();
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= g.nop) return;
  const size_t boff = (size_t)blockIdx.y * g.bstride;  // blockIdx.y = pair of a batched handle
  const float2* __restrict__ pflow = bshift(pflow0, boff);
  int2* __restrict__ anchor = bshift(anchor0, boff);
  float4* __restrict__ wbil = bshift(wbil0, boff);
  int* __restrict__ maxdisp = bshift(maxdisp0, boff);
  const int gx = ip / g.noph, gy = ip - gx * g.noph;
  const int cx = gx * o.steps + g.offw, cy = gy * o.steps + g.offh;
  const float2 p = pflow[ip];
  const float rx = (float)cx + p.x, ry = (float)cy + p.y;  // GetPointPos(): pt_ref + p_iter
  const int p0 = (int)ceil((double)rx + .00001), p1 = (int)ceil((double)ry + .00001);
  const int p2 = (int)floor((double)rx), p3 = (int)floor((double)ry);
  const float r0 = rx - (float)p2, r1 = ry - (float)p3;
  anchor[ip] = make_int2(p0, p1);
  wbil[ip] = make_float4(r0 * r1, (1 - r0) * r1, r0 * (1 - r1), (1 - r0) * (1 - r1));
  int dsp = max(abs(p0 - cx), abs(p1 - cy));
  dsp = min(max(dsp, 0), 1 << 14);  // NaN / overflow guard
  atomicMax(maxdisp, dsp);          // max is order independent: deterministic
}

template <bool FB>
__global__ void __launch_bounds__(256, 2) k_densify(const DensifyArgs a) {
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= a.g.w || y >= a.g.h) return;
  const size_t boff = (size_t)blockIdx.z * a.g.bstride;  // blockIdx.z = pair of a batched handle
  const float2* __restrict__ q_pflow = bshift_nn(a.pflow, boff);  // never null (the backward-grid pointers may be)
  const float* __restrict__ q_pweight = bshift_nn(a.pweight, boff);
  const float2* __restrict__ q_pflow_bw = bshift(a.pflow_bw, boff);
  const float* __restrict__ q_pweight_bw = bshift(a.pweight_bw, boff);
  const int2* __restrict__ q_anchor = bshift(a.anchor, boff);
  const float4* __restrict__ q_wbil = bshift(a.wbil, boff);
  const int* __restrict__ q_maxdisp = bshift(a.maxdisp, boff);
  float2* __restrict__ q_flow = bshift_nn(a.flow, boff);
  const int P = a.o.p, N = a.o.novals, steps = a.o.steps, half = P / 2;
  // patches whose footprint [c-half, c+half-1] contains the pixel
  // c = g*steps + off  =>  g in [ceil((x-off-half+1)/steps), floor((x-off+half)/steps)]
  const int ax = x - a.g.offw, ay = y - a.g.offh;
  int gx0 = ax - half + 1, gx1 = ax + half, gy0 = ay - half + 1, gy1 = ay + half;
  gx0 = gx0 <= 0 ? 0 : (gx0 + steps - 1) / steps;
  gy0 = gy0 <= 0 ? 0 : (gy0 + steps - 1) / steps;
  gx1 = gx1 < 0 ? -1 : min(gx1 / steps, a.g.nopw - 1);
  gy1 = gy1 < 0 ? -1 : min(gy1 / steps, a.g.noph - 1);
  float we = 0.0f, fu = 0.0f, fv = 0.0f;
  // common cases (p/steps <= 2 or <= 4): issue all loads first, then accumulate in patch-index order
  auto gather_fixed = [&](auto cov_c) {
    constexpr int C = decltype(cov_c)::value;
    float wv[C * C];
    float2 fl[C * C];
    const int nx = gx1 - gx0 + 1, ny = gy1 - gy0 + 1;
    const int lx0 = x - (gx0 * steps + a.g.offw) + half, ly0 = y - (gy0 * steps + a.g.offh) + half;
#pragma unroll
    for (int u = 0; u < C; ++u)
#pragma unroll
      for (int v = 0; v < C; ++v) {
        const bool ok = u < nx && v < ny;
        const int ip = ok ? (gx0 + u) * a.g.noph + gy0 + v : 0;
        const int off = ok ? (ly0 - v * steps) * P + (lx0 - u * steps) : 0;
        wv[u * C + v] = __ldg(q_pweight + (size_t)ip * N + off);
        fl[u * C + v] = __ldg(q_pflow + ip);
      }
#pragma unroll
    for (int u = 0; u < C; ++u)
#pragma unroll
      for (int v = 0; v < C; ++v)
        if (u < nx && v < ny) {
          const float aw = absw_of(wv[u * C + v]);
          we += aw;
          fu += fl[u * C + v].x * aw;
          fv += fl[u * C + v].y * aw;
        }
  };
  const int noc = a.o.noc;
  if (noc == 1 && a.cover <= 2) {
    gather_fixed(std::integral_constant<int, 2>{});
  } else if (noc == 1 && a.cover <= 4) {
    gather_fixed(std::integral_constant<int, 4>{});
  } else {
    for (int gx = gx0; gx <= gx1; ++gx) {
      const int pcx = gx * steps + a.g.offw;
      const int lx = x - pcx + half;
      const int xlo = max(0, half - pcx), xhi = min(P - 1, a.g.w - 1 - pcx + half);
      for (int gy = gy0; gy <= gy1; ++gy) {
        const int ip = gx * a.g.noph + gy;
        const int pcy = gy * steps + a.g.offh;
        const int ly = y - pcy + half;
        const float2 f = __ldg(q_pflow + ip);
        const float aw = patch_absw(q_pweight + (size_t)ip * N, P, noc, lx, ly, xlo, xhi, max(0, half - pcy));
        we += aw;
        fu += f.x * aw;
        fv += f.y * aw;
      }
    }
  }
  if (FB) {
    const int lb = -half, ub = half - 1, w = a.g.w, h = a.g.h;
    const int D = *q_maxdisp;
    // a patch can reach x iff its anchor p0 is in [x-ub, x-lb+1]; |p0 - centre| <= D
    int bx0 = x - ub - D - a.g.offw, bx1 = x - lb + 1 + D - a.g.offw;
    int by0 = y - ub - D - a.g.offh, by1 = y - lb + 1 + D - a.g.offh;
    bx0 = bx0 <= 0 ? 0 : (bx0 + steps - 1) / steps;
    by0 = by0 <= 0 ? 0 : (by0 + steps - 1) / steps;
    bx1 = bx1 < 0 ? -1 : min(bx1 / steps, a.g.nopw - 1);
    by1 = by1 < 0 ? -1 : min(by1 / steps, a.g.noph - 1);
    for (int gx = bx0; gx <= bx1; ++gx)
      for (int gy = by0; gy <= by1; ++gy) {
        const int ip = gx * a.g.noph + gy;
        const int2 an = __ldg(q_anchor + ip);
        const int dx = x - an.x, dy = y - an.y;
        if (dx < lb - 1 || dx > ub || dy < lb - 1 || dy > ub) continue;
        const float2 f = __ldg(q_pflow_bw + ip);
        const float4 wb = __ldg(q_wbil + ip);
        const float* pw = q_pweight_bw + (size_t)ip * N;
        // source element (ex,ey) of the patch sits at target (XT,YT) and feeds this pixel with weight wk
        // accepted source pixels of this patch (xt,yt in [1,w-2] x [1,h-2]) as a rectangle in patch coordinates
        const int bxlo = max(0, 1 - an.x - lb), bxhi = min(P - 1, w - 2 - an.x - lb), bylo = max(0, 1 - an.y - lb);
        auto contrib = [&](int ex, int ey, int XT, int YT, float wk) {
          if (ex < lb || ex > ub || ey < lb || ey > ub) return;
          if (!(XT >= 1 && YT >= 1 && XT < (w - 1) && YT < (h - 1))) return;
          const float aw = patch_absw(pw, P, a.o.noc, ex - lb, ey - lb, bxlo, bxhi, bylo);
          const float f0 = f.x * aw, f1 = f.y * aw;
          we += wk * aw;
          fu -= wk * f0;
          fv -= wk * f1;
        };
        contrib(dx, dy, x, y, wb.x);              // cc
        contrib(dx + 1, dy, x + 1, y, wb.y);      // fc
        contrib(dx, dy + 1, x, y + 1, wb.z);      // cf
        contrib(dx + 1, dy + 1, x + 1, y + 1, wb.w);  // ff
      }
  }
  float2 out = make_float2(fu, fv);
  if (we > 0.0f) {
    out.x = fu / we;
    out.y = fv / we;
  }
  q_flow[(size_t)y * a.g.w + x] = out;
}

}  // namespace

void launch_densify(const DensifyArgs& a, cudaStream_t st) {
  dim3 block(32, 8);
  dim3 grid((a.g.w + 31) / 32, (a.g.h + 7) / 8, a.g.nb);
  if (a.pflow_bw != nullptr) {
    cudaMemset2DAsync(a.maxdisp, a.g.bstride ? a.g.bstride : sizeof(int), 0, sizeof(int), a.g.nb, st);  // one counter per pair
    launch_pdl(k_bw_anchors, dim3((a.g.nop + 255) / 256, a.g.nb), dim3(256), 0, st, a.g, a.o, a.pflow_bw, a.anchor, a.wbil, a.maxdisp);
    launch_pdl(k_densify<true>, grid, block, 0, st, a);
  } else {
    launch_pdl(k_densify<false>, grid, block, 0, st, a);
  }
}

}  // namespace dis
