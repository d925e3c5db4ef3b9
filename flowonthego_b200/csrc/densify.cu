// densify.cu -- stage 3: densification (weighted average of patch displacements per pixel).
//
// Replaces PatGridClass::AggregateFlowDense (kroeger/patchgrid.cpp:213-397).  The reference
// splats patch after patch into two accumulators; here every pixel *gathers* from the patches
// covering it, visiting them in ascending patch index ip = gx*noph + gy, which is exactly the
// order in which the sequential splat adds to that pixel -- so the float sums are identical
// and no atomics are needed.  Per covering patch: a = 1/max(2,|r|) (std::max(minerrval,w),
// NaN -> 2), we += a, flow += p*a; finally flow /= we where we > 0.
#include "common.cuh"

namespace dis {
namespace {

__device__ __forceinline__ float absw_of(float w) { return 1.0f / (2.0f < w ? w : 2.0f); }

__global__ void __launch_bounds__(256) k_densify(const DensifyArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= a.g.w || y >= a.g.h) return;
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
  for (int gx = gx0; gx <= gx1; ++gx) {
    const int lx = x - (gx * steps + a.g.offw) + half;
    for (int gy = gy0; gy <= gy1; ++gy) {
      const int ip = gx * a.g.noph + gy;
      const int ly = y - (gy * steps + a.g.offh) + half;
      const float w = __ldg(a.pweight + (size_t)ip * N + ly * P + lx);
      const float2 f = __ldg(a.pflow + ip);
      const float aw = absw_of(w);
      we += aw;
      fu += f.x * aw;
      fv += f.y * aw;
    }
  }
  float2 out = make_float2(fu, fv);
  if (we > 0.0f) {
    out.x = fu / we;
    out.y = fv / we;
  }
  a.flow[(size_t)y * a.g.w + x] = out;
}

}  // namespace

void launch_densify(const DensifyArgs& a, cudaStream_t st) {
  dim3 block(32, 8);
  dim3 grid((a.g.w + 31) / 32, (a.g.h + 7) / 8);
  k_densify<<<grid, block, 0, st>>>(a);
}

}  // namespace dis
