// finish.cu -- final upsampling and cropping of the level-lv_l flow to the input resolution.
//
// Replaces kroeger/run_dense.cpp:407-414: `flowout *= 2^lv_l; cv::resize(x 2^lv_l, INTER_LINEAR);`
// then the crop that removes the divisibility padding.  OpenCV's bilinear resize: source
// coordinate (dst+0.5)/s - 0.5, floor, clamp to the edge, separable lerp horizontal then vertical.
#include "common.cuh"

namespace dis {
namespace {

__global__ void __launch_bounds__(256) k_finish(const float2* __restrict__ fl, int wl, int hl, int lv_l,
                                                int left, int top, int w_org, int h_org,
                                                const Mailbox* __restrict__ mb) {
  float2* __restrict__ out = mb->out;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w_org || y >= h_org) return;
  const int X = x + left, Y = y + top;
  float2 r;
  if (lv_l == 0) {
    r = fl[(size_t)Y * wl + X];
  } else {
    const float s = (float)(1 << lv_l), inv = 1.0f / s;
    float fx = ((float)X + 0.5f) * inv - 0.5f, fy = ((float)Y + 0.5f) * inv - 0.5f;
    int sx = (int)floorf(fx), sy = (int)floorf(fy);
    fx -= (float)sx;
    fy -= (float)sy;
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx >= wl - 1) { fx = 0.0f; sx = wl - 1; }
    if (sy < 0) { fy = 0.0f; sy = 0; }
    if (sy >= hl - 1) { fy = 0.0f; sy = hl - 1; }
    const int sx1 = min(sx + 1, wl - 1), sy1 = min(sy + 1, hl - 1);
    const float2 p00 = fl[(size_t)sy * wl + sx], p01 = fl[(size_t)sy * wl + sx1];
    const float2 p10 = fl[(size_t)sy1 * wl + sx], p11 = fl[(size_t)sy1 * wl + sx1];
    const float a0 = (p00.x * s) * (1.f - fx) + (p01.x * s) * fx;
    const float a1 = (p10.x * s) * (1.f - fx) + (p11.x * s) * fx;
    const float b0 = (p00.y * s) * (1.f - fx) + (p01.y * s) * fx;
    const float b1 = (p10.y * s) * (1.f - fx) + (p11.y * s) * fx;
    r.x = a0 * (1.f - fy) + a1 * fy;
    r.y = b0 * (1.f - fy) + b1 * fy;
  }
  out[(size_t)y * w_org + x] = r;
}

}  // namespace

void launch_finish(const float2* flow_l, int wl, int hl, int lv_l, int left, int top, int w_org, int h_org,
                   const Mailbox* mb, cudaStream_t st) {
  dim3 block(32, 8), grid((w_org + 31) / 32, (h_org + 7) / 8);
  k_finish<<<grid, block, 0, st>>>(flow_l, wl, hl, lv_l, left, top, w_org, h_org, mb);
}

}  // namespace dis
