// finish.cu -- final upsampling and cropping of the level-lv_l flow to the input resolution.
//
// Replaces kroeger/run_dense.cpp:407-414: `flowout *= 2^lv_l; cv::resize(x 2^lv_l, INTER_LINEAR);`
// then the crop that removes the divisibility padding.  OpenCV's bilinear resize: source
// coordinate (dst+0.5)/s - 0.5, floor, clamp to the edge, separable lerp horizontal then vertical.
#include "common.cuh"

namespace dis {
namespace {

// One thread = 4 consecutive output pixels of a row: the vertical weights and row pointers are shared, the four
// source taps of each pixel come from L1.  Same expressions, in the same order, as a per-pixel evaluation.
__global__ void __launch_bounds__(256) k_finish(const float2* __restrict__ fl, int wl, int hl, int lv_l,
                                                int left, int top, int w_org, int h_org,
                                                const Mailbox* __restrict__ mb) {
  float2* __restrict__ out = mb->out;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= w_org || y >= h_org) return;
  const int Y = y + top;
  float2 r[4];
  if (lv_l == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fl[(size_t)Y * wl + min(x0 + k, w_org - 1) + left];
  } else {
    const float s = (float)(1 << lv_l), inv = 1.0f / s;
    float fy = ((float)Y + 0.5f) * inv - 0.5f;
    int sy = (int)floorf(fy);
    fy -= (float)sy;
    if (sy < 0) { fy = 0.0f; sy = 0; }
    if (sy >= hl - 1) { fy = 0.0f; sy = hl - 1; }
    const float2* __restrict__ r0 = fl + (size_t)sy * wl;
    const float2* __restrict__ r1 = fl + (size_t)min(sy + 1, hl - 1) * wl;
    const float gy = 1.f - fy;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int X = min(x0 + k, w_org - 1) + left;
      float fx = ((float)X + 0.5f) * inv - 0.5f;
      int sx = (int)floorf(fx);
      fx -= (float)sx;
      if (sx < 0) { fx = 0.0f; sx = 0; }
      if (sx >= wl - 1) { fx = 0.0f; sx = wl - 1; }
      const int sx1 = min(sx + 1, wl - 1);
      const float2 p00 = __ldg(r0 + sx), p01 = __ldg(r0 + sx1);
      const float2 p10 = __ldg(r1 + sx), p11 = __ldg(r1 + sx1);
      const float gx = 1.f - fx;
      const float a0 = (p00.x * s) * gx + (p01.x * s) * fx;
      const float a1 = (p10.x * s) * gx + (p11.x * s) * fx;
      const float b0 = (p00.y * s) * gx + (p01.y * s) * fx;
      const float b1 = (p10.y * s) * gx + (p11.y * s) * fx;
      r[k].x = a0 * gy + a1 * fy;
      r[k].y = b0 * gy + b1 * fy;
    }
  }
  float2* o = out + (size_t)y * w_org + x0;
  if (x0 + 3 < w_org && ((size_t)o & 15) == 0) {
    reinterpret_cast<float4*>(o)[0] = make_float4(r[0].x, r[0].y, r[1].x, r[1].y);
    reinterpret_cast<float4*>(o)[1] = make_float4(r[2].x, r[2].y, r[3].x, r[3].y);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x0 + k < w_org) o[k] = r[k];
  }
}

}  // namespace

void launch_finish(const float2* flow_l, int wl, int hl, int lv_l, int left, int top, int w_org, int h_org,
                   const Mailbox* mb, cudaStream_t st) {
  dim3 block(32, 8), grid((w_org + 127) / 128, (h_org + 7) / 8);
  k_finish<<<grid, block, 0, st>>>(flow_l, wl, hl, lv_l, left, top, w_org, h_org, mb);
}

}  // namespace dis
