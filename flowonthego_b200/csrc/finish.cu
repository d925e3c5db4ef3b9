// finish.cu -- final upsampling and cropping of the level-lv_l flow to the input resolution.
//
// Replaces kroeger/run_dense.cpp:407-414: `flowout *= 2^lv_l; cv::resize(x 2^lv_l, INTER_LINEAR);`
// then the crop that removes the divisibility padding.  OpenCV's bilinear resize: source
// coordinate (dst+0.5)/s - 0.5, floor, clamp to the edge, separable lerp horizontal then vertical.
#include "common.cuh"

namespace dis {
namespace {

// Optional export of the engine's own output (OFC::OFClass outflow = the level-lv_l flow the finish kernel reads)
// to the per-run pointer in the mailbox: what a stream consumer or a multi-GPU result gather takes instead of the
// full-resolution field.  Folded into the finish kernels: a separate copy after the graph costs ~70 us of
// throughput per launch (measured), this costs one float2 per thread.
__device__ __forceinline__ void export_level(const float2* __restrict__ fl, int n_px, const Mailbox* __restrict__ mb) {
  float2* __restrict__ lo = mb->lvl_out;
  if (lo == nullptr) return;
  const int nthr = gridDim.x * gridDim.y * blockDim.x * blockDim.y;
  const int tid = (blockIdx.y * gridDim.x + blockIdx.x) * (blockDim.x * blockDim.y) + threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < n_px; i += nthr) lo[i] = fl[i];
}

// One thread = 4 consecutive output pixels of a row: the vertical weights and row pointers are shared, the four
// source taps of each pixel come from L1.  Same expressions, in the same order, as a per-pixel evaluation.
__global__ void __launch_bounds__(256) k_finish(const float2* __restrict__ fl, int wl, int hl, int lv_l,
                                                int left, int top, int w_org, int h_org,
                                                const Mailbox* __restrict__ mb, size_t bstride) {
  pdl_wait();
  fl = bshift_nn(fl, (size_t)blockIdx.z * bstride);  // blockIdx.z = pair of a batched handle
  mb = bshift_nn(mb, (size_t)blockIdx.z * bstride);
  float2* __restrict__ out = mb->out;
  export_level(fl, wl * hl, mb);
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

// Scale 4 (lv_l = 2, the 1080p operating points) with padding offsets that are multiples of 4: an aligned 4x4
// block of output pixels reads a 3x3 neighbourhood of the level flow; the four horizontal lerps of each source row
// are shared by the output rows that use it.  Every value is produced by the same expression as in k_finish
// (horizontal lerp of the pre-scaled samples, then vertical lerp), so the result is bit-identical.  Blocks that
// touch the border clamps or the crop edge fall back to the per-pixel form.
__global__ void __launch_bounds__(256) k_finish_x4(const float2* __restrict__ fl, int wl, int hl, int left, int top,
                                                   int w_org, int h_org, const Mailbox* __restrict__ mb, size_t bstride) {
  pdl_wait();
  fl = bshift_nn(fl, (size_t)blockIdx.z * bstride);  // blockIdx.z = pair of a batched handle
  mb = bshift_nn(mb, (size_t)blockIdx.z * bstride);
  float2* __restrict__ out = mb->out;
  export_level(fl, wl * hl, mb);
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
  const int x0 = bx * 4, y0 = by * 4;
  if (x0 >= w_org || y0 >= h_org) return;
  const int X0 = x0 + left, Y0 = y0 + top;  // multiples of 4
  const int cx = X0 >> 2, cy = Y0 >> 2;
  const float s = 4.0f, inv = 0.25f;
  const bool interior = cx >= 1 && cx + 1 <= wl - 2 && cy >= 1 && cy + 1 <= hl - 2 && x0 + 3 < w_org && y0 + 3 < h_org;
  if (interior) {
    float fx[4], fy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // sx = cx-1, cx-1, cx, cx for k = 0..3 (same for rows)
      const float px = ((float)(X0 + k) + 0.5f) * inv - 0.5f, py = ((float)(Y0 + k) + 0.5f) * inv - 0.5f;
      fx[k] = px - (float)(cx - 1 + (k >> 1));
      fy[k] = py - (float)(cy - 1 + (k >> 1));
    }
    float2 H[3][4];  // horizontally interpolated, per source row
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float2* row = fl + (size_t)(cy - 1 + r) * wl + (cx - 1);
      const float2 p0 = __ldg(row), p1 = __ldg(row + 1), p2 = __ldg(row + 2);
      const float2 q0 = make_float2(p0.x * s, p0.y * s), q1 = make_float2(p1.x * s, p1.y * s),
                   q2 = make_float2(p2.x * s, p2.y * s);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = (k >> 1) ? q1 : q0, b = (k >> 1) ? q2 : q1;
        H[r][k].x = a.x * (1.f - fx[k]) + b.x * fx[k];
        H[r][k].y = a.y * (1.f - fx[k]) + b.y * fx[k];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = j >> 1;
      float2 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k].x = H[r][k].x * (1.f - fy[j]) + H[r + 1][k].x * fy[j];
        v[k].y = H[r][k].y * (1.f - fy[j]) + H[r + 1][k].y * fy[j];
      }
      float2* o = out + (size_t)(y0 + j) * w_org + x0;
      if (((size_t)o & 15) == 0) {
        reinterpret_cast<float4*>(o)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        reinterpret_cast<float4*>(o)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = v[k];
      }
    }
    return;
  }
  // border block: per-pixel form with the clamps of cv::resize
  for (int j = 0; j < 4 && y0 + j < h_org; ++j) {
    float fyv = ((float)(Y0 + j) + 0.5f) * inv - 0.5f;
    int sy = (int)floorf(fyv);
    fyv -= (float)sy;
    if (sy < 0) { fyv = 0.0f; sy = 0; }
    if (sy >= hl - 1) { fyv = 0.0f; sy = hl - 1; }
    const float2* r0 = fl + (size_t)sy * wl;
    const float2* r1 = fl + (size_t)min(sy + 1, hl - 1) * wl;
    for (int k = 0; k < 4 && x0 + k < w_org; ++k) {
      float fxv = ((float)(X0 + k) + 0.5f) * inv - 0.5f;
      int sx = (int)floorf(fxv);
      fxv -= (float)sx;
      if (sx < 0) { fxv = 0.0f; sx = 0; }
      if (sx >= wl - 1) { fxv = 0.0f; sx = wl - 1; }
      const int sx1 = min(sx + 1, wl - 1);
      const float2 p00 = __ldg(r0 + sx), p01 = __ldg(r0 + sx1), p10 = __ldg(r1 + sx), p11 = __ldg(r1 + sx1);
      const float a0 = (p00.x * s) * (1.f - fxv) + (p01.x * s) * fxv;
      const float a1 = (p10.x * s) * (1.f - fxv) + (p11.x * s) * fxv;
      const float b0 = (p00.y * s) * (1.f - fxv) + (p01.y * s) * fxv;
      const float b1 = (p10.y * s) * (1.f - fxv) + (p11.y * s) * fxv;
      out[(size_t)(y0 + j) * w_org + x0 + k] = make_float2(a0 * (1.f - fyv) + a1 * fyv, b0 * (1.f - fyv) + b1 * fyv);
    }
  }
}

}  // namespace

void launch_finish(const float2* flow_l, int wl, int hl, int lv_l, int left, int top, int w_org, int h_org,
                   const Mailbox* mb, int nb, size_t bstride, cudaStream_t st) {
  if (lv_l == 2 && (left & 3) == 0 && (top & 3) == 0) {
    dim3 block(32, 8), grid((w_org + 127) / 128, (h_org + 31) / 32, nb);
    launch_pdl(k_finish_x4, grid, block, 0, st, flow_l, wl, hl, left, top, w_org, h_org, mb, bstride);
    return;
  }
  dim3 block(32, 8), grid((w_org + 127) / 128, (h_org + 7) / 8, nb);
  launch_pdl(k_finish, grid, block, 0, st, flow_l, wl, hl, lv_l, left, top, w_org, h_org, mb, bstride);
}

}  // namespace dis
