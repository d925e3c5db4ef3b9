// flowviz.cu -- the evaluation tools that sit behind the hot path: Middlebury colour coding of a flow field
// and endpoint-error statistics, both on the GPU.
//
// Colour coding replaces flow_code/C/color_flow.cpp:MotionToColor (:19-71) + colorcode.cpp:computeColor
// (:53-77): the reference's own known-answer pair kroeger/flows/alley_0001.flo -> alley_0001.png is
// reproduced byte for byte (tests/golden).  The reference mixes float and double arithmetic (C's usual
// promotions: sqrt/atan2 in double, `/ 255.0`, `* .75`, `255.0 * col` in double, everything assigned to
// float variables); the kernels below keep every one of those conversions.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace dis {
namespace {

constexpr int kNCols = 55;  // RY + YG + GC + CB + BM + MR (colorcode.cpp:33-39)
__constant__ int c_wheel[kNCols][3];

void make_wheel(int wheel[kNCols][3]) {  // colorcode.cpp:26-51
  const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
  int k = 0;
  auto set = [&](int r, int g, int b) { wheel[k][0] = r; wheel[k][1] = g; wheel[k][2] = b; ++k; };
  for (int i = 0; i < RY; i++) set(255, 255 * i / RY, 0);
  for (int i = 0; i < YG; i++) set(255 - 255 * i / YG, 255, 0);
  for (int i = 0; i < GC; i++) set(0, 255, 255 * i / GC);
  for (int i = 0; i < CB; i++) set(0, 255 - 255 * i / CB, 255);
  for (int i = 0; i < BM; i++) set(255 * i / BM, 0, 255);
  for (int i = 0; i < MR; i++) set(255, 0, 255 - 255 * i / MR);
}

// flowIO.cpp:35-39 (UNKNOWN_FLOW_THRESH 1e9, flowIO.h:5)
__device__ __forceinline__ bool unknown_flow(float u, float v) {
  return fabs((double)u) > 1e9 || fabs((double)v) > 1e9 || isnan(u) || isnan(v);
}

// order-preserving float <-> uint mapping for atomicMin/Max
__device__ __forceinline__ unsigned enc(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec(unsigned u) {
  const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// stats[0..4] = maxrad, minx, maxx, miny, maxy (encoded); reference initial values color_flow.cpp:27-29
__global__ void k_range_init(unsigned* stats) {
  stats[0] = enc(-1.0f);
  stats[1] = enc(999.0f);
  stats[2] = enc(-999.0f);
  stats[3] = enc(999.0f);
  stats[4] = enc(-999.0f);
}

__global__ void __launch_bounds__(256) k_range(const float2* __restrict__ fl, size_t n, unsigned* stats) {
  float mr = -1.0f, mnx = 999.0f, mxx = -999.0f, mny = 999.0f, mxy = -999.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float2 f = fl[i];
    if (unknown_flow(f.x, f.y)) continue;
    mxx = fmaxf(mxx, f.x);
    mxy = fmaxf(mxy, f.y);
    mnx = fminf(mnx, f.x);
    mny = fminf(mny, f.y);
    const float rad = (float)sqrt((double)(f.x * f.x + f.y * f.y));
    mr = fmaxf(mr, rad);
  }
  for (int o = 16; o; o >>= 1) {
    mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&stats[0], enc(mr));
    atomicMin(&stats[1], enc(mnx));
    atomicMax(&stats[2], enc(mxx));
    atomicMin(&stats[3], enc(mny));
    atomicMax(&stats[4], enc(mxy));
  }
}

__global__ void __launch_bounds__(256) k_color(const float2* __restrict__ fl, size_t n, float maxmotion,
                                               const unsigned* __restrict__ stats, uint8_t* __restrict__ bgr) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float maxrad = dec(stats[0]);
  if (maxmotion > 0) maxrad = maxmotion;  // color_flow.cpp:49-53
  if (maxrad == 0) maxrad = 1;
  const float2 f = fl[i];
  uint8_t* pix = bgr + 3 * i;
  if (unknown_flow(f.x, f.y)) {
    pix[0] = pix[1] = pix[2] = 0;
    return;
  }
  // computeColor(fx / maxrad, fy / maxrad, pix), colorcode.cpp:53-77
  const float fx = f.x / maxrad, fy = f.y / maxrad;
  const float rad = (float)sqrt((double)(fx * fx + fy * fy));
  // atan2(-fy, -fx) on floats is atan2f in the reference (C++ overload); its last bit depends on the host
  // libm, so the correctly rounded float arctangent is used (double atan2 rounded to float)
  const float at = (float)atan2((double)-fy, (double)-fx);
  const float a = (float)((double)at / M_PI);
  const float fk = (float)(((double)a + 1.0) / 2.0 * (double)(kNCols - 1));
  const int k0 = (int)fk;
  const int k1 = (k0 + 1) % kNCols;
  const float ff = fk - (float)k0;
  for (int b = 0; b < 3; b++) {
    const float col0 = (float)(c_wheel[k0][b] / 255.0);
    const float col1 = (float)(c_wheel[k1][b] / 255.0);
    float col = (1 - ff) * col0 + ff * col1;
    if (rad <= 1)
      col = 1 - rad * (1 - col);  // increase saturation with radius
    else
      col = (float)((double)col * .75);  // out of range
    pix[2 - b] = (uint8_t)(int)(255.0 * (double)col);
  }
}

// endpoint error |a - b| per pixel inside the margin; per-block partial sums in double, summed in block order
// by the host -> deterministic
__global__ void __launch_bounds__(256) k_epe(const float2* __restrict__ fa, const float2* __restrict__ fb, int w,
                                             int h, int margin, double* __restrict__ psum,
                                             float* __restrict__ pmax, unsigned long long* __restrict__ pcnt) {
  __shared__ double s_sum[8];
  __shared__ float s_max[8];
  __shared__ unsigned s_cnt[8];
  const int iw = w - 2 * margin, ih = h - 2 * margin;
  const size_t n = (size_t)iw * ih;
  double sum = 0.0;
  float mx = 0.0f;
  unsigned cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / iw) + margin, x = (int)(i % iw) + margin;
    const float2 a = fa[(size_t)y * w + x], b = fb[(size_t)y * w + x];
    if (unknown_flow(a.x, a.y) || unknown_flow(b.x, b.y)) continue;
    const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
    const double e = sqrt(dx * dx + dy * dy);
    sum += e;
    mx = fmaxf(mx, (float)e);
    ++cnt;
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_sum[wid] = sum;
    s_max[wid] = mx;
    s_cnt[wid] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float m = 0.0f;
    unsigned long long c = 0;
    for (int k = 0; k < 8; ++k) {
      t += s_sum[k];
      m = fmaxf(m, s_max[k]);
      c += s_cnt[k];
    }
    psum[blockIdx.x] = t;
    pmax[blockIdx.x] = m;
    pcnt[blockIdx.x] = c;
  }
}

}  // namespace

void flowviz_init_device() {
  int wheel[kNCols][3];
  make_wheel(wheel);
  cudaMemcpyToSymbol(c_wheel, wheel, sizeof(wheel));
}

// d_stats: 5 words of device scratch
void launch_flow_color(const float2* d_flow, int w, int h, float maxmotion, uint8_t* d_bgr, unsigned* d_stats,
                       cudaStream_t st) {
  const size_t n = (size_t)w * h;
  k_range_init<<<1, 1, 0, st>>>(d_stats);
  const int blocks = (int)min((n + 255) / 256, (size_t)148 * 8);
  k_range<<<blocks, 256, 0, st>>>(d_flow, n, d_stats);
  k_color<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_flow, n, maxmotion, d_stats, d_bgr);
}

void decode_flow_stats(const unsigned* h_stats, float out[5]) {
  for (int i = 0; i < 5; ++i) out[i] = dec(h_stats[i]);
}

int flow_epe_blocks() { return 148 * 4; }

void launch_flow_epe(const float2* d_a, const float2* d_b, int w, int h, int margin, double* d_psum, float* d_pmax,
                     unsigned long long* d_pcnt, cudaStream_t st) {
  k_epe<<<flow_epe_blocks(), 256, 0, st>>>(d_a, d_b, w, h, margin, d_psum, d_pmax, d_pcnt);
}

}  // namespace dis

// ---- C-ABI (include/dis_c.h) ---------------------------------------------------------------------------
#include "../../include/dis_c.h"
#include "imgio.h"

namespace {

struct DevBuf {  // scope-bound device allocation
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

#define CUG(call)                                                                                              \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess) {                                                                                   \
      dis::set_global_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);       \
      return DIS_ERR_CUDA;                                                                                     \
    }                                                                                                          \
  } while (0)

}  // namespace

extern "C" {

int dis_flow_to_color_device(const float* d_flow_uv, int w, int h, float maxmotion, uint8_t* d_bgr,
                             uint32_t* d_stats, void* stream) {
  if (!d_flow_uv || !d_bgr || !d_stats || w <= 0 || h <= 0) return DIS_ERR_INVALID_ARG;
  dis::launch_flow_color(reinterpret_cast<const float2*>(d_flow_uv), w, h, maxmotion, d_bgr, d_stats,
                         static_cast<cudaStream_t>(stream));
  return cudaGetLastError() == cudaSuccess ? DIS_OK : DIS_ERR_CUDA;
}

int dis_flow_to_color(const float* flow_uv, int w, int h, float maxmotion, int device, uint8_t* bgr_out,
                      float* stats_out) {
  if (!flow_uv || !bgr_out || w <= 0 || h <= 0) {
    dis::set_global_error("dis_flow_to_color: bad argument");
    return DIS_ERR_INVALID_ARG;
  }
  CUG(cudaSetDevice(device));
  dis::flowviz_init_device();
  const size_t n = (size_t)w * h;
  DevBuf d_flow, d_bgr, d_stats;
  CUG(d_flow.alloc(n * sizeof(float2)));
  CUG(d_bgr.alloc(n * 3));
  CUG(d_stats.alloc(5 * sizeof(unsigned)));
  CUG(cudaMemcpy(d_flow.p, flow_uv, n * sizeof(float2), cudaMemcpyHostToDevice));
  dis::launch_flow_color(d_flow.as<float2>(), w, h, maxmotion, d_bgr.as<uint8_t>(), d_stats.as<unsigned>(), 0);
  CUG(cudaGetLastError());
  CUG(cudaMemcpy(bgr_out, d_bgr.p, n * 3, cudaMemcpyDeviceToHost));
  if (stats_out) {
    unsigned hs[5];
    CUG(cudaMemcpy(hs, d_stats.p, sizeof hs, cudaMemcpyDeviceToHost));
    dis::decode_flow_stats(hs, stats_out);
  }
  return DIS_OK;
}

int dis_flow_epe(const float* flow_a, const float* flow_b, int w, int h, int margin, int device, double* mean_out,
                 double* max_out, long long* count_out) {
  if (!flow_a || !flow_b || w <= 0 || h <= 0 || margin < 0 || 2 * margin >= w || 2 * margin >= h) {
    dis::set_global_error("dis_flow_epe: bad argument");
    return DIS_ERR_INVALID_ARG;
  }
  CUG(cudaSetDevice(device));
  const size_t n = (size_t)w * h;
  const int nb = dis::flow_epe_blocks();
  DevBuf d_a, d_b, d_sum, d_max, d_cnt;
  CUG(d_a.alloc(n * sizeof(float2)));
  CUG(d_b.alloc(n * sizeof(float2)));
  CUG(d_sum.alloc(nb * sizeof(double)));
  CUG(d_max.alloc(nb * sizeof(float)));
  CUG(d_cnt.alloc(nb * sizeof(unsigned long long)));
  CUG(cudaMemcpy(d_a.p, flow_a, n * sizeof(float2), cudaMemcpyHostToDevice));
  CUG(cudaMemcpy(d_b.p, flow_b, n * sizeof(float2), cudaMemcpyHostToDevice));
  dis::launch_flow_epe(d_a.as<float2>(), d_b.as<float2>(), w, h, margin, d_sum.as<double>(), d_max.as<float>(),
                       d_cnt.as<unsigned long long>(), 0);
  CUG(cudaGetLastError());
  std::vector<double> ps(nb);
  std::vector<float> pm(nb);
  std::vector<unsigned long long> pc(nb);
  CUG(cudaMemcpy(ps.data(), d_sum.p, nb * sizeof(double), cudaMemcpyDeviceToHost));
  CUG(cudaMemcpy(pm.data(), d_max.p, nb * sizeof(float), cudaMemcpyDeviceToHost));
  CUG(cudaMemcpy(pc.data(), d_cnt.p, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double sum = 0.0, mx = 0.0;
  unsigned long long cnt = 0;
  for (int i = 0; i < nb; ++i) {
    sum += ps[i];
    mx = pm[i] > mx ? pm[i] : mx;
    cnt += pc[i];
  }
  if (mean_out) *mean_out = cnt ? sum / (double)cnt : 0.0;
  if (max_out) *max_out = mx;
  if (count_out) *count_out = (long long)cnt;
  return DIS_OK;
}

int dis_write_png_bgr(const char* path, const uint8_t* bgr, int w, int h) {
  if (!path || !bgr || w <= 0 || h <= 0) return DIS_ERR_INVALID_ARG;
  const std::string err = write_png_bgr(path, bgr, w, h);
  if (!err.empty()) {
    dis::set_global_error("%s", err.c_str());
    return DIS_ERR_IO;
  }
  return DIS_OK;
}

}  // extern "C"
