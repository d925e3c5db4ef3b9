// pyramid.cu -- stage 1: image pyramid, central-difference gradients and padding, on the GPU.
//
// Replaces the OpenCV call sequence of the reference's ConstructImgPyramide
// (kroeger/run_dense.cpp:130-178) plus the divisibility padding (:298-311) and the u8->f32
// conversion (:326-327):
//   level 0   = float(u8), replicate-padded to a multiple of 2^lv_f
//   level l   = cv::resize(level l-1, 0.5, INTER_LINEAR)  == 2x2 box mean ((a+b)+(c+d))*0.25
//   Ix, Iy    = cv::Sobel(ksize=1, BORDER_DEFAULT): I(x+1)-I(x-1) with reflect-101 borders
//   padding   = copyMakeBorder by `patchsz`: image REPLICATE, gradients CONSTANT 0
// One kernel per level writes the padded image and both padded gradients of both frames
// (blockIdx.z selects the frame); each thread produces 4 consecutive pixels (float4 stores,
// rows are 128-byte aligned because the pitch is a multiple of 32 floats).
// HBM-bound streaming stencil: per level-0 pixel 1 B in (u8) and 12 B out (I, Ix, Iy).
#include "common.cuh"

namespace dis {
namespace {

// NC = channels (1 grey, 3 interleaved BGR).  Every OpenCV call of ConstructImgPyramide works per channel, so
// the colour pyramid is the grey arithmetic applied to "float columns" f = x*NC + ch of the interleaved rows.
template <int NC>
struct SrcU8 {
  const Mailbox* mb;
  int which;
  const uint8_t* p;
  int w_org, h_org, pitch, left, top;
  __device__ __forceinline__ void resolve() {
    p = which ? mb->b : mb->a;
    pitch = mb->pitch;
  }
  __device__ __forceinline__ float at(int x, int y, int ch) const {
    int sx = min(max(x - left, 0), w_org - 1);
    int sy = min(max(y - top, 0), h_org - 1);
    return (float)__ldg(p + (size_t)sy * pitch + sx * NC + ch);
  }
};

template <int NC>
struct SrcDown {  // 2x2 mean of the finer level (padded array, pad offset applied)
  const float* p;
  int pitch, pad;
  __device__ __forceinline__ void resolve() {}
  __device__ __forceinline__ float at(int x, int y, int ch) const {
    const float* r0 = p + (size_t)(2 * y + pad) * pitch + (2 * x + pad) * NC + ch;
    const float2 a = make_float2(__ldg(r0), __ldg(r0 + NC));
    const float2 b = make_float2(__ldg(r0 + pitch), __ldg(r0 + pitch + NC));
    return ((a.x + a.y) + (b.x + b.y)) * 0.25f;
  }
};

template <int NC, typename Src>
__global__ void __launch_bounds__(256) k_pyr_level(Src sa, Src sb, int w, int h, int pad, int pitch,
                                                   int tw, int th, float* Ia, float* Iax, float* Iay,
                                                   float* Ib, float* Ibx, float* Iby) {
  const int F0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;  // first float column of this thread
  const int Y = blockIdx.y * blockDim.y + threadIdx.y;
  if (F0 >= pitch || Y >= th) return;
  const bool second = blockIdx.z == 1;
  Src s = second ? sb : sa;
  s.resolve();
  float* I = second ? Ib : Ia;
  float* Gx = second ? Ibx : Iax;
  float* Gy = second ? Iby : Iay;
  const int y = min(max(Y - pad, 0), h - 1);
  const bool yin = (Y >= pad) && (Y < pad + h);
  float vi[4], vx[4], vy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int F = F0 + k;
    const int X = (NC == 1) ? F : F / NC, ch = (NC == 1) ? 0 : F - X * NC;
    const int x = min(max(X - pad, 0), w - 1);
    vi[k] = s.at(x, y, ch);
    vx[k] = 0.0f;
    vy[k] = 0.0f;
    if (Gx != nullptr && yin && X >= pad && X < pad + w) {
      // reflect-101: -1 -> 1, w -> w-2
      const int xm = x == 0 ? (w > 1 ? 1 : 0) : x - 1;
      const int xq = x == w - 1 ? (w > 1 ? w - 2 : 0) : x + 1;
      const int ym = y == 0 ? (h > 1 ? 1 : 0) : y - 1;
      const int yq = y == h - 1 ? (h > 1 ? h - 2 : 0) : y + 1;
      vx[k] = s.at(xq, y, ch) - s.at(xm, y, ch);
      vy[k] = s.at(x, yq, ch) - s.at(x, ym, ch);
    }
  }
  const size_t o = (size_t)Y * pitch + F0;
  *reinterpret_cast<float4*>(I + o) = make_float4(vi[0], vi[1], vi[2], vi[3]);
  if (Gx != nullptr) {
    *reinterpret_cast<float4*>(Gx + o) = make_float4(vx[0], vx[1], vx[2], vx[3]);
    *reinterpret_cast<float4*>(Gy + o) = make_float4(vy[0], vy[1], vy[2], vy[3]);
  }
}

template <int NC, typename Src>
void launch(Src sa, Src sb, const LevelGeom& g, float* Ia, float* Iax, float* Iay, float* Ib,
            float* Ibx, float* Iby, cudaStream_t st) {
  dim3 block(64, 4);
  dim3 grid((g.pitch / 4 + block.x - 1) / block.x, (g.th + block.y - 1) / block.y, 2);
  k_pyr_level<NC, Src><<<grid, block, 0, st>>>(sa, sb, g.w, g.h, g.pad, g.pitch, g.tw, g.th, Ia, Iax, Iay,
                                               Ib, Ibx, Iby);
}

}  // namespace

__global__ void k_set_mailbox(Mailbox* mb, const uint8_t* a, const uint8_t* b, float2* out, int pitch) {
  mb->a = a;
  mb->b = b;
  mb->out = out;
  mb->pitch = pitch;
}

void launch_set_mailbox(Mailbox* mb, const uint8_t* a, const uint8_t* b, float2* out, int pitch, cudaStream_t st) {
  k_set_mailbox<<<1, 1, 0, st>>>(mb, a, b, out, pitch);
}

void launch_level0(const Mailbox* mb, int w_org, int h_org, int left, int top, const LevelGeom& g, float* Ia,
                   float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby, cudaStream_t st) {
  if (g.noc == 3) {
    SrcU8<3> sa{mb, 0, nullptr, w_org, h_org, 0, left, top}, sb{mb, 1, nullptr, w_org, h_org, 0, left, top};
    launch<3>(sa, sb, g, Ia, Iax, Iay, Ib, Ibx, Iby, st);
  } else {
    SrcU8<1> sa{mb, 0, nullptr, w_org, h_org, 0, left, top}, sb{mb, 1, nullptr, w_org, h_org, 0, left, top};
    launch<1>(sa, sb, g, Ia, Iax, Iay, Ib, Ibx, Iby, st);
  }
}

void launch_downsample(const LevelGeom& gf, const LevelGeom& gc, const float* Ia_f, const float* Ib_f,
                       float* Ia, float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby,
                       cudaStream_t st) {
  if (gc.noc == 3) {
    SrcDown<3> sa{Ia_f, gf.pitch, gf.pad}, sb{Ib_f, gf.pitch, gf.pad};
    launch<3>(sa, sb, gc, Ia, Iax, Iay, Ib, Ibx, Iby, st);
  } else {
    SrcDown<1> sa{Ia_f, gf.pitch, gf.pad}, sb{Ib_f, gf.pitch, gf.pad};
    launch<1>(sa, sb, gc, Ia, Iax, Iay, Ib, Ibx, Iby, st);
  }
}

}  // namespace dis
