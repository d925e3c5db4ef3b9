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
  __device__ __forceinline__ void shift(size_t off) { mb = bshift(mb, off); }
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
  __device__ __forceinline__ void shift(size_t off) { p = bshift(p, off); }
  __device__ __forceinline__ void resolve() {}
  __device__ __forceinline__ float at(int x, int y, int ch) const {
    const float* r0 = p + (size_t)(2 * y + pad) * pitch + (2 * x + pad) * NC + ch;
    const float2 a = make_float2(__ldg(r0), __ldg(r0 + NC));
    const float2 b = make_float2(__ldg(r0 + pitch), __ldg(r0 + pitch + NC));
    return ((a.x + a.y) + (b.x + b.y)) * 0.25f;
  }
};

template <int NC>
struct SrcPlain {  // unpadded level image (output of k_block_mean)
  const float* p;
  int w;
  __device__ __forceinline__ void shift(size_t off) { p = bshift(p, off); }
  __device__ __forceinline__ void resolve() {}
  __device__ __forceinline__ float at(int x, int y, int ch) const { return __ldg(p + ((size_t)y * w + x) * NC + ch); }
};

// Level L (1 <= L <= 8) straight from the u8 frames: the 2x2 box means of levels 1..L telescope into the mean of a
// 2^L x 2^L block, and every partial sum of the reference's ((a+b)+(c+d))*0.25 chain is exact in fp32 (values are
// multiples of 4^-(l-1) below 256: 8+2l-1 <= 24 significant bits), so float(sum of the block) * 4^-L is bit-identical
// to the level-by-level result.  The block is taken from the replicate-padded level-0 image (run_dense.cpp:298-311),
// i.e. source coordinates are clamped.  Saves writing and re-reading the levels below lv_l that nothing else uses.
template <int NC>
__global__ void __launch_bounds__(256) k_block_mean(const Mailbox* __restrict__ mb0, int L, int w_org, int h_org,
                                                    int left, int top, int w, int h, float* __restrict__ out_a,
                                                    float* __restrict__ out_b, size_t bstride, int only_b) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  // grid.z = 2 * pair + frame; only_b (frame a's pyramid is reused from the previous pair of a stream): grid.z = pair
  const size_t boff = (size_t)(only_b ? blockIdx.z : blockIdx.z >> 1) * bstride;
  const Mailbox* __restrict__ mb = bshift(mb0, boff);
  const bool second = only_b || (blockIdx.z & 1);
  const uint8_t* __restrict__ src = second ? mb->b : mb->a;
  float* __restrict__ out = bshift(second ? out_b : out_a, boff);
  const int pitch = mb->pitch, B = 1 << L;
  const int x0 = x * B - left, y0 = y * B - top;
  unsigned acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0;
  const bool inside = x0 >= 0 && y0 >= 0 && x0 + B <= w_org && y0 + B <= h_org;
  for (int dy = 0; dy < B; ++dy) {
    const int sy = min(max(y0 + dy, 0), h_org - 1);
    const uint8_t* row = src + (size_t)sy * pitch;
    if (inside && NC == 1 && B >= 4 && (((size_t)(row + x0)) & 3) == 0) {
      for (int dx = 0; dx < B; dx += 4) {  // aligned interior: 4 pixels per load
        const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(row + x0 + dx));
        acc[0] += (unsigned)v.x + v.y + v.z + v.w;
      }
    } else {
      for (int dx = 0; dx < B; ++dx) {
        const int sx = min(max(x0 + dx, 0), w_org - 1);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += __ldg(row + sx * NC + c);
      }
    }
  }
  const float scale = 1.0f / (float)(1u << (2 * L));  // power of two: exact
#pragma unroll
  for (int c = 0; c < NC; ++c) out[((size_t)y * w + x) * NC + c] = (float)acc[c] * scale;
}

template <int NC, typename Src>
__global__ void __launch_bounds__(256) k_pyr_level(Src sa, Src sb, int w, int h, int pad, int pitch,
                                                   int tw, int th, float* Ia, float* Iax, float* Iay,
                                                   float* Ib, float* Ibx, float* Iby, size_t bstride, int only_b) {
  const int F0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;  // first float column of this thread
  const int Y = blockIdx.y * blockDim.y + threadIdx.y;
  if (F0 >= pitch || Y >= th) return;
  const bool second = only_b || (blockIdx.z & 1);  // grid.z = 2 * pair + frame (only_b: grid.z = pair)
  const size_t boff = (size_t)(only_b ? blockIdx.z : blockIdx.z >> 1) * bstride;
  Src s = second ? sb : sa;
  s.shift(boff);
  s.resolve();
  float* I = bshift(second ? Ib : Ia, boff);
  float* Gx = bshift(second ? Ibx : Iax, boff);
  float* Gy = bshift(second ? Iby : Iay, boff);
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
            float* Ibx, float* Iby, cudaStream_t st, int only_b) {
  dim3 block(64, 4);
  dim3 grid((g.pitch / 4 + block.x - 1) / block.x, (g.th + block.y - 1) / block.y, (only_b ? 1 : 2) * g.nb);
  k_pyr_level<NC, Src><<<grid, block, 0, st>>>(sa, sb, g.w, g.h, g.pad, g.pitch, g.tw, g.th, Ia, Iax, Iay,
                                               Ib, Ibx, Iby, g.bstride, only_b);
}

}  // namespace

__global__ void k_set_mailboxes(Mailbox* mb0, size_t bstride, int nb, const MailboxBatch m, int pitch) {
  const int b = threadIdx.x;
  if (b >= nb) return;
  Mailbox* mb = bshift(mb0, (size_t)b * bstride);
  mb->a = m.a[b];
  mb->b = m.b[b];
  mb->out = m.out[b];
  mb->lvl_out = m.lvl[b];
  mb->pitch = pitch;
}

void launch_set_mailboxes(Mailbox* mb0, size_t bstride, int nb, const MailboxBatch& m, int pitch, cudaStream_t st) {
  k_set_mailboxes<<<1, kMaxBatch, 0, st>>>(mb0, bstride, nb, m, pitch);
}

void launch_level0(const Mailbox* mb, int w_org, int h_org, int left, int top, const LevelGeom& g, float* Ia,
                   float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby, cudaStream_t st, int only_b) {
  if (g.noc == 3) {
    SrcU8<3> sa{mb, 0, nullptr, w_org, h_org, 0, left, top}, sb{mb, 1, nullptr, w_org, h_org, 0, left, top};
    launch<3>(sa, sb, g, Ia, Iax, Iay, Ib, Ibx, Iby, st, only_b);
  } else {
    SrcU8<1> sa{mb, 0, nullptr, w_org, h_org, 0, left, top}, sb{mb, 1, nullptr, w_org, h_org, 0, left, top};
    launch<1>(sa, sb, g, Ia, Iax, Iay, Ib, Ibx, Iby, st, only_b);
  }
}

// First processed level L = lv_l > 0 of the product path: block means from the u8 frames into the scratch images
// bm_a / bm_b (g.w x g.h x noc floats), then the padded image + gradients from those.
void launch_first_level(const Mailbox* mb, int L, int w_org, int h_org, int left, int top, const LevelGeom& g,
                        float* bm_a, float* bm_b, float* Ia, float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby,
                        cudaStream_t st, int only_b) {
  dim3 block(32, 8), grid((g.w + 31) / 32, (g.h + 7) / 8, (only_b ? 1 : 2) * g.nb);
  if (g.noc == 3) {
    k_block_mean<3><<<grid, block, 0, st>>>(mb, L, w_org, h_org, left, top, g.w, g.h, bm_a, bm_b, g.bstride, only_b);
    SrcPlain<3> sa{bm_a, g.w}, sb{bm_b, g.w};
    launch<3>(sa, sb, g, Ia, Iax, Iay, Ib, Ibx, Iby, st, only_b);
  } else {
    k_block_mean<1><<<grid, block, 0, st>>>(mb, L, w_org, h_org, left, top, g.w, g.h, bm_a, bm_b, g.bstride, only_b);
    SrcPlain<1> sa{bm_a, g.w}, sb{bm_b, g.w};
    launch<1>(sa, sb, g, Ia, Iax, Iay, Ib, Ibx, Iby, st, only_b);
  }
}

void launch_downsample(const LevelGeom& gf, const LevelGeom& gc, const float* Ia_f, const float* Ib_f,
                       float* Ia, float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby,
                       cudaStream_t st, int only_b) {
  if (gc.noc == 3) {
    SrcDown<3> sa{Ia_f, gf.pitch, gf.pad}, sb{Ib_f, gf.pitch, gf.pad};
    launch<3>(sa, sb, gc, Ia, Iax, Iay, Ib, Ibx, Iby, st, only_b);
  } else {
    SrcDown<1> sa{Ia_f, gf.pitch, gf.pad}, sb{Ib_f, gf.pitch, gf.pad};
    launch<1>(sa, sb, gc, Ia, Iax, Iay, Ib, Ibx, Iby, st, only_b);
  }
}

}  // namespace dis
