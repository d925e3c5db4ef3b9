#include "imgio.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>

namespace {

bool read_file(const char* path, std::vector<uint8_t>* buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf->resize(n > 0 ? (size_t)n : 0);
  bool ok = n >= 0 && fread(buf->data(), 1, buf->size(), f) == buf->size();
  fclose(f);
  return ok;
}

// OpenCV's PxM reader -> grey (icvCvt_BGR2Gray_8u_C3C1R): (B*1868 + G*9617 + R*4899 + 8192) >> 14;
// checked against cv2.imread(IMREAD_GRAYSCALE) of a PPM
inline uint8_t gray_cv(int r, int g, int b) { return (uint8_t)((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14); }
// libpng png_do_rgb_to_gray as driven by OpenCV's PNG reader (png_set_rgb_to_gray(.., 0.299, 0.587)):
// 15-bit coefficients 9797/19234/3737, truncating.  Checked bit-for-bit against
// cv2.imread(IMREAD_GRAYSCALE) (OpenCV 4.13) on the reference's RGB frames images/alley_1/*.png.
inline uint8_t gray_png(int r, int g, int b) { return (uint8_t)((9797 * r + 19234 * g + 3737 * b) >> 15); }

// Largest accepted image edge (the .flo reader uses the same bound): keeps every size computation below far from
// overflow and a malformed header from requesting gigabytes.
constexpr long kMaxDim = 99999;
constexpr size_t kMaxPixels = (size_t)1 << 30;

std::string read_pnm(const std::vector<uint8_t>& d, int want, GrayImage* out) {
  size_t p = 2;
  auto next_int = [&](int* v) {
    while (p < d.size()) {
      if (d[p] == '#') {
        while (p < d.size() && d[p] != '\n') ++p;
      } else if (d[p] == ' ' || d[p] == '\n' || d[p] == '\r' || d[p] == '\t') {
        ++p;
      } else {
        break;
      }
    }
    if (p >= d.size() || d[p] < '0' || d[p] > '9') return false;
    long x = 0;
    while (p < d.size() && d[p] >= '0' && d[p] <= '9') {
      x = x * 10 + (d[p++] - '0');
      if (x > kMaxDim) return false;  // no header integer of a valid file is larger (also stops overflow)
    }
    *v = (int)x;
    return true;
  };
  const bool color = d[1] == '6';
  int w, h, mx;
  if (!next_int(&w) || !next_int(&h) || !next_int(&mx)) return "bad PNM header";
  if (mx != 255) return "only 8-bit PNM supported";
  ++p;  // single whitespace after maxval
  if (w <= 0 || h <= 0 || (size_t)w * h > kMaxPixels) return "bad PNM size";
  const size_t need = (size_t)w * h * (color ? 3 : 1);
  if (p + need > d.size()) return "truncated PNM";
  out->w = w;
  out->h = h;
  out->ch = want;
  out->px.resize((size_t)w * h * want);
  const uint8_t* q = d.data() + p;
  if (want == 1) {
    if (!color) {
      memcpy(out->px.data(), q, need);
    } else {
      for (size_t i = 0; i < (size_t)w * h; ++i) out->px[i] = gray_cv(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
    }
  } else {  // BGR like cv::imread(.., IMREAD_COLOR)
    for (size_t i = 0; i < (size_t)w * h; ++i)
      for (int k = 0; k < 3; ++k) out->px[3 * i + k] = color ? q[3 * i + 2 - k] : q[i];
  }
  return "";
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

std::string read_png(const std::vector<uint8_t>& d, int want, GrayImage* out) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (d.size() < 33 || memcmp(d.data(), sig, 8)) return "not a PNG";
  size_t p = 8;
  int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  while (p + 12 <= d.size()) {
    const uint32_t len = be32(&d[p]);
    const char* type = (const char*)&d[p + 4];
    if (p + 12 + len > d.size()) return "truncated PNG";
    const uint8_t* body = &d[p + 8];
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return "bad PNG header";
      if (be32(body) > (uint32_t)kMaxDim || be32(body + 4) > (uint32_t)kMaxDim) return "PNG too large";
      w = (int)be32(body);
      h = (int)be32(body + 4);
      depth = body[8];
      ctype = body[9];
      interlace = body[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    p += 12 + len;
  }
  if (w <= 0 || h <= 0 || (size_t)w * h > kMaxPixels) return "bad PNG header";
  if (depth != 8 || interlace != 0) return "only 8-bit non-interlaced PNG supported";
  int ch;
  switch (ctype) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: return "unsupported PNG colour type";
  }
  const size_t stride = (size_t)w * ch;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size())
    return "PNG inflate failed";
  // undo the scanline filters in place
  std::vector<uint8_t> img(stride * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* src = &raw[(stride + 1) * y + 1];
    uint8_t* cur = &img[stride * y];
    const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)ch ? cur[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
      int v = src[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: {
          const int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c);
          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: return "bad PNG filter";
      }
      cur[x] = (uint8_t)v;
    }
  }
  out->w = w;
  out->h = h;
  out->ch = want;
  out->px.resize((size_t)w * h * want);
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    const uint8_t* q = &img[i * ch];
    if (want == 3) {  // BGR like cv::imread(.., IMREAD_COLOR): alpha dropped, grey replicated
      uint8_t* o = &out->px[3 * i];
      switch (ctype) {
        case 0: case 4: o[0] = o[1] = o[2] = q[0]; break;
        case 2: case 6: o[0] = q[2]; o[1] = q[1]; o[2] = q[0]; break;
        case 3:
          if ((size_t)q[0] * 3 + 2 >= plte.size()) return "PNG palette index out of range";
          o[0] = plte[q[0] * 3 + 2]; o[1] = plte[q[0] * 3 + 1]; o[2] = plte[q[0] * 3];
          break;
      }
      continue;
    }
    switch (ctype) {
      case 0: case 4: out->px[i] = q[0]; break;
      case 2: case 6: out->px[i] = gray_png(q[0], q[1], q[2]); break;
      case 3: {
        if ((size_t)q[0] * 3 + 2 >= plte.size()) return "PNG palette index out of range";
        out->px[i] = gray_png(plte[q[0] * 3], plte[q[0] * 3 + 1], plte[q[0] * 3 + 2]);
        break;
      }
    }
  }
  return "";
}


// ---- baseline JPEG -> grey, the way cv::imread(.., IMREAD_GRAYSCALE) decodes it -------------------------------------
// OpenCV hands JPEG files to libjpeg(-turbo) with out_color_space = JCS_GRAYSCALE, which for a YCbCr (or grey) file is
// the luminance plane itself: Huffman decode, dequantise, "islow" inverse DCT (jidctint.c: 13-bit constants,
// PASS1_BITS = 2; the SIMD versions are bit-identical by design), +128, clamp.  Chroma is entropy-decoded only to
// keep the bit stream in step.  Baseline sequential (SOF0 / SOF1-less) 8-bit files with restart markers are covered --
// that is what the reference's images/road_HD.jpg (4:2:0) and images/yosemite_4k.jpg (4:4:4, DRI) are; progressive,
// arithmetic-coded, CMYK and RGB-transform files are refused.  Colour output would need libjpeg's fancy chroma
// upsampling and is not provided: run_dense (= run_OF_INT) is the grey build.
struct JpegHuff {
  uint8_t bits[17] = {};
  uint8_t vals[256] = {};
  int mincode[18] = {}, maxcode[18] = {}, valptr[18] = {};
  bool set = false;
  void build() {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l];
      k += bits[l];
      maxcode[l] = bits[l] ? code - 1 : -1;
      code <<= 1;
    }
    set = true;
  }
};

struct JpegBits {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  bool marker = false;  // ran into a marker (or the end): feed zeros
  void fill() {
    while (n <= 24) {
      int b = 0;
      if (!marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {
            marker = true;
            b = 0;
          }
        } else {
          ++p;
        }
      } else {
        marker = true;
      }
      acc |= (uint32_t)b << (24 - n);
      n += 8;
    }
  }
  int get(int k) {  // k <= 16
    if (k == 0) return 0;
    if (n < k) fill();
    const int v = (int)(acc >> (32 - k));
    acc <<= k;
    n -= k;
    return v;
  }
  int decode(const JpegHuff& h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
      code = (code << 1) | get(1);
      if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
  }
  static int extend(int v, int t) { return t == 0 ? 0 : (v < (1 << (t - 1)) ? v - (1 << t) + 1 : v); }
  bool restart() {  // byte-align and consume RSTn
    acc = 0;
    n = 0;
    marker = false;
    while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
    if (p + 1 >= end) return false;
    p += 2;
    return true;
  }
};

inline int jdescale(long x, int n) { return (int)((x + (1L << (n - 1))) >> n); }

// jidctint.c jpeg_idct_islow on one dequantised block (natural order) -> 8x8 samples
void jpeg_idct_islow(const int* in, uint8_t* out, size_t pitch) {
  constexpr int CB = 13, P1 = 2;
  constexpr long F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                 F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
  long ws[64];
  for (int c = 0; c < 8; ++c) {
    const int* ip = in + c;
    if (!(ip[8] | ip[16] | ip[24] | ip[32] | ip[40] | ip[48] | ip[56])) {
      const long dc = (long)ip[0] * (1 << P1);
      for (int r = 0; r < 8; ++r) ws[8 * r + c] = dc;
      continue;
    }
    long z2 = ip[16], z3 = ip[48];
    long z1 = (z2 + z3) * F_0_541;
    long tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    z2 = ip[0];
    z3 = ip[32];
    long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);
    const long t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
    tmp0 = ip[56];
    tmp1 = ip[40];
    tmp2 = ip[24];
    tmp3 = ip[8];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    long z4 = tmp1 + tmp3;
    const long z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
    z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    ws[0 * 8 + c] = jdescale(t10 + tmp3, CB - P1); ws[7 * 8 + c] = jdescale(t10 - tmp3, CB - P1);
    ws[1 * 8 + c] = jdescale(t11 + tmp2, CB - P1); ws[6 * 8 + c] = jdescale(t11 - tmp2, CB - P1);
    ws[2 * 8 + c] = jdescale(t12 + tmp1, CB - P1); ws[5 * 8 + c] = jdescale(t12 - tmp1, CB - P1);
    ws[3 * 8 + c] = jdescale(t13 + tmp0, CB - P1); ws[4 * 8 + c] = jdescale(t13 - tmp0, CB - P1);
  }
  auto limit = [](int x) {  // range_limit[x & RANGE_MASK]: 10-bit wrap, +128, clamp
    int v = ((x & 1023) ^ 512) - 512;
    v += 128;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  };
  for (int r = 0; r < 8; ++r) {
    const long* w = ws + 8 * r;
    uint8_t* o = out + (size_t)r * pitch;
    long z2 = w[2], z3 = w[6];
    long z1 = (z2 + z3) * F_0_541;
    long tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    long tmp0 = (w[0] + w[4]) * (1L << CB), tmp1 = (w[0] - w[4]) * (1L << CB);
    const long t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
    tmp0 = w[7];
    tmp1 = w[5];
    tmp2 = w[3];
    tmp3 = w[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    long z4 = tmp1 + tmp3;
    const long z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
    z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    constexpr int S = CB + P1 + 3;
    o[0] = limit(jdescale(t10 + tmp3, S)); o[7] = limit(jdescale(t10 - tmp3, S));
    o[1] = limit(jdescale(t11 + tmp2, S)); o[6] = limit(jdescale(t11 - tmp2, S));
    o[2] = limit(jdescale(t12 + tmp1, S)); o[5] = limit(jdescale(t12 - tmp1, S));
    o[3] = limit(jdescale(t13 + tmp0, S)); o[4] = limit(jdescale(t13 - tmp0, S));
  }
}

std::string read_jpeg(const std::vector<uint8_t>& d, int want, GrayImage* out) {
  if (want != 1) return "JPEG is decoded to grey only (the colour build needs PNG / PPM input)";
  static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  uint16_t qt[4][64] = {};
  bool qset[4] = {};
  JpegHuff hdc[4], hac[4];
  struct Comp { int id, h, v, tq, td, ta; };
  Comp comp[4] = {};
  int ncomp = 0, w = 0, h = 0, dri = 0, adobe_transform = -1;
  size_t p = 2;
  while (p + 4 <= d.size()) {
    if (d[p] != 0xFF) return "bad JPEG marker";
    const int m = d[p + 1];
    p += 2;
    if (m == 0xFF) { --p; continue; }  // fill byte
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (p + 2 > d.size()) return "truncated JPEG";
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    if (len < 2 || p + len > d.size()) return "truncated JPEG";
    const uint8_t* b = &d[p + 2];
    const size_t n = len - 2;
    if (m == 0xC0 || m == 0xC1) {
      if (n < 6 || b[0] != 8) return "only 8-bit JPEG supported";
      h = (b[1] << 8) | b[2];
      w = (b[3] << 8) | b[4];
      ncomp = b[5];
      if ((ncomp != 1 && ncomp != 3) || n < 6 + 3 * (size_t)ncomp) return "unsupported JPEG component count";
      for (int c = 0; c < ncomp; ++c) {
        comp[c].id = b[6 + 3 * c];
        comp[c].h = b[7 + 3 * c] >> 4;
        comp[c].v = b[7 + 3 * c] & 15;
        comp[c].tq = b[8 + 3 * c] & 3;
        if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4) return "bad JPEG sampling factors";
      }
    } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
      return "only baseline sequential JPEG supported (this one is progressive, lossless or arithmetic-coded)";
    } else if (m == 0xDB) {
      size_t q = 0;
      while (q < n) {
        const int pq = b[q] >> 4, tq = b[q] & 15;
        if (tq > 3 || q + 1 + (pq ? 128 : 64) > n) return "bad JPEG quantisation table";
        for (int k = 0; k < 64; ++k) qt[tq][zz[k]] = pq ? (uint16_t)((b[q + 1 + 2 * k] << 8) | b[q + 2 + 2 * k]) : b[q + 1 + k];
        qset[tq] = true;
        q += 1 + (pq ? 128 : 64);
      }
    } else if (m == 0xC4) {
      size_t q = 0;
      while (q + 17 <= n) {
        const int tc = b[q] >> 4, th = b[q] & 15;
        if (tc > 1 || th > 3) return "bad JPEG Huffman table";
        JpegHuff& t = tc ? hac[th] : hdc[th];
        int total = 0;
        for (int l = 1; l <= 16; ++l) total += (t.bits[l] = b[q + l]);
        if (total > 256 || q + 17 + total > n) return "bad JPEG Huffman table";
        memcpy(t.vals, b + q + 17, total);
        t.build();
        q += 17 + total;
      }
    } else if (m == 0xDD) {
      if (n >= 2) dri = (b[0] << 8) | b[1];
    } else if (m == 0xEE) {
      if (n >= 12 && !memcmp(b, "Adobe", 5)) adobe_transform = b[11];
    } else if (m == 0xDA) {
      if (!w || !h) return "JPEG scan before frame header";
      if (n < 1 || b[0] != ncomp || n < 1 + 2 * (size_t)ncomp + 3) return "only single-scan (interleaved) baseline JPEG supported";
      for (int c = 0; c < ncomp; ++c) {
        if (b[1 + 2 * c] != comp[c].id) return "unexpected JPEG scan component order";
        comp[c].td = b[2 + 2 * c] >> 4;
        comp[c].ta = b[2 + 2 * c] & 15;
        if (comp[c].td > 3 || comp[c].ta > 3 || !hdc[comp[c].td].set || !hac[comp[c].ta].set || !qset[comp[c].tq])
          return "JPEG scan refers to a missing table";
      }
      p += len;
      break;
    }
    p += len;
  }
  if (!w || !h || p >= d.size()) return "no JPEG scan found";
  if ((long)w > kMaxDim || (long)h > kMaxDim || (size_t)w * h > kMaxPixels) return "JPEG too large";
  if (ncomp == 3 && adobe_transform == 0) return "RGB-coded JPEG (Adobe transform 0) not supported";
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; ++c) {
    hmax = std::max(hmax, comp[c].h);
    vmax = std::max(vmax, comp[c].v);
  }
  if (ncomp == 1) comp[0].h = comp[0].v = hmax = vmax = 1;  // a single-component scan is never interleaved
  if (comp[0].h != hmax || comp[0].v != vmax) return "JPEG with a subsampled luminance plane not supported";
  const int mw = 8 * hmax, mh = 8 * vmax, mx = (w + mw - 1) / mw, my = (h + mh - 1) / mh;
  const size_t pw = (size_t)mx * mw, ph = (size_t)my * mh;
  std::vector<uint8_t> Y(pw * ph);
  JpegBits br{&d[p], d.data() + d.size()};
  int pred[4] = {};
  int count = 0;
  for (int y = 0; y < my; ++y)
    for (int x = 0; x < mx; ++x) {
      if (dri && count == dri) {
        if (!br.restart()) return "truncated JPEG (restart marker missing)";
        pred[0] = pred[1] = pred[2] = pred[3] = 0;
        count = 0;
      }
      ++count;
      for (int c = 0; c < ncomp; ++c)
        for (int by = 0; by < comp[c].v; ++by)
          for (int bx = 0; bx < comp[c].h; ++bx) {
            int blk[64] = {};
            int t = br.decode(hdc[comp[c].td]);
            if (t < 0 || t > 11) return "corrupt JPEG data";
            pred[c] += JpegBits::extend(br.get(t), t);
            blk[0] = pred[c] * qt[comp[c].tq][0];
            for (int k = 1; k < 64;) {
              const int rs = br.decode(hac[comp[c].ta]);
              if (rs < 0) return "corrupt JPEG data";
              const int r = rs >> 4, sz = rs & 15;
              if (sz == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) return "corrupt JPEG data";
              blk[zz[k]] = JpegBits::extend(br.get(sz), sz) * qt[comp[c].tq][zz[k]];
              ++k;
            }
            if (c == 0) jpeg_idct_islow(blk, &Y[((size_t)y * mh + 8 * by) * pw + (size_t)x * mw + 8 * bx], pw);
          }
    }
  out->w = w;
  out->h = h;
  out->ch = 1;
  out->px.resize((size_t)w * h);
  for (int y = 0; y < h; ++y) memcpy(&out->px[(size_t)y * w], &Y[(size_t)y * pw], w);
  return "";
}

}  // namespace

std::string write_png_bgr(const char* path, const uint8_t* bgr, int w, int h) {
  // scanlines: filter byte 0 + RGB
  const size_t stride = (size_t)w * 3 + 1;
  std::vector<uint8_t> raw(stride * h);
  for (int y = 0; y < h; ++y) {
    uint8_t* r = &raw[stride * y];
    *r++ = 0;
    const uint8_t* s = bgr + (size_t)y * w * 3;
    for (int x = 0; x < w; ++x, s += 3, r += 3) {
      r[0] = s[2];
      r[1] = s[1];
      r[2] = s[0];
    }
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return "PNG deflate failed";
  FILE* f = fopen(path, "wb");
  if (!f) return std::string("cannot write ") + path;
  auto put32 = [](uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
  bool ok = true;
  auto chunk = [&](const char* type, const uint8_t* data, size_t len) {
    uint8_t hd[8];
    put32(hd, (uint32_t)len);
    memcpy(hd + 4, type, 4);
    uint32_t crc = crc32(0, hd + 4, 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    uint8_t tail[4];
    put32(tail, crc);
    ok = ok && fwrite(hd, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
  };
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  ok = fwrite(sig, 1, 8, f) == 8;
  uint8_t ihdr[13];
  put32(ihdr, (uint32_t)w);
  put32(ihdr + 4, (uint32_t)h);
  ihdr[8] = 8;   // bit depth
  ihdr[9] = 2;   // colour type RGB
  ihdr[10] = ihdr[11] = ihdr[12] = 0;
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", z.data(), zlen);
  chunk("IEND", nullptr, 0);
  ok = (fclose(f) == 0) && ok;
  return ok ? "" : std::string("short write to ") + path;
}

std::string read_gray_image(const char* path, GrayImage* out) { return read_image(path, 1, out); }

std::string read_image(const char* path, int channels, GrayImage* out) {
  if (channels != 1 && channels != 3) return "channels must be 1 or 3";
  try {  // nothing may propagate through the extern "C" callers (std::bad_alloc on a hostile header)
    std::vector<uint8_t> d;
    if (!read_file(path, &d)) return std::string("cannot read ") + path;
    if (d.size() > 2 && d[0] == 'P' && (d[1] == '5' || d[1] == '6')) return read_pnm(d, channels, out);
    if (d.size() > 8 && d[0] == 0x89 && d[1] == 'P') return read_png(d, channels, out);
    if (d.size() > 4 && d[0] == 0xFF && d[1] == 0xD8) return read_jpeg(d, channels, out);
    return "unsupported image format (PNG, PGM, PPM and baseline JPEG are read natively; convert others first)";
  } catch (const std::exception& e) {
    return std::string("image decode failed: ") + e.what();
  }
}
