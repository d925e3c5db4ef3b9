#include "imgio.h"

#include <zlib.h>

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
    return "unsupported image format (PNG, PGM and PPM are read natively; convert others first)";
  } catch (const std::exception& e) {
    return std::string("image decode failed: ") + e.what();
  }
}
