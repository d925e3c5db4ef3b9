// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" entry into the *verbatim* colour-coding function of the reference
// (flow_code/C/colorcode.cpp:computeColor, compiled from where it lies) for
// oracle/_ref/libcolor_ref.so.  The per-image loop around it restates MotionToColor
// (flow_code/C/color_flow.cpp:19-71), which cannot be compiled here because it needs the
// Szeliski imageLib + libpng headers.
#include <cmath>

typedef unsigned char uchar;
void computeColor(float fx, float fy, uchar* pix);  // flow_code/C/colorcode.cpp:53

static bool unknown(float u, float v) {  // flow_code/C/flowIO.cpp:35-39
  return (fabs(u) > 1e9) || (fabs(v) > 1e9) || std::isnan(u) || std::isnan(v);
}

extern "C" void color_ref_pixel(float fx, float fy, uchar* pix) { computeColor(fx, fy, pix); }

extern "C" float color_ref_image(const float* flow, int width, int height, float maxmotion, uchar* out) {
  float maxrad = -1;
  for (int i = 0; i < width * height; ++i) {
    float fx = flow[2 * i], fy = flow[2 * i + 1];
    if (unknown(fx, fy)) continue;
    float rad = sqrt(fx * fx + fy * fy);
    maxrad = maxrad > rad ? maxrad : rad;
  }
  if (maxmotion > 0) maxrad = maxmotion;
  if (maxrad == 0) maxrad = 1;
  for (int i = 0; i < width * height; ++i) {
    float fx = flow[2 * i], fy = flow[2 * i + 1];
    uchar* pix = out + 3 * i;
    if (unknown(fx, fy))
      pix[0] = pix[1] = pix[2] = 0;
    else
      computeColor(fx / maxrad, fy / maxrad, pix);
  }
  return maxrad;
}
