// Test program for the header-only drop-in include/oflow.h: the call the reference makes at
// kroeger/run_dense.cpp:391-400, with pyramids read from a file written by the Python test.
//   ofclass_demo pyramids.bin out.bin
// pyramids.bin: int32 header {noc, width, height, lv_f, lv_l, patchsz, usefbcon, usetvref}, then for each of the six
// pyramids (a, a_dx, a_dy, b, b_dx, b_dy) and each level 0..lv_f the padded float image.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "oflow.h"

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hd[8];
  if (fread(hd, sizeof(int), 8, f) != 8) return 2;
  const int noc = hd[0], width = hd[1], height = hd[2], lv_f = hd[3], lv_l = hd[4], patchsz = hd[5];
  std::vector<std::vector<float>> store(6 * (lv_f + 1));
  const float* pyr[6][16];
  for (int k = 0; k < 6; ++k)
    for (int l = 0; l <= lv_f; ++l) {
      const size_t n = (size_t)((width >> l) + 2 * patchsz) * ((height >> l) + 2 * patchsz) * noc;
      std::vector<float>& v = store[k * (lv_f + 1) + l];
      v.resize(n);
      if (fread(v.data(), sizeof(float), n, f) != n) return 2;
      pyr[k][l] = v.data();
    }
  fclose(f);
  const int sc = 1 << lv_l;
  std::vector<float> out((size_t)(width / sc) * (height / sc) * 2, 0.0f);
  // operating point 2 of the reference (run_dense.cpp:246-250) except for the values taken from the header
  OFC::OFClass ofc(pyr[0], pyr[1], pyr[2], pyr[3], pyr[4], pyr[5], patchsz, out.data(), nullptr, width, height, lv_f, lv_l,
                   12, 12, 0.05f, 0.95f, 0.0f, patchsz, 0.4f, hd[6] != 0, 0, noc, 1, hd[7] != 0, 10.0f, 10.0f, 5.0f, 1, 3,
                   1.6f, 0);
  if (ofc.status() != 0) {
    fprintf(stderr, "OFClass failed (%d): %s\n", ofc.status(), ofc.error().c_str());
    return 1;
  }
  f = fopen(argv[2], "wb");
  if (!f || fwrite(out.data(), sizeof(float), out.size(), f) != out.size()) return 2;
  fclose(f);
  return 0;
}
