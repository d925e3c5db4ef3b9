// flow_epe -- endpoint error between two .flo files (mean / max |dflow| in px with
// a border margin excluded), computed on the GPU (dis_flow_epe):
//   flow_epe a.flo b.flo [margin]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dis_c.h"

static bool load(const char* name, std::vector<float>* f, int* w, int* h) {
  if (dis_read_flo(name, nullptr, 0, w, h) != DIS_OK) return false;
  f->resize((size_t)*w * *h * 2);
  return dis_read_flo(name, f->data(), f->size(), w, h) == DIS_OK;
}

int main(int argc, char** argv) {
  if (argc < 3 || argc > 4) {
    fprintf(stderr, "usage: %s a.flo b.flo [margin]\n", argv[0]);
    return 2;
  }
  std::vector<float> a, b;
  int wa, ha, wb, hb;
  if (!load(argv[1], &a, &wa, &ha) || !load(argv[2], &b, &wb, &hb)) {
    fprintf(stderr, "flow_epe: cannot read the flow files\n");
    return 1;
  }
  if (wa != wb || ha != hb) {
    fprintf(stderr, "flow_epe: sizes differ (%dx%d vs %dx%d)\n", wa, ha, wb, hb);
    return 1;
  }
  const int margin = argc == 4 ? atoi(argv[3]) : 0;
  double mean, mx;
  long long cnt;
  if (dis_flow_epe(a.data(), b.data(), wa, ha, margin, 0, &mean, &mx, &cnt) != DIS_OK) {
    fprintf(stderr, "flow_epe: %s\n", dis_last_error(nullptr));
    return 1;
  }
  printf("EPE mean %.9g  max %.9g  pixels %lld  (%dx%d, margin %d)\n", mean, mx, cnt, wa, ha, margin);
  return 0;
}
