// color_flow -- command line of the reference's visualisation tool (flow_code/C/color_flow.cpp:73-104;
// the pre-built tools/color_flow binary) on top of libdis_b200.so:
//   color_flow [-quiet] in.flo out.png [maxmotion]
// The colour coding runs on the GPU (dis_flow_to_color); same console lines as the reference.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dis_c.h"

int main(int argc, char** argv) {
  int verbose = 1, argn = 1;
  if (argc > 1 && argv[1][0] == '-' && argv[1][1] == 'q') {
    verbose = 0;
    argn++;
  }
  if (!(argn >= argc - 3 && argn <= argc - 2)) {
    fprintf(stderr, "\n  usage: %s [-quiet] in.flo out.png [maxmotion]\n\n", argv[0]);
    return -1;
  }
  const char* flowname = argv[argn++];
  const char* outname = argv[argn++];
  const float maxmotion = argn < argc ? (float)atof(argv[argn++]) : -1;
  int w = 0, h = 0;
  if (dis_read_flo(flowname, nullptr, 0, &w, &h) != DIS_OK) {
    fprintf(stderr, "ReadFlowFile: could not read %s\n", flowname);
    return -1;
  }
  std::vector<float> flow((size_t)w * h * 2);
  if (dis_read_flo(flowname, flow.data(), flow.size(), &w, &h) != DIS_OK) {
    fprintf(stderr, "ReadFlowFile: file %s is too short\n", flowname);
    return -1;
  }
  std::vector<uint8_t> bgr((size_t)w * h * 3);
  float st[5];
  if (dis_flow_to_color(flow.data(), w, h, maxmotion, 0, bgr.data(), st) != DIS_OK) {
    fprintf(stderr, "color_flow: %s\n", dis_last_error(nullptr));
    return -1;
  }
  printf("max motion: %.4f  motion range: u = %.3f .. %.3f;  v = %.3f .. %.3f\n", st[0], st[1], st[2], st[3], st[4]);
  float maxrad = st[0];
  if (maxmotion > 0) maxrad = maxmotion;
  if (maxrad == 0) maxrad = 1;
  if (verbose) fprintf(stderr, "normalizing by %g\n", maxrad);
  if (verbose) fprintf(stderr, "Writing image %s\n", outname);
  if (dis_write_png_bgr(outname, bgr.data(), w, h) != DIS_OK) {
    fprintf(stderr, "color_flow: %s\n", dis_last_error(nullptr));
    return -1;
  }
  return 0;
}
