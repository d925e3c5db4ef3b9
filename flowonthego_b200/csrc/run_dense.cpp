// run_dense -- command line of the reference (kroeger/run_dense.cpp:185-431, usage kroeger/README.md:48-88)
// on top of libdis_b200.so:
//   run_dense img1 img2 out.flo                      operating point 2, coarsest scale chosen automatically
//   run_dense img1 img2 out.flo X                    operating point X = 1..4
//   run_dense img1 img2 out.flo p1 ... p20           all parameters explicit (order of run_dense.cpp:271-291)
// Built twice like the reference (kroeger/CMakeLists.txt: run_OF_INT with SELECTCHANNEL=1, run_OF_RGB with
// SELECTCHANNEL=3): run_dense reads the images as grey, run_dense_rgb as BGR colour.
// Everything between image decode and SaveFlowFile runs on the GPU (dis_run_u8).  Timing lines follow
// the reference's format when verbosity > 0 / > 1.
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dis_c.h"
#include "imgio.h"

#ifndef SELECTCHANNEL
#define SELECTCHANNEL 1
#endif

static double now_ms() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

int main(int argc, char** argv) {
  if (argc < 4 || (argc > 5 && argc < 24)) {
    fprintf(stderr,
            "usage: %s image1 image2 out.flo [X | lv_f lv_l maxiter miniter mindprate mindrrate minimgerr patchsz poverl "
            "usefbcon patnorm costfct usetvref tv_alpha tv_gamma tv_delta tv_innerit tv_solverit tv_sor verbosity]\n",
            argv[0]);
    return 2;
  }
  double t0 = now_ms();
  GrayImage a, b;
  std::string err = read_image(argv[1], SELECTCHANNEL, &a);
  if (err.empty()) err = read_image(argv[2], SELECTCHANNEL, &b);
  if (!err.empty()) {
    fprintf(stderr, "run_dense: %s\n", err.c_str());
    return 1;
  }
  if (a.w != b.w || a.h != b.h) {
    fprintf(stderr, "run_dense: image sizes differ (%dx%d vs %dx%d)\n", a.w, a.h, b.w, b.h);
    return 1;
  }
  dis_params p;
  if (argc <= 5) {
    dis_params_preset(&p, argc == 5 ? atoi(argv[4]) : 2, a.w);
  } else if (dis_params_from_argv(&p, argc - 4, argv + 4) != DIS_OK) {
    fprintf(stderr, "run_dense: need 20 parameters\n");
    return 2;
  }
  if (p.verbosity > 1) printf("TIME (Image loading     ) (ms): %3g\n", now_ms() - t0);

  dis_handle* h = nullptr;
  if (dis_create_c(&p, SELECTCHANNEL, a.w, a.h, 0, &h) != DIS_OK) {
    fprintf(stderr, "run_dense: %s\n", dis_last_error(nullptr));
    return 1;
  }
  if (p.verbosity > 1) dis_enable_stage_timing(h, 1);
  std::vector<float> flow((size_t)a.w * a.h * 2);
  if (dis_run_u8(h, a.px.data(), b.px.data(), a.w, a.h, a.w * SELECTCHANNEL, flow.data()) != DIS_OK) {
    fprintf(stderr, "run_dense: %s\n", dis_last_error(h));
    dis_destroy(h);
    return 1;
  }
  if (p.verbosity > 1) {
    dis_timings tm;
    dis_get_timings(h, &tm);
    printf("TIME (Pyramide+Gradients) (ms): %3g\n", tm.pyramid_ms);
  }
  dis_destroy(h);
  t0 = now_ms();
  if (dis_write_flo(argv[3], flow.data(), a.w, a.h) != DIS_OK) {
    fprintf(stderr, "run_dense: could not write %s\n", argv[3]);
    return 1;
  }
  if (p.verbosity > 1) printf("TIME (Saving flow file  ) (ms): %3g\n", now_ms() - t0);
  return 0;
}
