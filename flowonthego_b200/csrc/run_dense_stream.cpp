// run_dense_stream -- flow for every consecutive pair of a frame sequence (dis_video_* of dis_c.h):
//   run_dense_stream [-rgb] [-depth N] [-op X | -params p1 ... p20] out_prefix frame0 frame1 [frame2 ...]
// writes out_prefix0001.flo (frame0 -> frame1), out_prefix0002.flo, ...  Parameters as in run_dense
// (kroeger/run_dense.cpp:241-291): -op X selects operating point X (default 2), -params the 20 explicit values.
// Each .flo is byte-identical to `run_dense frameK frameK+1 out.flo ...` on that pair; the pairs are pipelined
// on the GPU (N in flight, default 8) and every frame is decoded and uploaded once.
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dis_c.h"
#include "imgio.h"

static double now_ms() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

int main(int argc, char** argv) {
  int channels = 1, depth = 8, op = 2, a = 1;
  char** explicit_params = nullptr;
  for (; a < argc && argv[a][0] == '-'; ++a) {
    if (!strcmp(argv[a], "-rgb")) {
      channels = 3;
    } else if (!strcmp(argv[a], "-depth") && a + 1 < argc) {
      depth = atoi(argv[++a]);
    } else if (!strcmp(argv[a], "-op") && a + 1 < argc) {
      op = atoi(argv[++a]);
    } else if (!strcmp(argv[a], "-params") && a + 20 < argc) {
      explicit_params = argv + a + 1;
      a += 20;
    } else {
      break;
    }
  }
  if (argc - a < 3 || depth < 1) {
    fprintf(stderr, "usage: %s [-rgb] [-depth N] [-op X | -params p1 ... p20] out_prefix frame0 frame1 [frame2 ...]\n", argv[0]);
    return 2;
  }
  const std::string prefix = argv[a++];
  const int nframes = argc - a;
  const double t0 = now_ms();

  GrayImage first;
  std::string err = read_image(argv[a], channels, &first);
  if (!err.empty()) {
    fprintf(stderr, "run_dense_stream: %s\n", err.c_str());
    return 1;
  }
  dis_params p;
  if (explicit_params) {
    if (dis_params_from_argv(&p, 20, explicit_params) != DIS_OK) return 2;
  } else {
    dis_params_preset(&p, op, first.w);
  }
  const int verbosity = p.verbosity;
  p.verbosity = 0;  // per-pair TIME lines would interleave; a summary is printed instead
  dis_video* v = nullptr;
  if (dis_video_create(&p, channels, first.w, first.h, 0, depth, &v) != DIS_OK ||
      dis_video_set_output(v, DIS_VIDEO_OUT_FULL) != DIS_OK) {  // .flo files hold the full-resolution field
    fprintf(stderr, "run_dense_stream: %s\n", dis_last_error(nullptr));
    return 1;
  }
  const size_t fbytes = (size_t)first.w * first.h * channels, oflo = (size_t)first.w * first.h * 2;
  // pinned staging: depth+1 frames (a frame must outlive its upload), depth flow fields
  std::vector<uint8_t*> hf(depth + 1);
  std::vector<float*> ho(depth);
  for (auto& q : hf)
    if (dis_host_alloc((void**)&q, fbytes) != DIS_OK) return 1;
  for (auto& q : ho)
    if (dis_host_alloc((void**)&q, oflo * sizeof(float)) != DIS_OK) return 1;

  int written = 0, rc = 0;
  auto pop_and_write = [&]() {
    float* done = nullptr;
    if (dis_video_pop(v, &done) != DIS_OK) {
      fprintf(stderr, "run_dense_stream: %s\n", dis_last_error(nullptr));
      return false;
    }
    char name[32];
    snprintf(name, sizeof name, "%04d.flo", ++written);
    if (dis_write_flo((prefix + name).c_str(), done, first.w, first.h) != DIS_OK) {
      fprintf(stderr, "run_dense_stream: could not write %s%s\n", prefix.c_str(), name);
      return false;
    }
    return true;
  };
  for (int f = 0; f < nframes && rc == 0; ++f) {
    GrayImage img;
    if (f == 0) {
      img = first;
    } else if (!(err = read_image(argv[a + f], channels, &img)).empty()) {
      fprintf(stderr, "run_dense_stream: %s\n", err.c_str());
      rc = 1;
      break;
    }
    if (img.w != first.w || img.h != first.h) {
      fprintf(stderr, "run_dense_stream: %s is %dx%d, expected %dx%d\n", argv[a + f], img.w, img.h, first.w, first.h);
      rc = 1;
      break;
    }
    if (dis_video_pending(v) >= depth && !pop_and_write()) rc = 1;  // frees a flow buffer and a frame buffer
    uint8_t* stage = hf[f % (depth + 1)];
    memcpy(stage, img.px.data(), fbytes);
    if (rc == 0 && dis_video_push(v, stage, (int)(first.w * channels), f ? ho[(f - 1) % depth] : nullptr) != DIS_OK) {
      fprintf(stderr, "run_dense_stream: %s\n", dis_last_error(nullptr));
      rc = 1;
    }
  }
  while (rc == 0 && dis_video_pending(v) > 0)
    if (!pop_and_write()) rc = 1;
  dis_video_destroy(v);
  for (auto q : hf) dis_host_free(q);
  for (auto q : ho) dis_host_free(q);
  if (rc == 0 && verbosity > 0) {
    const double ms = now_ms() - t0;
    printf("TIME (%d pairs, decode + flow + save) (ms): %3g  (%.3g ms/pair)\n", written, ms, ms / (written ? written : 1));
  }
  return rc;
}
