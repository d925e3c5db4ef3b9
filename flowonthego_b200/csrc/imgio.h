// imgio.h -- minimal grey image input for the run_dense CLI: PGM/PPM (binary) and PNG (via zlib).
// The reference reads its inputs with cv::imread(..., CV_LOAD_IMAGE_GRAYSCALE) (kroeger/run_dense.cpp:208-209);
// OpenCV's C++ SDK is not available here, so the decode is done natively and reproduces OpenCV's
// grey conversion (libpng's rgb_to_gray for PNG, cvtColor's fixed-point BGR2GRAY for PPM).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct GrayImage {
  int w = 0, h = 0, ch = 1;
  std::vector<uint8_t> px;  // row-major, pitch == w * ch; ch == 3: interleaved BGR
};

// Returns empty string on success, otherwise the reason.
std::string read_gray_image(const char* path, GrayImage* out);
// channels = 1: as above; channels = 3: BGR as cv::imread(.., CV_LOAD_IMAGE_COLOR) delivers it for the
// reference's colour build (kroeger/run_dense.cpp:203-206): alpha dropped, grey files replicated.
std::string read_image(const char* path, int channels, GrayImage* out);
// 8-bit RGB PNG from interleaved BGR pixels (what flow_code/C/imageLib/ImageIOpng.cpp writes for color_flow).
std::string write_png_bgr(const char* path, const uint8_t* bgr, int w, int h);
