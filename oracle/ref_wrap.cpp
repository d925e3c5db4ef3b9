// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" entry into the *verbatim* reference engine (kroeger/oflow.cpp) for
// oracle/_ref/libdis_ref*.so.  All work happens in the OFC::OFClass constructor
// (kroeger/oflow.h:84-111, oflow.cpp:32-363); this shim only forwards arguments so
// that Python (ctypes) can call it.  The FDF1.0.1 C functions are exported by the
// same library under their own names and need no shim.
#include <iostream>
#include <vector>
#include <cstring>
#include <xmmintrin.h>
#include "oflow.h"

extern "C" int dis_ref_channels() {
#if (SELECTCHANNEL == 3)
  return 3;
#else
  return 1;
#endif
}

extern "C" void dis_ref_ofclass(const float** im_ao, const float** im_ao_dx, const float** im_ao_dy,
                                const float** im_bo, const float** im_bo_dx, const float** im_bo_dy,
                                int imgpadding, float* outflow, const float* initflow, int width,
                                int height, int sc_f, int sc_l, int max_iter, int min_iter,
                                float dp_thresh, float dr_thresh, float res_thresh, int p_samp_s,
                                float patove, int usefbcon, int costfct, int noc, int patnorm,
                                int usetvref, float tv_alpha, float tv_gamma, float tv_delta,
                                int tv_innerit, int tv_solverit, float tv_sor, int verbosity) {
  OFC::OFClass ofc(im_ao, im_ao_dx, im_ao_dy, im_bo, im_bo_dx, im_bo_dy, imgpadding, outflow,
                   initflow, width, height, sc_f, sc_l, max_iter, min_iter, dp_thresh, dr_thresh,
                   res_thresh, p_samp_s, patove, usefbcon != 0, costfct, noc, patnorm,
                   usetvref != 0, tv_alpha, tv_gamma, tv_delta, tv_innerit, tv_solverit, tv_sor,
                   verbosity);
}
