// oflow.h -- header-only drop-in for the reference's engine class.
//
// `OFC::OFClass` with exactly the constructor of the reference (kroeger/oflow.h:84-111; the
// definition names argument 19 p_samp_s_in, kroeger/oflow.cpp:46).  Like the reference, all the
// work happens inside the constructor and the result is written into `outflow`; the object keeps
// nothing.  The work itself runs on the B200 through the C-ABI of dis_c.h (dis_run_pyramids).
// Differences a caller can observe:
//   * errors do not pass silently: OFClass::status() / OFClass::error() report them (the reference
//     has no error channel at all);
//   * noc is a run-time argument here (1 = grey, 3 = interleaved BGR) whereas the reference fixes it at
//     compile time (SELECTCHANNEL, run_OF_INT / run_OF_RGB); imgpadding must equal the patch size, as in
//     the reference's only caller (kroeger/run_dense.cpp:391-400).
// Link with -ldis_b200.
#ifndef OFC_HEADER
#define OFC_HEADER

#include <string>

#include "dis_c.h"

namespace OFC {

class OFClass {
 public:
  OFClass(const float** im_ao_in, const float** im_ao_dx_in, const float** im_ao_dy_in,
          const float** im_bo_in, const float** im_bo_dx_in, const float** im_bo_dy_in,
          const int imgpadding_in,
          float* outflow,         // (width/2^sc_l) x (height/2^sc_l) x 2
          const float* initflow,  // optional, resolution of scale sc_f+1
          const int width_in, const int height_in, const int sc_f_in, const int sc_l_in,
          const int max_iter_in, const int min_iter_in, const float dp_thresh_in, const float dr_thresh_in,
          const float res_thresh_in, const int p_samp_s_in, const float patove_in, const bool usefbcon_in,
          const int costfct_in, const int noc_in, const int patnorm_in, const bool usetvref_in,
          const float tv_alpha_in, const float tv_gamma_in, const float tv_delta_in, const int tv_innerit_in,
          const int tv_solverit_in, const float tv_sor_in, const int verbosity_in, const int device = 0)
      : status_(DIS_OK) {
    if (noc_in != 1 && noc_in != 3) {
      status_ = DIS_ERR_UNSUPPORTED;
      error_ = "noc must be 1 (grey) or 3 (interleaved BGR)";
      return;
    }
    dis_params p;
    p.lv_f = sc_f_in;
    p.lv_l = sc_l_in;
    p.maxiter = max_iter_in;
    p.miniter = min_iter_in;
    p.mindprate = dp_thresh_in;
    p.mindrrate = dr_thresh_in;
    p.minimgerr = res_thresh_in;
    p.patchsz = p_samp_s_in;
    p.poverl = patove_in;
    p.usefbcon = usefbcon_in ? 1 : 0;
    p.patnorm = patnorm_in;
    p.costfct = costfct_in;
    p.usetvref = usetvref_in ? 1 : 0;
    p.tv_alpha = tv_alpha_in;
    p.tv_gamma = tv_gamma_in;
    p.tv_delta = tv_delta_in;
    p.tv_innerit = tv_innerit_in;
    p.tv_solverit = tv_solverit_in;
    p.tv_sor = tv_sor_in;
    p.verbosity = verbosity_in;
    dis_handle* h = nullptr;
    status_ = dis_create_c(&p, noc_in, width_in, height_in, device, &h);
    if (status_ != DIS_OK) {
      error_ = dis_last_error(nullptr);
      return;
    }
    status_ = dis_run_pyramids(h, im_ao_in, im_ao_dx_in, im_ao_dy_in, im_bo_in, im_bo_dx_in, im_bo_dy_in,
                               imgpadding_in, width_in, height_in, initflow, outflow);
    if (status_ != DIS_OK) error_ = dis_last_error(h);
    dis_destroy(h);
  }
  int status() const { return status_; }
  const std::string& error() const { return error_; }

 private:
  int status_;
  std::string error_;
};

}  // namespace OFC

#endif /* OFC_HEADER */
