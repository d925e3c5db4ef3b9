// common.cuh -- shared declarations for the sm_100a kernels of libdis_b200.so.
//
// Arithmetic contract: every translation unit is compiled with
// -fmad=false (no FMA contraction), IEEE div/sqrt (nvcc defaults -prec-div/-prec-sqrt=true),
// no flush-to-zero, so that `a*b + c` rounds twice exactly like the reference's -msse4 build.
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include <stdint.h>

struct dis_handle;

namespace dis {

// Geometry of one pyramid level (reference: camparam, kroeger/oflow.h:16-29; grid geometry
// kroeger/patchgrid.cpp:42-49).
struct LevelGeom {
  int lv;          // pyramid level
  int w, h;        // unpadded level size
  int pad;         // image padding (= patch size)
  int pitch;       // row pitch of the padded images, in floats (multiple of 32)
  int tw, th;      // padded size w+2*pad, h+2*pad
  float lb, ubw, ubh;  // valid region for patch centres (oflow.cpp:147-149)
  int nopw, noph, offw, offh, nop;  // patch grid
  int fpitch;      // pitch (in pixels) of per-level planar/float2 work images (= w)
  int noc;         // channels of the padded images: 1 grey, 3 interleaved BGR (SELECTCHANNEL=3); pitch counts floats
  // Batched handles (dis_create_batch): nb pairs per launch.  Every device buffer of pair b lives at the same
  // offset inside its own copy of the workspace, bstride bytes after pair b-1's, so a kernel serves pair b by
  // adding b * bstride to each pointer it was given (bshift below); the batch index rides on a grid dimension.
  int nb;          // pairs per launch (>= 1)
  size_t bstride;  // bytes between the workspaces of consecutive pairs
};

// pointer of pair b given pair 0's pointer (null stays null)
template <class T>
__host__ __device__ __forceinline__ T* bshift(T* p, size_t off) {
  return p ? reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) + off) : nullptr;
}

// the same for a pointer that is never null (no select: kernels shift a dozen pointers per thread)
template <class T>
__host__ __device__ __forceinline__ T* bshift_nn(T* p, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) + off);
}

// Parameters derived in OFClass::OFClass (kroeger/oflow.cpp:75-108)
struct OptParams {
  int p, novals, steps, max_iter, min_iter, patnorm, costfct, noc;
  float outlierthresh, dp_thresh, dr_thresh, res_thresh;
  float outlier_sq;  // largest x with sqrtf(x) <= outlierthresh: sqrtf(x) > thresh  <=>  x > outlier_sq, exactly
};

struct VarParams {  // kroeger/refine_variational.cpp:28-42
  float qa, hg, hd, omega;
  int n_inner, n_solver;
  int sor_group;  // 8 or 16: k_sor_wavefront instantiation (DIS_OPT_SOR_GROUP)
  int sor_small;  // levels of at most this many 32-row blocks take the one-CTA k_sor_small (DIS_OPT_SOR_SMALL)
  int sor_full;   // 1: one CTA per (sweep, row block) item -- lowest latency for a lone pair; 0: only as many CTAs as
                  // items are busy at a time, each warp taking ticket after ticket (best pairs/s)
};

// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while its predecessor in the
// stream (or captured graph) is still draining; it must call pdl_wait() before it touches anything the predecessor
// wrote -- here: as its first statement, so only the launch latency and the block scheduling overlap.  Every kernel of
// the chain calls pdl_wait(), which makes completion transitive (a kernel cannot finish before its predecessor has).
// (No early griddepcontrol.launch_dependents: measured, it makes dependents resident that only wait -- lone pair 1.60
// -> 1.67 ms, -2 % pairs/s, -27 % through the video front end.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// Optional per-kernel profiling hook (CUDA events around every launch; used by bench.py's roofline
// leg through dis_profile_kernels, never on the timed path).
struct Prof {
  virtual void begin(const char* kernel, int level, double alg_bytes) = 0;
  virtual void end() = 0;
};
struct ProfScope {
  Prof* p;
  ProfScope(Prof* p_, const char* k, int level, double bytes) : p(p_) { if (p) p->begin(k, level, bytes); }
  ~ProfScope() { if (p) p->end(); }
};

// Per-run pointers live in a small device-side mailbox so that the recorded CUDA graph does not
// depend on the caller's buffers: a one-thread kernel refreshes it before every graph launch.
struct Mailbox {
  const uint8_t* a;
  const uint8_t* b;
  float2* out;
  float2* lvl_out;  // optional: the finish kernel also copies the level-lv_l flow (the engine's own output) here
  int pitch;  // row pitch of a and b in bytes
};
constexpr int kMaxBatch = 8;
struct MailboxBatch {  // by-value kernel argument: the per-run pointers of up to kMaxBatch pairs
  const uint8_t* a[kMaxBatch];
  const uint8_t* b[kMaxBatch];
  float2* out[kMaxBatch];
  float2* lvl[kMaxBatch];
};
void launch_set_mailboxes(Mailbox* mb0, size_t bstride, int nb, const MailboxBatch& m, int pitch, cudaStream_t st);

// ---- launchers (defined in the .cu files) --------------------------------------------------
// pyramid.cu
void launch_level0(const Mailbox* mb, int w_org, int h_org,
                   int left, int top, const LevelGeom& g, float* Ia, float* Iax, float* Iay,
                   float* Ib, float* Ibx, float* Iby, cudaStream_t st, int only_b = 0);
void launch_first_level(const Mailbox* mb, int L, int w_org, int h_org, int left, int top, const LevelGeom& g,
                        float* bm_a, float* bm_b, float* Ia, float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby,
                        cudaStream_t st, int only_b = 0);
void launch_downsample(const LevelGeom& gf, const LevelGeom& gc, const float* Ia_f, const float* Ib_f,
                       float* Ia, float* Iax, float* Iay, float* Ib, float* Ibx, float* Iby,
                       cudaStream_t st, int only_b = 0);
// patch_search.cu
struct PatchSearchArgs {
  const float *I0, *I0x, *I0y, *I1;
  LevelGeom g;
  OptParams o;
  const float2* flow_coarse;  // level l+1 dense flow (or initflow); nullptr -> zero init
  float2* pflow;              // nop
  float* pweight;             // nop * novals
};
int launch_patch_search(const PatchSearchArgs& a, cudaStream_t st);
void patch_search_init_device();
void varref_init_device();
// tolerance mode (DIS_OPT_ARITH = 1): the same sources compiled with FMA contraction allowed
int launch_patch_search_fast(const PatchSearchArgs& a, cudaStream_t st);
void patch_search_init_device_fast();
void varref_init_device_fast();
// densify.cu
struct DensifyArgs {
  LevelGeom g;
  OptParams o;
  const float2* pflow;
  const float* pweight;
  const float2* pflow_bw;   // complementary grid (forward-backward merge) or nullptr
  const float* pweight_bw;
  float2* flow;             // w*h
  int2* anchor;             // forward-backward merge scratch: per-patch anchor,
  float4* wbil;             //   bilinear weights,
  int* maxdisp;             //   level-wide maximum anchor displacement
  int cover;                // ceil(p / steps): patches covering a pixel per axis
};
void launch_densify(const DensifyArgs& a, cudaStream_t st);
// varref.cu
struct VarRefBuffers {
  float *avg, *Iz, *mask, *Ix, *Iy, *Ixx, *Ixy, *Iyy, *Ixz, *Iyz;  // planar, w*h
  // the ten arrays above are one allocation: array k (order avg, Iz, Ix, Iy, Ixx, Ixy, Iyy, Ixz, Iyz, mask) starts at
  // stack + k * astride floats, so that a kernel needs one shifted base pointer instead of ten (k_assemble)
  float* stack;
  unsigned astride;
  float4 *coefA, *coefB;  // wavefront-major {a11,a12,a22,horiz} (inverted 2x2 blocks) and {b1,b2,vert,-}
  float4* du4;            // wavefront-major records {du, dv, tag, -}
  int* progress;          // SOR wavefront flags: counters, ticket, epoch
};
int launch_varref(const LevelGeom& g, const VarParams& v, const float* I0, const float* I1,
                  float2* flow, const VarRefBuffers& b, cudaStream_t st, Prof* prof = nullptr);
int launch_varref_fast(const LevelGeom& g, const VarParams& v, const float* I0, const float* I1,
                       float2* flow, const VarRefBuffers& b, cudaStream_t st, Prof* prof = nullptr);
void varref_sizes(int w, int h, int n_solver, size_t* n_coef4, size_t* n_du4, size_t* n_prog);
// engine.cu: error text returned by dis_last_error(NULL) (handle-less entry points)
void set_global_error(const char* fmt, ...);
// engine.cu, for the video front end (stream.cu): pyramid reuse between consecutive pairs of a stream.  A chained
// handle keeps the gradients of its second frame; with `reuse` set, its next run takes the first frame's pyramid from
// `prev`'s second-frame buffers instead of building it (only_b launches), and every run records `pyramid_event` after
// its pyramid kernels (inside the CUDA graph, as an external event-record node).
int engine_chain(dis_handle* h, dis_handle* prev);
void engine_set_reuse(dis_handle* h, bool reuse);
cudaEvent_t engine_pyramid_event(dis_handle* h);
// flowviz.cu
void flowviz_init_device();
void launch_flow_color(const float2* d_flow, int w, int h, float maxmotion, uint8_t* d_bgr, unsigned* d_stats,
                       cudaStream_t st);
void decode_flow_stats(const unsigned* h_stats, float out[5]);
int flow_epe_blocks();
void launch_flow_epe(const float2* d_a, const float2* d_b, int w, int h, int margin, double* d_psum, float* d_pmax,
                     unsigned long long* d_pcnt, cudaStream_t st);
// finish.cu
void launch_finish(const float2* flow_l, int wl, int hl, int lv_l, int left, int top, int w_org,
                   int h_org, const Mailbox* mb, int nb, size_t bstride, cudaStream_t st);

}  // namespace dis
