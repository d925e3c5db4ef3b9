// engine.cu -- host side of libdis_b200.so: the C-ABI of include/dis_c.h.
//
// Mirrors the control flow of the reference engine OFC::OFClass::OFClass
// (kroeger/oflow.cpp:32-363): derive the optimisation parameters (:75-108), lay out one
// `camparam` per scale (:138-160), then run the coarse-to-fine loop (:184-337)
//     search (InitializeGrid/SetTargetImage/InitializeFromCoarserOF/Optimize) -> densify -> refine
// with every stage a CUDA kernel on the handle's stream.  The driver part of the reference CLI
// that sits between imread and SaveFlowFile (kroeger/run_dense.cpp:298-414) is covered by
// dis_run_u8 / dis_submit_u8*: padding, pyramid and gradients, engine, upsampling, crop.
//
// All device memory is one slab planned per (image size, parameter set); a whole run is
// recorded once into a CUDA graph and replayed (the coarse levels are launch-latency bound).
// There is no CPU fallback anywhere: if CUDA is unavailable every entry point fails.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/dis_c.h"
#include "common.cuh"
#include "imgio.h"

using namespace dis;

namespace {

thread_local std::string g_create_error;

// Many handles = many streams, each interleaving copies and a graph launch.  With the driver's default of 8
// hardware work queues ("connections"), streams that share a queue pick up false dependencies -- measured on
// B200, C5 workload: 1500 -> 3000 pairs/s through the host-buffer path, 4570 -> 5990 device-resident, for
// 8 -> 32 queues.  The variable is read when the CUDA context is created, so it is set when the library is
// loaded (before main() for a linked host, at dlopen/ctypes load otherwise); an explicit setting wins.
struct ConnectionsDefault {
  ConnectionsDefault() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_default;

// NVTX ranges around the stages (tracing).  Host-side: they bracket the enqueue of a stage, i.e.
// the kernels themselves when a run is launched kernel by kernel (DIS_OPT_USE_GRAPH = 0, taps, profiling) and the
// capture when it is recorded into a graph; a replayed graph shows up as one "dis:run" range.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  NvtxRange(const char* fmt, int v) {
    char buf[48];
    snprintf(buf, sizeof buf, fmt, v);
    nvtxRangePushA(buf);
  }
  ~NvtxRange() { nvtxRangePop(); }
};

struct LevelBufs {
  LevelGeom g{};
  float *Ia = nullptr, *Iax = nullptr, *Iay = nullptr, *Ib = nullptr, *Ibx = nullptr, *Iby = nullptr;
  float2 *flow = nullptr, *flow_bw = nullptr;
};

struct Tap {
  std::vector<float> data;
};

struct KernelProf : Prof {
  cudaStream_t st = nullptr;
  std::vector<dis_kernel_time> recs;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  dis_kernel_time* cur = nullptr;
  void begin(const char* kernel, int level, double bytes) override {
    if (!e0) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
    }
    cur = nullptr;
    for (auto& r : recs)
      if (r.level == level && strcmp(r.name, kernel) == 0) cur = &r;
    if (!cur) {
      dis_kernel_time r{};
      snprintf(r.name, sizeof r.name, "%s", kernel);
      r.level = level;
      recs.push_back(r);
      cur = &recs.back();
    }
    cur->launches++;
    cur->alg_bytes += bytes;
    cudaEventRecord(e0, st);
  }
  void end() override {
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (cur) cur->ms += ms;
  }
  ~KernelProf() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

}  // namespace

struct dis_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  dis_params P{};
  int noc = 1;  // image channels: 1 grey (SELECTCHANNEL=1), 3 interleaved BGR (SELECTCHANNEL=3)
  OptParams opt{};
  int max_w = 0, max_h = 0;
  std::string err;

  // current plan
  int w_org = 0, h_org = 0, w_pad = 0, h_pad = 0, left = 0, top = 0;
  std::vector<LevelBufs> lv;  // index = level
  uint8_t *d_a = nullptr, *d_b = nullptr;
  float2 *pflow = nullptr, *pflow_bw = nullptr;
  float *pweight = nullptr, *pweight_bw = nullptr;
  int2* fb_anchor = nullptr;
  float4* fb_wbil = nullptr;
  int* fb_maxdisp = nullptr;
  VarRefBuffers vb{};
  float2* d_out = nullptr;
  Mailbox* mailbox = nullptr;
  float *bm_a = nullptr, *bm_b = nullptr;  // block-mean scratch of the first processed level (lv_l > 0)
  char* slab = nullptr;
  size_t slab_bytes = 0;

  // graph
  bool use_graph = true;
  int sor_group = 0;  // DIS_OPT_SOR_GROUP: 0 auto, 8, 16
  int sor_small = -1;  // DIS_OPT_SOR_SMALL: row blocks (of 32 rows) up to which a level takes the one-CTA SOR; -1 auto
  bool level_output = false;  // DIS_OPT_LEVEL_OUTPUT
  bool arith_fast = false;    // DIS_OPT_ARITH
  float2* lvl_export[kMaxBatch] = {};  // dis_set_level_export: per pair, device target of the level-lv_l flow (or null)
  int nb = 1;            // pairs per launch (dis_create_batch)
  size_t bstride = 0;    // bytes between the workspaces of consecutive pairs of the batch
  // two recorded variants: [0] both pyramids built, [1] first frame's pyramid reused from `chain_prev` (streams)
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  cudaGraph_t graph[2] = {nullptr, nullptr};          // the captured graphs (own mb_node)
  cudaGraphNode_t mb_node[2] = {nullptr, nullptr};    // root node: k_set_mailboxes, re-parameterised before every launch
  cudaKernelNodeParams mb_params[2] = {};
  int graph_w = 0, graph_h = 0;
  // pyramid reuse in a stream (engine_chain): see common.cuh
  dis_handle* chain_prev = nullptr;
  bool chained = false;      // keeps the gradients of the second frame
  bool reuse_a = false;      // the next run takes frame a's pyramid from chain_prev
  cudaEvent_t pyr_ev = nullptr;

  // timing / taps
  bool stage_timing = false;
  int taps = 0;  // 1: all pyramid levels built the reference's way; 2: the product pyramid path, taps of its levels
  bool kprof_on = false;
  KernelProf kprof;
  std::vector<std::vector<Tap>> tapdata;  // [tap][level]
  dis_timings tm{};
  cudaEvent_t ev[4] = {};  // dis_submit_u8: start, inputs uploaded, compute done, result copied
  bool ev_valid = false;
  int launches = 0;
  bool in_flight = false;
};

namespace {

int fail(dis_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return code;
}

}  // namespace

namespace dis {
int engine_chain(dis_handle* h, dis_handle* prev);
void set_global_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_create_error = buf;
}
}  // namespace dis

namespace {

#define CU(h, call)                                                                           \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(h, DIS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                  \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// kroeger/oflow.cpp:75-108
void derive_opt(const dis_params& q, int noc, OptParams* o) {
  o->noc = noc;
  o->p = q.patchsz;
  o->outlierthresh = (float)o->p / 2;
  {  // IEEE sqrtf is monotonic, so the reference's `norm() > outlierthresh` (patch.cpp:196) is a threshold on the
     // squared norm; find it with the host's correctly rounded sqrtf
    float x = o->outlierthresh * o->outlierthresh;
    while (sqrtf(nextafterf(x, INFINITY)) <= o->outlierthresh) x = nextafterf(x, INFINITY);
    while (sqrtf(x) > o->outlierthresh) x = nextafterf(x, 0.0f);
    o->outlier_sq = x;
  }
  o->max_iter = q.maxiter;
  o->min_iter = q.miniter;
  o->dp_thresh = q.mindprate * q.mindprate;
  o->dr_thresh = q.mindrrate;
  o->res_thresh = q.minimgerr;
  o->steps = std::max(1, (int)floor(o->p * (1 - q.poverl)));
  o->novals = noc * o->p * o->p;  // oflow.cpp:92
  o->patnorm = q.patnorm;
  o->costfct = q.costfct;
}

// kroeger/oflow.cpp:138-160 (camparam) and kroeger/patchgrid.cpp:42-49 (grid geometry)
void derive_level(const OptParams& o, int width, int height, int pad, int sl, LevelGeom* g) {
  g->noc = o.noc;
  const float sc_fct = (float)pow(2, -sl);
  g->lv = sl;
  g->h = (int)(height * sc_fct);
  g->w = (int)(width * sc_fct);
  g->pad = pad;
  g->tw = g->w + 2 * pad;
  g->th = g->h + 2 * pad;
  g->pitch = (int)align_up((size_t)g->tw * o.noc, 32);
  g->lb = -(float)o.p / 2;
  g->ubw = (float)(g->w + o.p / 2 - 2);
  g->ubh = (float)(g->h + o.p / 2 - 2);
  g->nopw = (int)ceil((float)g->w / (float)o.steps);
  g->noph = (int)ceil((float)g->h / (float)o.steps);
  g->offw = (int)floor((g->w - (g->nopw - 1) * o.steps) / 2);
  g->offh = (int)floor((g->h - (g->noph - 1) * o.steps) / 2);
  g->nop = g->nopw * g->noph;
  g->fpitch = g->w;
}

struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// Lays out every buffer of a run for input size (w,h); with base == nullptr only sizes it.
size_t carve(dis_handle* h, int w, int h_img, char* base, bool assign) {
  const dis_params& q = h->P;
  int wp, hp, left, top;
  dis_padded_size(w, h_img, q.lv_f, &wp, &hp, &left, &top);
  Carver c{base};
  std::vector<LevelBufs> lv(q.lv_f + 1);
  uint8_t* d_a = c.take<uint8_t>((size_t)w * h_img * h->noc);
  uint8_t* d_b = c.take<uint8_t>((size_t)w * h_img * h->noc);
  // the product path builds level lv_l straight from the u8 frames (enqueue_pyramids); the finer levels exist only
  // when every level is built the reference's way (taps == 1, or lv_l > 8 where the block mean is no longer exact)
  const bool fine_levels = h->taps == 1 || q.lv_l > 8;
  for (int l = 0; l <= q.lv_f; ++l) {
    LevelBufs& L = lv[l];
    derive_level(h->opt, wp, hp, q.patchsz, l, &L.g);
    const size_t n = (size_t)L.g.pitch * L.g.th;
    if (l >= q.lv_l || fine_levels) {
      L.Ia = c.take<float>(n);
      L.Ib = c.take<float>(n);
    }
    if (l >= q.lv_l) {
      L.Iax = c.take<float>(n);
      L.Iay = c.take<float>(n);
      if (q.usefbcon || h->chained) {  // chained: this pair's second frame is the next pair's first
        L.Ibx = c.take<float>(n);
        L.Iby = c.take<float>(n);
      }
      L.flow = c.take<float2>((size_t)L.g.w * L.g.h);
      if (q.usefbcon) L.flow_bw = c.take<float2>((size_t)L.g.w * L.g.h);
    }
  }
  const LevelGeom& gf = lv[q.lv_l].g;  // finest processed level = largest buffers
  float* bm_a = c.take<float>((size_t)gf.w * gf.h * h->noc);
  float* bm_b = c.take<float>((size_t)gf.w * gf.h * h->noc);
  float2* pflow = c.take<float2>(gf.nop);
  float* pweight = c.take<float>((size_t)gf.nop * h->opt.novals);
  float2* pflow_bw = nullptr;
  float* pweight_bw = nullptr;
  if (q.usefbcon) {
    pflow_bw = c.take<float2>(gf.nop);
    pweight_bw = c.take<float>((size_t)gf.nop * h->opt.novals);
  }
  int2* fb_anchor = q.usefbcon ? c.take<int2>(gf.nop) : nullptr;
  float4* fb_wbil = q.usefbcon ? c.take<float4>(gf.nop) : nullptr;
  int* fb_maxdisp = q.usefbcon ? c.take<int>(1) : nullptr;
  VarRefBuffers vb{};
  if (q.usetvref) {
    const size_t n = (size_t)gf.w * gf.h;
    float** planes[] = {&vb.avg, &vb.Iz, &vb.Ix, &vb.Iy, &vb.Ixx, &vb.Ixy, &vb.Iyy, &vb.Ixz, &vb.Iyz, &vb.mask};
    vb.astride = (unsigned)align_up(n * h->noc, 64);  // one plane per colour channel (the mask has one)
    vb.stack = c.take<float>((size_t)vb.astride * 10);
    for (int k = 0; k < 10; ++k) *planes[k] = vb.stack ? vb.stack + (size_t)k * vb.astride : nullptr;
    size_t n_coef4, n_du4, n_prog;
    varref_sizes(gf.w, gf.h, std::max(1, q.tv_solverit), &n_coef4, &n_du4, &n_prog);
    vb.coefA = c.take<float4>(n_coef4);
    vb.coefB = c.take<float4>(n_coef4);
    vb.du4 = c.take<float4>(n_du4);
    vb.progress = c.take<int>(n_prog);
  }
  float2* d_out = c.take<float2>((size_t)w * h_img);
  Mailbox* mailbox = c.take<Mailbox>(2);  // [1]: scratch target of the debug launches
  if (assign) {
    h->mailbox = mailbox;
    h->bm_a = bm_a;
    h->bm_b = bm_b;
    h->lv = lv;
    h->d_a = d_a;
    h->d_b = d_b;
    h->pflow = pflow;
    h->pweight = pweight;
    h->pflow_bw = pflow_bw;
    h->pweight_bw = pweight_bw;
    h->fb_anchor = fb_anchor;
    h->fb_wbil = fb_wbil;
    h->fb_maxdisp = fb_maxdisp;
    h->vb = vb;
    h->d_out = d_out;
    h->w_org = w;
    h->h_org = h_img;
    h->w_pad = wp;
    h->h_pad = hp;
    h->left = left;
    h->top = top;
  }
  return align_up(c.off, 256);
}

void drop_graph(dis_handle* h) {
  for (int v = 0; v < 2; ++v) {
    if (h->graph_exec[v]) {
      cudaGraphExecDestroy(h->graph_exec[v]);
      h->graph_exec[v] = nullptr;
    }
    if (h->graph[v]) {
      cudaGraphDestroy(h->graph[v]);
      h->graph[v] = nullptr;
    }
    h->mb_node[v] = nullptr;
  }
}

int plan(dis_handle* h, int w, int h_img) {
  if (w <= 0 || h_img <= 0) return fail(h, DIS_ERR_INVALID_ARG, "non-positive image size %dx%d", w, h_img);
  if (w == h->w_org && h_img == h->h_org && !h->lv.empty()) return DIS_OK;
  {  // validate the geometry before anything is committed: a failed plan must not leave a usable-looking handle
    int wp, hp, left, top;
    dis_padded_size(w, h_img, h->P.lv_f, &wp, &hp, &left, &top);
    LevelGeom gc{};
    derive_level(h->opt, wp, hp, h->P.patchsz, h->P.lv_f, &gc);
    // coarsest level must be large enough for the kernels (and for the reference itself)
    if (gc.w < 2 || gc.h < 4) {
      h->lv.clear();
      h->w_org = h->h_org = 0;
      drop_graph(h);
      return fail(h, DIS_ERR_UNSUPPORTED, "coarsest level %dx%d too small (lv_f=%d)", gc.w, gc.h, h->P.lv_f);
    }
  }
  const size_t one = carve(h, w, h_img, nullptr, false);  // workspace of one pair (multiple of 256 bytes)
  const size_t need = one * (size_t)h->nb;
  if (need > h->slab_bytes) {  // grow the slab (sizes above the create-time maximum re-allocate)
    drop_graph(h);
    if (h->slab) {
      CU(h, cudaStreamSynchronize(h->stream));
      cudaFree(h->slab);
    }
    h->slab = nullptr;
    h->slab_bytes = 0;
    CU(h, cudaMalloc(&h->slab, need));
    h->slab_bytes = need;
  }
  drop_graph(h);
  carve(h, w, h_img, h->slab, true);  // pointers of pair 0; pair b's are bstride * b bytes further (bshift)
  h->bstride = h->nb > 1 ? one : 0;
  for (LevelBufs& L : h->lv) {
    L.g.nb = h->nb;
    L.g.bstride = h->bstride;
  }
  // SOR hand-off tags/epoch start from a clean slate (stale bytes could alias a tag)
  CU(h, cudaMemsetAsync(h->slab, 0, h->slab_bytes, h->stream));
  return DIS_OK;
}

void tap_store(dis_handle* h, int tap, int level, const void* dptr, size_t n_floats) {
  if (!h->taps) return;
  if ((int)h->tapdata.size() <= tap) h->tapdata.resize(tap + 1);
  if ((int)h->tapdata[tap].size() <= level) h->tapdata[tap].resize(level + 1);
  auto& v = h->tapdata[tap][level].data;
  v.resize(n_floats);
  cudaMemcpyAsync(v.data(), dptr, n_floats * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
}

void tap_image(dis_handle* h, int tap, int level, const float* d, const LevelGeom& g) {
  if (!h->taps || d == nullptr) return;
  if ((int)h->tapdata.size() <= tap) h->tapdata.resize(tap + 1);
  if ((int)h->tapdata[tap].size() <= level) h->tapdata[tap].resize(level + 1);
  auto& v = h->tapdata[tap][level].data;
  v.resize((size_t)g.tw * g.th * g.noc);
  cudaMemcpy2DAsync(v.data(), sizeof(float) * g.tw * g.noc, d, sizeof(float) * g.pitch, sizeof(float) * g.tw * g.noc, g.th,
                    cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
}

// stage 1 (kroeger/run_dense.cpp:298-344): pyramids of both frames from the u8 inputs
int enqueue_pyramids(dis_handle* h) {
  NvtxRange nv("dis:pyramids");
  const dis_params& q = h->P;
  // Product path: nothing reads the levels below lv_l (the engine starts at lv_l, oflow.cpp:199), so the first
  // processed level is built straight from the u8 frames (bit-identical, see k_block_mean) and the finer levels are
  // not materialised.  With taps enabled (tests) every level is built the reference's way so that it can be compared.
  const int first = (h->taps == 1 || q.lv_l > 8) ? 0 : q.lv_l;
  const int only_b = h->reuse_a ? 1 : 0;  // frame a's pyramid lives in chain_prev's second-frame buffers
  for (int l = first; l <= q.lv_f; ++l) {
    LevelBufs& L = h->lv[l];
    // SURVEY 8(d) B_P: u8 in (level 0 only), I of both frames out, Ix/Iy of frame a on used levels
    double np_ = (double)L.g.tw * L.g.th;
    if (l == first)
      for (int k = 0; k < first; ++k) np_ += (double)h->lv[k].g.tw * h->lv[k].g.th;  // the model counts all levels
    ProfScope ps(h->kprof_on ? &h->kprof : nullptr, l == first ? "k_pyr_level0" : "k_pyr_down", l,
                 (l == first ? 2.0 * h->w_org * h->h_org : 0.0) + 8.0 * np_ + (l >= q.lv_l ? 8.0 * np_ : 0.0));
    if (l == first && first > 0) {
      launch_first_level(h->mailbox, first, h->w_org, h->h_org, h->left, h->top, L.g, h->bm_a, h->bm_b, L.Ia, L.Iax,
                         L.Iay, L.Ib, L.Ibx, L.Iby, h->stream, only_b);
      h->launches++;
    } else if (l == 0)
      launch_level0(h->mailbox, h->w_org, h->h_org, h->left, h->top, L.g, L.Ia, L.Iax, L.Iay, L.Ib,
                    L.Ibx, L.Iby, h->stream, only_b);
    else
      launch_downsample(h->lv[l - 1].g, L.g, h->lv[l - 1].Ia, h->lv[l - 1].Ib, L.Ia, L.Iax, L.Iay, L.Ib,
                        L.Ibx, L.Iby, h->stream, only_b);
    h->launches++;
  }
  if (h->chained && h->pyr_ev) {  // the next pair of the stream may read this pair's second-frame pyramid from here on
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(h->stream, &cs);
    CU(h, cudaEventRecordWithFlags(h->pyr_ev, h->stream,
                                   cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
  }
  if (h->taps)
    for (int l = first; l <= q.lv_f; ++l) {
      LevelBufs& L = h->lv[l];
      tap_image(h, DIS_TAP_IMG_A, l, L.Ia, L.g);
      tap_image(h, DIS_TAP_IMG_A_DX, l, L.Iax, L.g);
      tap_image(h, DIS_TAP_IMG_A_DY, l, L.Iay, L.g);
      tap_image(h, DIS_TAP_IMG_B, l, L.Ib, L.g);
      tap_image(h, DIS_TAP_IMG_B_DX, l, L.Ibx, L.g);
      tap_image(h, DIS_TAP_IMG_B_DY, l, L.Iby, L.g);
    }
  return DIS_OK;
}

struct StageClock {
  dis_handle* h;
  bool on;
  cudaEvent_t a = nullptr, b = nullptr;
  float* acc;
  StageClock(dis_handle* h_, float* acc_) : h(h_), on(h_->stage_timing), acc(acc_) {
    if (on) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, h->stream);
    }
  }
  float stop() {
    float ms = 0.f;
    if (on) {
      cudaEventRecord(b, h->stream);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
      *acc += ms;
      cudaEventDestroy(a);
      cudaEventDestroy(b);
      on = false;
    }
    return ms;
  }
  ~StageClock() { stop(); }
};

// the coarse-to-fine loop, kroeger/oflow.cpp:184-337
int enqueue_engine(dis_handle* h, const float2* d_initflow) {
  const dis_params& q = h->P;
  VarParams vp{};
  vp.qa = 0.25f * q.tv_alpha;               // refine_variational.cpp:40-42
  vp.hg = q.tv_gamma * 0.5f / 3.0f;
  vp.hd = q.tv_delta * 0.5f / 3.0f;
  vp.omega = q.tv_sor;
  vp.n_solver = q.tv_solverit;
  {  // auto: a fine level that fills the GPU by itself (4K at lv_l = 0) prefers the low-latency instantiation
    const LevelGeom& gfin = h->lv[q.lv_l].g;
    vp.sor_group = h->sor_group ? h->sor_group : ((size_t)gfin.w * gfin.h >= (1u << 20) ? 16 : 8);
    vp.sor_full = h->sor_group == 16;  // asked for explicitly: the latency setting
    vp.sor_small = h->sor_small >= 0 ? h->sor_small : (vp.sor_full ? 4 : 5);  // measured: latency / throughput optimum at 1080p
  }
  Prof* prof = h->kprof_on ? &h->kprof : nullptr;
  // Tolerance mode: FMA-contracted builds of the search and the refinement.  Rounding differences grow with the
  // number of Gauss-Newton iterations (2e-3 px at 16, 0.25 px at 128), so it is refused beyond 32.
  const bool fast = h->arith_fast;
  if (fast && q.maxiter > 32)
    return fail(h, DIS_ERR_UNSUPPORTED, "DIS_OPT_ARITH = 1 (tolerance mode) is limited to maxiter <= 32 (got %d)", q.maxiter);
  for (int sl = q.lv_f; sl >= q.lv_l; --sl) {
    LevelBufs L = h->lv[sl];
    if (h->reuse_a) {  // first frame = second frame of the previous pair of the stream
      const LevelBufs& P = h->chain_prev->lv[sl];
      L.Ia = P.Ib;
      L.Iax = P.Ibx;
      L.Iay = P.Iby;
    }
    float t_search = 0, t_dens = 0, t_var = 0;
    const float2* coarse = nullptr;
    if (sl < q.lv_f)
      coarse = h->lv[sl + 1].flow;
    else if (d_initflow)
      coarse = d_initflow;
    {
      NvtxRange nv("dis:search L%d", sl);
      StageClock ck(h, &h->tm.search_ms);
      PatchSearchArgs pa{L.Ia, L.Iax, L.Iay, L.Ib, L.g, h->opt, coarse, h->pflow, h->pweight};
      // SURVEY 8(d) B_D: 4 padded arrays in, coarse flow in, patch results out
      const double np_ = (double)L.g.tw * L.g.th;
      ProfScope ps(prof, "k_patch_search", sl,
                   16.0 * np_ + (sl < q.lv_f ? 8.0 * (double)(L.g.w / 2) * (L.g.h / 2) : 0.0) + 16.0 * L.g.nop);
      if ((fast ? launch_patch_search_fast : launch_patch_search)(pa, h->stream)) return fail(h, DIS_ERR_UNSUPPORTED, "patch size %d", h->opt.p);
      h->launches++;
      t_search = ck.stop();
    }
    const bool fb = q.usefbcon != 0;
    if (fb) {  // backward grid: roles of the two frames swapped (kroeger/oflow.cpp:193-197, 214-215, 234-235)
      StageClock ck(h, &h->tm.search_ms);
      const float2* coarse_bw = (sl < q.lv_f) ? h->lv[sl + 1].flow_bw : nullptr;
      PatchSearchArgs pb{L.Ib, L.Ibx, L.Iby, L.Ia, L.g, h->opt, coarse_bw, h->pflow_bw, h->pweight_bw};
      const double np_ = (double)L.g.tw * L.g.th;
      ProfScope ps(prof, "k_patch_search", sl,
                   16.0 * np_ + (sl < q.lv_f ? 8.0 * (double)(L.g.w / 2) * (L.g.h / 2) : 0.0) + 16.0 * L.g.nop);
      if ((fast ? launch_patch_search_fast : launch_patch_search)(pb, h->stream)) return fail(h, DIS_ERR_UNSUPPORTED, "patch size %d", h->opt.p);
      h->launches++;
      t_search += ck.stop();
    }
    tap_store(h, DIS_TAP_PATCH_FLOW, sl, h->pflow, (size_t)L.g.nop * 2);
    {
      NvtxRange nv("dis:densify L%d", sl);
      StageClock ck(h, &h->tm.densify_ms);
      DensifyArgs da{L.g, h->opt, h->pflow, h->pweight, fb ? h->pflow_bw : nullptr, fb ? h->pweight_bw : nullptr,
                     L.flow, h->fb_anchor, h->fb_wbil, h->fb_maxdisp, (h->opt.p + h->opt.steps - 1) / h->opt.steps};
      // SURVEY 8(d) B_A: patch results + I0,I1 in, flow out
      ProfScope ps(prof, "k_densify", sl, 16.0 * L.g.nop + 8.0 * (double)L.g.tw * L.g.th + 8.0 * (double)L.g.w * L.g.h);
      launch_densify(da, h->stream);
      h->launches += fb ? 2 : 1;
      if (fb && sl > q.lv_l) {  // backward flow is only needed to seed the next finer scale (oflow.cpp:269-270)
        DensifyArgs db{L.g, h->opt, h->pflow_bw, h->pweight_bw, h->pflow, h->pweight, L.flow_bw,
                       h->fb_anchor, h->fb_wbil, h->fb_maxdisp, (h->opt.p + h->opt.steps - 1) / h->opt.steps};
        launch_densify(db, h->stream);
        h->launches += 2;
      }
      t_dens = ck.stop();
    }
    tap_store(h, DIS_TAP_FLOW_DENSE, sl, L.flow, (size_t)L.g.w * L.g.h * 2);
    if (q.usetvref) {
      NvtxRange nv("dis:refine L%d", sl);
      StageClock ck(h, &h->tm.varref_ms);
      vp.n_inner = q.tv_innerit * (sl + 1);  // refine_variational.cpp:36
      int n = (fast ? launch_varref_fast : launch_varref)(L.g, vp, L.Ia, L.Ib, L.flow, h->vb, h->stream, prof);
      if (n < 0)
        return fail(h, DIS_ERR_UNSUPPORTED, "refinement of level %d (%dx%d, %d sweeps): need width >= 2, height >= 4, 1..256 sweeps",
                    sl, L.g.w, L.g.h, vp.n_solver);
      h->launches += n;
      if (fb && sl > q.lv_l) {  // oflow.cpp:291-294
        n = (fast ? launch_varref_fast : launch_varref)(L.g, vp, L.Ib, L.Ia, L.flow_bw, h->vb, h->stream, prof);
        h->launches += n;
      }
      t_var = ck.stop();
    }
    tap_store(h, DIS_TAP_FLOW_REFINED, sl, L.flow, (size_t)L.g.w * L.g.h * 2);
    if (q.verbosity > 1 && h->stage_timing)  // format of kroeger/oflow.cpp:303
      printf("TIME (Sc: %i, #p:%6i, pconst, pinit, poptim, cflow, tvopt, total): %8.2f %8.2f %8.2f %8.2f %8.2f -> %8.2f ms.\n",
             sl, L.g.nop, 0.0, 0.0, t_search, t_dens, t_var, t_search + t_dens + t_var);
  }
  CU(h, cudaGetLastError());
  return DIS_OK;
}

int enqueue_finish(dis_handle* h) {
  NvtxRange nv("dis:finish");
  const LevelBufs& L = h->lv[h->P.lv_l];
  ProfScope ps(h->kprof_on ? &h->kprof : nullptr, "k_finish", h->P.lv_l,
               8.0 * (double)L.g.w * L.g.h + 8.0 * (double)h->w_org * h->h_org);
  launch_finish(L.flow, L.g.w, L.g.h, h->P.lv_l, h->left, h->top, h->w_org, h->h_org, h->mailbox, h->nb, h->bstride,
                h->stream);
  h->launches++;
  CU(h, cudaGetLastError());
  return DIS_OK;
}

// Enqueue the whole device-side run (stage 1 .. finish); replayed from a CUDA graph when possible.
int enqueue_run_device_batch(dis_handle* h, int n, const uint8_t* const* d_a, const uint8_t* const* d_b, int pitch,
                             float2* const* d_out);

int enqueue_run_device(dis_handle* h, const uint8_t* d_a, const uint8_t* d_b, int pitch, float2* d_out) {
  return enqueue_run_device_batch(h, 1, &d_a, &d_b, pitch, &d_out);
}

// n <= nb pairs; the unused slots of a batched handle recompute pair 0 into their own scratch output
int enqueue_run_device_batch(dis_handle* h, int n, const uint8_t* const* d_a, const uint8_t* const* d_b, int pitch,
                             float2* const* d_out) {
  NvtxRange nv("dis:run");
  // Per-run pointers go to the device mailbox through a one-block kernel.  On the graph path that kernel is the
  // graph's ROOT node and its arguments are patched before every launch (cudaGraphExecKernelNodeSetParams): a
  // separate launch in front of (or behind) each graph launch costs ~70 us of throughput per pair-run when many
  // handles share the GPU (measured: bench.py device arm, 16.9 -> 15.7 ms per 128 pairs for one launch less).
  MailboxBatch m{};
  for (int b = 0; b < h->nb; ++b) {
    m.a[b] = d_a[b < n ? b : 0];
    m.b[b] = d_b[b < n ? b : 0];
    m.out[b] = b < n ? d_out[b] : bshift(h->d_out, (size_t)b * h->bstride);
    m.lvl[b] = b < n ? h->lvl_export[b] : nullptr;
  }
  auto set_mailbox_now = [&]() { launch_set_mailboxes(h->mailbox, h->bstride, h->nb, m, pitch, h->stream); };
  const bool graphable = h->use_graph && !h->taps && !h->stage_timing && !h->kprof_on;
  const int gv = h->reuse_a ? 1 : 0;
  if (graphable && (h->graph_w != h->w_org || h->graph_h != h->h_org)) drop_graph(h);
  if (graphable && h->graph_exec[gv]) {
    Mailbox* mb0 = h->mailbox;
    size_t bs = h->bstride;
    int nb = h->nb, pt = pitch;
    void* args[] = {&mb0, &bs, &nb, &m, &pt};
    cudaKernelNodeParams kp = h->mb_params[gv];
    kp.kernelParams = args;
    kp.extra = nullptr;
    CU(h, cudaGraphExecKernelNodeSetParams(h->graph_exec[gv], h->mb_node[gv], &kp));
    CU(h, cudaGraphLaunch(h->graph_exec[gv], h->stream));
    return DIS_OK;
  }
  if (graphable) {
    h->launches = 1;
    CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    set_mailbox_now();  // captured: the root node
    int rc = enqueue_pyramids(h);
    if (rc == DIS_OK) rc = enqueue_engine(h, nullptr);
    if (rc == DIS_OK) rc = enqueue_finish(h);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (rc != DIS_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) return fail(h, DIS_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    size_t n_roots = 1;
    cudaGraphNode_t root = nullptr;
    cudaGraphNodeType ty = cudaGraphNodeTypeEmpty;
    if ((e = cudaGraphGetRootNodes(g, &root, &n_roots)) != cudaSuccess || n_roots != 1 ||
        (e = cudaGraphNodeGetType(root, &ty)) != cudaSuccess || ty != cudaGraphNodeTypeKernel ||
        (e = cudaGraphKernelNodeGetParams(root, &h->mb_params[gv])) != cudaSuccess) {
      cudaGraphDestroy(g);
      return fail(h, DIS_ERR_CUDA, "graph capture: mailbox root node not found (%s)", cudaGetErrorString(e));
    }
    h->mb_node[gv] = root;
    e = cudaGraphInstantiate(&h->graph_exec[gv], g, 0);
    if (e != cudaSuccess) {
      cudaGraphDestroy(g);
      return fail(h, DIS_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    }
    h->graph[gv] = g;  // kept: mb_node is a node of this graph
    h->graph_w = h->w_org;
    h->graph_h = h->h_org;
    h->tm.launches = h->launches;
    CU(h, cudaGraphLaunch(h->graph_exec[gv], h->stream));  // the captured arguments are this run's
    return DIS_OK;
  }
  set_mailbox_now();
  h->launches = 1;
  int rc;
  {
    StageClock ck(h, &h->tm.pyramid_ms);
    rc = enqueue_pyramids(h);
  }
  if (rc == DIS_OK) rc = enqueue_engine(h, nullptr);
  if (rc == DIS_OK) {
    StageClock ck(h, &h->tm.finish_ms);
    rc = enqueue_finish(h);
  }
  h->tm.launches = h->launches;
  return rc;
}

void reset_timings(dis_handle* h) {
  const int l = h->tm.launches;
  h->tm = dis_timings{};
  h->tm.launches = l;
}

}  // namespace

// ---- pyramid reuse between the consecutive pairs of a stream (used by stream.cu) -------------------------------
namespace dis {
int engine_chain(dis_handle* h, dis_handle* prev) {
  if (!h || !prev || h == prev || h->nb != 1 || prev->nb != 1 || h->device != prev->device) return DIS_ERR_INVALID_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  if (!h->pyr_ev) CU(h, cudaEventCreateWithFlags(&h->pyr_ev, cudaEventDisableTiming));
  h->chain_prev = prev;
  if (!h->chained) {  // second-frame gradients join the workspace: re-plan
    h->chained = true;
    const int w = h->w_org, hh = h->h_org;
    h->lv.clear();
    h->w_org = h->h_org = 0;
    return plan(h, w, hh);
  }
  return DIS_OK;
}
void engine_set_reuse(dis_handle* h, bool reuse) { h->reuse_a = reuse && h->chain_prev != nullptr; }
cudaEvent_t engine_pyramid_event(dis_handle* h) { return h->pyr_ev; }
}  // namespace dis

// ================================================================================ C-ABI
extern "C" {

const char* dis_version(void) { return "dis_b200 0.1 sm_100a"; }

int dis_auto_first_scale(int imgwidth, int fratio, int patchsize) {
  return std::max(0, (int)std::floor(log2((2.0f * (float)imgwidth) / ((float)fratio * (float)patchsize))));
}

int dis_params_preset(dis_params* p, int preset, int width_org) {
  if (!p || width_org <= 0) return DIS_ERR_INVALID_ARG;
  // defaults, kroeger/run_dense.cpp:227-233
  p->mindprate = 0.05f;
  p->mindrrate = 0.95f;
  p->minimgerr = 0.0f;
  p->usefbcon = 0;
  p->patnorm = 1;
  p->costfct = 0;
  p->tv_alpha = 10.0f;
  p->tv_gamma = 10.0f;
  p->tv_delta = 5.0f;
  p->tv_innerit = 1;
  p->tv_solverit = 3;
  p->tv_sor = 1.6f;
  p->verbosity = 2;
  const int fratio = 5;
  switch (preset) {  // run_dense.cpp:239-267
    case 1:
      p->patchsz = 8; p->poverl = 0.3f;
      p->lv_f = dis_auto_first_scale(width_org, fratio, p->patchsz);
      p->lv_l = std::max(p->lv_f - 2, 0); p->maxiter = 16; p->miniter = 16; p->usetvref = 0;
      break;
    case 3:
      p->patchsz = 12; p->poverl = 0.75f;
      p->lv_f = dis_auto_first_scale(width_org, fratio, p->patchsz);
      p->lv_l = std::max(p->lv_f - 4, 0); p->maxiter = 16; p->miniter = 16; p->usetvref = 1;
      break;
    case 4:
      p->patchsz = 12; p->poverl = 0.75f;
      p->lv_f = dis_auto_first_scale(width_org, fratio, p->patchsz);
      p->lv_l = std::max(p->lv_f - 5, 0); p->maxiter = 128; p->miniter = 128; p->usetvref = 1;
      break;
    case 2:
    default:
      p->patchsz = 8; p->poverl = 0.4f;
      p->lv_f = dis_auto_first_scale(width_org, fratio, p->patchsz);
      p->lv_l = std::max(p->lv_f - 2, 0); p->maxiter = 12; p->miniter = 12; p->usetvref = 1;
      break;
  }
  return DIS_OK;
}

int dis_params_from_argv(dis_params* p, int n, const char* const* a) {
  if (!p || !a || n < 20) return DIS_ERR_INVALID_ARG;
  int k = 0;  // order of kroeger/run_dense.cpp:271-291
  p->lv_f = atoi(a[k++]);
  p->lv_l = atoi(a[k++]);
  p->maxiter = atoi(a[k++]);
  p->miniter = atoi(a[k++]);
  p->mindprate = (float)atof(a[k++]);
  p->mindrrate = (float)atof(a[k++]);
  p->minimgerr = (float)atof(a[k++]);
  p->patchsz = atoi(a[k++]);
  p->poverl = (float)atof(a[k++]);
  p->usefbcon = atoi(a[k++]) != 0;
  p->patnorm = atoi(a[k++]);
  p->costfct = atoi(a[k++]);
  p->usetvref = atoi(a[k++]) != 0;
  p->tv_alpha = (float)atof(a[k++]);
  p->tv_gamma = (float)atof(a[k++]);
  p->tv_delta = (float)atof(a[k++]);
  p->tv_innerit = atoi(a[k++]);
  p->tv_solverit = atoi(a[k++]);
  p->tv_sor = (float)atof(a[k++]);
  p->verbosity = atoi(a[k++]);
  return DIS_OK;
}

int dis_params_validate(const dis_params* p, char* why, size_t why_len) {
  const char* msg = nullptr;
  if (!p) msg = "null params";
  else if (p->lv_l < 0 || p->lv_f < p->lv_l || p->lv_f > 14) msg = "need 0 <= lv_l <= lv_f <= 14";
  else if (p->patchsz < 4 || p->patchsz > 16 || (p->patchsz & 1)) msg = "patchsz must be even and in 4..16";
  else if (!(p->poverl >= 0.0f && p->poverl < 1.0f)) msg = "poverl must be in [0,1)";
  else if (p->costfct < 0 || p->costfct > 2) msg = "costfct must be 0 (L2), 1 (L1) or 2 (Huber)";
  else if (p->maxiter < 0 || p->miniter < 0) msg = "negative iteration count";
  else if (p->usetvref && (p->tv_solverit < 1 || p->tv_solverit > 256)) msg = "tv_solverit must be in 1..256";
  else if (p->usetvref && p->tv_innerit < 0) msg = "tv_innerit must be >= 0";
  if (msg) {
    if (why && why_len) snprintf(why, why_len, "%s", msg);
    return DIS_ERR_INVALID_ARG;
  }
  if (why && why_len) why[0] = 0;
  return DIS_OK;
}

int dis_padded_size(int w, int h, int lv_f, int* w_pad, int* h_pad, int* left, int* top) {
  if (w <= 0 || h <= 0 || lv_f < 0 || lv_f > 14) return DIS_ERR_INVALID_ARG;
  const int sc = 1 << lv_f;  // kroeger/run_dense.cpp:298-311
  int padw = 0, padh = 0;
  if (w % sc) padw = sc - w % sc;
  if (h % sc) padh = sc - h % sc;
  if (w_pad) *w_pad = w + padw;
  if (h_pad) *h_pad = h + padh;
  if (left) *left = (int)floorf((float)padw / 2.0f);
  if (top) *top = (int)floorf((float)padh / 2.0f);
  return DIS_OK;
}

const char* dis_last_error(const dis_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dis_create(const dis_params* params, int max_w, int max_h, int device, dis_handle** out) {
  return dis_create_c(params, 1, max_w, max_h, device, out);
}

int dis_create_c(const dis_params* params, int channels, int max_w, int max_h, int device, dis_handle** out) {
  return dis_create_batch(params, channels, max_w, max_h, device, 1, out);
}

int dis_create_batch(const dis_params* params, int channels, int max_w, int max_h, int device, int batch, dis_handle** out) {
  if (!out) return fail(nullptr, DIS_ERR_INVALID_ARG, "out is null");
  if (batch < 1 || batch > kMaxBatch) return fail(nullptr, DIS_ERR_INVALID_ARG, "batch must be in 1..%d", kMaxBatch);
  *out = nullptr;
  if (channels != 1 && channels != 3) return fail(nullptr, DIS_ERR_UNSUPPORTED, "channels must be 1 (grey) or 3 (BGR)");
  char why[128];
  int rc = dis_params_validate(params, why, sizeof why);
  if (rc != DIS_OK) return fail(nullptr, rc, "invalid parameters: %s", why);
  if (max_w <= 0 || max_h <= 0) return fail(nullptr, DIS_ERR_INVALID_ARG, "non-positive max size");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, DIS_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, DIS_ERR_INVALID_ARG, "device %d out of range", device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, DIS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, DIS_ERR_UNSUPPORTED, "device %d is sm_%d%d; kernels are built for sm_100a only", device,
                prop.major, prop.minor);
  dis_handle* h = new dis_handle;
  h->device = device;
  h->P = *params;
  h->noc = channels;
  h->nb = batch;
  derive_opt(h->P, h->noc, &h->opt);
  h->max_w = max_w;
  h->max_h = max_h;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    fail(nullptr, DIS_ERR_CUDA, "stream creation: %s", cudaGetErrorString(e));
    delete h;
    return DIS_ERR_CUDA;
  }
  patch_search_init_device();
  patch_search_init_device_fast();
  flowviz_init_device();
  varref_init_device();
  varref_init_device_fast();
  rc = plan(h, max_w, max_h);
  if (rc != DIS_OK) {
    g_create_error = h->err;
    dis_destroy(h);
    return rc;
  }
  *out = h;
  return DIS_OK;
}

int dis_destroy(dis_handle* h) {
  if (!h) return DIS_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  drop_graph(h);
  for (cudaEvent_t e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->pyr_ev) cudaEventDestroy(h->pyr_ev);
  if (h->slab) cudaFree(h->slab);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DIS_OK;
}

int dis_set_option(dis_handle* h, int option, int value) {
  if (!h) return DIS_ERR_INVALID_ARG;
  switch (option) {
    case DIS_OPT_SOR_GROUP:
      if (value != 0 && value != 8 && value != 16) return fail(h, DIS_ERR_INVALID_ARG, "DIS_OPT_SOR_GROUP must be 0, 8 or 16");
      if (value != h->sor_group) {
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaStreamSynchronize(h->stream));
        drop_graph(h);
        h->sor_group = value;
      }
      return DIS_OK;
    case DIS_OPT_SOR_SMALL:
      if (value < -1 || value > 9) return fail(h, DIS_ERR_INVALID_ARG, "DIS_OPT_SOR_SMALL must be -1 (auto) or 0 ... 9 (row blocks of 32 rows)");
      if (value != h->sor_small) {
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaStreamSynchronize(h->stream));
        drop_graph(h);
        h->sor_small = value;
      }
      return DIS_OK;
    case DIS_OPT_LEVEL_OUTPUT:
      h->level_output = value != 0;
      return DIS_OK;
    case DIS_OPT_USE_GRAPH:
      h->use_graph = value != 0;
      return DIS_OK;
    case DIS_OPT_ARITH:
      if (value != 0 && value != 1) return fail(h, DIS_ERR_INVALID_ARG, "DIS_OPT_ARITH must be 0 (exact) or 1 (fast)");
      if ((value != 0) != h->arith_fast) {
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaStreamSynchronize(h->stream));
        drop_graph(h);
        h->arith_fast = value != 0;
      }
      return DIS_OK;
    default:
      return fail(h, DIS_ERR_INVALID_ARG, "unknown option %d", option);
  }
}

int dis_set_params(dis_handle* h, const dis_params* params) {
  if (!h) return DIS_ERR_INVALID_ARG;
  char why[128];
  int rc = dis_params_validate(params, why, sizeof why);
  if (rc != DIS_OK) return fail(h, rc, "invalid parameters: %s", why);
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  h->P = *params;
  derive_opt(h->P, h->noc, &h->opt);
  const int w = h->w_org, hh = h->h_org;
  h->lv.clear();
  h->w_org = h->h_org = 0;
  return plan(h, w, hh);
}

void* dis_stream(dis_handle* h) { return h ? (void*)h->stream : nullptr; }

int dis_enable_stage_timing(dis_handle* h, int on) {
  if (!h) return DIS_ERR_INVALID_ARG;
  h->stage_timing = on != 0;
  return DIS_OK;
}

int dis_enable_taps(dis_handle* h, int on) {
  if (!h) return DIS_ERR_INVALID_ARG;
  const int was = h->taps;
  h->taps = on == 2 ? 2 : (on != 0);
  if (!on) h->tapdata.clear();
  if ((was == 1) != (h->taps == 1) && !h->lv.empty()) {  // the fine pyramid levels exist only for taps == 1: re-plan
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    const int w = h->w_org, hh = h->h_org;
    h->lv.clear();
    h->w_org = h->h_org = 0;
    return plan(h, w, hh);
  }
  return DIS_OK;
}

int dis_fetch_tap(dis_handle* h, int tap, int level, float* out, size_t n_floats, size_t* n_written) {
  if (!h || tap < 0 || level < 0) return DIS_ERR_INVALID_ARG;
  if (tap >= (int)h->tapdata.size() || level >= (int)h->tapdata[tap].size() ||
      h->tapdata[tap][level].data.empty())
    return fail(h, DIS_ERR_INVALID_ARG, "tap %d level %d not recorded", tap, level);
  const auto& v = h->tapdata[tap][level].data;
  if (n_written) *n_written = v.size();
  if (out) {
    if (n_floats < v.size()) return fail(h, DIS_ERR_INVALID_ARG, "tap buffer too small");
    memcpy(out, v.data(), v.size() * sizeof(float));
  }
  return DIS_OK;
}

int dis_enable_kernel_profile(dis_handle* h, int on) {
  if (!h) return DIS_ERR_INVALID_ARG;
  h->kprof_on = on != 0;
  h->kprof.st = h->stream;
  h->kprof.recs.clear();
  return DIS_OK;
}

int dis_get_kernel_profile(dis_handle* h, dis_kernel_time* out, int cap, int* n) {
  if (!h || !n) return DIS_ERR_INVALID_ARG;
  *n = (int)h->kprof.recs.size();
  if (out)
    for (int i = 0; i < *n && i < cap; ++i) out[i] = h->kprof.recs[i];
  return DIS_OK;
}

int dis_get_timings(dis_handle* h, dis_timings* out) {
  if (!h || !out) return DIS_ERR_INVALID_ARG;
  *out = h->tm;
  return DIS_OK;
}

int dis_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return DIS_ERR_INVALID_ARG;
  return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? DIS_OK : DIS_ERR_NOMEM;
}
int dis_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? DIS_OK : DIS_ERR_CUDA; }

int dis_submit_u8_device(dis_handle* h, const uint8_t* d_a, const uint8_t* d_b, int w, int h_img, int pitch,
                         float* d_flow) {
  if (!h || !d_a || !d_b || !d_flow || pitch < w * h->noc) return h ? fail(h, DIS_ERR_INVALID_ARG, "bad argument") : DIS_ERR_INVALID_ARG;
  CU(h, cudaSetDevice(h->device));
  int rc = plan(h, w, h_img);
  if (rc != DIS_OK) return rc;
  reset_timings(h);
  rc = enqueue_run_device(h, d_a, d_b, pitch, reinterpret_cast<float2*>(d_flow));
  h->in_flight = rc == DIS_OK;
  return rc;
}

int dis_batch_size(const dis_handle* h) { return h ? h->nb : 0; }

int dis_submit_u8_device_batch(dis_handle* h, int n_pairs, const uint8_t* const* d_a, const uint8_t* const* d_b, int w,
                               int h_img, int pitch, float* const* d_flow) {
  if (!h || !d_a || !d_b || !d_flow || n_pairs < 1 || n_pairs > h->nb || pitch < w * h->noc)
    return h ? fail(h, DIS_ERR_INVALID_ARG, "bad argument (1..%d pairs)", h->nb) : DIS_ERR_INVALID_ARG;
  for (int i = 0; i < n_pairs; ++i)
    if (!d_a[i] || !d_b[i] || !d_flow[i]) return fail(h, DIS_ERR_INVALID_ARG, "null pointer in pair %d", i);
  CU(h, cudaSetDevice(h->device));
  int rc = plan(h, w, h_img);
  if (rc != DIS_OK) return rc;
  reset_timings(h);
  rc = enqueue_run_device_batch(h, n_pairs, d_a, d_b, pitch, reinterpret_cast<float2* const*>(d_flow));
  h->in_flight = rc == DIS_OK;
  return rc;
}

int dis_submit_u8(dis_handle* h, const uint8_t* a, const uint8_t* b, int w, int h_img, int pitch, float* flow_out) {
  if (!h || !a || !b || !flow_out || pitch < w * h->noc) return h ? fail(h, DIS_ERR_INVALID_ARG, "bad argument") : DIS_ERR_INVALID_ARG;
  CU(h, cudaSetDevice(h->device));
  int rc = plan(h, w, h_img);
  if (rc != DIS_OK) return rc;
  reset_timings(h);
  if (h->ev_valid) return fail(h, DIS_ERR_INVALID_ARG, "a dis_submit_u8 is already in flight on this handle: dis_wait() first");
  for (int i = 0; i < 4; ++i)  // created once per handle, destroyed in dis_destroy
    if (!h->ev[i]) CU(h, cudaEventCreate(&h->ev[i]));
  h->ev_valid = false;
  CU(h, cudaEventRecord(h->ev[0], h->stream));
  const size_t rowb = (size_t)w * h->noc;
  CU(h, cudaMemcpy2DAsync(h->d_a, rowb, a, pitch, rowb, h_img, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpy2DAsync(h->d_b, rowb, b, pitch, rowb, h_img, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaEventRecord(h->ev[1], h->stream));
  rc = enqueue_run_device(h, h->d_a, h->d_b, (int)rowb, h->d_out);
  if (rc != DIS_OK) return rc;
  CU(h, cudaEventRecord(h->ev[2], h->stream));
  if (h->level_output) {  // the engine's own output (OFClass outflow): level lv_l, padded size
    const LevelBufs& L = h->lv[h->P.lv_l];
    CU(h, cudaMemcpyAsync(flow_out, L.flow, sizeof(float2) * (size_t)L.g.w * L.g.h, cudaMemcpyDeviceToHost, h->stream));
  } else {
    CU(h, cudaMemcpyAsync(flow_out, h->d_out, sizeof(float2) * (size_t)w * h_img, cudaMemcpyDeviceToHost, h->stream));
  }
  CU(h, cudaEventRecord(h->ev[3], h->stream));
  h->ev_valid = true;
  h->in_flight = true;
  return DIS_OK;
}

int dis_wait(dis_handle* h) {
  if (!h) return DIS_ERR_INVALID_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  if (h->ev_valid) {
    cudaEventElapsedTime(&h->tm.h2d_ms, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->tm.d2h_ms, h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&h->tm.total_ms, h->ev[0], h->ev[3]);
    h->ev_valid = false;
    if (h->P.verbosity > 0)  // format of kroeger/oflow.cpp:359
      printf("TIME (O.Flow Run-Time   ) (ms): %3g\n", h->tm.total_ms - h->tm.h2d_ms - h->tm.d2h_ms);
  }
  h->in_flight = false;
  CU(h, cudaGetLastError());
  return DIS_OK;
}

int dis_run_u8(dis_handle* h, const uint8_t* a, const uint8_t* b, int w, int h_img, int pitch, float* flow_out) {
  int rc = dis_submit_u8(h, a, b, w, h_img, pitch, flow_out);
  if (rc != DIS_OK) return rc;
  return dis_wait(h);
}

int dis_fetch_level_flow(dis_handle* h, float* out, size_t n_floats) {
  if (!h || !out || h->lv.empty()) return DIS_ERR_INVALID_ARG;
  const LevelBufs& L = h->lv[h->P.lv_l];
  const size_t n = (size_t)L.g.w * L.g.h * 2;
  if (n_floats < n) return fail(h, DIS_ERR_INVALID_ARG, "buffer too small: need %zu floats", n);
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(out, L.flow, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return DIS_OK;
}

int dis_level_flow_size(const dis_handle* h, int* w_l, int* h_l) {
  if (!h || h->lv.empty()) return DIS_ERR_INVALID_ARG;
  const LevelBufs& L = h->lv[h->P.lv_l];
  if (w_l) *w_l = L.g.w;
  if (h_l) *h_l = L.g.h;
  return DIS_OK;
}

const float* dis_level_flow_ptr(const dis_handle* h, int pair) {
  if (!h || h->lv.empty() || pair < 0 || pair >= h->nb) return nullptr;
  return reinterpret_cast<const float*>(bshift(h->lv[h->P.lv_l].flow, (size_t)pair * h->bstride));
}

int dis_copy_level_flow_device(dis_handle* h, int pair, float* d_dst) {
  if (!h || !d_dst || h->lv.empty() || pair < 0 || pair >= h->nb)
    return h ? fail(h, DIS_ERR_INVALID_ARG, "dis_copy_level_flow_device: bad argument") : DIS_ERR_INVALID_ARG;
  const LevelBufs& L = h->lv[h->P.lv_l];
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(d_dst, bshift(L.flow, (size_t)pair * h->bstride), sizeof(float2) * (size_t)L.g.w * L.g.h,
                        cudaMemcpyDeviceToDevice, h->stream));
  return DIS_OK;
}

int dis_set_level_export(dis_handle* h, int n_pairs, float* const* d_level) {
  if (!h || n_pairs < 0 || n_pairs > h->nb || (n_pairs > 0 && !d_level))
    return h ? fail(h, DIS_ERR_INVALID_ARG, "dis_set_level_export: 0..%d pointers", h->nb) : DIS_ERR_INVALID_ARG;
  for (int b = 0; b < kMaxBatch; ++b) h->lvl_export[b] = b < n_pairs ? reinterpret_cast<float2*>(d_level[b]) : nullptr;
  return DIS_OK;
}

int dis_run_pyramids(dis_handle* h, const float* const* im_ao, const float* const* im_ao_dx,
                     const float* const* im_ao_dy, const float* const* im_bo, const float* const* im_bo_dx,
                     const float* const* im_bo_dy, int imgpadding, int width, int height, const float* initflow,
                     float* outflow) {
  if (!h || !im_ao || !im_ao_dx || !im_ao_dy || !im_bo || !outflow)
    return h ? fail(h, DIS_ERR_INVALID_ARG, "null pyramid/outflow pointer") : DIS_ERR_INVALID_ARG;
  const dis_params& q = h->P;
  if (imgpadding != q.patchsz)
    return fail(h, DIS_ERR_UNSUPPORTED, "imgpadding (%d) must equal patchsz (%d) as in kroeger/run_dense.cpp:393",
                imgpadding, q.patchsz);
  const int sc = 1 << q.lv_f;
  if (width <= 0 || height <= 0 || width % sc || height % sc)
    return fail(h, DIS_ERR_INVALID_ARG, "width/height must be positive multiples of 2^lv_f (kroeger/oflow.h:87)");
  CU(h, cudaSetDevice(h->device));
  int rc = plan(h, width, height);  // already padded: plan() adds no further padding
  if (rc != DIS_OK) return rc;
  reset_timings(h);
  h->launches = 0;
  cudaEvent_t e0, e1, e2, e3;
  CU(h, cudaEventCreate(&e0));
  CU(h, cudaEventCreate(&e1));
  CU(h, cudaEventCreate(&e2));
  CU(h, cudaEventCreate(&e3));
  CU(h, cudaEventRecord(e0, h->stream));
  for (int l = q.lv_l; l <= q.lv_f; ++l) {
    LevelBufs& L = h->lv[l];
    const float* src[6] = {im_ao[l], im_ao_dx[l], im_ao_dy[l], im_bo[l], im_bo_dx ? im_bo_dx[l] : nullptr,
                           im_bo_dy ? im_bo_dy[l] : nullptr};
    float* dst[6] = {L.Ia, L.Iax, L.Iay, L.Ib, L.Ibx, L.Iby};
    for (int k = 0; k < 6; ++k) {
      if (!dst[k]) continue;
      if (!src[k]) return fail(h, DIS_ERR_INVALID_ARG, "pyramid %d level %d is null", k, l);
      CU(h, cudaMemcpy2DAsync(dst[k], sizeof(float) * L.g.pitch, src[k], sizeof(float) * L.g.tw * L.g.noc,
                              sizeof(float) * L.g.tw * L.g.noc, L.g.th, cudaMemcpyHostToDevice, h->stream));
    }
  }
  float2* d_init = nullptr;
  if (initflow) {  // resolution of level lv_f+1 (kroeger/oflow.h:91); staged in the output buffer
    const LevelGeom& gc = h->lv[q.lv_f].g;
    const size_t n = (size_t)(gc.w / 2) * (gc.h / 2);
    d_init = h->d_out;
    CU(h, cudaMemcpyAsync(d_init, initflow, n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  }
  CU(h, cudaEventRecord(e1, h->stream));
  rc = enqueue_engine(h, d_init);
  if (rc != DIS_OK) return rc;
  CU(h, cudaEventRecord(e2, h->stream));
  const LevelBufs& L = h->lv[q.lv_l];
  CU(h, cudaMemcpyAsync(outflow, L.flow, sizeof(float2) * (size_t)L.g.w * L.g.h, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaEventRecord(e3, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->tm.h2d_ms, e0, e1);
  cudaEventElapsedTime(&h->tm.d2h_ms, e2, e3);
  cudaEventElapsedTime(&h->tm.total_ms, e0, e3);
  h->tm.launches = h->launches;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  cudaEventDestroy(e3);
  if (q.verbosity > 0) printf("TIME (O.Flow Run-Time   ) (ms): %3g\n", h->tm.total_ms);
  CU(h, cudaGetLastError());
  return DIS_OK;
}

int dis_read_image_gray(const char* path, uint8_t* out, size_t cap, int* w, int* h) {
  if (!path || !w || !h) return DIS_ERR_INVALID_ARG;
  GrayImage g;
  const std::string err = read_gray_image(path, &g);
  if (!err.empty()) {
    g_create_error = err;
    return DIS_ERR_IO;
  }
  *w = g.w;
  *h = g.h;
  if (out) {
    if (cap < g.px.size()) return DIS_ERR_INVALID_ARG;
    memcpy(out, g.px.data(), g.px.size());
  }
  return DIS_OK;
}

int dis_read_image_bgr(const char* path, uint8_t* out, size_t cap, int* w, int* h) {
  if (!path || !w || !h) return DIS_ERR_INVALID_ARG;
  GrayImage g;
  const std::string err = read_image(path, 3, &g);
  if (!err.empty()) {
    g_create_error = err;
    return DIS_ERR_IO;
  }
  *w = g.w;
  *h = g.h;
  if (out) {
    if (cap < g.px.size()) return DIS_ERR_INVALID_ARG;
    memcpy(out, g.px.data(), g.px.size());
  }
  return DIS_OK;
}

int dis_write_flo(const char* path, const float* flow_uv, int w, int h) {
  if (!path || !flow_uv || w <= 0 || h <= 0) return DIS_ERR_INVALID_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) return DIS_ERR_IO;
  int ok = fwrite("PIEH", 1, 4, f) == 4;  // kroeger/run_dense.cpp:28-31
  ok = ok && fwrite(&w, sizeof(int), 1, f) == 1 && fwrite(&h, sizeof(int), 1, f) == 1;
  ok = ok && fwrite(flow_uv, sizeof(float), (size_t)2 * w * h, f) == (size_t)2 * w * h;
  ok = (fclose(f) == 0) && ok;
  return ok ? DIS_OK : DIS_ERR_IO;
}

int dis_read_flo(const char* path, float* flow_uv, size_t n_floats, int* w, int* h) {
  if (!path || !w || !h) return DIS_ERR_INVALID_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return DIS_ERR_IO;
  float tag = 0;
  int rc = DIS_OK;
  if (fread(&tag, sizeof(float), 1, f) != 1 || tag != 202021.25f ||  // flow_code/C/flowIO.cpp:27,66-72
      fread(w, sizeof(int), 1, f) != 1 || fread(h, sizeof(int), 1, f) != 1 || *w < 1 || *h < 1 || *w > 99999 ||
      *h > 99999)
    rc = DIS_ERR_IO;
  if (rc == DIS_OK && flow_uv) {
    const size_t n = (size_t)2 * *w * *h;
    if (n_floats < n)
      rc = DIS_ERR_INVALID_ARG;
    else if (fread(flow_uv, sizeof(float), n, f) != n)
      rc = DIS_ERR_IO;
  }
  fclose(f);
  return rc;
}

}  // extern "C"
