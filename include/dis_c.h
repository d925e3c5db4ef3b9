/* dis_c.h -- C-ABI of libdis_b200.so, the B200-native Dense Inverse Search optical-flow engine.
 *
 * This is the drop-in boundary for the reference's hot path (zhaorz/FlowOnTheGo, CPU tree
 * `kroeger/`): frame pair in -> dense flow out.  Every entry point names the reference
 * interface it replaces.  Plain pointers and sizes only; every function returns a dis_status
 * (0 = ok) unless stated otherwise.  There is no CPU fallback: without a CUDA device (or with a
 * device that is not sm_100) dis_create fails with DIS_ERR_CUDA / DIS_ERR_UNSUPPORTED.
 *
 * Scope: SELECTMODE=1 (optical flow), SELECTCHANNEL=1 (grey) -- the reference's `run_OF_INT`
 * binary (kroeger/CMakeLists.txt:44-48).
 */
#ifndef DIS_C_H
#define DIS_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dis_status {
  DIS_OK = 0,
  DIS_ERR_INVALID_ARG = 1, /* null pointer, non-positive size, size larger than the handle was created for */
  DIS_ERR_UNSUPPORTED = 2, /* parameter combination outside the scoped mode (see dis_params) */
  DIS_ERR_CUDA = 3,        /* CUDA runtime error; text in dis_last_error() */
  DIS_ERR_IO = 4,          /* file could not be read / written / parsed */
  DIS_ERR_NOMEM = 5
} dis_status;

/* The 20 explicit command-line parameters of the reference, in CLI order
 * (kroeger/run_dense.cpp:271-291, kroeger/README.md:71-88).  Meaning and defaults are the
 * reference's (kroeger/oflow.h:31-76). */
typedef struct dis_params {
  int32_t lv_f;        /*  1 coarsest scale (sc_f)                                   */
  int32_t lv_l;        /*  2 finest scale (sc_l)                                     */
  int32_t maxiter;     /*  3 max. Gauss-Newton iterations per patch and scale        */
  int32_t miniter;     /*  4 min. iterations                                         */
  float mindprate;     /*  5 dp_thresh  (squared internally, oflow.cpp:88)           */
  float mindrrate;     /*  6 dr_thresh                                               */
  float minimgerr;     /*  7 res_thresh                                              */
  int32_t patchsz;     /*  8 patch edge length p (even, 4..16)                       */
  float poverl;        /*  9 patch overlap in [0,1): steps = max(1, floor(p*(1-ov))) */
  int32_t usefbcon;    /* 10 forward-backward consistency merge                      */
  int32_t patnorm;     /* 11 mean-normalise patches                                  */
  int32_t costfct;     /* 12 0 = L2, 1 = L1, 2 = pseudo-Huber                        */
  int32_t usetvref;    /* 13 variational refinement on/off                           */
  float tv_alpha;      /* 14 */
  float tv_gamma;      /* 15 */
  float tv_delta;      /* 16 */
  int32_t tv_innerit;  /* 17 inner fixed-point iterations = tv_innerit*(level+1)     */
  int32_t tv_solverit; /* 18 SOR sweeps per inner iteration                          */
  float tv_sor;        /* 19 SOR omega                                               */
  int32_t verbosity;   /* 20 0 silent, 1 total run time line, 2 per-scale TIME lines */
} dis_params;

/* Device-side stage timings of the last run (CUDA events), the counterpart of the reference's
 * gettimeofday brackets (kroeger/oflow.cpp:112-128, 199-304, 354-360). Milliseconds. */
typedef struct dis_timings {
  float total_ms;   /* whole run incl. copies when host buffers are used */
  float h2d_ms;     /* host -> device copy of the inputs                  */
  float pyramid_ms; /* stage 1: pyramid + gradients (0 for dis_run_pyramids) */
  float search_ms;  /* stage 2: template/Hessian + coarse init + inverse search, all scales */
  float densify_ms; /* stage 3: densification, all scales */
  float varref_ms;  /* stage 4: variational refinement, all scales */
  float finish_ms;  /* upsample + crop */
  float d2h_ms;     /* device -> host copy of the result */
  int32_t launches; /* kernels launched by the last run */
} dis_timings;

typedef struct dis_handle dis_handle;

/* ---- parameter helpers (host only, no CUDA) -------------------------------------------- */

/* kroeger/run_dense.cpp:180-183 AutoFirstScaleSelect */
int dis_auto_first_scale(int imgwidth, int fratio, int patchsize);

/* Operating points 1..4 of kroeger/run_dense.cpp:225-267 for an image of width `width_org`
 * (anything else selects 2, like the reference's `default:` label).  verbosity is set to 2. */
int dis_params_preset(dis_params* out, int preset, int width_org);

/* The 20 explicit parameters as strings in CLI order (argv[4..23] of the reference,
 * kroeger/run_dense.cpp:271-291; parsed with atoi/atof like the reference). */
int dis_params_from_argv(dis_params* out, int n, const char* const* args);

/* Checks a parameter set against the scoped mode; on failure writes a reason into `why`. */
int dis_params_validate(const dis_params* p, char* why, size_t why_len);

/* Padded size the reference would work on: multiple of 2^lv_f (kroeger/run_dense.cpp:298-311). */
int dis_padded_size(int w, int h, int lv_f, int* w_pad, int* h_pad, int* left, int* top);

/* ---- engine lifetime ------------------------------------------------------------------- */

/* Creates an engine on CUDA device `device` with workspace for images up to max_w x max_h
 * (unpadded input size).  One handle owns one CUDA stream; several handles run concurrently. */
int dis_create(const dis_params* params, int max_w, int max_h, int device, dis_handle** out);
/* Same with an explicit channel count: 1 = grey (the reference's run_OF_INT build, SELECTCHANNEL=1), 3 =
 * interleaved BGR as cv::imread(.., COLOR) delivers it (run_OF_RGB, SELECTCHANNEL=3; kroeger/run_dense.cpp:203-206,
 * noc=3 at the engine boundary).  For 3 channels every u8 / float image argument below is interleaved and
 * `pitch` still counts bytes per row. */
int dis_create_c(const dis_params* params, int channels, int max_w, int max_h, int device, dis_handle** out);
int dis_destroy(dis_handle* h);
/* Replaces the parameter set (workspace is re-planned; fails if it no longer fits). */
int dis_set_params(dis_handle* h, const dis_params* params);
/* Execution options; only DIS_OPT_ARITH changes results.
 *   DIS_OPT_SOR_GROUP  8: smaller shared-memory footprint of the SOR wavefront kernel, best pairs/s when many
 *                      handles share the GPU; 16: lowest latency for a lone pair (about 10 % at 1080p: larger groups
 *                      and one warp per (sweep, row block) item instead of persistent warps); 0 (default): persistent
 *                      warps, groups of 16 when the finest processed level has >= 2^20 pixels, else of 8.
 *   DIS_OPT_SOR_SMALL  n: pyramid levels of at most n blocks of 32 rows run their SOR sweeps in one CTA per pair, one
 *                      warp per (sweep, row block), two columns per barrier-separated step (k_sor_small) instead of
 *                      the wavefront pipeline, if max(tv_solverit, 3) * blocks <= 16 and the rings fit in shared memory;
 *                      0 = never; -1 (default) = 5, or 4 when DIS_OPT_SOR_GROUP is 16 (the latency setting).
 *   DIS_OPT_USE_GRAPH  1 (default): record a run into a CUDA graph and replay it; 0: launch kernel by kernel.
 * DIS_OPT_LEVEL_OUTPUT changes WHAT dis_run_u8 / dis_submit_u8 copy back, not how it is computed: 1 = the engine's
 * own output as the OFC::OFClass constructor delivers it (level lv_l, (w_pad/2^lv_l) x (h_pad/2^lv_l) x 2 floats,
 * see dis_padded_size), leaving the x2^lv_l resize and crop of kroeger/run_dense.cpp:407-414 to the caller -- 16x
 * less device-to-host traffic at lv_l = 2; 0 (default) = full-resolution flow.
 *
 * DIS_OPT_ARITH is the one option that DOES change results: 0 (default) = the exact engine, bit-identical to the
 * reference's -msse4 build (no FMA contraction, reference reduction order); 1 = tolerance mode: the inverse search
 * and the refinement kernels compiled with FMA contraction allowed (same reduction order).  Its output stays within
 * the stated tolerance of the reference (mean |dflow| <= 1e-3 px, max <= 1e-2 px outside the border margin) for
 * operating points with maxiter <= 32, which is enforced: a run with more iterations fails with
 * DIS_ERR_UNSUPPORTED (rounding differences grow with the iteration count, 0.25 px at 128).  It is never part of a
 * parity claim; bench.py reports it as a separate line (--arith fast). */
typedef enum dis_option {
  DIS_OPT_SOR_GROUP = 1, DIS_OPT_USE_GRAPH = 2, DIS_OPT_LEVEL_OUTPUT = 3, DIS_OPT_ARITH = 4, DIS_OPT_SOR_SMALL = 5
} dis_option;
int dis_set_option(dis_handle* h, int option, int value);
/* Last error text of this handle (or of dis_create when h is NULL). Never NULL. */
const char* dis_last_error(const dis_handle* h);

/* ---- the reference's engine boundary --------------------------------------------------- */

/* Replaces `OFC::OFClass::OFClass(...)` (kroeger/oflow.h:84-111, oflow.cpp:32-363; called at
 * kroeger/run_dense.cpp:391-400).  Same argument meaning: six host pyramids indexed by level
 * [0..lv_f] (entries below lv_l may be NULL; the gradient pyramids of image b are only read
 * when usefbcon is set), each a row-major float image of (width/2^l + 2*imgpadding) x
 * (height/2^l + 2*imgpadding); width/height are the padded sizes (multiples of 2^lv_f);
 * outflow receives (width/2^lv_l) x (height/2^lv_l) interleaved (u,v); initflow is optional
 * (resolution of level lv_f+1).  imgpadding must equal params.patchsz as in the reference's
 * only caller.  Synchronous. */
int dis_run_pyramids(dis_handle* h, const float* const* im_ao, const float* const* im_ao_dx,
                     const float* const* im_ao_dy, const float* const* im_bo,
                     const float* const* im_bo_dx, const float* const* im_bo_dy, int imgpadding,
                     int width, int height, const float* initflow, float* outflow);

/* ---- the whole `run_dense` data path --------------------------------------------------- */

/* Replaces kroeger/run_dense.cpp:298-414 between imread and SaveFlowFile: divisibility padding,
 * u8->f32, pyramid + gradients (ConstructImgPyramide :130-178) -- all on the GPU -- then the
 * engine, then x2^lv_l upsampling and cropping.  a,b: host, row pitch in bytes, w x h grey u8.
 * flow_out: host, w*h*2 floats, interleaved (u,v), full resolution.  Synchronous. */
int dis_run_u8(dis_handle* h, const uint8_t* a, const uint8_t* b, int w, int h_img, int pitch,
               float* flow_out);

/* Asynchronous variant: enqueues copy-in, compute and copy-out on the handle's stream and
 * returns.  Buffers must stay valid (and should be pinned, see dis_host_alloc) until
 * dis_wait().  At most one submission may be in flight per handle. */
int dis_submit_u8(dis_handle* h, const uint8_t* a, const uint8_t* b, int w, int h_img, int pitch,
                  float* flow_out);
int dis_wait(dis_handle* h);

/* Device-resident variant for batched streams: d_a, d_b and d_flow are device pointers on the
 * handle's device; work is enqueued on the handle's stream (asynchronous; use dis_wait). */
int dis_submit_u8_device(dis_handle* h, const uint8_t* d_a, const uint8_t* d_b, int w, int h_img,
                         int pitch, float* d_flow);

/* Raw engine output of the last run_u8 (level lv_l, padded size), for parity tests against the
 * reference's OFClass output before the OpenCV upsampling: (w_pad/2^lv_l)*(h_pad/2^lv_l)*2. */
int dis_fetch_level_flow(dis_handle* h, float* out, size_t n_floats);
/* Size of that output for the current plan: (w_pad/2^lv_l) x (h_pad/2^lv_l). */
int dis_level_flow_size(const dis_handle* h, int* w_l, int* h_l);
/* Device-to-device copy of the level-lv_l flow of pair `pair` (0 for a single-pair handle, < batch for a batched
 * one) into d_dst (w_l*h_l*2 floats on the handle's device), enqueued on the handle's stream behind the run that
 * produces it.  This is what a multi-GPU job gathers over NCCL: the engine's output stays in HBM. */
int dis_copy_level_flow_device(dis_handle* h, int pair, float* d_dst);
/* Export without an extra launch: the following dis_submit_u8_device[_batch] calls on this handle also write the
 * level-lv_l flow of pair b to d_level[b] (device, w_l*h_l*2 floats each; done by the final kernel of the run).
 * Stays in force until changed; n_pairs = 0 turns it off. */
int dis_set_level_export(dis_handle* h, int n_pairs, float* const* d_level);
/* Device address of that flow inside the handle's workspace (valid until the next re-plan; rewritten by the
 * handle's next run), or NULL. */
const float* dis_level_flow_ptr(const dis_handle* h, int pair);

/* CUDA stream of the handle as a cudaStream_t (void* here to keep CUDA out of the header). */
void* dis_stream(dis_handle* h);
int dis_get_timings(dis_handle* h, dis_timings* out);
/* Enables per-stage event timing (off by default; stage timers add synchronisation points). */
int dis_enable_stage_timing(dis_handle* h, int on);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
int dis_host_alloc(void** ptr, size_t bytes);
int dis_host_free(void* ptr);

/* ---- .flo I/O (kroeger/run_dense.cpp:16-57 SaveFlowFile; flow_code/C/flowIO.cpp:46-133) -- */
int dis_write_flo(const char* path, const float* flow_uv, int w, int h);
/* Reads the header into *w,*h; if flow_uv is non-NULL (capacity n_floats) also the data. */
int dis_read_flo(const char* path, float* flow_uv, size_t n_floats, int* w, int* h);

/* ---- image input of the CLI (kroeger/run_dense.cpp:208-209 cv::imread(.., GRAYSCALE)) ------- */
/* Reads an 8-bit PNG / PGM / PPM as grey (OpenCV's grey conversion reproduced bit for bit).  With
 * out == NULL only *w,*h are filled. */
int dis_read_image_gray(const char* path, uint8_t* out, size_t cap, int* w, int* h);
/* Colour build (kroeger/run_dense.cpp:203-206 cv::imread(.., COLOR)): interleaved BGR, 3*w*h bytes. */
int dis_read_image_bgr(const char* path, uint8_t* out, size_t cap, int* w, int* h);

/* ---- batched handles: several pairs per kernel launch ----------------------------------------------------- */
/* dis_create_batch makes a handle whose every kernel launch serves `batch` (1..8) pairs at once: the workspace is
 * replicated `batch` times and the pair index rides on a grid dimension, so a pair costs 1/batch of the launches
 * (88 at 1080p) and the latency-bound coarse levels get `batch` times the parallelism.  Results per pair are
 * bit-identical to dis_run_u8.  dis_submit_u8_device_batch takes arrays of n_pairs <= batch device pointers (same
 * w, h, pitch for all); asynchronous on the handle's stream, dis_wait() waits for all pairs.  The single-pair
 * entry points also work on such a handle (they use slot 0; the other slots idle along). */
int dis_create_batch(const dis_params* params, int channels, int max_w, int max_h, int device, int batch, dis_handle** out);
int dis_batch_size(const dis_handle* h);
int dis_submit_u8_device_batch(dis_handle* h, int n_pairs, const uint8_t* const* d_a, const uint8_t* const* d_b, int w,
                               int h_img, int pitch, float* const* d_flow);

/* ---- video-stream front end ------------------------------------------------------------------ */
/* Consecutive frames in, one flow field per consecutive pair out (the loop the reference's CUDA twin runs over a
 * video, src/main.cpp; kroeger/run_dense.cpp itself handles one pair per process).  `depth` pairs are in flight
 * on `depth` engine handles; every frame is uploaded once.  Each flow is bit-identical to dis_run_u8 on that
 * pair.  Usage: push frame 0 (flow_out ignored), then for every further frame push(frame, flow_out) -- when
 * dis_video_pending() == depth, pop first.  dis_video_pop waits for the OLDEST pair in flight and returns the
 * flow_out pointer it was given.  Host buffers should be pinned (dis_host_alloc) and must stay valid until
 * popped.  Not thread-safe per object. */
typedef struct dis_video dis_video;
int dis_video_create(const dis_params* params, int channels, int w, int h, int device, int depth, dis_video** out);
/* Throughput variant: the pairs of `pairs_per_launch` (1 ... 8, dividing depth) consecutive pushes go out as ONE
 * launch chain on a batched handle (dis_create_batch), depth / pairs_per_launch handles in all.  Same results bit
 * for bit, 1 / pairs_per_launch of the kernel launches per pair; a pair's flow is ready only after the last frame
 * of its batch was pushed (dis_video_pop submits a partial batch rather than wait for frames that may never come;
 * a partial batch still occupies its handle, so after such pops dis_video_push can ask for another pop before
 * `depth` pairs are in flight).  No pyramid reuse in this mode. */
int dis_video_create_batched(const dis_params* params, int channels, int w, int h, int device, int depth,
                             int pairs_per_launch, dis_video** out);
void dis_video_destroy(dis_video* v);
/* What dis_video_push copies back per pair.  DIS_VIDEO_OUT_LEVEL (the default for streams): the engine's own
 * output as OFC::OFClass delivers it -- level lv_l, dis_video_flow_size() floats, 1.04 MB per 1080p pair at
 * lv_l = 2 -- leaving the x2^lv_l resize + crop of kroeger/run_dense.cpp:407-414 to the consumer;
 * DIS_VIDEO_OUT_FULL: the full-resolution w x h x 2 flow (what run_dense writes to the .flo file; 16.6 MB per
 * 1080p pair).  The computation is the same either way.  Only while no pair is in flight. */
typedef enum dis_video_output { DIS_VIDEO_OUT_LEVEL = 0, DIS_VIDEO_OUT_FULL = 1 } dis_video_output;
int dis_video_set_output(dis_video* v, int mode);
/* Pyramid reuse between consecutive pairs: pair k builds only its second frame's pyramid and takes the
 * first frame's from pair k-1.  Same results bit for bit; shortens a pair's critical path but makes pair k wait for
 * pair k-1's pyramid, which costs throughput in deep pipelines.  Default: on for 2 <= depth <= 8, off otherwise.
 * Only before the first frame is pushed; needs depth >= 2. */
int dis_video_set_reuse(dis_video* v, int on);
int dis_video_reuse(const dis_video* v);
/* Floats per flow field handed back in the current output mode, and its width / height. */
size_t dis_video_flow_size(const dis_video* v, int* w_out, int* h_out);
/* k-th engine handle (0 <= k < dis_video_handles() = depth / pairs_per_launch) -- e.g. for dis_stream(); owned by
 * the video object. */
dis_handle* dis_video_handle(dis_video* v, int k);
int dis_video_handles(const dis_video* v);
int dis_video_push(dis_video* v, const uint8_t* frame, int pitch, float* flow_out);
int dis_video_pop(dis_video* v, float** flow_out);
int dis_video_pending(const dis_video* v);

/* ---- evaluation tools behind the path (flow_code/C) ----------------------------------------- */
/* Middlebury colour coding of a flow field on the GPU: replaces MotionToColor (flow_code/C/color_flow.cpp:19-71)
 * with computeColor (flow_code/C/colorcode.cpp:53-77).  maxmotion <= 0: normalise by the largest motion present.
 * bgr_out: w*h*3 bytes, interleaved B,G,R (the byte order computeColor stores).  stats_out (optional, 5 floats):
 * max radius, min u, max u, min v, max v -- the numbers color_flow prints.  Unknown flow (|u| or |v| > 1e9, NaN;
 * flow_code/C/flowIO.cpp:35-39) is black.  Host buffers; runs on CUDA device `device`. */
int dis_flow_to_color(const float* flow_uv, int w, int h, float maxmotion, int device, uint8_t* bgr_out,
                      float* stats_out);
/* Device-resident variant for pipelines: d_stats is 5 words of device scratch; asynchronous on `stream`
 * (a cudaStream_t).  dis_create must have run on the device (it uploads the colour wheel). */
int dis_flow_to_color_device(const float* d_flow_uv, int w, int h, float maxmotion, uint8_t* d_bgr,
                             uint32_t* d_stats, void* stream);
/* Endpoint error between two flow fields on the GPU: mean and max of |a - b|_2 over the pixels at least
 * `margin` away from the border where both flows are known; computed in double, deterministic. */
int dis_flow_epe(const float* flow_a, const float* flow_b, int w, int h, int margin, int device, double* mean_out,
                 double* max_out, long long* count_out);
/* 8-bit RGB PNG from interleaved BGR pixels (the output format of color_flow). */
int dis_write_png_bgr(const char* path, const uint8_t* bgr, int w, int h);

/* ---- stage-level debug taps (tests only; never on the timed path) ---------------------- */
typedef enum dis_tap {
  DIS_TAP_IMG_A = 0,   /* padded pyramid image a, level l: (w_l+2p) x (h_l+2p)         */
  DIS_TAP_IMG_A_DX = 1,
  DIS_TAP_IMG_A_DY = 2,
  DIS_TAP_IMG_B = 3,
  DIS_TAP_PATCH_FLOW = 4,   /* per patch (u,v) after the inverse search, patch index x*noph+y */
  DIS_TAP_FLOW_DENSE = 5,   /* level flow after densification (before variational refinement) */
  DIS_TAP_FLOW_REFINED = 6, /* level flow after variational refinement                      */
  DIS_TAP_IMG_B_DX = 7,
  DIS_TAP_IMG_B_DY = 8
} dis_tap;
/* Keeps per-level copies of the tapped buffers during the next runs (costs memory and time).  on = 1: every
 * pyramid level 0..lv_f is built level by level like ConstructImgPyramide; on = 2: the pyramid is built as on the
 * untapped path (levels below lv_l are skipped, see pyramid.cu) and only levels >= lv_l can be fetched. */
int dis_enable_taps(dis_handle* h, int on);
int dis_fetch_tap(dis_handle* h, int tap, int level, float* out, size_t n_floats, size_t* n_written);

/* ---- per-kernel device timing (bench.py's roofline leg; never on the timed path) ---------- */
typedef struct dis_kernel_time {
  char name[32];     /* kernel name, e.g. "k_sor_wavefront"                                  */
  int32_t level;     /* pyramid level                                                        */
  int32_t launches;  /* launches accumulated                                                 */
  float ms;          /* sum of CUDA-event durations over those launches                      */
  double alg_bytes;  /* sum of ALGORITHMIC bytes (SURVEY.md section 8(d) model) over them     */
} dis_kernel_time;
/* While on, runs are not replayed from the CUDA graph and every launch is bracketed by events. */
int dis_enable_kernel_profile(dis_handle* h, int on);
/* Copies up to `cap` accumulated records (one per kernel and level); *n receives the count. */
int dis_get_kernel_profile(dis_handle* h, dis_kernel_time* out, int cap, int* n);

/* Library identification: "dis_b200 <version> sm_100a". */
const char* dis_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DIS_C_H */
