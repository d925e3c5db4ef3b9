// stream.cu -- video-stream front end: consecutive frames in, one flow field per consecutive pair out
// (SURVEY.md section 8(e) "within a GPU: several pairs in flight", 8(f)-4 "video-stream front end").
//
// The reference has no such caller on the CPU side: kroeger/run_dense.cpp handles exactly one pair per
// process (its CUDA twin src/ loops over a video in main).  Each pair still goes through the same engine as
// dis_run_u8, so every flow field is bit-identical to a separate run_dense call on that pair.
//
// Layout: `depth` engine handles (one CUDA stream + workspace + graph each) and a ring of depth+2 device
// frame slots.  Frame k is uploaded ONCE, on the stream of the pair that consumes it as its second image
// (pair k-1); pair k, on the next handle, waits for that upload through an event.  So a pair costs one u8
// frame of H2D traffic instead of two, and up to `depth` pairs overlap their copies and kernels.
//
// Batched mode (dis_video_create_batched, pairs_per_launch = nb > 1): depth / nb batched handles (dis_create_batch);
// the pairs of nb consecutive pushes are collected and go out as ONE launch chain, so a pair costs 1/nb of the
// launches (the device-resident rate of bench.py instead of the single-pair one) at the price of up to nb - 1 frame
// times of latency; a pop of a pair whose batch is still collecting submits the partial batch.
//
// Pyramid reuse ("frame k+1's pyramid is reused as the next pair's first image"): with depth >= 2
// the handles are chained (engine_chain): pair k builds only the pyramid of its second frame (with gradients) and
// reads its first frame's pyramid from the workspace of the handle that ran pair k-1, once that handle's pyramid
// event has fired.  A handle may overwrite its second-frame pyramid only after the pair that reads it as first frame
// has finished (the `done` event of the next handle).  The arithmetic per frame is unchanged, so is every flow.
#include <cuda_runtime.h>

#include <cstdio>
#include <deque>
#include <vector>

#include "../../include/dis_c.h"
#include "common.cuh"

struct dis_video {
  int w = 0, h = 0, noc = 1, device = 0, depth = 0;
  int nb = 1;                               // pairs per kernel launch (batched handles), depth = eng.size() * nb
  size_t frame_bytes = 0, flow_floats = 0;  // flow_floats: full-resolution field
  int out_mode = DIS_VIDEO_OUT_LEVEL;
  int lw = 0, lh = 0;                       // size of the level-lv_l flow (the engine's own output)
  std::vector<dis_handle*> eng;
  std::vector<uint8_t*> d_frame;       // ring, depth + 2 slots
  std::vector<cudaEvent_t> uploaded;   // per slot: upload finished
  std::vector<float*> d_flow;          // per pair in flight (handle * nb + position in its batch)
  std::vector<cudaEvent_t> done;       // per handle: D2H finished
  std::vector<float*> host_out;        // per pair in flight: destination
  // batched handles: the pairs of a handle's next launch, collected push by push
  std::vector<std::vector<const uint8_t*>> bat_a, bat_b;
  std::vector<int> fill;               // per handle: pairs collected and not yet submitted
  std::vector<long long> first_pair;   // per handle: index of the first pair of its current / latest batch
  std::vector<long long> last_pair;    // per handle: index of the last pair of its latest batch (-1: never used)
  std::vector<int> pair_handle;        // per pair in flight: the handle it went to
  int cur = 0;                         // the handle that is collecting (handles are used round-robin)
  long long pushed = 0;                // frames pushed so far
  long long popped = 0;                // pairs handed back so far
  bool chained = false;                // engine_chain done (workspaces hold the second frame's gradients)
  bool reuse = false;                  // pyramids shared between consecutive pairs
};

namespace {

#define CUV(call)                                                                                            \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess) {                                                                                 \
      dis::set_global_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
      return DIS_ERR_CUDA;                                                                                   \
    }                                                                                                        \
  } while (0)

// Launches the pairs collected on handle k (one pair when nb = 1) and queues their copies to the host.
int submit_batch(dis_video* v, int k) {
  const int n = v->fill[k], nH = (int)v->eng.size(), nslots = v->depth + 2;
  const long long p0 = v->first_pair[k];
  cudaStream_t st = static_cast<cudaStream_t>(dis_stream(v->eng[k]));
  const size_t rowb = (size_t)v->w * v->noc;
  const bool reuse = v->reuse && p0 >= 1;  // (nb = 1 only) pair - 1 built this pair's first frame as its second one
  if (reuse) {
    // its pyramid (and with it the upload) is ready when the previous handle's pyramid event fires ...
    CUV(cudaStreamWaitEvent(st, dis::engine_pyramid_event(v->eng[(k + nH - 1) % nH]), 0));
    // ... and this handle's own second-frame pyramid, about to be overwritten, was the first frame of the pair that
    // ran on the next handle depth-1 pairs ago
    if (p0 >= v->depth) CUV(cudaStreamWaitEvent(st, v->done[(k + 1) % nH], 0));
  } else {
    // the first image of the first pair was uploaded on the previous handle's stream (the other frames on this one)
    CUV(cudaStreamWaitEvent(st, v->uploaded[p0 % nslots], 0));
  }
  int rc;
  if (v->nb == 1) {
    dis::engine_set_reuse(v->eng[k], reuse);
    rc = dis_submit_u8_device(v->eng[k], v->bat_a[k][0], v->bat_b[k][0], v->w, v->h, (int)rowb, v->d_flow[k]);
  } else {
    rc = dis_submit_u8_device_batch(v->eng[k], n, v->bat_a[k].data(), v->bat_b[k].data(), v->w, v->h, (int)rowb,
                                    &v->d_flow[(size_t)k * v->nb]);
  }
  if (rc != DIS_OK) {
    dis::set_global_error("%s", dis_last_error(v->eng[k]));
    return rc;
  }
  for (int j = 0; j < n; ++j) {
    float* dst = v->host_out[(p0 + j) % v->depth];
    if (v->out_mode == DIS_VIDEO_OUT_LEVEL) {  // the engine's own output, straight from the handle's workspace
      CUV(cudaMemcpyAsync(dst, dis_level_flow_ptr(v->eng[k], j), (size_t)v->lw * v->lh * 2 * sizeof(float),
                          cudaMemcpyDeviceToHost, st));
    } else {
      CUV(cudaMemcpyAsync(dst, v->d_flow[(size_t)k * v->nb + j], v->flow_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
  }
  CUV(cudaEventRecord(v->done[k], st));
  v->fill[k] = 0;
  v->cur = (k + 1) % nH;  // the next pair starts a batch on the next handle
  return DIS_OK;
}

}  // namespace

extern "C" {

int dis_video_create(const dis_params* params, int channels, int w, int h, int device, int depth, dis_video** out) {
  return dis_video_create_batched(params, channels, w, h, device, depth, 1, out);
}

int dis_video_create_batched(const dis_params* params, int channels, int w, int h, int device, int depth,
                             int pairs_per_launch, dis_video** out) {
  const int nb = pairs_per_launch;
  if (!params || !out || w <= 0 || h <= 0 || depth < 1 || depth > 256 || nb < 1 || nb > 8 || depth % nb != 0) {
    dis::set_global_error("dis_video_create: bad argument (depth 1 ... 256, pairs_per_launch 1 ... 8 dividing depth)");
    return DIS_ERR_INVALID_ARG;
  }
  *out = nullptr;
  dis_video* v = new dis_video;
  v->w = w;
  v->h = h;
  v->noc = channels;
  v->device = device;
  v->depth = depth;
  v->nb = nb;
  v->frame_bytes = (size_t)w * h * channels;
  v->flow_floats = (size_t)w * h * 2;
  auto bail = [&](int rc) {
    dis_video_destroy(v);
    return rc;
  };
  for (int i = 0; i < depth / nb; ++i) {
    dis_handle* e = nullptr;
    const int rc = nb > 1 ? dis_create_batch(params, channels, w, h, device, nb, &e) : dis_create_c(params, channels, w, h, device, &e);
    if (rc != DIS_OK) return bail(rc);
    v->eng.push_back(e);
    v->bat_a.emplace_back(nb, nullptr);
    v->bat_b.emplace_back(nb, nullptr);
    v->fill.push_back(0);
    v->first_pair.push_back(0);
    v->last_pair.push_back(-1);
  }
  v->pair_handle.assign(depth, 0);
  dis_level_flow_size(v->eng[0], &v->lw, &v->lh);
  // Pyramid reuse is on by default for shallow pipelines (live streams, where it shortens a pair's critical path) and
  // off for deep ones: measured on the C5 stream at depth 64, the dependency between consecutive pairs costs more
  // (6 980 pairs/s) than the saved pyramid work brings (7 180 without).  dis_video_set_reuse() overrides.
  if (nb == 1 && depth >= 2 && depth <= 8) {
    const int rc = dis_video_set_reuse(v, 1);
    if (rc != DIS_OK) return bail(rc);
  }
  if (cudaSetDevice(device) != cudaSuccess) return bail(DIS_ERR_CUDA);
  for (int i = 0; i < depth + 2; ++i) {
    uint8_t* p = nullptr;
    cudaEvent_t ev = nullptr;
    if (cudaMalloc(&p, v->frame_bytes) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
      return bail(DIS_ERR_CUDA);
    v->d_frame.push_back(p);
    v->uploaded.push_back(ev);
  }
  for (int i = 0; i < depth; ++i) {
    float* p = nullptr;
    if (cudaMalloc(&p, v->flow_floats * sizeof(float)) != cudaSuccess) return bail(DIS_ERR_CUDA);
    v->d_flow.push_back(p);
    v->host_out.push_back(nullptr);
  }
  for (int i = 0; i < depth / nb; ++i) {
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(DIS_ERR_CUDA);
    v->done.push_back(ev);
  }
  *out = v;
  return DIS_OK;
}

void dis_video_destroy(dis_video* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  for (dis_handle* e : v->eng) {
    dis_wait(e);
    dis_destroy(e);
  }
  for (uint8_t* p : v->d_frame) cudaFree(p);
  for (float* p : v->d_flow) cudaFree(p);
  for (cudaEvent_t e : v->uploaded) cudaEventDestroy(e);
  for (cudaEvent_t e : v->done) cudaEventDestroy(e);
  delete v;
}

int dis_video_set_reuse(dis_video* v, int on) {
  if (!v || dis_video_pending(v) > 0 || v->pushed > 0) {
    dis::set_global_error("dis_video_set_reuse: only before the first frame is pushed");
    return DIS_ERR_INVALID_ARG;
  }
  if (on && (v->depth < 2 || v->nb > 1)) {
    dis::set_global_error("dis_video_set_reuse: needs depth >= 2 and one pair per launch");
    return DIS_ERR_INVALID_ARG;
  }
  if (on && !v->chained) {  // pair k reads its first frame's pyramid from the handle of pair k-1
    for (int i = 0; i < v->depth; ++i) {
      const int rc = dis::engine_chain(v->eng[i], v->eng[(i + v->depth - 1) % v->depth]);
      if (rc != DIS_OK) {
        dis::set_global_error("%s", dis_last_error(v->eng[i]));
        return rc;
      }
    }
    v->chained = true;
  }
  v->reuse = on != 0;
  return DIS_OK;
}

int dis_video_reuse(const dis_video* v) { return v && v->reuse ? 1 : 0; }

int dis_video_set_output(dis_video* v, int mode) {
  if (!v || (mode != DIS_VIDEO_OUT_LEVEL && mode != DIS_VIDEO_OUT_FULL) || dis_video_pending(v) > 0) {
    dis::set_global_error("dis_video_set_output: bad mode, or pairs in flight");
    return DIS_ERR_INVALID_ARG;
  }
  v->out_mode = mode;
  return DIS_OK;
}

size_t dis_video_flow_size(const dis_video* v, int* w_out, int* h_out) {
  if (!v) return 0;
  const bool lvl = v->out_mode == DIS_VIDEO_OUT_LEVEL;
  if (w_out) *w_out = lvl ? v->lw : v->w;
  if (h_out) *h_out = lvl ? v->lh : v->h;
  return lvl ? (size_t)v->lw * v->lh * 2 : v->flow_floats;
}

dis_handle* dis_video_handle(dis_video* v, int k) { return (v && k >= 0 && k < (int)v->eng.size()) ? v->eng[k] : nullptr; }

int dis_video_handles(const dis_video* v) { return v ? (int)v->eng.size() : 0; }

int dis_video_pending(const dis_video* v) { return v ? (int)((v->pushed > 0 ? v->pushed - 1 : 0) - v->popped) : 0; }

int dis_video_pop(dis_video* v, float** flow_out) {
  if (!v) return DIS_ERR_INVALID_ARG;
  if (dis_video_pending(v) <= 0) {
    dis::set_global_error("dis_video_pop: no pair in flight");
    return DIS_ERR_INVALID_ARG;
  }
  const int k = v->pair_handle[v->popped % v->depth];
  CUV(cudaSetDevice(v->device));
  if (v->fill[k] > 0 && v->first_pair[k] <= v->popped) {  // its batch is still collecting: send what there is
    const int rc = submit_batch(v, k);
    if (rc != DIS_OK) return rc;
  }
  const int rc = dis_wait(v->eng[k]);  // the handle's stream carries the graph and the copy-out
  if (rc != DIS_OK) {
    dis::set_global_error("%s", dis_last_error(v->eng[k]));
    return rc;
  }
  const int slot = (int)(v->popped % v->depth);
  if (flow_out) *flow_out = v->host_out[slot];
  v->host_out[slot] = nullptr;
  ++v->popped;
  return DIS_OK;
}

int dis_video_push(dis_video* v, const uint8_t* frame, int pitch, float* flow_out) {
  if (!v || !frame || pitch < v->w * v->noc) {
    dis::set_global_error("dis_video_push: bad argument");
    return DIS_ERR_INVALID_ARG;
  }
  const long long f = v->pushed;  // index of this frame
  if (f > 0 && !flow_out) {
    dis::set_global_error("dis_video_push: flow_out is required from the second frame on");
    return DIS_ERR_INVALID_ARG;
  }
  if (f > 0 && dis_video_pending(v) >= v->depth) {
    dis::set_global_error("dis_video_push: %d pairs in flight, pop one first", v->depth);
    return DIS_ERR_INVALID_ARG;
  }
  CUV(cudaSetDevice(v->device));
  const int nslots = v->depth + 2;
  const int slot = (int)(f % nslots);
  // the handle that carries this frame's upload: that of pair f-1 (this frame is its second image); the very first
  // frame rides on handle 0's stream
  const long long pair = f > 0 ? f - 1 : 0;
  const int k = v->cur;  // the collecting handle
  if (f > 0 && v->fill[k] == 0 && v->last_pair[k] >= 0 && v->first_pair[k] >= v->popped) {
    // the handle's latest batch was never waited for (the pop of its first pair does that for the whole batch); only
    // possible after pops forced partial batches out
    dis::set_global_error("dis_video_push: all %d handles hold pairs in flight, pop one first", (int)v->eng.size());
    return DIS_ERR_INVALID_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(dis_stream(v->eng[k]));
  // slot reuse: its previous tenant, frame f - nslots, was last read by pair f - nslots, which was popped
  // before pair f - 1 could be admitted (pending < depth) -- nothing to wait for
  const size_t rowb = (size_t)v->w * v->noc;
  CUV(cudaMemcpy2DAsync(v->d_frame[slot], rowb, frame, pitch, rowb, v->h, cudaMemcpyHostToDevice, st));
  CUV(cudaEventRecord(v->uploaded[slot], st));
  ++v->pushed;
  if (f == 0) return DIS_OK;
  const int prev = (int)((f - 1) % nslots);
  const int j = v->fill[k];  // position in the handle's batch
  if (j == 0) v->first_pair[k] = pair;
  v->last_pair[k] = pair;
  v->bat_a[k][j] = v->d_frame[prev];
  v->bat_b[k][j] = v->d_frame[slot];
  v->host_out[pair % v->depth] = flow_out;
  v->pair_handle[pair % v->depth] = k;
  v->fill[k] = j + 1;
  return v->fill[k] == v->nb ? submit_batch(v, k) : DIS_OK;
}

}  // extern "C"
