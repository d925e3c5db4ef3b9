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
  size_t frame_bytes = 0, flow_floats = 0;  // flow_floats: full-resolution field
  int out_mode = DIS_VIDEO_OUT_LEVEL;
  int lw = 0, lh = 0;                       // size of the level-lv_l flow (the engine's own output)
  std::vector<dis_handle*> eng;
  std::vector<uint8_t*> d_frame;       // ring, depth + 2 slots
  std::vector<cudaEvent_t> uploaded;   // per slot: upload finished
  std::vector<float*> d_flow;          // per handle
  std::vector<cudaEvent_t> done;       // per handle: D2H finished
  std::vector<float*> host_out;        // per handle: destination of the pair in flight
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

}  // namespace

extern "C" {

int dis_video_create(const dis_params* params, int channels, int w, int h, int device, int depth, dis_video** out) {
  if (!params || !out || w <= 0 || h <= 0 || depth < 1 || depth > 256) {
    dis::set_global_error("dis_video_create: bad argument");
    return DIS_ERR_INVALID_ARG;
  }
  *out = nullptr;
  dis_video* v = new dis_video;
  v->w = w;
  v->h = h;
  v->noc = channels;
  v->device = device;
  v->depth = depth;
  v->frame_bytes = (size_t)w * h * channels;
  v->flow_floats = (size_t)w * h * 2;
  auto bail = [&](int rc) {
    dis_video_destroy(v);
    return rc;
  };
  for (int i = 0; i < depth; ++i) {
    dis_handle* e = nullptr;
    const int rc = dis_create_c(params, channels, w, h, device, &e);
    if (rc != DIS_OK) return bail(rc);
    v->eng.push_back(e);
  }
  dis_level_flow_size(v->eng[0], &v->lw, &v->lh);
  // Pyramid reuse is on by default for shallow pipelines (live streams, where it shortens a pair's critical path) and
  // off for deep ones: measured on the C5 stream at depth 64, the dependency between consecutive pairs costs more
  // (6 980 pairs/s) than the saved pyramid work brings (7 180 without).  dis_video_set_reuse() overrides.
  if (depth >= 2 && depth <= 8) {
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
    cudaEvent_t ev = nullptr;
    if (cudaMalloc(&p, v->flow_floats * sizeof(float)) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
      return bail(DIS_ERR_CUDA);
    v->d_flow.push_back(p);
    v->done.push_back(ev);
    v->host_out.push_back(nullptr);
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
  if (on && v->depth < 2) {
    dis::set_global_error("dis_video_set_reuse: needs depth >= 2");
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

dis_handle* dis_video_handle(dis_video* v, int k) { return (v && k >= 0 && k < v->depth) ? v->eng[k] : nullptr; }

int dis_video_pending(const dis_video* v) { return v ? (int)((v->pushed > 0 ? v->pushed - 1 : 0) - v->popped) : 0; }

int dis_video_pop(dis_video* v, float** flow_out) {
  if (!v) return DIS_ERR_INVALID_ARG;
  if (dis_video_pending(v) <= 0) {
    dis::set_global_error("dis_video_pop: no pair in flight");
    return DIS_ERR_INVALID_ARG;
  }
  const int k = (int)(v->popped % v->depth);
  CUV(cudaSetDevice(v->device));
  const int rc = dis_wait(v->eng[k]);  // the handle's stream carries the graph and the copy-out
  if (rc != DIS_OK) {
    dis::set_global_error("%s", dis_last_error(v->eng[k]));
    return rc;
  }
  if (flow_out) *flow_out = v->host_out[k];
  v->host_out[k] = nullptr;
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
  // the pair that carries this frame's upload: pair f-1 (this frame is its second image); the very first frame
  // rides on handle 0's stream
  const int k = (int)((f > 0 ? f - 1 : 0) % v->depth);
  cudaStream_t st = static_cast<cudaStream_t>(dis_stream(v->eng[k]));
  // slot reuse: its previous tenant, frame f - nslots, was last read by pair f - nslots, which was popped
  // before pair f - 1 could be admitted (pending < depth) -- nothing to wait for
  const size_t rowb = (size_t)v->w * v->noc;
  CUV(cudaMemcpy2DAsync(v->d_frame[slot], rowb, frame, pitch, rowb, v->h, cudaMemcpyHostToDevice, st));
  CUV(cudaEventRecord(v->uploaded[slot], st));
  ++v->pushed;
  if (f == 0) return DIS_OK;
  const int prev = (int)((f - 1) % nslots);
  const long long pair = f - 1;
  const bool reuse = v->reuse && pair >= 1;  // pair - 1 built this pair's first frame as its second one
  if (reuse) {
    // its pyramid (and with it the upload) is ready when the previous handle's pyramid event fires ...
    CUV(cudaStreamWaitEvent(st, dis::engine_pyramid_event(v->eng[(k + v->depth - 1) % v->depth]), 0));
    // ... and this handle's own second-frame pyramid, about to be overwritten, was the first frame of the pair that
    // ran on the next handle depth-1 pairs ago
    if (pair >= v->depth) CUV(cudaStreamWaitEvent(st, v->done[(k + 1) % v->depth], 0));
  } else {
    CUV(cudaStreamWaitEvent(st, v->uploaded[prev], 0));  // first image was uploaded on the previous pair's stream
  }
  dis::engine_set_reuse(v->eng[k], reuse);
  const int rc = dis_submit_u8_device(v->eng[k], v->d_frame[prev], v->d_frame[slot], v->w, v->h, (int)rowb, v->d_flow[k]);
  if (rc != DIS_OK) {
    dis::set_global_error("%s", dis_last_error(v->eng[k]));
    return rc;
  }
  if (v->out_mode == DIS_VIDEO_OUT_LEVEL) {  // the engine's own output, straight from the handle's workspace
    CUV(cudaMemcpyAsync(flow_out, dis_level_flow_ptr(v->eng[k], 0), (size_t)v->lw * v->lh * 2 * sizeof(float),
                        cudaMemcpyDeviceToHost, st));
  } else {
    CUV(cudaMemcpyAsync(flow_out, v->d_flow[k], v->flow_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CUV(cudaEventRecord(v->done[k], st));
  v->host_out[k] = flow_out;
  return DIS_OK;
}

}  // extern "C"
