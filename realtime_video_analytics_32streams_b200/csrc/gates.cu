// a12 on the device: the per-stream gates of StreamWorker._process_packet (pipeline.py:143-170) and the adaptive-FPS
// state machine (pipeline.py:104-116, 242-262) as two one-warp kernels, so that a tick whose streams use the motion
// gate or adaptive FPS needs no host round trip between the motion kernel and the letterbox: the decisions stay in
// HBM as a uint8 skip mask that the letterbox, decode, NMS and tracker kernels read (b200va_set_skip_mask), and the
// host reads them back with the result tables.
//
//   k_gates_decide   frame_index += 1; motion: count / pixels >= threshold in float64 (frame_filter.py:38-40; the
//                    first frame of a stream, count -1, is always processed: :33-35); then the adaptive-FPS test
//                    (frame_index - 1) % process_every (pipeline.py:165-170).  Motion is tested first, like the
//                    reference, which returns from the motion branch before it looks at process_every.
//   k_gates_commit   _adjust_adaptive_state(len(filtered), len(tracks)) (pipeline.py:242-262) from the detection and
//                    track counts the tick's kernels left in HBM; a skipped frame counts as zero detections
//                    (_skip_frame, pipeline.py:214-222).
#include "common.cuh"

namespace {

struct GateState {
  int32_t* frame_index;    // [max_streams]
  int32_t* idle_frames;    // [max_streams]
  int32_t* process_every;  // [max_streams], stored as process_every - 1 so that zeroed memory is the initial state
  void* base = nullptr;
};

constexpr int kGateChunk = 64;  // descriptors per launch: they travel by value in the kernel parameter block
struct GateParams {
  b200va_gate g[kGateChunk];
};
static_assert(sizeof(GateParams) <= 3800, "kernel parameter block too large");

__global__ void k_gates_decide(GateState S, const __grid_constant__ GateParams g, int batch, const int32_t* __restrict__ changed,
                               uint8_t* __restrict__ skip_out) {
  const int i = threadIdx.x;
  if (i >= batch) return;
  const b200va_gate& c = g.g[i];
  const int fi = S.frame_index[c.slot] + 1;  // self._frame_index += 1 (pipeline.py:144)
  S.frame_index[c.slot] = fi;
  int skip = B200VA_GATE_PROCESS;
  if (c.motion && c.changed_index >= 0) {
    const int cnt = changed[c.changed_index];
    // float(np.count_nonzero(thresh)) / float(thresh.size) >= threshold, IEEE double (frame_filter.py:38-40)
    if (cnt >= 0 && !(__ddiv_rn((double)cnt, (double)c.pixels) >= c.motion_threshold)) skip = B200VA_GATE_SKIP_MOTION;
  }
  if (skip == B200VA_GATE_PROCESS && c.adaptive) {
    const int pe = S.process_every[c.slot] + 1;
    if (pe > 1 && (fi - 1) % pe != 0) skip = B200VA_GATE_SKIP_ADAPTIVE;
  }
  skip_out[i] = (uint8_t)skip;
}

__global__ void k_gates_commit(GateState S, const __grid_constant__ GateParams g, int batch, const int32_t* __restrict__ det_count,
                               const int32_t* __restrict__ trk_count, const uint8_t* __restrict__ skip,
                               int32_t* __restrict__ state_out) {
  const int i = threadIdx.x;
  if (i >= batch) return;
  const b200va_gate& c = g.g[i];
  const int sk = skip ? skip[i] : 0;
  const int n_det = sk ? 0 : det_count[i];
  const int n_trk = trk_count[i];
  int idle = S.idle_frames[c.slot], pe = S.process_every[c.slot] + 1;
  if (c.adaptive) {
    if (n_det > 0 || n_trk > 0) {
      idle = 0;
      pe = 1;
    } else {
      idle += 1;
      if (idle >= c.idle_tolerance) pe = max(c.max_process_every, 1);
    }
    S.idle_frames[c.slot] = idle;
    S.process_every[c.slot] = pe - 1;
  }
  if (state_out) {
    int4 o;
    o.x = sk;
    o.y = pe;
    o.z = idle;
    o.w = S.frame_index[c.slot];
    reinterpret_cast<int4*>(state_out)[i] = o;
  }
}

__global__ void k_gates_reset(GateState S, int slot) {
  S.frame_index[slot] = 0;
  S.idle_frames[slot] = 0;
  S.process_every[slot] = 0;
}

GateState* state(b200va_ctx* h) { return (GateState*)h->gates; }

int check(b200va_ctx* h, const b200va_gate* gates, int batch, bool need_changed, const int32_t* changed) {
  REQUIRE(h, gates != nullptr, "NULL gate descriptors");
  REQUIRE(h, batch >= 0 && batch <= h->cfg.max_batch && batch <= B200VA_MAX_BATCH, "batch %d outside [0, %d]", batch, h->cfg.max_batch);
  for (int i = 0; i < batch; ++i) {
    const b200va_gate& g = gates[i];
    REQUIRE(h, g.slot >= 0 && g.slot < h->cfg.max_streams, "gate slot %d outside [0, %d)", g.slot, h->cfg.max_streams);
    if (g.motion && g.changed_index >= 0) {
      REQUIRE(h, g.pixels > 0, "gate %d: pixels must be positive", i);
      REQUIRE(h, !need_changed || changed != nullptr, "gate %d uses the motion gate but `changed` is NULL", i);
    }
    if (g.adaptive) REQUIRE(h, g.max_process_every >= 1 && g.idle_tolerance >= 1, "gate %d: bad adaptive-FPS parameters", i);
  }
  return B200VA_OK;
}

}  // namespace

int gates_create(b200va_ctx* h) {
  GateState* S = new GateState();
  h->gates = S;
  const size_t n = (size_t)h->cfg.max_streams * sizeof(int32_t);
  const size_t stride = (n + 255) & ~(size_t)255;
  CUDA_TRY(h, cudaMalloc(&S->base, 3 * stride));
  CUDA_TRY(h, cudaMemset(S->base, 0, 3 * stride));
  uint8_t* b = (uint8_t*)S->base;
  S->frame_index = (int32_t*)b;
  S->idle_frames = (int32_t*)(b + stride);
  S->process_every = (int32_t*)(b + 2 * stride);
  return B200VA_OK;
}

void gates_destroy(b200va_ctx* h) {
  GateState* S = state(h);
  if (!S) return;
  if (S->base) cudaFree(S->base);
  delete S;
  h->gates = nullptr;
}

extern "C" int b200va_set_skip_mask(b200va_handle h, const uint8_t* skip) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  h->skip_dev = skip;
  return B200VA_OK;
}

extern "C" int b200va_gates_decide(b200va_handle h, const b200va_gate* gates, int batch, const int32_t* changed,
                                   uint8_t* skip_out, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  REQUIRE(h, skip_out != nullptr, "NULL skip_out");
  const int rc = check(h, gates, batch, true, changed);
  if (rc != B200VA_OK || batch == 0) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < batch; base += kGateChunk) {
    const int n = std::min(kGateChunk, batch - base);
    GateParams g;
    memcpy(g.g, gates + base, sizeof(b200va_gate) * n);
    k_gates_decide<<<1, kGateChunk, 0, st>>>(*state(h), g, n, changed, skip_out + base);
    LAUNCH_CHECK(h);
  }
  return B200VA_OK;
}

extern "C" int b200va_gates_commit(b200va_handle h, const b200va_gate* gates, int batch, const int32_t* det_count,
                                   const int32_t* trk_count, const uint8_t* skip, int32_t* state_out, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  REQUIRE(h, det_count && trk_count, "NULL count arrays");
  const int rc = check(h, gates, batch, false, nullptr);
  if (rc != B200VA_OK || batch == 0) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < batch; base += kGateChunk) {
    const int n = std::min(kGateChunk, batch - base);
    GateParams g;
    memcpy(g.g, gates + base, sizeof(b200va_gate) * n);
    k_gates_commit<<<1, kGateChunk, 0, st>>>(*state(h), g, n, det_count + base, trk_count + base, skip ? skip + base : nullptr,
                                             state_out ? state_out + 4 * base : nullptr);
    LAUNCH_CHECK(h);
  }
  return B200VA_OK;
}

extern "C" int b200va_gates_reset(b200va_handle h, int slot, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  REQUIRE(h, slot >= 0 && slot < h->cfg.max_streams, "gate slot %d outside [0, %d)", slot, h->cfg.max_streams);
  k_gates_reset<<<1, 1, 0, (cudaStream_t)stream>>>(*state(h), slot);
  LAUNCH_CHECK(h);
  return B200VA_OK;
}
