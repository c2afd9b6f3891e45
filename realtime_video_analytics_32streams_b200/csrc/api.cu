// Library entry points that are not tied to one kernel family: handle life cycle, error text,
// launch accounting and the capacity flags the kernels raise.
#include <nvtx3/nvToolsExt.h>

#include <cstdlib>

#include "common.cuh"

int postprocess_configure(b200va_ctx* h);  // postprocess.cu
void postprocess_release(b200va_ctx* h);   // postprocess.cu
void egress_destroy(b200va_ctx* h);        // egress.cu
int preprocess_configure(b200va_ctx* h);   // preprocess.cu
int filters_configure(b200va_ctx* h);      // filters.cu
int gates_create(b200va_ctx* h);           // gates.cu
void gates_destroy(b200va_ctx* h);         // gates.cu

extern "C" int b200va_version(void) { return B200VA_VERSION; }

extern "C" const char* b200va_error_string(int status) {
  switch (status) {
    case B200VA_OK: return "ok";
    case B200VA_ERR_INVALID: return "invalid argument";
    case B200VA_ERR_CUDA: return "CUDA runtime error";
    case B200VA_ERR_CAPACITY: return "configured capacity exceeded";
    case B200VA_ERR_STATE: return "invalid call order";
    default: return "unknown status";
  }
}

const char* phase_name(int phase) {
  static const char* names[B200VA_PHASE_COUNT] = {"b200va:upload", "b200va:roi", "b200va:resize", "b200va:motion",
                                                  "b200va:preprocess", "b200va:decode", "b200va:nms", "b200va:tracker",
                                                  "b200va:dfl", "b200va:tick", "b200va:egress"};
  return phase >= 0 && phase < B200VA_PHASE_COUNT ? names[phase] : "b200va:?";
}

PhaseScope::PhaseScope(b200va_ctx* h_, int phase_, cudaStream_t st_) : h(h_), phase(phase_), st(st_) {
  nvtxRangePushA(phase_name(phase));
  if (!h->profiling) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return;
  }
  timed = cudaEventRecord(h->prof_ev[phase][0], st) == cudaSuccess;
}

PhaseScope::~PhaseScope() {
  if (timed && cudaEventRecord(h->prof_ev[phase][1], st) == cudaSuccess) h->prof_rec[phase] = true;
  nvtxRangePop();
}

extern "C" int b200va_set_profiling(b200va_handle h, int enable) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  if (enable && !h->prof_ev[0][0])
    for (int p = 0; p < B200VA_PHASE_COUNT; ++p)
      for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaEventCreate(&h->prof_ev[p][k]));
  h->profiling = enable != 0;
  for (int p = 0; p < B200VA_PHASE_COUNT; ++p) h->prof_rec[p] = false;
  return B200VA_OK;
}

extern "C" int b200va_get_phase_times(b200va_handle h, float* ms) {
  if (!h || !ms) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  for (int p = 0; p < B200VA_PHASE_COUNT; ++p) {
    ms[p] = -1.f;
    if (!h->prof_rec[p]) continue;
    h->prof_rec[p] = false;
    CUDA_TRY(h, cudaEventSynchronize(h->prof_ev[p][1]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[p], h->prof_ev[p][0], h->prof_ev[p][1]));
  }
  return B200VA_OK;
}

static int env_int(const char* name) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : 0;
}

static int create_impl(b200va_ctx* h) {
  const b200va_config& c = h->cfg;
  h->tune.decode_impl = env_int("B200VA_DECODE_IMPL");
  h->tune.decode_ta = env_int("B200VA_DECODE_TA");
  h->tune.decode_rows = env_int("B200VA_DECODE_ROWS");
  h->tune.decode_stages = env_int("B200VA_DECODE_STAGES");
  h->tune.decode_ctas_per_sm = env_int("B200VA_DECODE_CTAS_PER_SM");
  if (getenv("B200VA_FUSE_POST_TRACK")) h->tune.fuse_post_track = env_int("B200VA_FUSE_POST_TRACK");
  if (getenv("B200VA_PDL")) h->tune.pdl = env_int("B200VA_PDL");
  h->tune.uniform_carveout = env_int("B200VA_UNIFORM_CARVEOUT");
  h->tune.trk_smem_tracks = env_int("B200VA_TRK_SMEM_TRACKS");
  h->tune.trk_max_threads = env_int("B200VA_TRK_THREADS");
  h->tune.dense_impl = env_int("B200VA_DENSE_IMPL");
  h->tune.dense_ctas_per_sm = env_int("B200VA_DENSE_CTAS");
  if (getenv("B200VA_LB_SMEM_FLOOR")) h->tune.lb_smem_floor = env_int("B200VA_LB_SMEM_FLOOR");
  if (getenv("B200VA_DENSE_CARVEOUT")) h->tune.dense_carveout = env_int("B200VA_DENSE_CARVEOUT");
  if (getenv("B200VA_POST_CARVEOUT")) h->tune.post_carveout = env_int("B200VA_POST_CARVEOUT");
  REQUIRE(h, c.max_batch >= 1 && c.max_batch <= B200VA_MAX_BATCH, "max_batch must be in [1, %d]", B200VA_MAX_BATCH);
  REQUIRE(h, c.max_anchors >= 1 && c.max_anchors <= 262144, "max_anchors must be in [1, 262144]");
  REQUIRE(h, c.max_candidates >= 1 && c.max_candidates <= 8192, "max_candidates must be in [1, 8192]");
  REQUIRE(h, c.max_dets >= 1 && c.max_dets <= c.max_candidates, "max_dets must be in [1, max_candidates]");
  REQUIRE(h, c.max_streams >= 1 && c.max_streams <= 4096, "max_streams must be in [1, 4096]");
  REQUIRE(h, c.max_tracks >= 1 && c.max_tracks <= 4096, "max_tracks must be in [1, 4096]");
  int ndev = 0;
  CUDA_TRY(h, cudaGetDeviceCount(&ndev));
  REQUIRE(h, c.device >= 0 && c.device < ndev, "device %d not present (%d visible)", c.device, ndev);
  CUDA_TRY(h, cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, c.device));
  REQUIRE(h, prop.major >= 10, "device %d is sm_%d%d; this library is built for sm_100a only", c.device, prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  const size_t frames = c.max_batch < B200VA_LAUNCH_FRAMES ? c.max_batch : B200VA_LAUNCH_FRAMES;
  const size_t n = frames * (size_t)c.max_candidates;
  h->cand_set_elems = n;
  h->cand_set_frames = (int)frames;
  CUDA_TRY(h, cudaMalloc(&h->cand_key, 2 * n * sizeof(unsigned long long)));
  CUDA_TRY(h, cudaMalloc(&h->cand_box, 2 * n * sizeof(float4)));
  CUDA_TRY(h, cudaMalloc(&h->cand_cls, 2 * n * sizeof(int32_t)));
  CUDA_TRY(h, cudaMalloc(&h->cand_count, 2 * frames * sizeof(int32_t)));
  CUDA_TRY(h, cudaMemset(h->cand_count, 0, 2 * frames * sizeof(int32_t)));
  CUDA_TRY(h, cudaMalloc(&h->status_flags, FLAG_COUNT * sizeof(int32_t)));
  CUDA_TRY(h, cudaMemset(h->status_flags, 0, FLAG_COUNT * sizeof(int32_t)));
  CUDA_TRY(h, cudaMalloc(&h->roi_scratch, ROI_SCRATCH_BYTES));
  CUDA_TRY(h, cudaMalloc(&h->dbg, DBG_SLOTS * sizeof(long long)));
  CUDA_TRY(h, cudaMemset(h->dbg, 0, DBG_SLOTS * sizeof(long long)));
  int rc = tap_cache_create(h);
  if (rc) return rc;
  rc = preprocess_configure(h);
  if (rc) return rc;
  rc = postprocess_configure(h);
  if (rc) return rc;
  rc = filters_configure(h);
  if (rc) return rc;
  rc = tracker_state_create(h);
  if (rc) return rc;
  rc = gates_create(h);
  if (rc) return rc;
  if (cudaHostAlloc((void**)&h->nms_stats_host, 64, cudaHostAllocMapped) == cudaSuccess) {
    memset(h->nms_stats_host, 0, 64);
    if (cudaHostGetDevicePointer((void**)&h->nms_stats_dev, h->nms_stats_host, 0) != cudaSuccess) h->nms_stats_dev = nullptr;
  } else {
    cudaGetLastError();
    h->nms_stats_host = nullptr;
  }
  int prio_lo = 0, prio_hi = 0;
  CUDA_TRY(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CUDA_TRY(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_hi));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_decoded, cudaEventDisableTiming));
  CUDA_TRY(h, cudaDeviceSynchronize());
  return B200VA_OK;
}

static std::string g_create_error;

extern "C" int b200va_create(const b200va_config* cfg, b200va_handle* out) {
  if (!cfg || !out) return B200VA_ERR_INVALID;
  *out = nullptr;
  b200va_ctx* h = new b200va_ctx();
  h->cfg = *cfg;
  int prev = -1;
  cudaGetDevice(&prev);
  int rc = create_impl(h);
  if (prev >= 0) cudaSetDevice(prev);
  if (rc != B200VA_OK) {
    g_create_error = h->last_error;
    b200va_destroy(h);
    return rc;
  }
  *out = h;
  return B200VA_OK;
}

extern "C" int b200va_destroy(b200va_handle h) {
  if (!h) return B200VA_OK;
  {
    DeviceGuard guard(h->cfg.device);
    cudaDeviceSynchronize();
    postprocess_release(h);
    egress_destroy(h);
    tracker_state_destroy(h);
    gates_destroy(h);
    tap_cache_destroy(h);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_decoded) cudaEventDestroy(h->ev_decoded);
    for (int p = 0; p < B200VA_PHASE_COUNT; ++p)
      for (int k = 0; k < 2; ++k)
        if (h->prof_ev[p][k]) cudaEventDestroy(h->prof_ev[p][k]);
    if (h->nms_stats_host) cudaFreeHost(h->nms_stats_host);
    if (h->cand_key) cudaFree(h->cand_key);
    if (h->cand_box) cudaFree(h->cand_box);
    if (h->cand_cls) cudaFree(h->cand_cls);
    if (h->cand_count) cudaFree(h->cand_count);
    if (h->status_flags) cudaFree(h->status_flags);
    if (h->roi_scratch) cudaFree(h->roi_scratch);
    if (h->dbg) cudaFree(h->dbg);
  }
  delete h;
  return B200VA_OK;
}

extern "C" const char* b200va_last_error(b200va_handle h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

extern "C" int64_t b200va_launch_count(b200va_handle h) { return h ? h->launches.load() : 0; }

extern "C" int b200va_poll_status(b200va_handle h, void* stream) {
  if (!h) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  int32_t flags[FLAG_COUNT];
  CUDA_TRY(h, cudaMemcpyAsync(flags, h->status_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  if (flags[FLAG_CAND_OVERFLOW] || flags[FLAG_DET_OVERFLOW] || flags[FLAG_TRACK_OVERFLOW] || flags[FLAG_TRACK_ROWS]) {
    CUDA_TRY(h, cudaMemsetAsync(h->status_flags, 0, sizeof(flags), st));
    return set_error(h, B200VA_ERR_CAPACITY, "capacity exceeded since the last poll:%s%s%s%s",
                     flags[FLAG_CAND_OVERFLOW] ? " candidates>max_candidates" : "",
                     flags[FLAG_DET_OVERFLOW] ? " detections>max_dets" : "",
                     flags[FLAG_TRACK_OVERFLOW] ? " tracks>max_tracks" : "",
                     flags[FLAG_TRACK_ROWS] ? " tracks>output rows" : "");
  }
  return B200VA_OK;
}

static_assert(B200VA_STATUS_WORDS == FLAG_COUNT, "status word count");
extern "C" int b200va_read_status_async(b200va_handle h, int32_t* host_out, int clear, void* stream) {
  if (!h || !host_out) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(h, cudaMemcpyAsync(host_out, h->status_flags, sizeof(int32_t) * FLAG_COUNT, cudaMemcpyDeviceToHost, st));
  if (clear) CUDA_TRY(h, cudaMemsetAsync(h->status_flags, 0, sizeof(int32_t) * FLAG_COUNT, st));
  return B200VA_OK;
}

// Developer aid: SM-clock stamps written by builds compiled with -DB200VA_PHASE_TIMING (zeros otherwise).
extern "C" B200VA_API int b200va_debug_read(b200va_handle h, int64_t* out, int n) {
  if (!h || !out || n < 0 || n > DBG_SLOTS) return B200VA_ERR_INVALID;
  DeviceGuard guard(h->cfg.device);
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaMemcpy(out, h->dbg, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  return B200VA_OK;
}

// Developer aid: arm the timeline slots (min-start slots to +inf, max-end slots to 0) before a measured tick.
extern "C" B200VA_API int b200va_debug_reset(b200va_handle h) {
  if (!h) return B200VA_ERR_INVALID;
  DeviceGuard guard(h->cfg.device);
  CUDA_TRY(h, cudaDeviceSynchronize());
  long long host[DBG_SLOTS];
  memset(host, 0, sizeof(host));
  for (int s = 40; s < 48; s += 2) host[s] = -1ll;  // all ones = UINT64_MAX for atomicMin
  CUDA_TRY(h, cudaMemcpy(h->dbg, host, sizeof(host), cudaMemcpyHostToDevice));
  return B200VA_OK;
}
