/* A plain C consumer of libb200va.so: what a non-Python host (or a cgo / JNI shim) links against.
 * Reads one BGR frame and one decoded head from files, runs letterbox -> post-process -> tracker
 * through the C ABI only (include/b200va.h; CUDA runtime just for cudaMalloc / cudaMemcpy), and writes
 * the results to files for tests/test_c_consumer.py to compare with the oracle.
 *
 *   cabi_consumer frame.bin H W head.bin C A out_dir
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200va.h"

#define CHECK_CUDA(x)                                                              \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                     \
      return 2;                                                                    \
    }                                                                              \
  } while (0)
#define CHECK_VA(h, x)                                                             \
  do {                                                                             \
    int rc_ = (x);                                                                 \
    if (rc_ != B200VA_OK) {                                                        \
      fprintf(stderr, "%s: %s (%s)\n", #x, b200va_error_string(rc_), b200va_last_error(h)); \
      return 3;                                                                    \
    }                                                                              \
  } while (0)

static void* read_file(const char* path, size_t bytes) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  void* p = malloc(bytes);
  if (p && fread(p, 1, bytes, f) != bytes) {
    free(p);
    p = NULL;
  }
  fclose(f);
  return p;
}

static int write_file(const char* dir, const char* name, const void* data, size_t bytes) {
  char path[1024];
  snprintf(path, sizeof(path), "%s/%s", dir, name);
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  const int bad = fwrite(data, 1, bytes, f) != bytes;
  fclose(f);
  return bad;
}

int main(int argc, char** argv) {
  if (argc != 8) {
    fprintf(stderr, "usage: %s frame.bin H W head.bin C A out_dir\n", argv[0]);
    return 1;
  }
  const int H = atoi(argv[2]), W = atoi(argv[3]), C = atoi(argv[5]), A = atoi(argv[6]);
  const char* out_dir = argv[7];
  const size_t frame_bytes = (size_t)H * W * 3, head_bytes = (size_t)C * A * sizeof(float);
  uint8_t* frame = (uint8_t*)read_file(argv[1], frame_bytes);
  float* head = (float*)read_file(argv[4], head_bytes);
  if (!frame || !head) {
    fprintf(stderr, "cannot read inputs\n");
    return 1;
  }
  enum { MAX_DETS = 256, MAX_TRACKS = 256, IN = 640 };
  b200va_config cfg = {0, 1, 0, 2048, MAX_DETS, 1, MAX_TRACKS};
  cfg.max_anchors = A;
  b200va_handle h = NULL;
  int rc = b200va_create(&cfg, &h);
  if (rc != B200VA_OK) {
    fprintf(stderr, "b200va_create: %s (%s)\n", b200va_error_string(rc), b200va_last_error(NULL));
    return 3;
  }

  uint8_t* d_frame;
  float *d_head, *d_net;
  CHECK_CUDA(cudaMalloc((void**)&d_frame, frame_bytes));
  CHECK_CUDA(cudaMalloc((void**)&d_head, head_bytes));
  CHECK_CUDA(cudaMalloc((void**)&d_net, (size_t)3 * IN * IN * sizeof(float)));
  CHECK_CUDA(cudaMemcpy(d_frame, frame, frame_bytes, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_head, head, head_bytes, cudaMemcpyHostToDevice));

  /* a1: letterbox */
  const uint8_t* frames[1] = {d_frame};
  b200va_letterbox meta;
  CHECK_VA(h, b200va_preprocess(h, frames, &H, &W, NULL, 1, NULL, d_net, IN, IN, B200VA_OUT_F32_RGB_NCHW, &meta, NULL));

  /* a3-a7: decode + NMS (+ the float64 re-threshold of pipeline.py:182) */
  b200va_dets dets;
  CHECK_CUDA(cudaMalloc((void**)&dets.bbox_xyxy, MAX_DETS * 4 * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dets.conf, MAX_DETS * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dets.cls, MAX_DETS * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc((void**)&dets.count, sizeof(int32_t)));
  CHECK_VA(h, b200va_postprocess(h, d_head, B200VA_HEAD_CHANNEL_MAJOR, 1, C, A, &meta, 0.35, 0.5, NULL, 0,
                                 B200VA_SCORE_REF_COMPAT, B200VA_NMS_AGNOSTIC, 0.35, 1, &dets, NULL));

  /* a8: tracker, two updates with the same detections (the second one matches every track) */
  b200va_tracks trk;
  memset(&trk, 0, sizeof(trk)); /* rows = 0: the arrays hold max_tracks rows per stream */
  int32_t* d_new;
  CHECK_CUDA(cudaMalloc((void**)&trk.track_id, MAX_TRACKS * sizeof(int64_t)));
  CHECK_CUDA(cudaMalloc((void**)&trk.cls, MAX_TRACKS * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc((void**)&trk.conf, MAX_TRACKS * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&trk.bbox_xyxy, MAX_TRACKS * 4 * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&trk.age, MAX_TRACKS * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc((void**)&trk.hits, MAX_TRACKS * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc((void**)&trk.count, sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc((void**)&d_new, sizeof(int32_t)));
  const int slot = 0;
  const b200va_tracker_cfg tcfg = {30, 1, 0.5};
  for (int rep = 0; rep < 2; ++rep)
    CHECK_VA(h, b200va_tracker_update(h, &slot, 1, &dets, MAX_DETS, NULL, NULL, &tcfg, NULL, &trk, d_new, NULL));
  CHECK_VA(h, b200va_poll_status(h, NULL));

  /* results back to the host, into files */
  float* net = (float*)malloc((size_t)3 * IN * IN * sizeof(float));
  float box[MAX_DETS * 4], conf[MAX_DETS];
  int32_t cls[MAX_DETS], n_det = 0, n_trk = 0, hits[MAX_TRACKS];
  int64_t ids[MAX_TRACKS];
  double tbox[MAX_TRACKS * 4];
  CHECK_CUDA(cudaMemcpy(net, d_net, (size_t)3 * IN * IN * sizeof(float), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(&n_det, dets.count, sizeof(n_det), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(box, dets.bbox_xyxy, sizeof(box), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(conf, dets.conf, sizeof(conf), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(cls, dets.cls, sizeof(cls), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(&n_trk, trk.count, sizeof(n_trk), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(ids, trk.track_id, sizeof(ids), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(hits, trk.hits, sizeof(hits), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(tbox, trk.bbox_xyxy, sizeof(tbox), cudaMemcpyDeviceToHost));
  int bad = 0;
  bad |= write_file(out_dir, "net.bin", net, (size_t)3 * IN * IN * sizeof(float));
  bad |= write_file(out_dir, "det_box.bin", box, (size_t)n_det * 4 * sizeof(float));
  bad |= write_file(out_dir, "det_conf.bin", conf, (size_t)n_det * sizeof(float));
  bad |= write_file(out_dir, "det_cls.bin", cls, (size_t)n_det * sizeof(int32_t));
  bad |= write_file(out_dir, "trk_id.bin", ids, (size_t)n_trk * sizeof(int64_t));
  bad |= write_file(out_dir, "trk_hits.bin", hits, (size_t)n_trk * sizeof(int32_t));
  bad |= write_file(out_dir, "trk_box.bin", tbox, (size_t)n_trk * 4 * sizeof(double));
  printf("meta new=%dx%d pad=(%d,%d) scale=%.17g dets=%d tracks=%d launches=%lld\n", meta.new_w, meta.new_h, meta.pad_left,
         meta.pad_top, meta.scale, n_det, n_trk, (long long)b200va_launch_count(h));
  b200va_destroy(h);
  return bad ? 4 : 0;
}
