// b200va_tick: the letterbox of the next detector batch and the post-process + tracker update of
// the last head tensor as two branches of one call (fork / join on events, capturable in a graph).
#include "common.cuh"

extern "C" int b200va_tick(b200va_handle h, const b200va_tick_args* a, void* stream) {
  if (!h || !a) return B200VA_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lock(h->mu);
  DeviceGuard guard(h->cfg.device);
  REQUIRE(h, a->schedule >= 0 && a->schedule <= 2, "unknown schedule %d", a->schedule);
  cudaStream_t main_st = (cudaStream_t)stream;
  const bool has_pre = a->frames != nullptr && a->batch > 0;
  const bool has_post = a->head != nullptr && a->head_batch > 0;
  const bool has_trk = a->stream_slots != nullptr && a->trk_batch > 0;
  const bool fork = a->schedule != 0 && has_pre && (has_post || has_trk);
  cudaStream_t post_st = fork ? h->side_stream : main_st;
  int rc = B200VA_OK;
  if (fork) {
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, main_st));
    CUDA_TRY(h, cudaStreamWaitEvent(post_st, h->ev_fork, 0));
  }
  if (has_post) {
    h->hook_after_decode = (fork && a->schedule == 1) ? h->ev_decoded : nullptr;
    rc = b200va_postprocess(h, a->head, a->layout, a->head_batch, a->channels, a->anchors, a->meta, a->conf_thr,
                            a->iou_thr, a->classes, a->n_classes, a->score_mode, a->nms_mode, a->filter_conf_thr_f64,
                            a->use_filter, a->dets, post_st);
    const bool hooked = h->hook_after_decode != nullptr;
    h->hook_after_decode = nullptr;
    if (rc == B200VA_OK && hooked && a->channels >= 5 && a->anchors > 0)
      CUDA_TRY(h, cudaStreamWaitEvent(main_st, h->ev_decoded, 0));
  }
  if (rc == B200VA_OK && has_trk)
    rc = b200va_tracker_update(h, a->stream_slots, a->trk_batch, a->dets, a->max_dets, a->det_scale, a->skip, a->trk_cfg,
                               a->id_base, a->tracks, a->new_counts, post_st);
  if (fork) CUDA_TRY(h, cudaEventRecord(h->ev_join, post_st));  // always rejoin, also after an error
  if (rc == B200VA_OK && has_pre) {
    if (a->ev_pre_begin) CUDA_TRY(h, cudaEventRecord((cudaEvent_t)a->ev_pre_begin, main_st));
    rc = b200va_preprocess(h, a->frames, a->src_h, a->src_w, a->src_pitch, a->batch, a->roi_masks, a->net_out, a->dst_h,
                           a->dst_w, a->out_format, a->meta_out, main_st);
    if (rc == B200VA_OK && a->ev_pre_end) CUDA_TRY(h, cudaEventRecord((cudaEvent_t)a->ev_pre_end, main_st));
  }
  if (fork) CUDA_TRY(h, cudaStreamWaitEvent(main_st, h->ev_join, 0));
  return rc;
}
