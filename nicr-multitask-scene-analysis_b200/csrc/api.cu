// api.cu -- C-ABI glue: error strings, launch-error bookkeeping, the one-call pipeline
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "finalize.cuh"

namespace npb {

static thread_local char g_last_cuda_error[256] = "";

bool dependent_launch_enabled()
{
    static const bool on = [] {
        const char *e = getenv("NPB_NO_PDL");
        return !(e && e[0] && e[0] != '0');
    }();
    return on;
}

#ifdef NPB_TIMELINE
TimelineSlot *timeline_buffer()
{
    static TimelineSlot *buf = nullptr;
    if (!buf) {
        cudaMalloc(&buf, 16 * sizeof(TimelineSlot));
        TimelineSlot init[16];
        for (int i = 0; i < 16; ++i) init[i] = TimelineSlot{~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
        cudaMemcpy(buf, init, sizeof(init), cudaMemcpyHostToDevice);
    }
    return buf;
}
#endif

int record_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return NPB_OK;
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
    return NPB_ERR_CUDA;
}


}  // namespace npb

using namespace npb;

extern "C" int npb_abi_version(void) { return 3; }

extern "C" const char *npb_last_cuda_error(void) { return g_last_cuda_error; }

#ifdef NPB_TIMELINE
// debug builds only: copy the 16 x 6 timeline words to the host and reset them (synchronises)
extern "C" int npb_timeline_read(unsigned long long *h_out)
{
    TimelineSlot *buf = timeline_buffer();
    cudaDeviceSynchronize();
    cudaMemcpy(h_out, buf, 16 * sizeof(TimelineSlot), cudaMemcpyDeviceToHost);
    TimelineSlot init[16];
    for (int i = 0; i < 16; ++i) init[i] = TimelineSlot{~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
    cudaMemcpy(buf, init, sizeof(init), cudaMemcpyHostToDevice);
    return NPB_OK;
}
#endif

extern "C" const char *npb_build_info(void)
{
    static char info[96];
    snprintf(info, sizeof(info), "sm_100a pdl=%d debug=%d", dependent_launch_enabled() ? 1 : 0,
#ifdef NPB_DEBUG
             1
#else
             0
#endif
    );
    return info;
}

extern "C" const char *npb_error_string(int code)
{
    switch (code) {
        case NPB_OK: return "ok";
        case NPB_ERR_ARG: return "invalid argument";
        case NPB_ERR_TOO_MANY_CENTERS: return "more than 255 instance centres in a frame";
        case NPB_ERR_ZERO_DIVISION: return "division by zero (union of segments is empty)";
        case NPB_ERR_CATEGORY_RANGE: return "class / category id out of range";
        case NPB_ERR_CAPACITY: return "fixed-capacity table overflowed";
        case NPB_ERR_CUDA: return "CUDA error";
        default: return "unknown error";
    }
}

// workspace layout of npb_panoptic_forward: [centres scratch | vote_hist | ori_sum]
extern "C" size_t npb_panoptic_forward_workspace_bytes(int B, int C, int H, int W,
                                                       int nms_kernel_size)
{
    size_t bytes = align256(npb_instance_centers_workspace_bytes(B, H, W, nms_kernel_size));
    bytes += align256((size_t)B * kMaxInst * C * sizeof(uint32_t));
    bytes += align256((size_t)B * kMaxInst * 2 * sizeof(double));
    return bytes;
}

extern "C" int npb_panoptic_forward_workspace_init(void *workspace, int B, int C, int H, int W,
                                                   int nms_kernel_size, void *stream)
{
    if (!workspace || B < 1) return NPB_ERR_ARG;
    cudaMemsetAsync(workspace, 0, npb_panoptic_forward_workspace_bytes(B, C, H, W, nms_kernel_size),
                    (cudaStream_t)stream);
    return record_launch("npb_panoptic_forward_workspace_init");
}

static int panoptic_forward_impl(
    const float *logits, const float *heat, const float *offset, const float *orientation, int B,
    int C, int H, int W, const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
    float threshold, int nms_kernel_size, int top_k, int apply_fg_mask, int normalized_offset,
    int use_distance_threshold, float distance_threshold, int64_t max_instances_per_category,
    void *workspace, uint8_t *sem_out, uint8_t *inst_out, int64_t *pan_out, uint8_t *pan_sem_out,
    int32_t *centers_yx, int32_t *n_centers, float *center_score, int32_t *inst_class,
    int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle, int32_t *status,
    const npb_eval_args *eval, const npb_eval_args *pending, int pending_B, bool pipelined,
    void *stream)
{
    if (!logits || !heat || !offset || !workspace || !sem_out || !inst_out || !pan_out ||
        !centers_yx || !n_centers || !center_score || !inst_class || !inst_pan_id || !inst_area ||
        !status || !h_thing_lut)
        return NPB_ERR_ARG;
    if (C < 1 || C > 255) return NPB_ERR_ARG;
    if (orientation && (!inst_angle || !h_orientation_lut)) return NPB_ERR_ARG;
    char *ws = (char *)workspace;
    void *ws_centers = ws;
    ws += align256(npb_instance_centers_workspace_bytes(B, H, W, nms_kernel_size));
    uint32_t *vote_hist = (uint32_t *)ws;
    ws += align256((size_t)B * kMaxInst * C * sizeof(uint32_t));
    double *ori_sum = orientation ? (double *)ws : nullptr;

    // No memset in the chain: the counters of the centre detection are left at zero by every
    // call (npb_panoptic_forward_workspace_init zeroes them once), the vote histograms /
    // orientation sums and the cleared part of the evaluation workspace are zeroed by the NMS
    // pass itself -- nothing reads or writes them before that grid has completed.  Every kernel
    // of the chain is launched as a programmatic dependent of its predecessor (common.cuh).
    ScratchToClear scratch;
    scratch.p0 = vote_hist;
    scratch.bytes0 = align256((size_t)B * kMaxInst * C * sizeof(uint32_t)) +
                     (orientation ? (size_t)B * kMaxInst * 2 * sizeof(double) : 0);
    scratch.p1 = nullptr;
    scratch.bytes1 = 0;
    if (eval) {
        if (!eval->workspace) return NPB_ERR_ARG;
        // pipelined: the hand-over tables are zero at rest (the matcher cleans up behind itself) and
        // the matcher of the PREVIOUS call may still be reading them while this call's NMS runs
        if (!pipelined)
            pq_cleared_range(eval->workspace, B, eval->num_categories, max_instances_per_category,
                             &scratch.p1, &scratch.bytes1);
    }
    int rc;
    // Pipelined evaluation: the matcher of the PREVIOUS call (1 CTA per frame of pure latency,
    // nothing for the other SMs to do) is the first kernel of this chain.  It lets its dependents
    // start as soon as its own dependency wait has passed, and the NMS pass -- which needs nothing
    // from it -- then runs beside it and only waits for it at its END, so that "NMS complete"
    // still implies "matcher complete" for the pixel pass of this call, the next writer of the
    // hand-over tables.  A plain chain of programmatic dependent launches: eager or captured.
    bool nms_late_wait = false;
    if (pending) {
        rc = pq_match_impl(pending, pending_B, max_instances_per_category, stream);
        if (rc != NPB_OK) return rc;
        nms_late_wait = !apply_fg_mask;     // (the arg-max pass of that option is a plain launch)
    }
    const uint8_t *fg = nullptr;
    const float *group_logits = logits;
    const uint8_t *group_sem = nullptr;
    if (apply_fg_mask) {
        // the centre mask needs the thing mask first (instance.py:142-143): run the arg-max
        // as its own pass and reuse inst_out as the temporary foreground map
        rc = npb_semantic_argmax(logits, B, C, H, W, sem_out, nullptr, stream);
        if (rc != NPB_OK) return rc;
        rc = npb_thing_mask(sem_out, (int64_t)B * H * W, C, h_thing_lut, inst_out, stream);
        if (rc != NPB_OK) return rc;
        fg = inst_out;
        group_logits = nullptr;
        group_sem = sem_out;
    }
    rc = instance_centers_impl(heat, B, H, W, threshold, nms_kernel_size, top_k, fg, apply_fg_mask,
                               ws_centers, centers_yx, n_centers, center_score, status, true, true,
                               &scratch, nms_late_wait, stream);
    if (rc != NPB_OK) return rc;
    rc = group_pixels_impl(group_logits, group_sem, nullptr, offset, orientation, B, C, H, W,
                           h_thing_lut, centers_yx, n_centers, normalized_offset,
                           use_distance_threshold, distance_threshold, sem_out, inst_out, vote_hist,
                           ori_sum, true, stream);
    if (rc != NPB_OK) return rc;
    if (eval) {     // instance tables, ids written and evaluated in one pass
        FinalizeParams f;
        f.vote_hist = vote_hist; f.ori_sum = ori_sum; f.n_centers = n_centers; f.C = C;
        f.class_offset = 1; f.L = (long long)max_instances_per_category; f.void_label = 0;
        f.orient = orientation_class_set(h_orientation_lut, C, 1);
        f.inst_class = inst_class; f.inst_pan_id = inst_pan_id; f.inst_area = inst_area;
        f.inst_angle = inst_angle;
        return write_panoptic_eval_impl(sem_out, inst_out, inst_pan_id, inst_class, B, C, H, W,
                                        h_thing_lut, max_instances_per_category, pan_out,
                                        pan_sem_out, eval, &f, true, pipelined, stream);
    }
    rc = npb_finalize_instances(vote_hist, ori_sum, n_centers, B, C, 1, max_instances_per_category,
                                0, h_orientation_lut, inst_class, inst_pan_id, inst_area,
                                inst_angle, stream);
    if (rc != NPB_OK) return rc;
    return npb_write_panoptic(sem_out, inst_out, inst_pan_id, inst_class, B, C, H, W, h_thing_lut,
                              max_instances_per_category, pan_out, pan_sem_out, stream);
}

extern "C" int npb_panoptic_forward(
    const float *logits, const float *heat, const float *offset, const float *orientation, int B,
    int C, int H, int W, const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
    float threshold, int nms_kernel_size, int top_k, int apply_fg_mask, int normalized_offset,
    int use_distance_threshold, float distance_threshold, int64_t max_instances_per_category,
    void *workspace, uint8_t *sem_out, uint8_t *inst_out, int64_t *pan_out, uint8_t *pan_sem_out,
    int32_t *centers_yx, int32_t *n_centers, float *center_score, int32_t *inst_class,
    int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle, int32_t *status, void *stream)
{
    return panoptic_forward_impl(logits, heat, offset, orientation, B, C, H, W, h_thing_lut,
                                 h_orientation_lut, threshold, nms_kernel_size, top_k, apply_fg_mask,
                                 normalized_offset, use_distance_threshold, distance_threshold,
                                 max_instances_per_category, workspace, sem_out, inst_out, pan_out,
                                 pan_sem_out, centers_yx, n_centers, center_score, inst_class,
                                 inst_pan_id, inst_area, inst_angle, status, nullptr, nullptr, 0, false,
                                 stream);
}

extern "C" int npb_panoptic_forward_eval(
    const float *logits, const float *heat, const float *offset, const float *orientation, int B,
    int C, int H, int W, const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
    float threshold, int nms_kernel_size, int top_k, int apply_fg_mask, int normalized_offset,
    int use_distance_threshold, float distance_threshold, int64_t max_instances_per_category,
    void *workspace, uint8_t *sem_out, uint8_t *inst_out, int64_t *pan_out, uint8_t *pan_sem_out,
    int32_t *centers_yx, int32_t *n_centers, float *center_score, int32_t *inst_class,
    int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle, int32_t *status,
    const npb_eval_args *eval, void *stream)
{
    if (!eval) return NPB_ERR_ARG;
    return panoptic_forward_impl(logits, heat, offset, orientation, B, C, H, W, h_thing_lut,
                                 h_orientation_lut, threshold, nms_kernel_size, top_k, apply_fg_mask,
                                 normalized_offset, use_distance_threshold, distance_threshold,
                                 max_instances_per_category, workspace, sem_out, inst_out, pan_out,
                                 pan_sem_out, centers_yx, n_centers, center_score, inst_class,
                                 inst_pan_id, inst_area, inst_angle, status, eval, nullptr, 0, false,
                                 stream);
}

extern "C" int npb_panoptic_forward_eval_pipelined(
    const float *logits, const float *heat, const float *offset, const float *orientation, int B,
    int C, int H, int W, const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
    float threshold, int nms_kernel_size, int top_k, int apply_fg_mask, int normalized_offset,
    int use_distance_threshold, float distance_threshold, int64_t max_instances_per_category,
    void *workspace, uint8_t *sem_out, uint8_t *inst_out, int64_t *pan_out, uint8_t *pan_sem_out,
    int32_t *centers_yx, int32_t *n_centers, float *center_score, int32_t *inst_class,
    int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle, int32_t *status,
    const npb_eval_args *eval, const npb_eval_args *pending, int pending_B, void *stream)
{
    if (!eval) return NPB_ERR_ARG;
    if (pending && pending_B < 1) return NPB_ERR_ARG;
    return panoptic_forward_impl(logits, heat, offset, orientation, B, C, H, W, h_thing_lut,
                                 h_orientation_lut, threshold, nms_kernel_size, top_k, apply_fg_mask,
                                 normalized_offset, use_distance_threshold, distance_threshold,
                                 max_instances_per_category, workspace, sem_out, inst_out, pan_out,
                                 pan_sem_out, centers_yx, n_centers, center_score, inst_class,
                                 inst_pan_id, inst_area, inst_angle, status, eval, pending, pending_B,
                                 true, stream);
}
