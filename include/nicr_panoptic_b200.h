/*
 * nicr_panoptic_b200.h -- C ABI of libnicr_panoptic_b200.so
 *
 * B200 (sm_100a) implementation of the dense panoptic post-processing and
 * evaluation hot path of TUI-NICR/nicr-multitask-scene-analysis v0.3.0.
 * Reference citations are relative to src/nicr_mt_scene_analysis/ of that repo.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name starts with `h_`
 *    (host pointer, read synchronously before the call returns);
 *  - dense tensors are contiguous NCHW / (B,H,W); P = H*W;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *    the calls never synchronise;
 *  - return value: NPB_OK or a negative NPB_ERR_* for errors detectable on the
 *    host (bad arguments, launch failure).  Data-dependent errors (too many
 *    centres, table overflow, class id out of range ...) are reported through
 *    the `status` device words (one int32 per frame or per call, see each
 *    function), using the same NPB_ERR_* codes;
 *  - `status` words must be zero (NPB_OK) on entry; errors are merged with atomicMin so
 *    several calls may share one status buffer;
 *  - no function allocates device memory: scratch is passed in as `workspace`
 *    (size from the matching *_workspace_bytes function, 256-byte aligned).
 *
 * Per-instance tables use a fixed row length NPB_MAX_INST = 256 (instance ids
 * are uint8 in the reference, model/postprocessing/instance.py:236; id 0 = no
 * instance, so at most 255 centres per frame are representable).
 */
#ifndef NICR_PANOPTIC_B200_H
#define NICR_PANOPTIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPB_MAX_INST 256
#define NPB_MAX_WIDE_CENTERS 8192   /* centres of a frame kept for npb_overflow_centers */

#define NPB_OK 0
#define NPB_ERR_ARG (-1)               /* invalid argument / unsupported size            */
#define NPB_ERR_TOO_MANY_CENTERS (-2)  /* > 255 centres in a frame (uint8 ids would wrap) */
#define NPB_ERR_ZERO_DIVISION (-3)     /* PQ: union == 0 (reference raises ZeroDivisionError) */
#define NPB_ERR_CATEGORY_RANGE (-4)    /* class / category id outside [0, n)              */
#define NPB_ERR_CAPACITY (-5)          /* a fixed-capacity table overflowed               */
#define NPB_ERR_CUDA (-6)              /* CUDA launch error (see npb_last_cuda_error)     */

/* integer dtypes accepted where the reference accepts "any int tensor" */
#define NPB_U8 0
#define NPB_I16 1
#define NPB_I32 2
#define NPB_I64 3
#define NPB_BOOL 4

int npb_abi_version(void);
const char *npb_error_string(int code);
const char *npb_last_cuda_error(void);
/* "sm_100a pdl=<0|1> debug=<0|1>": programmatic dependent launches on (NPB_NO_PDL=1 in the
 * environment turns them off), library built with -DNPB_DEBUG (table-bound assertions). */
const char *npb_build_info(void);

/* ---------------------------------------------------------------------------
 * Semantic arg-max.
 * Replaces: model/postprocessing/semantic.py:52-53  (softmax(dim=1); max(dim=1)).
 * sem_out[b][p]   = first index of the maximal logit (== reference arg-max of the
 *                   soft-max, see DESIGN.md "soft-max caveat"), uint8, 0..C-1
 * score_out[b][p] = soft-max probability of that class (nullable)
 * ------------------------------------------------------------------------- */
int npb_semantic_argmax(const float *logits, int B, int C, int H, int W,
                        uint8_t *sem_out, float *score_out, void *stream);

/* probs = softmax(logits, dim=1), (B,C,H,W) f32: the `semantic_softmax_scores` entry of
 * semantic.py:52, 55 (materialised on demand only). */
int npb_softmax(const float *logits, int B, int C, int H, int W, float *probs, void *stream);

/* ---------------------------------------------------------------------------
 * Crop to the valid region [y0, y0+Hc) x [x0, x0+Wc) and resize to (Hout, Wout).
 * Replaces: DensePostprocessingBase._crop_to_valid_region_and_resize_prediction,
 *           model/postprocessing/dense_base.py:15-58 (F.interpolate).
 * npb_resize_nearest: `planes` planes of elem_size 1 / 4 / 8 bytes (index maps, masks, score
 *   maps; panoptic.py:246-291, instance.py:385-397) -- exact.
 * npb_resize_bilinear: f32 planes, align_corners=False (semantic.py:63-66, lazily, only when
 *   the resized logits themselves are read).
 * npb_semantic_argmax_resized: bilinear resize of the C logit planes on the fly + arg-max
 *   (+ soft-max score), semantic.py:63-72 without materialising the resized logits.
 * ------------------------------------------------------------------------- */
int npb_resize_nearest(const void *src, int elem_size, int planes, int Hin, int Win, int y0,
                       int x0, int Hc, int Wc, int Hout, int Wout, void *dst, void *stream);
int npb_resize_bilinear(const float *src, int planes, int Hin, int Win, int y0, int x0, int Hc,
                        int Wc, int Hout, int Wout, float *dst, void *stream);
int npb_semantic_argmax_resized(const float *logits, int B, int C, int Hin, int Win, int y0,
                                int x0, int Hc, int Wc, int Hout, int Wout, uint8_t *sem_out,
                                float *score_out, void *stream);

/* ---------------------------------------------------------------------------
 * Class-set mask: mask_out[i] = h_class_lut[sem[i]] (uint8 0/1), N = number of pixels.
 * Replaces: torch.isin(semantic_idx, thing_class_ids), model/postprocessing/panoptic.py:123-127
 *           (and :296-300 for the orientation classes).
 * ------------------------------------------------------------------------- */
int npb_thing_mask(const uint8_t *sem, int64_t N, int C, const uint8_t *h_class_lut,
                   uint8_t *mask_out, void *stream);

/* out[i] = (int64) in[i] + add : the reference's int64 index maps (semantic.py:53 idx,
 * panoptic.py:160 `pan // L`) from the compact uint8 maps the kernels produce. */
int npb_widen_u8(const uint8_t *in, int64_t N, int64_t add, int64_t *out, void *stream);

/* ---------------------------------------------------------------------------
 * Instance-centre detection: threshold, k x k NMS (first maximum wins), top-k value,
 * raster ordered centre list.
 * Replaces: InstancePostprocessing._get_instance_centers,
 *           model/postprocessing/instance.py:78-168.
 * heat (B,1,H,W) f32.  fg (B,H,W) u8 is only read when apply_fg_mask != 0
 * (instance.py:142-143).  Outputs:
 *   centers_yx [B][256][2] int32 (y, x) in raster order, n_centers [B],
 *   center_score [B][256] f32 = heat[y][x] (instance.py:264), status [B].
 * ------------------------------------------------------------------------- */
size_t npb_instance_centers_workspace_bytes(int B, int H, int W, int nms_kernel_size);
int npb_instance_centers(const float *heat, int B, int H, int W, float threshold,
                         int nms_kernel_size, int top_k, const uint8_t *fg, int apply_fg_mask,
                         void *workspace, int32_t *centers_yx, int32_t *n_centers,
                         float *center_score, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Offset grouping (+ fused semantic arg-max, class votes and orientation sums).
 * Replaces: InstancePostprocessing._get_instance_segmentation (instance.py:170-268),
 *           the thing mask of PanopticPostprocessing (panoptic.py:118-128), the
 *           offset de-normalisation (panoptic.py:105-111 / instance.py:361-367), the
 *           per-instance histogram half of deeplab_merge (utils/panoptic_merge.py:194-199)
 *           and the sums of _get_instance_orientation (instance.py:301-310).
 * Exactly one of {logits, sem_in, fg_in} selects where "foreground" comes from:
 *   logits (B,C,H,W) f32 : arg-max here (sem_out (B,H,W) u8 is written), fg = h_thing_lut[class]
 *   sem_in (B,H,W) u8    : classes given,                                 fg = h_thing_lut[class]
 *   fg_in  (B,H,W) u8    : foreground mask given, no classes (C must be 1)
 * offset (B,2,H,W) f32 (ch0 = y, ch1 = x); orientation (B,2,H,W) f32 or NULL.
 * Outputs: inst_out (B,H,W) u8 (centre index + 1, 0 = none),
 *   vote_hist [B][256][C] u32 (pixels of instance i with class c; zeroed by the call),
 *   ori_sum  [B][256][2] f64 (sum cos, sum sin per instance; NULL iff orientation NULL).
 * ------------------------------------------------------------------------- */
int npb_group_pixels(const float *logits, const uint8_t *sem_in, const uint8_t *fg_in,
                     const float *offset, const float *orientation, int B, int C, int H, int W,
                     const uint8_t *h_thing_lut, const int32_t *centers_yx,
                     const int32_t *n_centers, int normalized_offset, int use_distance_threshold,
                     float distance_threshold, uint8_t *sem_out, uint8_t *inst_out,
                     uint32_t *vote_hist, double *ori_sum, void *stream);

/* ---------------------------------------------------------------------------
 * More than 255 centres in a frame: the reference's uint8 wrap, on request.
 * Replaces: the silent wrap of `instance_id = (instance_id + 1).type(torch.uint8)`,
 *           model/postprocessing/instance.py:231-236, with the meta dict of :253-266 that keeps
 *           growing past 255 entries (areas of the entries beyond 255 are zero).
 * A frame whose centre selection finds more than 255 centres (k-th-value ties of a saturated
 * heat-map) reports NPB_ERR_TOO_MANY_CENTERS and gets no instances.  Its complete centre list
 * (up to NPB_MAX_WIDE_CENTERS) stays in the `workspace` of that npb_instance_centers /
 * npb_panoptic_forward call until the next call on it:
 *   npb_overflow_centers   copies it out: n_out [1] (-1: beyond NPB_MAX_WIDE_CENTERS),
 *                          centers_yx [cap][2] (y, x) raster order, center_score [cap];
 *                          `heat` (B,1,H,W), B, H, W, nms_kernel_size as in the call that failed
 *                          (for npb_panoptic_forward: the head of its workspace);
 *   npb_group_pixels_wide  redoes the grouping of that ONE frame with all centres,
 *                          id = (arg-min + 1) mod 256 like the reference: inst_out (H,W) u8,
 *                          vote_hist [256][C] / ori_sum [256][2] of the frame (zeroed by the
 *                          call), n_rows_out [1] = min(n, 255) (the `n_centers` entry for
 *                          npb_finalize_instances), status [1] put back to NPB_OK.
 *                          sem_in / fg_in / offset / orientation: the frame's planes.
 * npb_finalize_instances + npb_write_panoptic (B = 1, the frame's rows) complete the frame.
 * ------------------------------------------------------------------------- */
int npb_overflow_centers(const void *workspace, const float *heat, int B, int H, int W,
                         int nms_kernel_size, int frame, int32_t *n_out, int32_t *centers_yx,
                         float *center_score, int cap, void *stream);
int npb_group_pixels_wide(const uint8_t *sem_in, const uint8_t *fg_in, const float *offset,
                          const float *orientation, int C, int H, int W,
                          const uint8_t *h_thing_lut, const int32_t *centers_yx,
                          const int32_t *n_centers, int normalized_offset,
                          int use_distance_threshold, float distance_threshold,
                          uint8_t *inst_out, uint32_t *vote_hist, double *ori_sum,
                          int32_t *n_rows_out, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Per-frame instance table: majority class (smallest class on ties), per-class running
 * instance number in ascending instance id, panoptic id, area, mean orientation.
 * Replaces: the loop of deeplab_merge_semantic_and_instance,
 *           utils/panoptic_merge.py:192-210, the area bincount instance.py:253 and
 *           atan2 of instance.py:313 / utils/_orientation.py:39-42.
 * class_offset is added to the histogram column to obtain the panoptic class
 * (1 for network classes without void, 0 when the votes already include void);
 * an instance whose panoptic class is 0 is skipped (panoptic_merge.py:201-202).
 * Outputs (rows of 256, row 0 unused):
 *   inst_class  i32 : panoptic class of the instance, -1 = no pixel / skipped
 *   inst_pan_id i64 : class * L + running number, `void_label` if skipped
 *   inst_area   i32 ; inst_angle f32 (NaN when the class has no orientation or ori_sum NULL)
 * ------------------------------------------------------------------------- */
int npb_finalize_instances(const uint32_t *vote_hist, const double *ori_sum,
                           const int32_t *n_centers, int B, int C, int class_offset,
                           int64_t max_instances_per_category, int64_t void_label,
                           const uint8_t *h_orientation_lut, int32_t *inst_class,
                           int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle,
                           void *stream);

/* ---------------------------------------------------------------------------
 * Panoptic id map.
 * Replaces: the masked assignments of utils/panoptic_merge.py:210, 213-223 and
 *           `pan // L` of panoptic.py:160.
 *   inst > 0            -> inst_pan_id[inst]
 *   inst == 0, stuff c  -> (c + 1) * L
 *   inst == 0, thing c  -> 0 (void)
 * sem (B,H,W) u8 network classes, inst (B,H,W) u8; pan_out (B,H,W) i64;
 * pan_sem_out (B,H,W) u8 = pan // L (nullable; needs inst_class from npb_finalize_instances).
 * ------------------------------------------------------------------------- */
int npb_write_panoptic(const uint8_t *sem, const uint8_t *inst, const int64_t *inst_pan_id,
                       const int32_t *inst_class, int B, int C, int H, int W,
                       const uint8_t *h_thing_lut,
                       int64_t max_instances_per_category, int64_t *pan_out,
                       uint8_t *pan_sem_out, void *stream);

/* ---------------------------------------------------------------------------
 * Whole post-processing of a batch in one call (centres -> grouping -> table -> ids).
 * Replaces: PanopticPostprocessing._postprocess_inference, panoptic.py:77-167, 294-314.
 * Same arguments as the stage functions above; `status` [B] collects stage errors.
 * The chain contains no memset (every kernel is a programmatic dependent of its predecessor):
 * `workspace` must have been prepared ONCE with npb_panoptic_forward_workspace_init for the
 * same (B, C, H, W, nms_kernel_size); every call leaves it ready for the next one.
 * ------------------------------------------------------------------------- */
size_t npb_panoptic_forward_workspace_bytes(int B, int C, int H, int W, int nms_kernel_size);
int npb_panoptic_forward_workspace_init(void *workspace, int B, int C, int H, int W,
                                        int nms_kernel_size, void *stream);
int npb_panoptic_forward(const float *logits, const float *heat, const float *offset,
                         const float *orientation, int B, int C, int H, int W,
                         const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
                         float threshold, int nms_kernel_size, int top_k, int apply_fg_mask,
                         int normalized_offset, int use_distance_threshold,
                         float distance_threshold, int64_t max_instances_per_category,
                         void *workspace, uint8_t *sem_out, uint8_t *inst_out, int64_t *pan_out,
                         uint8_t *pan_sem_out, int32_t *centers_yx, int32_t *n_centers,
                         float *center_score, int32_t *inst_class, int64_t *inst_pan_id,
                         int32_t *inst_area, float *inst_angle, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Dense score maps of `compute_scores=True`.
 * Replaces: model/postprocessing/panoptic.py:171-239.
 * Inputs: logits (B,C,H,W); pan_sem / inst (B,H,W) u8 and the per-instance tables from
 * npb_panoptic_forward.  Outputs (B,H,W) f32: semantic score (soft-max probability of the
 * pixel's panoptic class, 0 for void), instance score (centre heat of the pixel's instance),
 * panoptic score (semantic score for stuff, mean semantic score x instance score for things);
 * per instance [B][256] f32: mean semantic score, panoptic score (-1 = instance dropped).
 * inst_sum [B][256] f64 is scratch (zeroed by the call).
 * ------------------------------------------------------------------------- */
int npb_panoptic_scores(const float *logits, const uint8_t *pan_sem, const uint8_t *inst,
                        const int32_t *inst_class, const int32_t *inst_area,
                        const float *center_score, int B, int C, int H, int W, double *inst_sum,
                        float *sem_score, float *inst_score, float *pan_score,
                        float *inst_mean_sem, float *inst_pan_score, void *stream);

/* ---------------------------------------------------------------------------
 * Stand-alone deeplab merge for arbitrary semantic / instance / foreground maps.
 * Replaces: deeplab_merge_batch, utils/panoptic_merge.py:18-40, 172-225.
 * sem (B,P) int64 in [0, n_classes) (0 = void), ins (B,P) u8, fg (B,P) u8.
 * h_thing_lut [n_classes].  Outputs as npb_finalize_instances + pan_out (B,P) i64.
 * workspace: npb_deeplab_merge_workspace_bytes.  status [1].
 * ------------------------------------------------------------------------- */
size_t npb_deeplab_merge_workspace_bytes(int B, int n_classes);
int npb_deeplab_merge(const int64_t *sem, const uint8_t *ins, const uint8_t *fg, int B, int64_t P,
                      int n_classes, int64_t max_instances_per_category,
                      const uint8_t *h_thing_lut, int64_t void_label, void *workspace,
                      int64_t *pan_out, int32_t *inst_class, int64_t *inst_pan_id,
                      int32_t *inst_area, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Ground-truth panoptic targets ("naive" merge: every (instance, class) part gets its own id).
 * Replaces: naive_merge_semantic_and_instance_np, utils/panoptic_merge.py:43-107, i.e. the body
 *           of PanopticTargetGenerator (data/preprocessing/panoptic.py:16-85).
 * sem (B,P) u8 (0 = void), ins (B,P) i32 ids in [0, 65535]; h_thing_lut [n_classes].
 * Outputs: pan_out (B,P) i64; per frame the sorted parts: part_keys_out [B][4096] u32
 * (instance << 16 | class), part_pan_out [B][4096] i64 (their panoptic ids), n_parts [B].
 * status [1].
 * ------------------------------------------------------------------------- */
size_t npb_naive_merge_workspace_bytes(int B);
int npb_naive_merge(const uint8_t *sem, const int32_t *ins, int B, int64_t P,
                    int64_t max_instances_per_category, const uint8_t *h_thing_lut, int n_classes,
                    int64_t void_label, void *workspace, int64_t *pan_out, uint32_t *part_keys_out,
                    int64_t *part_pan_out, int32_t *n_parts, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Instance targets from ground-truth maps (centre heat-map, offsets, masks).
 * Replaces: InstanceTargetGenerator._preprocess, data/preprocessing/instance.py:151-286
 *           (the step before the hot path; per sample there, per batch here).
 * sem (B,H,W) u8 semantic labels WITH void (0), ins (B,H,W) i32 ids in [0, 65535];
 * h_thing_lut [n_classes] (index = label, void included); gauss [(6*sigma+3)^2] f32 = the
 * reference's precomputed stamp (instance.py:140-147), supplied by the caller.
 * Outputs: center (B,H,W) f32; offset (B,2,H,W) f32 if normalized_offset else int16 (ch0 = y,
 * ch1 = x); fg, center_mask (B,H,W) u8; per frame the ids of the encoded instances and of those
 * skipped because their majority class is stuff ([B][list_cap], unordered) and their counts.
 * status [1]: NPB_ERR_ARG if a skipped instance still covers pixels (the reference asserts,
 * instance.py:260), NPB_ERR_CAPACITY on table overflow.
 * ------------------------------------------------------------------------- */
size_t npb_instance_targets_workspace_bytes(int B);
int npb_instance_targets(const uint8_t *sem, const int32_t *ins, int B, int H, int W,
                         const uint8_t *h_thing_lut, int n_classes, int sigma, const float *gauss,
                         int normalized_offset, void *workspace, float *center_out,
                         void *offset_out, uint8_t *fg_out, uint8_t *center_mask_out,
                         int32_t *encoded_ids, int32_t *skipped_ids, int32_t *n_encoded,
                         int32_t *n_skipped, int list_cap, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Stand-alone per-instance orientation for arbitrary instance maps.
 * Replaces: InstancePostprocessing._get_instance_orientation, instance.py:270-319.
 * orientation (B,2,P) f32, seg (B,P) of dtype seg_dtype (NPB_U8 / NPB_I32 / NPB_I64),
 * mask (B,P) u8 or NULL.  ids must lie in [0, max_id].
 * Outputs: count [B][max_id+1] i32 (pixels of the id inside the mask),
 *          angle [B][max_id+1] f32, sums [B][max_id+1][2] f64 (zeroed by the call).
 * ------------------------------------------------------------------------- */
int npb_instance_orientation(const float *orientation, const void *seg, int seg_dtype,
                             const uint8_t *mask, int B, int64_t P, int max_id, int32_t *count,
                             float *angle, double *sums, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * mIoU confusion matrix:  confmat[target][pred] += 1  (int64, accumulated in place).
 * Replaces: MeanIntersectionOverUnion.update, metric/miou.py:44-56.
 * status [1] : NPB_ERR_CATEGORY_RANGE if a value lies outside [0, n_classes).
 * ------------------------------------------------------------------------- */
int npb_confmat_update(const void *preds, int preds_dtype, const void *target, int target_dtype,
                       int64_t N, int n_classes, int64_t *confmat, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * The same for the validation step of the semantic task: elements whose target is 0 (void)
 * are skipped, the others count as confmat[target - 1][pred]; no masked copies of the maps.
 * Replaces: mask = target != 0; preds[mask]; target[mask] - 1; MeanIntersectionOverUnion.update
 *           of SemanticTaskHelper.validation_step, task_helper/semantic.py:126-131.
 * status [1] : NPB_ERR_CATEGORY_RANGE if pred or target - 1 of a non-void element lies
 *              outside [0, n_classes) (predictions at void elements are not looked at).
 * ------------------------------------------------------------------------- */
int npb_confmat_update_nonvoid(const void *preds, int preds_dtype, const void *target,
                               int target_dtype, int64_t N, int n_classes, int64_t *confmat,
                               int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * PQ segment matching + accumulation for a batch of frames (and, fused, the mIoU
 * confusion matrix of `pred // L` against a semantic target when confmat != NULL).
 * Replaces: compare_and_accumulate metric/pq.py:60-179 and the accumulation of
 *           PanopticQuality.update pq.py:298-303 (frames are added in frame order, the
 *           IoU sums of one frame in ascending (target*offset + pred) order, so the float64
 *           states are bit-identical to the reference's);
 *           with confmat: MeanIntersectionOverUnion.update of task_helper/panoptic.py:123-126.
 * pred, target (B,P) int64 panoptic ids (>= 0).  sem_target (B,P) u8 or NULL.
 * State (accumulated in place): iou/tp/fn/fp [num_categories] f64, confmat [n][n] i64.
 * Per-frame outputs (nullable): frame_stats [B + 1][4][num_categories] f64 -- rows 0..B-1 the
 *   (iou, tp, fn, fp) of every frame (zero for a frame that reports NPB_ERR_CAPACITY), row B the
 *   states as they were BEFORE this update: with the result of npb_pq_update_big_frame put in
 *   place of a failed frame's row, `row B + row 0 + ... + row B-1` (in that order) is the state
 *   the reference reaches frame by frame (pq.py:298-303), bit for bit;
 *   matches [B][match_cap][2] i64 (gt_id, pred_id), n_matches [B].
 * status [B]: NPB_ERR_ZERO_DIVISION / _CATEGORY_RANGE / _CAPACITY per frame; must be zero on
 *   entry.  A frame that reports NPB_ERR_CAPACITY (more than 4096 distinct (gt, pred) pairs,
 *   1536 segments on one side or 1024 matches) has added NOTHING to iou/tp/fn/fp (its pixels are
 *   in confmat): evaluate it with npb_pq_update_big_frame.
 * ------------------------------------------------------------------------- */
size_t npb_pq_update_workspace_bytes(int B, int num_categories);
int npb_pq_update(const int64_t *pred, const int64_t *target, const uint8_t *sem_target, int B,
                  int64_t P, int num_categories, int64_t ignored_label,
                  int64_t max_instances_per_category, int64_t offset, int64_t void_segment_id,
                  void *workspace, double *iou, double *tp, double *fn, double *fp,
                  int64_t *confmat, int confmat_n, double *frame_stats, int64_t *matches,
                  int match_cap, int32_t *n_matches, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Fall-back of npb_pq_update for ONE frame that exceeded the fixed capacities of the
 * shared-memory matcher (status NPB_ERR_CAPACITY: more than 4096 distinct (gt, pred) pairs or
 * 1536 segments, e.g. per-pixel random ids of an untrained network).  Such a frame contributed
 * nothing to the states; this call evaluates it with every table in global memory (up to one
 * pair per pixel, ~100 000 segments per side, 65536 matches) and adds its result.
 * Same arguments as npb_pq_update for B = 1, without the confusion matrix (npb_pq_update has
 * already counted the frame there).
 * ------------------------------------------------------------------------- */
size_t npb_pq_update_big_frame_workspace_bytes(int64_t P, int num_categories);
int npb_pq_update_big_frame(const int64_t *pred, const int64_t *target, int64_t P,
                            int num_categories, int64_t ignored_label,
                            int64_t max_instances_per_category, int64_t offset,
                            int64_t void_segment_id, void *workspace, double *iou, double *tp,
                            double *fn, double *fp, int64_t *matches, int match_cap,
                            int32_t *n_matches, int32_t *status, void *stream);

/* ---------------------------------------------------------------------------
 * Optional stuff-area filter on finished panoptic maps (NOT part of the reference: its merge
 * keeps every stuff class present in a frame, utils/panoptic_merge.py:213-223; Panoptic-DeepLab's
 * original merge drops stuff regions below `stuff_area` pixels to void).  A stuff segment is the
 * pixel set of one id with `pan > 0 && pan % L == 0`; segments with fewer than `stuff_area`
 * pixels become `void_label` (pan_sem, nullable, becomes `void_label / L`).
 * pan (B,P) i64 in/out, workspace [B][n_classes] u32 (n_classes counts void, <= 256).
 * ------------------------------------------------------------------------- */
int npb_filter_stuff_area(int64_t *pan, uint8_t *pan_sem, int B, int64_t P, int n_classes,
                          int64_t max_instances_per_category, int64_t stuff_area,
                          int64_t void_label, uint32_t *workspace, void *stream);

/* ---------------------------------------------------------------------------
 * Fused validation step: panoptic ids written AND evaluated in one pass.
 * Replaces: the tail of PanopticPostprocessing._postprocess_inference (panoptic.py:139-167)
 *           followed by PanopticTaskHelper.validation_step (task_helper/panoptic.py:104-126),
 *           which reads the freshly written ids back (8 B/px) for PQ and mIoU.
 * `npb_eval_args` carries the evaluation half: the arguments of npb_pq_update without `pred`
 * (the prediction is produced on the fly).  The fused kernel needs the reference's id geometry
 * (max_instances_per_category = 65536, offset = 256^3) and H*W % 4 == 0; otherwise the call runs
 * npb_write_panoptic followed by npb_pq_update -- the results are identical either way.
 * ------------------------------------------------------------------------- */
typedef struct npb_eval_args {
    const int64_t *target;      /* (B,H,W) panoptic target ids                              */
    const uint8_t *sem_target;  /* (B,H,W) semantic target, nullable together with confmat  */
    int num_categories;
    int confmat_n;
    int64_t ignored_label;
    int64_t offset;
    int64_t void_segment_id;
    void *workspace;            /* npb_pq_update_workspace_bytes(B, num_categories)         */
    double *iou, *tp, *fn, *fp; /* [num_categories] accumulated states                      */
    int64_t *confmat;           /* [confmat_n][confmat_n] accumulated, nullable             */
    double *frame_stats;        /* [B + 1][4][num_categories], nullable (see npb_pq_update) */
    int64_t *matches;           /* nullable */
    int match_cap;
    int32_t *n_matches;         /* nullable */
    int32_t *status;            /* [B] status words of the evaluation                       */
} npb_eval_args;

int npb_write_panoptic_eval(const uint8_t *sem, const uint8_t *inst, const int64_t *inst_pan_id,
                            const int32_t *inst_class, int B, int C, int H, int W,
                            const uint8_t *h_thing_lut, int64_t max_instances_per_category,
                            int64_t *pan_out, uint8_t *pan_sem_out, const npb_eval_args *eval,
                            void *stream);

/* npb_panoptic_forward with npb_write_panoptic_eval as its last stage. */
int npb_panoptic_forward_eval(const float *logits, const float *heat, const float *offset,
                              const float *orientation, int B, int C, int H, int W,
                              const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
                              float threshold, int nms_kernel_size, int top_k, int apply_fg_mask,
                              int normalized_offset, int use_distance_threshold,
                              float distance_threshold, int64_t max_instances_per_category,
                              void *workspace, uint8_t *sem_out, uint8_t *inst_out,
                              int64_t *pan_out, uint8_t *pan_sem_out, int32_t *centers_yx,
                              int32_t *n_centers, float *center_score, int32_t *inst_class,
                              int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle,
                              int32_t *status, const npb_eval_args *eval, void *stream);

/* ---------------------------------------------------------------------------
 * Validation LOOP: npb_panoptic_forward_eval with the matcher pipelined over consecutive calls.
 * Replaces: the same reference code as npb_panoptic_forward_eval; what changes is WHEN the PQ
 *           matcher of a batch runs (metric/pq.py:264-303 only defines the states after the loop).
 * The matcher of a batch is 1 CTA per frame of pure latency (hash merges, barriers) that leaves
 * the other SMs idle at the end of every call.  Here a call issues the pixel pass of ITS batch but
 * not the matcher; the next call starts with that matcher (`pending`, with its own batch size),
 * which lets the centre detection + grouping of the new batch run beside it, and the pixel pass of
 * the new batch is ordered behind it (all on `stream`, a chain of programmatic dependent
 * launches, capturable).  The states therefore lag one batch behind until npb_pq_match_pending
 * has been called for the last batch (the Python layer does that before anything reads the
 * states).  Frame order, and with it every float64 sum, is unchanged.
 *   eval       evaluation half of THIS batch (its matcher is left pending)
 *   pending    evaluation half of the previous pipelined call on the same eval->workspace whose
 *              matcher has not run yet, or NULL; pending_B its batch size
 * eval->workspace must have been zeroed ONCE by the caller when it was created: the hand-over
 * tables in it are zero at rest (the matcher cleans up behind itself), no call clears them.
 * ------------------------------------------------------------------------- */
int npb_panoptic_forward_eval_pipelined(
    const float *logits, const float *heat, const float *offset, const float *orientation, int B,
    int C, int H, int W, const uint8_t *h_thing_lut, const uint8_t *h_orientation_lut,
    float threshold, int nms_kernel_size, int top_k, int apply_fg_mask, int normalized_offset,
    int use_distance_threshold, float distance_threshold, int64_t max_instances_per_category,
    void *workspace, uint8_t *sem_out, uint8_t *inst_out, int64_t *pan_out, uint8_t *pan_sem_out,
    int32_t *centers_yx, int32_t *n_centers, float *center_score, int32_t *inst_class,
    int64_t *inst_pan_id, int32_t *inst_area, float *inst_angle, int32_t *status,
    const npb_eval_args *eval, const npb_eval_args *pending, int pending_B, void *stream);

/* The matcher + frame accumulation of a batch whose pixel pass was issued by
 * npb_panoptic_forward_eval_pipelined (call it on a stream ordered after that call). */
int npb_pq_match_pending(const npb_eval_args *pending, int B, int64_t max_instances_per_category,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NICR_PANOPTIC_B200_H */
