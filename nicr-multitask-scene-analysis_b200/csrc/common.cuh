// common.cuh -- shared device helpers for libnicr_panoptic_b200 (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nicr_panoptic_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnicr_panoptic_b200 is written for sm_100a (B200) only"
#endif

namespace npb {

constexpr int kMaxInst = NPB_MAX_INST;  // row length of every per-instance table
constexpr unsigned kFullMask = 0xffffffffu;

// 256-bit class set passed by value in kernel parameters (classes are uint8)
struct ClassSet {
    uint32_t w[8];
    __host__ __device__ bool has(int c) const { return (w[(c >> 5) & 7] >> (c & 31)) & 1u; }
};

inline ClassSet make_class_set(const uint8_t *h_lut, int n)
{
    ClassSet s;
    for (int i = 0; i < 8; ++i) s.w[i] = 0;
    if (h_lut)
        for (int c = 0; c < n && c < 256; ++c)
            if (h_lut[c]) s.w[c >> 5] |= (1u << (c & 31));
    return s;
}

// record the first (most negative wins is irrelevant: any) error of a frame / call
__device__ __forceinline__ void set_status(int32_t *status, int code)
{
    if (status) atomicMin(status, code);
}

// streaming (read-once) loads: keep them out of L1, mark evict-first in L2
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream_f1(const float *p) { return __ldcs(p); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

int record_launch(const char *what);  // api.cu: cudaGetLastError -> NPB_ERR_CUDA

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// counters at the end of the centres workspace: candidates per frame, finished NMS CTAs per frame
inline size_t centers_counter_bytes(int B) { return align256(2 * (size_t)B * sizeof(int32_t)); }
// behind them: the complete centre lists of frames with more than 255 centres (centers.cu)
inline size_t centers_wide_bytes(int B)
{
    return align256((size_t)B * sizeof(int32_t)) +
           align256((size_t)B * 2 * NPB_MAX_WIDE_CENTERS * sizeof(unsigned));
}

// internal forms of npb_instance_centers / npb_group_pixels for npb_panoptic_forward, which clears
// the scratch of both stages with ONE memset (centers.cu, group.cu)
struct ScratchToClear {     // two byte ranges (multiples of 4) zeroed by the NMS pass
    void *p0; size_t bytes0;
    void *p1; size_t bytes1;
};
int instance_centers_impl(const float *heat, int B, int H, int W, float threshold,
                          int nms_kernel_size, int top_k, const uint8_t *fg, int apply_fg_mask,
                          void *workspace, int32_t *centers_yx, int32_t *n_centers,
                          float *center_score, int32_t *status, bool cleared, bool reset_status,
                          const ScratchToClear *downstream, bool late_wait, void *stream);
int group_pixels_impl(const float *logits, const uint8_t *sem_in, const uint8_t *fg_in,
                      const float *offset, const float *orientation, int B, int C, int H, int W,
                      const uint8_t *h_thing_lut, const int32_t *centers_yx,
                      const int32_t *n_centers, int normalized_offset, int use_distance_threshold,
                      float distance_threshold, uint8_t *sem_out, uint8_t *inst_out,
                      uint32_t *vote_hist, double *ori_sum, bool cleared, void *stream);

struct FinalizeParams;      // finalize.cuh
// eval.cu: the memset of an evaluation update, and the id writer + evaluation with the instance
// tables optionally derived inside the pass (see there)
void pq_clear_workspace(void *workspace, int B, int num_categories,
                        int64_t max_instances_per_category, void *stream);
// the same range, for a kernel that clears it instead (see ScratchToClear)
void pq_cleared_range(void *workspace, int B, int num_categories,
                      int64_t max_instances_per_category, void **p, size_t *bytes);
int write_panoptic_eval_impl(const uint8_t *sem, const uint8_t *inst, int64_t *inst_pan_id,
                             int32_t *inst_class, int B, int C, int H, int W,
                             const uint8_t *h_thing_lut, int64_t max_instances_per_category,
                             int64_t *pan_out, uint8_t *pan_sem_out, const npb_eval_args *ev,
                             const FinalizeParams *fold, bool cleared, bool skip_match,
                             void *stream);
// the matcher + frame accumulation of an update whose pixel pass was issued with `skip_match`
int pq_match_impl(const npb_eval_args *ev, int B, int64_t max_instances_per_category, void *stream);
// merge.cu: launches finalize_instances_kernel with launch_dependent()
int launch_finalize(const FinalizeParams &f, int B, void *stream);

// ---- programmatic dependent launch ----------------------------------------------------------
// A kernel launched with launch_dependent() may start while its predecessor on the stream is
// still draining: its CTAs are scheduled and run their prologue (shared-memory set-up) early and
// block in grid_dependency_wait() until the predecessor has completed and its memory is visible.
// Every kernel launched that way MUST call grid_dependency_wait() before it touches global
// memory a predecessor writes.  (Without the launch attribute the instruction is a no-op.)
// Its inputs from predecessors must NOT be `const T *__restrict__` parameters: those loads
// compile to LDG.CONSTANT (ld.global.nc), which may be scheduled above the wait.
__device__ __forceinline__ void grid_dependency_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Allows the NEXT kernel on the stream (if it was launched with launch_dependent()) to be
// scheduled as soon as every CTA of this grid has executed this instruction (or exited): its
// CTAs fill the slots this grid leaves free and run up to their own grid_dependency_wait().
// Used by the NMS pass only, whose successor (the grouping kernel) has most of its work -- the
// arg-max over the logits -- in front of its wait.  The other kernels rely on the implicit
// trigger at CTA exit: the successor is staged while the grid drains (no launch gap), but its
// CTAs are not placed early.  Rule of the chain: EVERY kernel executes grid_dependency_wait() on
// every path, so "my predecessor completed" implies "everything before it completed".
__device__ __forceinline__ void grid_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- bounds / overflow asserts of the shared-memory tables (only with -DNPB_DEBUG) ------------
// compute-sanitizer is not available on the GPU pool, so the histogram / hash-table / queue
// kernels carry their own checks: an -DNPB_DEBUG build (csrc/build.py -DNPB_DEBUG --out=...,
// loaded through NPB_LIB_PATH) traps with file:line on the first violated bound, and the GPU test
// suite is run once against it per round (scripts/round_evidence.sh).
#ifdef NPB_DEBUG
#include <stdio.h>
#define NPB_ASSERT(cond)                                                                      \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            printf("NPB_ASSERT failed: %s  (%s:%d, block %d,%d thread %d)\n", #cond, __FILE__, \
                   __LINE__, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);              \
            __trap();                                                                         \
        }                                                                                     \
    } while (0)
#else
#define NPB_ASSERT(cond) ((void)0)
#endif

// ---- in-situ timeline (only with -DNPB_TIMELINE: scripts/probes/timeline.py) -----------------
// Every CTA stamps the global timer at its start, after its grid_dependency_wait() and at its
// end into slot `id` of a small device buffer (min / max), which shows where the kernels of a
// replayed step really start, wait and end relative to each other.
#ifdef NPB_TIMELINE
struct TimelineSlot { unsigned long long start_min, start_max, wait_min, wait_max, end_min, end_max; };
TimelineSlot *timeline_buffer();    // api.cu (lazily allocated, 16 slots)
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define NPB_TL_FIELD TimelineSlot *tl;
#define NPB_TL_SET(prm) (prm).tl = timeline_buffer()
#define NPB_TL(prm, id, what)                                                        \
    do {                                                                             \
        if (threadIdx.x == 0 && (prm).tl) {                                          \
            const unsigned long long t_ = global_ns();                               \
            atomicMin(&(prm).tl[id].what##_min, t_);                                 \
            atomicMax(&(prm).tl[id].what##_max, t_);                                 \
        }                                                                            \
    } while (0)
#else
#define NPB_TL_FIELD
#define NPB_TL_SET(prm) ((void)0)
#define NPB_TL(prm, id, what) ((void)0)
#endif

bool dependent_launch_enabled();    // api.cu: NPB_NO_PDL=1 turns the attribute off (A/B, debugging)

template <typename... KArgs, typename... Args>
inline void launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = dependent_launch_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace npb
