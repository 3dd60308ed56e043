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

// internal forms of npb_instance_centers / npb_group_pixels for npb_panoptic_forward, which clears
// the scratch of both stages with ONE memset (centers.cu, group.cu)
int instance_centers_impl(const float *heat, int B, int H, int W, float threshold,
                          int nms_kernel_size, int top_k, const uint8_t *fg, int apply_fg_mask,
                          void *workspace, int32_t *centers_yx, int32_t *n_centers,
                          float *center_score, int32_t *status, bool cleared, bool reset_status,
                          void *stream);
int group_pixels_impl(const float *logits, const uint8_t *sem_in, const uint8_t *fg_in,
                      const float *offset, const float *orientation, int B, int C, int H, int W,
                      const uint8_t *h_thing_lut, const int32_t *centers_yx,
                      const int32_t *n_centers, int normalized_offset, int use_distance_threshold,
                      float distance_threshold, uint8_t *sem_out, uint8_t *inst_out,
                      uint32_t *vote_hist, double *ori_sum, bool cleared, void *stream);

// ---- programmatic dependent launch ----------------------------------------------------------
// A kernel launched with launch_dependent() may start while its predecessor on the stream is
// still draining: its CTAs are scheduled and run their prologue (shared-memory set-up) early and
// block in grid_dependency_wait() until the predecessor has completed and its memory is visible.
// Every kernel launched that way MUST call grid_dependency_wait() before it touches global
// memory.  (Without the launch attribute the instruction is a no-op.)
__device__ __forceinline__ void grid_dependency_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

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
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace npb
