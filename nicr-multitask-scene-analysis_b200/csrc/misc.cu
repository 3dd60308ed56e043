// misc.cu -- small stand-alone pieces of the API surface
//   * thing / class-set mask            torch.isin(...)          panoptic.py:123-127, 296-300
//   * uint8 -> int64 widening           the reference's int64 index outputs (semantic.py:53,
//                                       panoptic.py:160)
//   * per-instance orientation for arbitrary instance maps       instance.py:270-319
//   * optional stuff-area filter on a finished panoptic map (not in the reference: off by default)
#include <math.h>

#include "common.cuh"

namespace npb {

__global__ void __launch_bounds__(256)
class_mask_kernel(const uint8_t *__restrict__ sem, long long N, ClassSet set,
                  uint8_t *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += stride)
        out[i] = set.has(sem[i]) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
widen_u8_kernel(const uint8_t *__restrict__ in, long long N, long long add,
                long long *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += stride)
        out[i] = (long long)in[i] + add;
}

__device__ __forceinline__ long long load_seg(const void *p, int dtype, size_t i)
{
    switch (dtype) {
        case NPB_U8: return ((const uint8_t *)p)[i];
        case NPB_I16: return ((const int16_t *)p)[i];
        case NPB_I32: return ((const int32_t *)p)[i];
        default: return ((const long long *)p)[i];
    }
}

__global__ void __launch_bounds__(256)
orientation_sums_kernel(const float *__restrict__ ori, const void *__restrict__ seg, int seg_dtype,
                        const uint8_t *__restrict__ mask, long long P, int max_id,
                        int32_t *__restrict__ count, double *__restrict__ sums,
                        int32_t *__restrict__ status)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const float *oc = ori + (size_t)b * 2 * P, *os = oc + P;
    int32_t *cnt = count + (size_t)b * (max_id + 1);
    double *sm = sums + (size_t)b * (max_id + 1) * 2;
    const long long stride = (long long)gridDim.x * 256;
    const long long n_round = ((P + 31) / 32) * 32;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n_round; p += stride) {
        int id = 0;
        float c = 0.0f, s = 0.0f;
        if (p < P) {
            const size_t q = (size_t)b * P + p;
            if (!mask || mask[q]) {
                const long long v = load_seg(seg, seg_dtype, q);
                if (v < 0 || v > max_id) set_status(status, NPB_ERR_CATEGORY_RANGE);
                else id = (int)v;
                if (id) { c = oc[p]; s = os[p]; }
            }
        }
        unsigned pending = __ballot_sync(kFullMask, id != 0);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int cur = __shfl_sync(kFullMask, id, leader);
            const bool mine = (id == cur);
            const unsigned same = __ballot_sync(kFullMask, mine);
            const float sc = warp_sum(mine ? c : 0.0f), ss = warp_sum(mine ? s : 0.0f);
            if (lane == leader) {
                atomicAdd(cnt + cur, __popc(same));
                atomicAdd(sm + 2 * cur, (double)sc);
                atomicAdd(sm + 2 * cur + 1, (double)ss);
            }
            pending &= ~same;
        }
    }
}

__global__ void orientation_angle_kernel(const int32_t *__restrict__ count,
                                         const double *__restrict__ sums, long long n,
                                         float *__restrict__ angle)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    angle[i] = count[i] > 0
                   ? (float)atan2((double)(float)sums[2 * i + 1], (double)(float)sums[2 * i])
                   : nanf("");
}

static int grid_for(long long n, int per_thread)
{
    long long blocks = (n + 256ll * per_thread - 1) / (256ll * per_thread);
    if (blocks > 148 * 16) blocks = 148 * 16;
    return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace npb

using namespace npb;

extern "C" int npb_thing_mask(const uint8_t *sem, int64_t N, int C, const uint8_t *h_class_lut,
                              uint8_t *mask_out, void *stream)
{
    if (!sem || !mask_out || !h_class_lut || N < 0 || C < 1 || C > 256) return NPB_ERR_ARG;
    if (N == 0) return NPB_OK;
    class_mask_kernel<<<grid_for(N, 8), 256, 0, (cudaStream_t)stream>>>(
        sem, (long long)N, make_class_set(h_class_lut, C), mask_out);
    return record_launch("npb_thing_mask");
}

extern "C" int npb_widen_u8(const uint8_t *in, int64_t N, int64_t add, int64_t *out, void *stream)
{
    if (!in || !out || N < 0) return NPB_ERR_ARG;
    if (N == 0) return NPB_OK;
    widen_u8_kernel<<<grid_for(N, 8), 256, 0, (cudaStream_t)stream>>>(in, (long long)N,
                                                                     (long long)add,
                                                                     (long long *)out);
    return record_launch("npb_widen_u8");
}

extern "C" int npb_instance_orientation(const float *orientation, const void *seg, int seg_dtype,
                                        const uint8_t *mask, int B, int64_t P, int max_id,
                                        int32_t *count, float *angle, double *sums,
                                        int32_t *status, void *stream)
{
    if (!orientation || !seg || !count || !angle || !sums || !status) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || P < 1 || max_id < 1) return NPB_ERR_ARG;
    if (seg_dtype != NPB_U8 && seg_dtype != NPB_I16 && seg_dtype != NPB_I32 && seg_dtype != NPB_I64)
        return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const long long rows = (long long)B * (max_id + 1);
    cudaMemsetAsync(count, 0, (size_t)rows * sizeof(int32_t), s);
    cudaMemsetAsync(sums, 0, (size_t)rows * 2 * sizeof(double), s);
    dim3 grid(grid_for(P, 8), B);
    orientation_sums_kernel<<<grid, 256, 0, s>>>(orientation, seg, seg_dtype, mask, (long long)P,
                                                 max_id, count, sums, status);
    orientation_angle_kernel<<<(int)((rows + 255) / 256), 256, 0, s>>>(count, sums, rows, angle);
    return record_launch("npb_instance_orientation");
}

// ---- optional stuff-area filter ---------------------------------------------------------------
namespace npb {

// The reference's merge gives EVERY stuff class present in a frame its id `class * L`
// (utils/panoptic_merge.py:213-223); Panoptic-DeepLab's original merge drops stuff regions
// smaller than `stuff_area` pixels to void.  Offered as an option (default off = the reference):
// a stuff segment is a pixel set with `pan > 0 && pan % L == 0`; pass 1 counts the pixels of
// every such segment per frame (privatised shared-memory histogram over the <= 256 classes),
// pass 2 rewrites the segments below the limit to `void_label`.
__global__ void __launch_bounds__(256)
stuff_area_count_kernel(const long long *__restrict__ pan, long long P, int n_classes, long long L,
                        int L_shift, unsigned *__restrict__ hist)
{
    __shared__ unsigned s_hist[256];
    const int b = blockIdx.y;
    s_hist[threadIdx.x] = 0u;
    __syncthreads();
    const long long *pb = pan + (size_t)b * P;
    const long long stride = (long long)gridDim.x * 256;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < P; p += stride) {
        const long long v = pb[p];
        if (v <= 0) continue;
        const long long c = L_shift >= 0 ? (v >> L_shift) : (v / L);
        if (v - c * L == 0 && c < n_classes) atomicAdd(&s_hist[(int)c], 1u);
    }
    __syncthreads();
    if ((int)threadIdx.x < n_classes && s_hist[threadIdx.x])
        atomicAdd(hist + (size_t)b * n_classes + threadIdx.x, s_hist[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
stuff_area_apply_kernel(long long *__restrict__ pan, uint8_t *__restrict__ pan_sem, long long P,
                        int n_classes, long long L, int L_shift, long long stuff_area,
                        long long void_label, const unsigned *__restrict__ hist)
{
    __shared__ unsigned char s_drop[256];
    const int b = blockIdx.y;
    s_drop[threadIdx.x] = ((int)threadIdx.x < n_classes &&
                           (long long)hist[(size_t)b * n_classes + threadIdx.x] < stuff_area) ? 1 : 0;
    __syncthreads();
    long long *pb = pan + (size_t)b * P;
    uint8_t *sb = pan_sem ? pan_sem + (size_t)b * P : nullptr;
    const long long stride = (long long)gridDim.x * 256;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < P; p += stride) {
        const long long v = pb[p];
        if (v <= 0) continue;
        const long long c = L_shift >= 0 ? (v >> L_shift) : (v / L);
        if (v - c * L == 0 && c < n_classes && s_drop[(int)c]) {
            pb[p] = void_label;
            if (sb) sb[p] = (uint8_t)(void_label / L);
        }
    }
}

}  // namespace npb

using namespace npb;

extern "C" int npb_filter_stuff_area(int64_t *pan, uint8_t *pan_sem, int B, int64_t P, int n_classes,
                                     int64_t max_instances_per_category, int64_t stuff_area,
                                     int64_t void_label, uint32_t *workspace, void *stream)
{
    if (!pan || !workspace) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || P < 1 || n_classes < 1 || n_classes > 256 ||
        max_instances_per_category < 1 || stuff_area < 0 || void_label < 0)
        return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int L_shift = -1;
    for (int sh = 0; sh < 62; ++sh)
        if ((1ll << sh) == max_instances_per_category) L_shift = sh;
    cudaMemsetAsync(workspace, 0, (size_t)B * n_classes * sizeof(uint32_t), s);
    long long blocks = (P + 256 * 8 - 1) / (256 * 8);
    if (blocks > 1024) blocks = 1024;
    dim3 grid((unsigned)blocks, B);
    stuff_area_count_kernel<<<grid, 256, 0, s>>>((const long long *)pan, P, n_classes,
                                                 max_instances_per_category, L_shift, workspace);
    stuff_area_apply_kernel<<<grid, 256, 0, s>>>((long long *)pan, pan_sem, P, n_classes,
                                                 max_instances_per_category, L_shift, stuff_area,
                                                 void_label, workspace);
    return record_launch("npb_filter_stuff_area");
}
