// scores.cu -- dense semantic / instance / panoptic score maps (compute_scores=True)
//
// Replaces the score section of PanopticPostprocessing._postprocess_inference
// (reference: model/postprocessing/panoptic.py:171-239):
//   semantic score  = soft-max probability of the pixel's PANOPTIC class (0 for void)
//   instance score  = heat-map value at the centre of the pixel's instance (0 elsewhere)
//   panoptic score  = semantic score for stuff, mean semantic score of the instance times its
//                     instance score for things (like YOLO / Panoptic-DeepLab)
// Kernel 1 streams the logits once (online soft-max: running max + rescaled sum, and the
// logit of the wanted class), writes the semantic score and accumulates its per-instance sum
// (f32 shuffle tree per distinct instance of a warp, f64 RED).  Kernel 2 writes the other two
// maps from the per-instance tables.  Tolerance-checked (1e-5 relative), not bit-exact: the
// reference's f32 soft-max / mean orders are ATen's.
#include "common.cuh"

namespace npb {

// VEC = 4: four consecutive pixels per thread (128-bit loads, 8 planes in flight, branch-free
// online soft-max -- the loop of semantic_argmax_kernel<4, true>); VEC = 1 for maps whose size or
// alignment rules that out.
template <int VEC>
__global__ void __launch_bounds__(256)
semantic_score_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ pan_sem,
                      const uint8_t *__restrict__ inst, int C, int P,
                      float *__restrict__ sem_score, double *__restrict__ inst_sum)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int p = (blockIdx.x * 256 + threadIdx.x) * VEC;
    const bool act = p < P;         // P % VEC == 0 is guaranteed by the launcher
    float score[VEC];
    int ii[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { score[j] = 0.0f; ii[j] = 0; }
    if (act) {
        const size_t q = (size_t)b * P + p;
        int want[VEC];              // network class of the panoptic label (-1: void)
        if (VEC == 4) {
            const uint32_t w = *(const uint32_t *)(pan_sem + q), i4 = *(const uint32_t *)(inst + q);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                want[j] = (int)((w >> (8 * j)) & 255u) - 1;
                ii[j] = (int)((i4 >> (8 * j)) & 255u);
            }
        } else {
            want[0] = (int)pan_sem[q] - 1;
            ii[0] = inst[q];
        }
        const float *lp = logits + (size_t)b * C * P + p;
        float mx[VEC], sum[VEC], sel[VEC];
        auto load = [&](int c, float (&v)[VEC]) {
            if (VEC == 4) {
                const float4 t = ld_stream_f4((const float4 *)(lp + (size_t)c * P));
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            } else {
                v[0] = ld_stream_f1(lp + (size_t)c * P);
            }
        };
        load(0, mx);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sum[j] = 1.0f; sel[j] = mx[j]; }
        if (VEC == 4) {
#pragma unroll 8
            for (int c = 1; c < C; ++c) {
                float v[VEC];
                load(c, v);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    sel[j] = (c == want[j]) ? v[j] : sel[j];
                    const bool gt = v[j] > mx[j];
                    const float e = __expf(gt ? mx[j] - v[j] : v[j] - mx[j]);
                    sum[j] = gt ? sum[j] * e + 1.0f : sum[j] + e;
                    mx[j] = gt ? v[j] : mx[j];
                }
            }
        } else {
            for (int c = 1; c < C; ++c) {
                float v[VEC];
                load(c, v);
                if (c == want[0]) sel[0] = v[0];
                if (v[0] > mx[0]) { sum[0] = sum[0] * __expf(mx[0] - v[0]) + 1.0f; mx[0] = v[0]; }
                else sum[0] += __expf(v[0] - mx[0]);
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)       // void has no valid score
            score[j] = want[j] >= 0 ? __expf(sel[j] - mx[j]) / sum[j] : 0.0f;
        if (VEC == 4) *(float4 *)(sem_score + q) = make_float4(score[0], score[1 % VEC], score[2 % VEC], score[3 % VEC]);
        else sem_score[q] = score[0];
    }
    // per-instance sums: one round per distinct instance of the warp; a round adds EVERY pixel
    // of that instance in the warp (lane-local sum, f32 shuffle tree, one f64 RED) and retires them
    double *sums = inst_sum + (size_t)b * kMaxInst;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        unsigned pending = __ballot_sync(kFullMask, ii[k] > 0);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int cur = __shfl_sync(kFullMask, ii[k], leader);
            float mine = 0.0f;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (ii[j] == cur) { mine += score[j]; ii[j] = 0; }
            const float s = warp_sum(mine);
            if (lane == leader) atomicAdd(sums + cur, (double)s);
            pending = __ballot_sync(kFullMask, ii[k] > 0);
        }
    }
}

__global__ void __launch_bounds__(256)
instance_panoptic_score_kernel(const float *__restrict__ sem_score,
                               const uint8_t *__restrict__ inst,
                               const int32_t *__restrict__ inst_class,
                               const int32_t *__restrict__ inst_area,
                               const float *__restrict__ center_score,
                               const double *__restrict__ inst_sum, int P,
                               float *__restrict__ inst_score_out, float *__restrict__ pan_score_out,
                               float *__restrict__ inst_mean_sem, float *__restrict__ inst_pan_score)
{
    __shared__ float s_score[kMaxInst], s_pan[kMaxInst];
    const int b = blockIdx.y, i = threadIdx.x;
    {   // per-instance values (row i of the tables; row 0 and dropped instances stay 0 / -1)
        const size_t o = (size_t)b * kMaxInst + i;
        float sc = 0.0f, ps = -1.0f, mean = 0.0f;
        if (i >= 1 && inst_class[o] >= 0 && inst_area[o] > 0) {
            sc = center_score[(size_t)b * kMaxInst + i - 1];    // centre i-1 <-> instance id i
            mean = (float)(inst_sum[o] / (double)inst_area[o]);
            ps = mean * sc;
        }
        s_score[i] = sc;
        s_pan[i] = ps;
        if (blockIdx.x == 0) { inst_mean_sem[o] = mean; inst_pan_score[o] = ps; }
    }
    __syncthreads();
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const size_t q = (size_t)b * P + p;
    const int ii = inst[q];
    const float ps = s_pan[ii];
    inst_score_out[q] = s_score[ii];
    pan_score_out[q] = (ii > 0 && ps >= 0.0f) ? ps : sem_score[q];
}

}  // namespace npb

using namespace npb;

extern "C" int npb_panoptic_scores(const float *logits, const uint8_t *pan_sem, const uint8_t *inst,
                                   const int32_t *inst_class, const int32_t *inst_area,
                                   const float *center_score, int B, int C, int H, int W,
                                   double *inst_sum, float *sem_score, float *inst_score,
                                   float *pan_score, float *inst_mean_sem, float *inst_pan_score,
                                   void *stream)
{
    if (!logits || !pan_sem || !inst || !inst_class || !inst_area || !center_score || !inst_sum ||
        !sem_score || !inst_score || !pan_score || !inst_mean_sem || !inst_pan_score)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || C < 1 || C > 255 || H < 1 || W < 1) return NPB_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return NPB_ERR_ARG;
    const int P = H * W;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(inst_sum, 0, (size_t)B * kMaxInst * sizeof(double), s);
    dim3 grid((P + 255) / 256, B);
    const bool vec4 = P % 4 == 0 && (((uintptr_t)logits | (uintptr_t)sem_score) & 15u) == 0 &&
                      (((uintptr_t)pan_sem | (uintptr_t)inst) & 3u) == 0;
    if (vec4)
        semantic_score_kernel<4><<<dim3((P / 4 + 255) / 256, B), 256, 0, s>>>(logits, pan_sem, inst, C, P,
                                                                          sem_score, inst_sum);
    else
        semantic_score_kernel<1><<<grid, 256, 0, s>>>(logits, pan_sem, inst, C, P, sem_score, inst_sum);
    instance_panoptic_score_kernel<<<grid, 256, 0, s>>>(sem_score, inst, inst_class, inst_area,
                                                        center_score, inst_sum, P, inst_score,
                                                        pan_score, inst_mean_sem, inst_pan_score);
    return record_launch("npb_panoptic_scores");
}
