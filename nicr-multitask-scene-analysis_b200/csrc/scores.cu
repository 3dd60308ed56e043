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

__global__ void __launch_bounds__(256)
semantic_score_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ pan_sem,
                      const uint8_t *__restrict__ inst, int C, int P,
                      float *__restrict__ sem_score, double *__restrict__ inst_sum)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool act = p < P;
    float score = 0.0f;
    int ii = 0;
    if (act) {
        const size_t q = (size_t)b * P + p;
        const int want = (int)pan_sem[q] - 1;       // network class of the panoptic label
        ii = inst[q];
        const float *lp = logits + (size_t)b * C * P + p;
        float mx = ld_stream_f1(lp), sum = 1.0f, sel = mx;
        for (int c = 1; c < C; ++c) {
            const float v = ld_stream_f1(lp + (size_t)c * P);
            if (c == want) sel = v;
            if (v > mx) { sum = sum * __expf(mx - v) + 1.0f; mx = v; }
            else sum += __expf(v - mx);
        }
        score = want >= 0 ? __expf(sel - mx) / sum : 0.0f;     // void has no valid score
        sem_score[q] = score;
    }
    // per-instance sums: loop over the distinct instances of the warp
    unsigned pending = __ballot_sync(kFullMask, act && ii > 0);
    double *sums = inst_sum + (size_t)b * kMaxInst;
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int cur = __shfl_sync(kFullMask, ii, leader);
        const bool mine = act && ii == cur;
        const float s = warp_sum(mine ? score : 0.0f);
        if (lane == leader) atomicAdd(sums + cur, (double)s);
        pending &= ~__ballot_sync(kFullMask, mine);
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
    semantic_score_kernel<<<grid, 256, 0, s>>>(logits, pan_sem, inst, C, P, sem_score, inst_sum);
    instance_panoptic_score_kernel<<<grid, 256, 0, s>>>(sem_score, inst, inst_class, inst_area,
                                                        center_score, inst_sum, P, inst_score,
                                                        pan_score, inst_mean_sem, inst_pan_score);
    return record_launch("npb_panoptic_scores");
}
