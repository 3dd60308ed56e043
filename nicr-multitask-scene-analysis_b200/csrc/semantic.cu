// semantic.cu -- stand-alone semantic arg-max (+ optional soft-max score of the winner)
//
// Replaces SemanticPostprocessing._postprocess_inference's softmax + max
// (reference: model/postprocessing/semantic.py:52-53).  One streaming pass over the C logit
// planes; the score uses an online soft-max (running max + rescaled running sum) so the
// logits are read exactly once:  score = exp(max - max) / sum_c exp(x_c - max) = 1 / sum.
#include "common.cuh"

namespace npb {

template <int VEC, bool SCORE>
__global__ void __launch_bounds__(256)
semantic_argmax_kernel(const float *__restrict__ logits, int C, int P,
                       uint8_t *__restrict__ sem_out, float *__restrict__ score_out)
{
    const int b = blockIdx.y;
    const int p0 = (blockIdx.x * 256 + threadIdx.x) * VEC;
    if (p0 >= P) return;
    const float *lp = logits + (size_t)b * C * P + p0;
    float best[VEC], sum[VEC], taint[VEC];
    int cls[VEC];
    if (VEC == 4) {
        const float4 t = ld_stream_f4((const float4 *)lp);
        best[0] = t.x; best[1 % VEC] = t.y; best[2 % VEC] = t.z; best[3 % VEC] = t.w;
    } else {
        best[0] = ld_stream_f1(lp);
    }
    // taint: +0 while every logit of the pixel is finite, NaN otherwise (x * 0 + taint)
#pragma unroll
    for (int j = 0; j < VEC; ++j) { cls[j] = 0; sum[j] = 1.0f; taint[j] = __fmul_rn(best[j], 0.0f); }
#pragma unroll 8
    for (int c = 1; c < C; ++c) {
        float v[VEC];
        if (VEC == 4) {
            const float4 t = ld_stream_f4((const float4 *)(lp + (size_t)c * P));
            v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
        } else {
            v[0] = ld_stream_f1(lp + (size_t)c * P);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const bool gt = v[j] > best[j];
            taint[j] = __fmaf_rn(v[j], 0.0f, taint[j]);
            if (SCORE) {
                // online soft-max, branch free: exp(best - v) for a new maximum, else exp(v - best)
                const float e = __expf(gt ? best[j] - v[j] : v[j] - best[j]);
                sum[j] = gt ? sum[j] * e + 1.0f : sum[j] + e;
            }
            best[j] = gt ? v[j] : best[j];
            cls[j] = gt ? c : cls[j];
        }
    }
    // Non-finite logits (rare path).  The reference takes max / arg-max of softmax(logits)
    // (semantic.py:52-53): a NaN or +Inf logit, or nothing but -Inf, makes every probability NaN
    // and torch.max answers (NaN, index 0); -Inf next to finite logits has probability 0.
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        if (taint[j] != taint[j]) {
            const float kInf = __int_as_float(0x7f800000);
            bool poisoned = false;
            float mx = -kInf;
            int arg = 0;
            for (int k = 0; k < C; ++k) {
                const float v = lp[(size_t)k * P + j];
                poisoned |= (v != v) || v == kInf;
                if (v > mx) { mx = v; arg = k; }
            }
            if (poisoned || mx == -kInf) {
                cls[j] = 0;
                sum[j] = __int_as_float(0x7fc00000);        // score = NaN
            } else {
                cls[j] = arg;
                float acc = 0.0f;
                for (int k = 0; k < C; ++k) acc += __expf(lp[(size_t)k * P + j] - mx);
                sum[j] = acc;
            }
        }
    }
    const size_t fb = (size_t)b * P + p0;
    if (VEC == 4) {
        *(uint32_t *)(sem_out + fb) = (uint32_t)cls[0] | ((uint32_t)cls[1 % VEC] << 8) |
                                      ((uint32_t)cls[2 % VEC] << 16) | ((uint32_t)cls[3 % VEC] << 24);
        if (SCORE)
            *(float4 *)(score_out + fb) = make_float4(1.0f / sum[0], 1.0f / sum[1 % VEC],
                                                      1.0f / sum[2 % VEC], 1.0f / sum[3 % VEC]);
    } else {
        sem_out[fb] = (uint8_t)cls[0];
        if (SCORE) score_out[fb] = 1.0f / sum[0];
    }
}

// soft-max over the class planes (the reference's `semantic_softmax_scores`, semantic.py:52):
// pass 1 running max + rescaled sum (one exponential per element, streaming loads in batches of
// 8 planes), pass 2 re-reads the (L2 resident) logits and writes exp(x - max) * (1 / sum).
// Fast exponentials (ex2.approx, <= 2 + 1.2 |x| ulp): with IEEE expf + a division per element
// the kernel was instruction bound (0.52 of the HBM peak); scores are compared at 1e-5.
// Only launched when somebody reads that (B,C,H,W) entry of the result dict.
template <int VEC>
__device__ __forceinline__ void load_plane(const float *p, float (&v)[VEC], bool stream)
{
    if (VEC == 4) {
        const float4 t = stream ? ld_stream_f4((const float4 *)p) : *(const float4 *)p;
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = stream ? ld_stream_f1(p) : *p;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256)
softmax_kernel(const float *__restrict__ logits, int C, int P, float *__restrict__ probs)
{
    const int b = blockIdx.y;
    const int p0 = (blockIdx.x * 256 + threadIdx.x) * VEC;
    if (p0 >= P) return;
    const float *lp = logits + (size_t)b * C * P + p0;
    float *op = probs + (size_t)b * C * P + p0;
    constexpr int U = 8;
    float mx[VEC], sum[VEC];
    load_plane<VEC>(lp, mx, false);
#pragma unroll
    for (int j = 0; j < VEC; ++j) sum[j] = 1.0f;
    auto update = [&](const float (&v)[VEC]) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float d = v[j] - mx[j];
            const float e = __expf(-fabsf(d));      // NaN logits propagate into the sum
            if (d > 0.0f) { sum[j] = __fmaf_rn(sum[j], e, 1.0f); mx[j] = v[j]; }
            else sum[j] += e;
        }
    };
    int c = 1;
    for (; c + U <= C; c += U) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) load_plane<VEC>(lp + (size_t)(c + u) * P, v[u], false);
#pragma unroll
        for (int u = 0; u < U; ++u) update(v[u]);
    }
    for (; c < C; ++c) {
        float v[VEC];
        load_plane<VEC>(lp + (size_t)c * P, v, false);
        update(v);
    }
    float inv[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) inv[j] = 1.0f / sum[j];
    auto store = [&](int plane, const float (&v)[VEC]) {
        float r[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) r[j] = __expf(v[j] - mx[j]) * inv[j];
        if (VEC == 4) __stcs((float4 *)(op + (size_t)plane * P), make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]));
        else op[(size_t)plane * P] = r[0];
    };
    c = 0;
    for (; c + U <= C; c += U) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) load_plane<VEC>(lp + (size_t)(c + u) * P, v[u], false);
#pragma unroll
        for (int u = 0; u < U; ++u) store(c + u, v[u]);
    }
    for (; c < C; ++c) {
        float v[VEC];
        load_plane<VEC>(lp + (size_t)c * P, v, false);
        store(c, v);
    }
}

}  // namespace npb

using namespace npb;

extern "C" int npb_softmax(const float *logits, int B, int C, int H, int W, float *probs,
                           void *stream)
{
    if (!logits || !probs) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || C < 1 || H < 1 || W < 1) return NPB_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return NPB_ERR_ARG;
    const int P = H * W;
    cudaStream_t s = (cudaStream_t)stream;
    if (P % 4 == 0 && (((uintptr_t)logits | (uintptr_t)probs) & 15u) == 0) {
        dim3 grid((P / 4 + 255) / 256, B);
        softmax_kernel<4><<<grid, 256, 0, s>>>(logits, C, P, probs);
    } else {
        dim3 grid((P + 255) / 256, B);
        softmax_kernel<1><<<grid, 256, 0, s>>>(logits, C, P, probs);
    }
    return record_launch("npb_softmax");
}

extern "C" int npb_semantic_argmax(const float *logits, int B, int C, int H, int W,
                                   uint8_t *sem_out, float *score_out, void *stream)
{
    if (!logits || !sem_out) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || C < 1 || C > 256 || H < 1 || W < 1) return NPB_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return NPB_ERR_ARG;
    const int P = H * W;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec4 = (P % 4 == 0) && (((uintptr_t)logits | (uintptr_t)score_out) & 15u) == 0 &&
                      ((uintptr_t)sem_out & 3u) == 0;
    if (vec4) {
        dim3 grid((P / 4 + 255) / 256, B);
        if (score_out) semantic_argmax_kernel<4, true><<<grid, 256, 0, s>>>(logits, C, P, sem_out, score_out);
        else semantic_argmax_kernel<4, false><<<grid, 256, 0, s>>>(logits, C, P, sem_out, score_out);
    } else {
        dim3 grid((P + 255) / 256, B);
        if (score_out) semantic_argmax_kernel<1, true><<<grid, 256, 0, s>>>(logits, C, P, sem_out, score_out);
        else semantic_argmax_kernel<1, false><<<grid, 256, 0, s>>>(logits, C, P, sem_out, score_out);
    }
    return record_launch("npb_semantic_argmax");
}
