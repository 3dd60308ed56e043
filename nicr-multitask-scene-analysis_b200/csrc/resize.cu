// resize.cu -- crop to the valid region + resize to the dataset resolution
//
// Replaces DensePostprocessingBase._crop_to_valid_region_and_resize_prediction
// (reference: model/postprocessing/dense_base.py:15-58, i.e. F.interpolate of ATen):
//   * nearest  (index maps; the reference round-trips them through float32, exact < 2^24):
//       src = min(floor(dst * scale), in - 1),  scale = (float)in / out
//   * bilinear, align_corners=False (semantic logits, semantic.py:63-72):
//       real = max(scale * (dst + 0.5) - 0.5, 0) (one fused multiply-add, as ATen's build
//       contracts it -- probed), i0 = floor(real), i1 = i0 + (i0 < in - 1), w1 = real - i0
// plus the fused variant the semantic path actually needs: bilinear resize of the C logit
// planes ON THE FLY + arg-max (+ soft-max score of the winner), without materialising the
// (B, C, h, w) resized logits (semantic.py:67-72).
// Interpolated values agree with ATen to f32 rounding (not bit for bit: ATen's summation order
// is build dependent), so the full-resolution class map can differ from the reference's where
// the two best resized logits are closer than ~1e-6 relative.
#include "common.cuh"

namespace npb {

struct ResizeGeom {
    int Hin, Win;        // full input plane
    int y0, x0, Hc, Wc;  // crop (valid region)
    int Hout, Wout;
    float sy, sx;        // (float)Hc / Hout, (float)Wc / Wout
};

template <typename T>
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const T *__restrict__ src, ResizeGeom g, T *__restrict__ dst)
{
    const int plane = blockIdx.y;
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= g.Hout * g.Wout) return;
    const int oy = o / g.Wout, ox = o - oy * g.Wout;
    const int iy = min((int)floorf(__fmul_rn((float)oy, g.sy)), g.Hc - 1);
    const int ix = min((int)floorf(__fmul_rn((float)ox, g.sx)), g.Wc - 1);
    dst[(size_t)plane * g.Hout * g.Wout + o] =
        src[(size_t)plane * g.Hin * g.Win + (size_t)(g.y0 + iy) * g.Win + g.x0 + ix];
}

struct LinCoord { int i0, i1; float w0, w1; };

__device__ __forceinline__ LinCoord lin_coord(int dst, float scale, int in_size)
{
    float real = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
    real = fmaxf(real, 0.0f);
    LinCoord c;
    c.i0 = min((int)real, in_size - 1);
    c.i1 = c.i0 + (c.i0 < in_size - 1 ? 1 : 0);
    c.w1 = fminf(fmaxf(real - (float)c.i0, 0.0f), 1.0f);
    c.w0 = 1.0f - c.w1;
    return c;
}

__device__ __forceinline__ float bilerp(const float *__restrict__ p, const ResizeGeom &g,
                                        const LinCoord &cy, const LinCoord &cx)
{
    const float *r0 = p + (size_t)(g.y0 + cy.i0) * g.Win + g.x0;
    const float *r1 = p + (size_t)(g.y0 + cy.i1) * g.Win + g.x0;
    const float top = __fadd_rn(__fmul_rn(__ldg(r0 + cx.i0), cx.w0), __fmul_rn(__ldg(r0 + cx.i1), cx.w1));
    const float bot = __fadd_rn(__fmul_rn(__ldg(r1 + cx.i0), cx.w0), __fmul_rn(__ldg(r1 + cx.i1), cx.w1));
    return __fadd_rn(__fmul_rn(top, cy.w0), __fmul_rn(bot, cy.w1));
}

__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const float *__restrict__ src, ResizeGeom g, float *__restrict__ dst)
{
    const int plane = blockIdx.y;
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= g.Hout * g.Wout) return;
    const int oy = o / g.Wout, ox = o - oy * g.Wout;
    const LinCoord cy = lin_coord(oy, g.sy, g.Hc), cx = lin_coord(ox, g.sx, g.Wc);
    dst[(size_t)plane * g.Hout * g.Wout + o] = bilerp(src + (size_t)plane * g.Hin * g.Win, g, cy, cx);
}

// The four taps of an output pixel are the same offsets in every class plane: they are computed
// once as four pointers that advance by one plane per class: a plane costs four pointer
// increments + four loads instead of the 64-bit row / column arithmetic of bilerp() per tap.
template <bool SCORE>
__global__ void __launch_bounds__(256)
argmax_bilinear_kernel(const float *__restrict__ logits, int C, ResizeGeom g,
                       uint8_t *__restrict__ sem_out, float *__restrict__ score_out)
{
    const int b = blockIdx.y;
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= g.Hout * g.Wout) return;
    const int oy = o / g.Wout, ox = o - oy * g.Wout;
    const LinCoord cy = lin_coord(oy, g.sy, g.Hc), cx = lin_coord(ox, g.sx, g.Wc);
    const size_t plane = (size_t)g.Hin * g.Win;
    const float *lp = logits + (size_t)b * C * plane;
    const float *r0 = lp + (size_t)(g.y0 + cy.i0) * g.Win + g.x0;
    const float *r1 = lp + (size_t)(g.y0 + cy.i1) * g.Win + g.x0;
    const float *p00 = r0 + cx.i0, *p01 = r0 + cx.i1, *p10 = r1 + cx.i0, *p11 = r1 + cx.i1;
    auto tap = [&]() {          // same operation order as bilerp()
        const float top = __fadd_rn(__fmul_rn(__ldg(p00), cx.w0), __fmul_rn(__ldg(p01), cx.w1));
        const float bot = __fadd_rn(__fmul_rn(__ldg(p10), cx.w0), __fmul_rn(__ldg(p11), cx.w1));
        p00 += plane; p01 += plane; p10 += plane; p11 += plane;
        return __fadd_rn(__fmul_rn(top, cy.w0), __fmul_rn(bot, cy.w1));
    };
    float best = tap(), sum = 1.0f;
    float taint = __fmul_rn(best, 0.0f);    // +0 while every resized logit is finite, else NaN
    int cls = 0;
    for (int c = 1; c < C; ++c) {
        const float v = tap();
        const bool gt = v > best;
        taint = __fmaf_rn(v, 0.0f, taint);
        if (SCORE) {
            // online soft-max, branch free: exp(best - v) if v is the new maximum, else exp(v - best)
            const float e = __expf(gt ? best - v : v - best);
            sum = gt ? sum * e + 1.0f : sum + e;
        }
        best = gt ? v : best;
        cls = gt ? c : cls;
    }
    if (taint != taint) {
        // non-finite resized logits (rare path): a NaN or +Inf value, or nothing but -Inf, makes the
        // reference's soft-max NaN in every class and its arg-max 0 (semantic.py:73-74)
        const float kInf = __int_as_float(0x7f800000);
        p00 -= (size_t)C * plane; p01 -= (size_t)C * plane; p10 -= (size_t)C * plane; p11 -= (size_t)C * plane;
        bool poisoned = false;
        float mx = -kInf;
        int arg = 0;
        for (int k = 0; k < C; ++k) {
            const float v = tap();
            poisoned |= (v != v) || v == kInf;
            if (v > mx) { mx = v; arg = k; }
        }
        if (poisoned || mx == -kInf) {
            cls = 0;
            sum = __int_as_float(0x7fc00000);
        } else {
            cls = arg;
            p00 -= (size_t)C * plane; p01 -= (size_t)C * plane; p10 -= (size_t)C * plane; p11 -= (size_t)C * plane;
            float acc = 0.0f;
            for (int k = 0; k < C; ++k) acc += __expf(tap() - mx);
            sum = acc;
        }
    }
    const size_t q = (size_t)b * g.Hout * g.Wout + o;
    sem_out[q] = (uint8_t)cls;
    if (SCORE) score_out[q] = 1.0f / sum;
}

static bool make_geom(int Hin, int Win, int y0, int x0, int Hc, int Wc, int Hout, int Wout,
                      ResizeGeom *g)
{
    if (Hin < 1 || Win < 1 || Hc < 1 || Wc < 1 || Hout < 1 || Wout < 1) return false;
    if (y0 < 0 || x0 < 0 || y0 + Hc > Hin || x0 + Wc > Win) return false;
    if ((long long)Hout * Wout >= (1ll << 31) || (long long)Hin * Win >= (1ll << 31)) return false;
    g->Hin = Hin; g->Win = Win; g->y0 = y0; g->x0 = x0; g->Hc = Hc; g->Wc = Wc;
    g->Hout = Hout; g->Wout = Wout;
    g->sy = (float)Hc / (float)Hout;
    g->sx = (float)Wc / (float)Wout;
    return true;
}

}  // namespace npb

using namespace npb;

extern "C" int npb_resize_nearest(const void *src, int elem_size, int planes, int Hin, int Win,
                                  int y0, int x0, int Hc, int Wc, int Hout, int Wout, void *dst,
                                  void *stream)
{
    ResizeGeom g;
    if (!src || !dst || planes < 1 || planes > 65535 ||
        !make_geom(Hin, Win, y0, x0, Hc, Wc, Hout, Wout, &g))
        return NPB_ERR_ARG;
    dim3 grid((Hout * Wout + 255) / 256, planes);
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_size) {
        case 1: resize_nearest_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)src, g, (uint8_t *)dst); break;
        case 4: resize_nearest_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t *)src, g, (uint32_t *)dst); break;
        case 8: resize_nearest_kernel<unsigned long long><<<grid, 256, 0, s>>>(
                    (const unsigned long long *)src, g, (unsigned long long *)dst); break;
        default: return NPB_ERR_ARG;
    }
    return record_launch("npb_resize_nearest");
}

extern "C" int npb_resize_bilinear(const float *src, int planes, int Hin, int Win, int y0, int x0,
                                   int Hc, int Wc, int Hout, int Wout, float *dst, void *stream)
{
    ResizeGeom g;
    if (!src || !dst || planes < 1 || planes > 65535 ||
        !make_geom(Hin, Win, y0, x0, Hc, Wc, Hout, Wout, &g))
        return NPB_ERR_ARG;
    dim3 grid((Hout * Wout + 255) / 256, planes);
    resize_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, g, dst);
    return record_launch("npb_resize_bilinear");
}

extern "C" int npb_semantic_argmax_resized(const float *logits, int B, int C, int Hin, int Win,
                                           int y0, int x0, int Hc, int Wc, int Hout, int Wout,
                                           uint8_t *sem_out, float *score_out, void *stream)
{
    ResizeGeom g;
    if (!logits || !sem_out || B < 1 || B > 65535 || C < 1 || C > 256 ||
        !make_geom(Hin, Win, y0, x0, Hc, Wc, Hout, Wout, &g))
        return NPB_ERR_ARG;
    dim3 grid((Hout * Wout + 255) / 256, B);
    cudaStream_t s = (cudaStream_t)stream;
    if (score_out) argmax_bilinear_kernel<true><<<grid, 256, 0, s>>>(logits, C, g, sem_out, score_out);
    else argmax_bilinear_kernel<false><<<grid, 256, 0, s>>>(logits, C, g, sem_out, score_out);
    return record_launch("npb_semantic_argmax_resized");
}
