// group.cu -- fused semantic arg-max + offset grouping + class votes + orientation sums
//
// One streaming pass over the decoder outputs of a batch.  Per pixel (4 consecutive pixels
// per thread, 128-bit loads):
//   1. arg-max over the C logit planes (first maximal index)      semantic.py:52-53
//   2. foreground = class is a thing                              panoptic.py:118-128
//   3. loc = (y, x) + offset * (H, W)   (one f32 mul, one f32 add) panoptic.py:105-111,
//                                                                  instance.py:194
//   4. nearest centre: argmin_i sqrtf(fmaf(dx,dx,dy*dy)), first index on ties
//                                                                  instance.py:223-236
//   5. optional distance threshold                                 instance.py:246-247
//   6. votes hist[instance][class] += 1 and sum(cos), sum(sin) per instance
//                                      panoptic_merge.py:194-199, instance.py:301-310
// The centre list of the frame (<= 255 (y, x) pairs) is staged in shared memory.
//
// Exactness of step 4: the reference compares the SQRT'ed f32
// distances.  sqrtf is monotone, so a later centre can only win if its squared distance is
// smaller; it is accepted without a square root when it is smaller by more than 2^-20
// relative (then the rounded roots differ for sure) and otherwise the two IEEE sqrtf values
// are compared -- identical decisions to the reference at ~5 flops per (pixel, centre).
//
// Votes and sums are warp-aggregated: pixels of a warp almost always share (instance,
// class), so a warp issues one RED per distinct key instead of one per pixel.
#include "common.cuh"

namespace npb {

// CTA sizes of the grouping kernel: 256 threads x 4 resident CTAs per SM, or 64 threads x 18
// (same 56 registers per thread, 1152 instead of 1024 threads per SM).  The smaller CTAs make the
// ramp and the last, partial wave of a small batch finer grained (8 frames 480x640: 0.865 of the
// HBM peak with 256-thread CTAs, 0.915 with 128 x 9, 0.936 with 64 x 18); the launcher picks the
// variant whose grid fills its waves best.
constexpr int kGroupThreads = 256;
#ifndef NPB_GROUP_THREADS_SMALL
#define NPB_GROUP_THREADS_SMALL 64
#endif
constexpr int kGroupThreadsSmall = NPB_GROUP_THREADS_SMALL;
#ifndef NPB_GROUP_CTAS
#define NPB_GROUP_CTAS 4
#endif
#ifndef NPB_GROUP_CTAS_SMALL
#define NPB_GROUP_CTAS_SMALL 18
#endif
#ifndef NPB_GROUP_U
#define NPB_GROUP_U 8
#endif
constexpr int kGroupCtasPerSm = NPB_GROUP_CTAS, kGroupCtasPerSmSmall = NPB_GROUP_CTAS_SMALL;

enum SemSource { kFromLogits = 0, kFromSemMap = 1, kFromFgMask = 2 };

struct GroupParams {
    const float *logits;
    const uint8_t *sem_in;
    const uint8_t *fg_in;
    const float *offset;
    const float *orientation;
    const int32_t *centers_yx;
    const int32_t *n_centers;
    uint8_t *sem_out;
    uint8_t *inst_out;
    uint32_t *vote_hist;
    double *ori_sum;
    int C, H, W, P;
    float fH, fW;
    float inv_W;                // 1 / W rounded to nearest (row estimate, see kernel)
    int normalized, use_thr;
    float dist_thr;
    ClassSet thing;
    NPB_TL_FIELD
};

template <int VEC>
struct PixVec;
template <>
struct PixVec<4> {
    using F = float4;
    __device__ static void loadf(const float *p, float (&v)[4], bool stream)
    {
        const float4 t = stream ? ld_stream_f4((const float4 *)p) : *(const float4 *)p;
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static void loadb(const uint8_t *p, int (&v)[4])
    {
        const uint32_t t = *(const uint32_t *)p;
        v[0] = t & 255u; v[1] = (t >> 8) & 255u; v[2] = (t >> 16) & 255u; v[3] = t >> 24;
    }
    __device__ static void storeb(uint8_t *p, const int (&v)[4])
    {
        *(uint32_t *)p = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) |
                         ((uint32_t)v[3] << 24);
    }
};
template <>
struct PixVec<1> {
    __device__ static void loadf(const float *p, float (&v)[1], bool stream)
    {
        v[0] = stream ? ld_stream_f1(p) : *p;
    }
    __device__ static void loadb(const uint8_t *p, int (&v)[1]) { v[0] = *p; }
    __device__ static void storeb(uint8_t *p, const int (&v)[1]) { *p = (uint8_t)v[0]; }
};

// max(a, b), NaN if either is NaN (PTX max.NaN.f32)
__device__ __forceinline__ float fmax_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// exact "does centre i beat the current best" test on squared distances (see header)
__device__ __forceinline__ void consider_center(float s, int i, float &sbest, int &ibest)
{
    const float kSafe = 0.99999905f;  // 1 - 2^-20
    if (s < sbest) {
        if (s < __fmul_rn(sbest, kSafe) || __fsqrt_rn(s) < __fsqrt_rn(sbest)) {
            sbest = s;
            ibest = i;
        }
    }
}

template <int VEC, int MODE, bool ORI, int NT = kGroupThreads>
__global__ void __launch_bounds__(NT, NT == kGroupThreads ? kGroupCtasPerSm : kGroupCtasPerSmSmall)
group_pixels_kernel(const GroupParams prm)
{
    // centres of the frame as (cy, cy, cx, cx): one LDS.128 feeds two packed f32x2 operands
    __shared__ float4 s_centers[kMaxInst];
    // thing flag per class: one LDS.U8 per pixel instead of an indexed constant load + shifts
    __shared__ unsigned char s_thing[256];
    NPB_TL(prm, 1, start);
    // the successor (id writer / evaluation pixel pass) is staged while the last wave of this grid
    // runs; its residency is limited by its (padded) shared memory, so its one wave of CTAs
    // spreads evenly however early it is placed
    grid_launch_dependents();
    for (int c = threadIdx.x; c < 256; c += NT) s_thing[c] = prm.thing.has(c) ? 1 : 0;

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int P = prm.P, W = prm.W, C = prm.C;

    const int p0 = (blockIdx.x * NT + tid) * VEC;
    const bool active = p0 < P;  // P % VEC == 0 is guaranteed by the launcher
    const size_t fb = (size_t)b * P + p0;
    // The arg-max over the logits (most of this kernel's bytes) depends on nothing the chain
    // produces: with a programmatic launch it runs while the centre detection is still busy.
    // Class / foreground maps of the other modes come from kernels of the chain: wait first.
    if (MODE != kFromLogits) grid_dependency_wait();

    // ---- 1. semantic class ------------------------------------------------------------
    int cls[VEC];
    bool fg[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { cls[j] = 0; fg[j] = false; }
    if (active) {
        if (MODE == kFromLogits) {
            const float *lp = prm.logits + (size_t)b * C * P + p0;
            // Batches of 8 independent 128-bit loads in flight per thread before the first
            // compare: the loop is latency bound, bytes in flight are what buys bandwidth.  All
            // C planes go through ceil(C / 8) batches (plane 0 included, the last batch with its
            // loads predicated): every batch is one DRAM latency in the life of the CTA.
            constexpr int U = NPB_GROUP_U;
            // Non-finite logits: the reference takes the arg-max of softmax(logits), which is NaN
            // in every class -- index 0 -- as soon as a logit is NaN or +Inf, or all are -Inf
            // (semantic.py:52-53); -Inf next to finite logits just has probability 0.  The
            // running maximum is kept with max.NaN (NaN-propagating, same cost as a select): once
            // a NaN is in, no later logit compares greater and it stays to the end.
            float best[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) best[j] = __int_as_float(0xff800000);
            auto consume = [&](const float (&v)[VEC], int c) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (v[j] > best[j]) cls[j] = c;                 // first maximum wins
                    best[j] = fmax_nan(best[j], v[j]);
                }
            };
            int c = 0;
            for (; c + U <= C; c += U) {
                float v[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) PixVec<VEC>::loadf(lp + (size_t)(c + u) * P, v[u], true);
#pragma unroll
                for (int u = 0; u < U; ++u) consume(v[u], c + u);
            }
            if (c < C) {        // up to 7 planes, loads of planes >= C predicated off
                float v[U - 1][VEC];
#pragma unroll
                for (int u = 0; u < U - 1; ++u) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[u][j] = __int_as_float(0xff800000);   // -inf: neutral
                    if (c + u < C) PixVec<VEC>::loadf(lp + (size_t)(c + u) * P, v[u], true);
                }
#pragma unroll
                for (int u = 0; u < U - 1; ++u) consume(v[u], c + u);
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (!(best[j] < __int_as_float(0x7f800000))) cls[j] = 0;    // NaN or +Inf came by
            PixVec<VEC>::storeb(prm.sem_out + fb, cls);
        } else if (MODE == kFromSemMap) {
            PixVec<VEC>::loadb(prm.sem_in + fb, cls);
        }
        if (MODE == kFromFgMask) {
            int m[VEC];
            PixVec<VEC>::loadb(prm.fg_in + fb, m);
#pragma unroll
            for (int j = 0; j < VEC; ++j) fg[j] = (m[j] != 0);
        }
    }
    // centres of the frame (the centre detection has completed after this wait)
    if (MODE == kFromLogits) grid_dependency_wait();
    NPB_TL(prm, 1, wait);
    const int n = prm.n_centers[b];
    NPB_ASSERT(n >= 0 && n < kMaxInst);
    for (int i = tid; i < n; i += NT) {
        const int32_t *c = prm.centers_yx + ((size_t)b * kMaxInst + i) * 2;
        const float cy = (float)c[0], cx = (float)c[1];
        s_centers[i] = make_float4(cy, cy, cx, cx);
    }
    __syncthreads();        // also orders s_thing
    if (MODE != kFromFgMask) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) fg[j] = active && s_thing[cls[j] & 255] != 0;
    }

    // ---- 2. nearest centre ------------------------------------------------------------
    int inst[VEC];
    bool any_fg = false;
#pragma unroll
    for (int j = 0; j < VEC; ++j) { inst[j] = 0; any_fg |= fg[j]; }
    any_fg = any_fg && n > 0;

    float oc[VEC], os[VEC];
    if (ORI) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { oc[j] = 0.0f; os[j] = 0.0f; }
    }
    float nly[VEC], nlx[VEC];   // NEGATED locations: c - l == c + (-l) exactly
#pragma unroll
    for (int j = 0; j < VEC; ++j) { nly[j] = 0.0f; nlx[j] = 0.0f; }
    if (any_fg) {
        float oy[VEC], ox[VEC];
        PixVec<VEC>::loadf(prm.offset + (size_t)b * 2 * P + p0, oy, true);
        PixVec<VEC>::loadf(prm.offset + (size_t)b * 2 * P + P + p0, ox, true);
        if (ORI) {   // issued here so that the distance loop hides their latency
            PixVec<VEC>::loadf(prm.orientation + (size_t)b * 2 * P + p0, oc, true);
            PixVec<VEC>::loadf(prm.orientation + (size_t)b * 2 * P + P + p0, os, true);
        }
        // row / column of the first pixel: float estimate of p0 / W, corrected by at most one
        // (exact for p0 < 2^24; larger frames take the integer division)
        int y, x;
        if (P <= (1 << 24) && W >= 4) {
            y = __float2int_rd(__fmul_rn((float)p0, prm.inv_W));
            x = p0 - y * W;
            if (x < 0) { x += W; --y; } else if (x >= W) { x -= W; ++y; }
        } else {
            y = p0 / W;
            x = p0 - y * W;
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float dy = oy[j], dx = ox[j];
            if (prm.normalized) { dy = __fmul_rn(dy, prm.fH); dx = __fmul_rn(dx, prm.fW); }
            // -(y + dy) == (-y) + (-dy) exactly (IEEE negation is symmetric)
            nly[j] = __fadd_rn(-(float)y, -dy);
            nlx[j] = __fadd_rn(-(float)x, -dx);
            if (++x == W) { x = 0; ++y; }
        }
    }
    // The centre loop is warp-cooperative (it prunes centres per warp), so every lane of a
    // warp that holds a thing pixel runs it; lanes without one just carry neutral values.
    if (__any_sync(kFullMask, any_fg)) {
        float sbest[VEC];
        int ibest[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sbest[j] = __int_as_float(0x7f800000); ibest[j] = 0; }

        // ---- centre pruning.  Bounding box of the warp's thing-pixel locations; for a centre c
        // lb(c) / ub(c) bound the squared distance from ANY point of the box.  Every pixel's
        // nearest centre is at most U = min_c ub(c) away, so a centre with lb(c) > U can neither
        // win nor tie for any pixel of the warp and is skipped.  The 1e-5 slack dwarfs the
        // f32 rounding of the bounds and of the exact distances (< 1e-6 relative), so the
        // surviving set always contains every centre the exact arg-min could pick; survivors
        // are visited in ascending index, which keeps the first-index tie rule.
        const float kInf = __int_as_float(0x7f800000);
        float ymin = kInf, ymax = -kInf, xmin = kInf, xmax = -kInf;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (fg[j]) {
                ymin = fminf(ymin, -nly[j]); ymax = fmaxf(ymax, -nly[j]);
                xmin = fminf(xmin, -nlx[j]); xmax = fmaxf(xmax, -nlx[j]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ymin = fminf(ymin, __shfl_xor_sync(kFullMask, ymin, o));
            ymax = fmaxf(ymax, __shfl_xor_sync(kFullMask, ymax, o));
            xmin = fminf(xmin, __shfl_xor_sync(kFullMask, xmin, o));
            xmax = fmaxf(xmax, __shfl_xor_sync(kFullMask, xmax, o));
        }
        float U = kInf;
        for (int i = lane; i < n; i += 32) {
            const float4 c = s_centers[i];
            const float dy = fmaxf(fabsf(c.x - ymin), fabsf(c.x - ymax));
            const float dx = fmaxf(fabsf(c.z - xmin), fabsf(c.z - xmax));
            U = fminf(U, dy * dy + dx * dx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) U = fminf(U, __shfl_xor_sync(kFullMask, U, o));
        // NaN / Inf boxes (garbage offsets) make U = Inf or NaN: then nothing is pruned
        const float prune_above = (U == U) ? U * 1.00001f : kInf;

        if (VEC == 4) {
            // Blackwell packed FP32 (FADD2 / FMUL2 / FFMA2): two pixels per instruction, each
            // lane individually IEEE-rounded exactly like the scalar sequence
            const float2 nlyA = make_float2(nly[0], nly[1 % VEC]), nlyB = make_float2(nly[2 % VEC], nly[3 % VEC]);
            const float2 nlxA = make_float2(nlx[0], nlx[1 % VEC]), nlxB = make_float2(nlx[2 % VEC], nlx[3 % VEC]);
            for (int base = 0; base < n; base += 32) {
                bool keep = false;
                if (base + lane < n) {
                    const float4 c = s_centers[base + lane];
                    const float dy = fmaxf(fmaxf(ymin - c.x, c.x - ymax), 0.0f);
                    const float dx = fmaxf(fmaxf(xmin - c.z, c.z - xmax), 0.0f);
                    keep = !(dy * dy + dx * dx > prune_above);
                }
                unsigned todo = __ballot_sync(kFullMask, keep);
                while (todo) {
                    const int i = base + __ffs(todo) - 1;
                    todo &= todo - 1;
                    const float4 c = s_centers[i];
                    const float2 cy2 = make_float2(c.x, c.y), cx2 = make_float2(c.z, c.w);
                    const float2 a0 = __fadd2_rn(cy2, nlyA), a1 = __fadd2_rn(cx2, nlxA);
                    const float2 b0 = __fadd2_rn(cy2, nlyB), b1 = __fadd2_rn(cx2, nlxB);
                    const float2 sA = __ffma2_rn(a1, a1, __fmul2_rn(a0, a0));
                    const float2 sB = __ffma2_rn(b1, b1, __fmul2_rn(b0, b0));
                    if ((sA.x < sbest[0]) | (sA.y < sbest[1 % VEC]) | (sB.x < sbest[2 % VEC]) |
                        (sB.y < sbest[3 % VEC])) {
                        consider_center(sA.x, i, sbest[0], ibest[0]);
                        consider_center(sA.y, i, sbest[1 % VEC], ibest[1 % VEC]);
                        consider_center(sB.x, i, sbest[2 % VEC], ibest[2 % VEC]);
                        consider_center(sB.y, i, sbest[3 % VEC], ibest[3 % VEC]);
                    }
                }
            }
        } else {
            for (int i = 0; i < n; ++i) {
                const float4 c = s_centers[i];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float d0 = __fadd_rn(c.x, nly[j]);
                    const float d1 = __fadd_rn(c.z, nlx[j]);
                    consider_center(__fmaf_rn(d1, d1, __fmul_rn(d0, d0)), i, sbest[j], ibest[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (fg[j]) {
                inst[j] = ibest[j] + 1;
                if (prm.use_thr && __fsqrt_rn(sbest[j]) > prm.dist_thr) inst[j] = 0;
            }
        }
    }
    if (active) PixVec<VEC>::storeb(prm.inst_out + fb, inst);
    NPB_TL(prm, 1, end);

    // ---- 3. class votes (warp aggregated: MATCH.ANY + REDUX, one RED per group) -----------
    const int CH = (MODE == kFromFgMask) ? 1 : C;
    uint32_t *hist = prm.vote_hist + (size_t)b * kMaxInst * CH;
    int key[VEC];
    bool any_inst = false;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        key[j] = inst[j] > 0 ? inst[j] * CH + (MODE == kFromFgMask ? 0 : cls[j]) : -1;
        any_inst |= (inst[j] > 0);
    }
    if (!__any_sync(kFullMask, any_inst)) return;
    {
        // the thread's first instance pixel leads; pixels with another key go individually
        int k0 = -1;
#pragma unroll
        for (int j = VEC - 1; j >= 0; --j)
            if (key[j] >= 0) k0 = key[j];
        int cnt0 = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) cnt0 += (key[j] == k0 && k0 >= 0) ? 1 : 0;
        const unsigned peers = __match_any_sync(kFullMask, k0);
        const int total = __reduce_add_sync(peers, cnt0);
        if (k0 >= 0) {
            NPB_ASSERT(k0 < kMaxInst * CH && total >= 1 && total <= 32 * VEC);
            if (lane == __ffs(peers) - 1) atomicAdd(hist + k0, (uint32_t)total);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (key[j] >= 0 && key[j] != k0) atomicAdd(hist + key[j], 1u);
        }
    }

    // ---- 4. orientation sums per instance: loop over the (few) distinct instances of the warp,
    //         f32 shuffle tree inside the warp, one f64 RED pair per instance ----------------
    if (ORI) {
        double *osum = prm.ori_sum + (size_t)b * kMaxInst * 2;
        unsigned pending = 0u;
#pragma unroll
        for (int j = 0; j < VEC; ++j) pending |= (inst[j] > 0 ? 1u : 0u) << j;
        while (true) {
            const unsigned has = __ballot_sync(kFullMask, pending != 0u);
            if (!has) break;
            const int leader = __ffs(has) - 1;
            int mine = 0;
#pragma unroll
            for (int j = VEC - 1; j >= 0; --j)
                if ((pending >> j) & 1u) mine = inst[j];
            const int cur = __shfl_sync(kFullMask, mine, leader);
            float sc = 0.0f, ss = 0.0f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (((pending >> j) & 1u) && inst[j] == cur) {
                    sc += oc[j];
                    ss += os[j];
                    pending &= ~(1u << j);
                }
            }
            sc = warp_sum(sc);
            ss = warp_sum(ss);
            if (lane == 0) {
                atomicAdd(osum + 2 * cur, (double)sc);
                atomicAdd(osum + 2 * cur + 1, (double)ss);
            }
        }
    }
}

static int group_sm_count()
{
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// fraction of the CTA slots of all waves that a grid of `ctas` CTAs keeps busy
static double wave_fill(long long ctas, long long slots)
{
    const long long waves = (ctas + slots - 1) / slots;
    return (double)ctas / (double)(waves * slots);
}

template <int VEC, int MODE, bool ORI>
static void launch_group_nt(const GroupParams &prm, int B, cudaStream_t s)
{
    const long long groups = prm.P / VEC;       // threads needed per frame
    const long long big = (groups + kGroupThreads - 1) / kGroupThreads * B;
    const long long small = (groups + kGroupThreadsSmall - 1) / kGroupThreadsSmall * B;
    static int n_sm = group_sm_count();
    // the small CTAs only pay where the batch is a handful of waves: their last wave is
    // finer grained (and 18 x 64 threads are resident per SM instead of 4 x 256)
    const bool use_small = VEC == 4 && big <= 16ll * n_sm * kGroupCtasPerSm &&
                           wave_fill(small, (long long)n_sm * kGroupCtasPerSmSmall) >
                               wave_fill(big, (long long)n_sm * kGroupCtasPerSm) + 0.02;
    if (use_small) {
        dim3 grid((unsigned)((groups + kGroupThreadsSmall - 1) / kGroupThreadsSmall), B);
        launch_dependent(group_pixels_kernel<VEC, MODE, ORI, kGroupThreadsSmall>, grid,
                         dim3(kGroupThreadsSmall), 0, s, prm);
    } else {
        dim3 grid((unsigned)((groups + kGroupThreads - 1) / kGroupThreads), B);
        launch_dependent(group_pixels_kernel<VEC, MODE, ORI, kGroupThreads>, grid, dim3(kGroupThreads),
                         0, s, prm);
    }
}

template <int VEC, int MODE>
static void launch_group(const GroupParams &prm, int B, bool ori, cudaStream_t s)
{
    if (ori) launch_group_nt<VEC, MODE, true>(prm, B, s);
    else launch_group_nt<VEC, MODE, false>(prm, B, s);
}

// ---- more than 255 centres: the reference's uint8 wrap ----------------------------------------
// One frame whose centre list does not fit the uint8 ids (npb_overflow_centers).  The reference
// goes on regardless (instance.py:231-236): arg-min over ALL centres, id = uint8(arg + 1), so
// centre 256 becomes "no instance", centre 257 joins instance 1, ...  Correctness path for a
// pathological regime (hundreds of exactly tied heat-map peaks): one pixel per thread, the
// centres streamed from global memory, the reference's distance formula as written (one sqrtf
// per pair, first minimum wins, the oracle's loop), plain atomics for votes / orientation sums.
struct WideGroupParams {
    const uint8_t *sem_in, *fg_in;
    const float *offset, *orientation;
    const int32_t *centers_yx, *n_centers;
    int P, W, C, normalized, use_thr;
    float fH, fW, dist_thr;
    ClassSet thing;
    uint8_t *inst_out;
    uint32_t *vote_hist;
    double *ori_sum;
    int32_t *n_rows_out, *status;
};

__global__ void __launch_bounds__(256) group_pixels_wide_kernel(const WideGroupParams prm)
{
    const int n = *prm.n_centers;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *prm.n_rows_out = n < kMaxInst - 1 ? n : kMaxInst - 1;     // rows of the instance tables
        *prm.status = NPB_OK;                                       // the frame has been redone
    }
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= prm.P) return;
    const int cls = prm.sem_in ? prm.sem_in[p] : 0;
    const bool fg = prm.sem_in ? prm.thing.has(cls) : prm.fg_in[p] != 0;
    int id = 0;
    if (fg && n > 0) {
        const int y = p / prm.W, x = p - y * prm.W;
        float oy = prm.offset[p], ox = prm.offset[(size_t)prm.P + p];
        if (prm.normalized) { oy = __fmul_rn(oy, prm.fH); ox = __fmul_rn(ox, prm.fW); }
        const float ly = __fadd_rn((float)y, oy), lx = __fadd_rn((float)x, ox);
        float best = 0.0f;
        int arg = 0;
        for (int i = 0; i < n; ++i) {
            const float d0 = __fsub_rn((float)__ldg(prm.centers_yx + 2 * i), ly);
            const float d1 = __fsub_rn((float)__ldg(prm.centers_yx + 2 * i + 1), lx);
            const float d = __fsqrt_rn(__fmaf_rn(d1, d1, __fmul_rn(d0, d0)));
            if (i == 0 || d < best) { best = d; arg = i; }
        }
        id = (arg + 1) & 255;                                        // instance.py:236
        if (prm.use_thr && best > prm.dist_thr) id = 0;             // instance.py:246
    }
    prm.inst_out[p] = (uint8_t)id;
    if (id > 0) {
        const int CH = prm.sem_in ? prm.C : 1;
        atomicAdd(prm.vote_hist + id * CH + (prm.sem_in ? cls : 0), 1u);
        if (prm.ori_sum) {
            atomicAdd(prm.ori_sum + 2 * id, (double)prm.orientation[p]);
            atomicAdd(prm.ori_sum + 2 * id + 1, (double)prm.orientation[(size_t)prm.P + p]);
        }
    }
}

}  // namespace npb

using namespace npb;

static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// internal form: `cleared` = the caller has already zeroed vote_hist and ori_sum
int npb::group_pixels_impl(const float *logits, const uint8_t *sem_in, const uint8_t *fg_in,
                           const float *offset, const float *orientation, int B, int C, int H, int W,
                           const uint8_t *h_thing_lut, const int32_t *centers_yx,
                           const int32_t *n_centers, int normalized_offset,
                           int use_distance_threshold, float distance_threshold, uint8_t *sem_out,
                           uint8_t *inst_out, uint32_t *vote_hist, double *ori_sum, bool cleared,
                           void *stream)
{
    const int n_src = (logits != nullptr) + (sem_in != nullptr) + (fg_in != nullptr);
    if (n_src != 1 || !offset || !centers_yx || !n_centers || !inst_out || !vote_hist)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || C < 1 || C > 256 || H < 1 || W < 1) return NPB_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return NPB_ERR_ARG;
    if (logits && !sem_out) return NPB_ERR_ARG;
    if (!fg_in && !h_thing_lut) return NPB_ERR_ARG;
    if (fg_in && C != 1) return NPB_ERR_ARG;
    if ((orientation != nullptr) != (ori_sum != nullptr)) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;

    GroupParams prm;
    prm.logits = logits; prm.sem_in = sem_in; prm.fg_in = fg_in;
    prm.offset = offset; prm.orientation = orientation;
    prm.centers_yx = centers_yx; prm.n_centers = n_centers;
    prm.sem_out = sem_out; prm.inst_out = inst_out;
    prm.vote_hist = vote_hist; prm.ori_sum = ori_sum;
    prm.C = C; prm.H = H; prm.W = W; prm.P = H * W;
    prm.fH = (float)H; prm.fW = (float)W;
    prm.inv_W = 1.0f / (float)W;
    prm.normalized = normalized_offset; prm.use_thr = use_distance_threshold;
    prm.dist_thr = distance_threshold;
    prm.thing = make_class_set(h_thing_lut, fg_in ? 0 : C);
    NPB_TL_SET(prm);

    const int CH = fg_in ? 1 : C;
    if (!cleared) {
        cudaMemsetAsync(vote_hist, 0, (size_t)B * kMaxInst * CH * sizeof(uint32_t), s);
        if (ori_sum) cudaMemsetAsync(ori_sum, 0, (size_t)B * kMaxInst * 2 * sizeof(double), s);
    }

    const bool vec4 = (prm.P % 4 == 0) && W >= 4 && aligned16(logits) && aligned16(offset) &&
                      aligned16(orientation) && (((uintptr_t)sem_in | (uintptr_t)fg_in |
                                                  (uintptr_t)sem_out | (uintptr_t)inst_out) & 3u) == 0;
    const bool ori = orientation != nullptr;
    if (logits) {
        if (vec4) launch_group<4, kFromLogits>(prm, B, ori, s);
        else launch_group<1, kFromLogits>(prm, B, ori, s);
    } else if (sem_in) {
        if (vec4) launch_group<4, kFromSemMap>(prm, B, ori, s);
        else launch_group<1, kFromSemMap>(prm, B, ori, s);
    } else {
        if (vec4) launch_group<4, kFromFgMask>(prm, B, ori, s);
        else launch_group<1, kFromFgMask>(prm, B, ori, s);
    }
    return record_launch("npb_group_pixels");
}

extern "C" int npb_group_pixels(const float *logits, const uint8_t *sem_in, const uint8_t *fg_in,
                                const float *offset, const float *orientation, int B, int C,
                                int H, int W, const uint8_t *h_thing_lut,
                                const int32_t *centers_yx, const int32_t *n_centers,
                                int normalized_offset, int use_distance_threshold,
                                float distance_threshold, uint8_t *sem_out, uint8_t *inst_out,
                                uint32_t *vote_hist, double *ori_sum, void *stream)
{
    return group_pixels_impl(logits, sem_in, fg_in, offset, orientation, B, C, H, W, h_thing_lut,
                             centers_yx, n_centers, normalized_offset, use_distance_threshold,
                             distance_threshold, sem_out, inst_out, vote_hist, ori_sum, false, stream);
}

extern "C" int npb_group_pixels_wide(const uint8_t *sem_in, const uint8_t *fg_in, const float *offset,
                                     const float *orientation, int C, int H, int W,
                                     const uint8_t *h_thing_lut, const int32_t *centers_yx,
                                     const int32_t *n_centers, int normalized_offset,
                                     int use_distance_threshold, float distance_threshold,
                                     uint8_t *inst_out, uint32_t *vote_hist, double *ori_sum,
                                     int32_t *n_rows_out, int32_t *status, void *stream)
{
    if (((sem_in != nullptr) + (fg_in != nullptr)) != 1 || !offset || !centers_yx || !n_centers ||
        !inst_out || !vote_hist || !n_rows_out || !status)
        return NPB_ERR_ARG;
    if (C < 1 || C > 256 || H < 1 || W < 1 || (long long)H * W >= (1ll << 30)) return NPB_ERR_ARG;
    if (sem_in && !h_thing_lut) return NPB_ERR_ARG;
    if (fg_in && C != 1) return NPB_ERR_ARG;
    if ((orientation != nullptr) != (ori_sum != nullptr)) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    WideGroupParams prm;
    prm.sem_in = sem_in; prm.fg_in = fg_in; prm.offset = offset; prm.orientation = orientation;
    prm.centers_yx = centers_yx; prm.n_centers = n_centers;
    prm.P = H * W; prm.W = W; prm.C = C; prm.normalized = normalized_offset;
    prm.use_thr = use_distance_threshold; prm.fH = (float)H; prm.fW = (float)W;
    prm.dist_thr = distance_threshold;
    prm.thing = make_class_set(h_thing_lut, fg_in ? 0 : C);
    prm.inst_out = inst_out; prm.vote_hist = vote_hist; prm.ori_sum = ori_sum;
    prm.n_rows_out = n_rows_out; prm.status = status;
    const int CH = fg_in ? 1 : C;
    cudaMemsetAsync(vote_hist, 0, (size_t)kMaxInst * CH * sizeof(uint32_t), s);
    if (ori_sum) cudaMemsetAsync(ori_sum, 0, (size_t)kMaxInst * 2 * sizeof(double), s);
    group_pixels_wide_kernel<<<(prm.P + 255) / 256, 256, 0, s>>>(prm);
    return record_launch("npb_group_pixels_wide");
}
