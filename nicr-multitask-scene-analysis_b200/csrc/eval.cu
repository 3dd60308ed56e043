// eval.cu -- mIoU confusion matrix and PQ segment matching / accumulation
//
// Replaces (reference: metric/miou.py:44-56, metric/pq.py:60-179, 298-303):
//   * confmat[target][pred] += 1 over all pixels
//   * per frame: areas of ground-truth / predicted segments and of their intersections
//     (three torch.unique(return_counts) sorts of H*W int64 in the reference), IoU matching,
//     TP / FN / FP counting, float64 IoU sums
//   * state += per-frame result, in frame order
//
// Design: every pixel belongs to exactly one (gt segment, pred segment) pair, so ONE
// contingency table per frame (pair -> pixel count) carries all three area tables
// (gt area = sum over pairs of that gt id, pred area likewise).  The pixel pass
// (pair_count_kernel) streams pred / target once (17 B/px with the fused semantic target),
// counts pairs in a per-CTA shared-memory hash table after warp-level aggregation
// (neighbouring pixels nearly always share the pair) and flushes it into a small per-frame
// global hash table.  match_frames_kernel (one CTA per frame) sorts the <= 2048 pairs by
// `target*offset + pred` -- the reference's visiting order, which fixes the float64
// summation order -- and does the matching; accumulate_frames_kernel adds the frames to
// the running state in frame order.  Both float64 orders equal the reference's, so the
// states are bit-identical, not merely close.
#include "common.cuh"

namespace npb {

constexpr int kPairThreads = 256;
constexpr int kSmemSlots = 512;        // per-CTA pair hash table
constexpr int kFrameSlots = 8192;      // per-frame global pair hash table
constexpr int kMaxPairs = 2048;        // pairs per frame handled by the matcher
constexpr int kMatchThreads = 512;
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kSmemConfmatMaxN = 96;   // n*n*4 B <= 36 KB privatised in shared memory

__device__ __forceinline__ unsigned hash64(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return (unsigned)k;
}

// insert (key, cnt) into an open-addressing table; returns false when no slot was found
__device__ __forceinline__ bool table_add(unsigned long long *keys, unsigned *cnts, int slots,
                                          int max_probe, unsigned long long key, unsigned cnt)
{
    unsigned h = hash64(key) & (unsigned)(slots - 1);
    for (int probe = 0; probe < max_probe; ++probe) {
        unsigned long long k = keys[h];
        if (k == kEmptyKey) k = atomicCAS(keys + h, kEmptyKey, key);
        if (k == kEmptyKey || k == key) {
            atomicAdd(cnts + h, cnt);
            return true;
        }
        h = (h + 1) & (unsigned)(slots - 1);
    }
    return false;
}

struct PairParams {
    const long long *pred;
    const long long *target;
    const uint8_t *sem_target;  // nullable
    long long P;
    long long offset, L;
    int L_shift;                // >= 0 when L is a power of two
    int n;                      // confusion-matrix size (0 = no confmat)
    unsigned long long *frame_keys;  // [B][kFrameSlots]
    unsigned *frame_cnts;
    unsigned long long *confmat;     // [n][n] int64, accumulated
    int32_t *status;                 // [B]
};

template <int VEC, bool CONFMAT>
__global__ void __launch_bounds__(kPairThreads) pair_count_kernel(const PairParams prm)
{
    __shared__ unsigned long long s_keys[kSmemSlots];
    __shared__ unsigned s_cnts[kSmemSlots];
    extern __shared__ unsigned s_cm[];  // n*n when privatised

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = prm.n;
    const bool cm_smem = CONFMAT && n <= kSmemConfmatMaxN;
    for (int i = tid; i < kSmemSlots; i += kPairThreads) { s_keys[i] = kEmptyKey; s_cnts[i] = 0; }
    if (cm_smem)
        for (int i = tid; i < n * n; i += kPairThreads) s_cm[i] = 0;
    __syncthreads();

    unsigned long long *fkeys = prm.frame_keys + (size_t)b * kFrameSlots;
    unsigned *fcnts = prm.frame_cnts + (size_t)b * kFrameSlots;
    const long long P = prm.P;
    const long long chunk = (long long)kPairThreads * VEC;
    const long long n_chunks = (P + chunk - 1) / chunk;

    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long long p0 = ch * chunk + (long long)tid * VEC;
        const size_t fb = (size_t)b * P + p0;
        unsigned long long key[VEC];
        int ckey[VEC];
        bool valid[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { valid[j] = false; key[j] = 0; ckey[j] = -1; }
        if (p0 < P) {  // P % VEC == 0 guaranteed by the launcher
            long long pv[VEC], tv[VEC];
            if (VEC == 4) {
                const longlong2 a0 = __ldcs((const longlong2 *)(prm.pred + fb));
                const longlong2 a1 = __ldcs((const longlong2 *)(prm.pred + fb) + 1);
                const longlong2 t0 = __ldcs((const longlong2 *)(prm.target + fb));
                const longlong2 t1 = __ldcs((const longlong2 *)(prm.target + fb) + 1);
                pv[0] = a0.x; pv[1 % VEC] = a0.y; pv[2 % VEC] = a1.x; pv[3 % VEC] = a1.y;
                tv[0] = t0.x; tv[1 % VEC] = t0.y; tv[2 % VEC] = t1.x; tv[3 % VEC] = t1.y;
            } else {
                pv[0] = __ldcs(prm.pred + fb);
                tv[0] = __ldcs(prm.target + fb);
            }
            unsigned sw = 0;
            if (CONFMAT) {
                if (VEC == 4) sw = *(const unsigned *)(prm.sem_target + fb);
                else sw = prm.sem_target[fb];
            }
            long long last_p = -1;
            int last_c = 0;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                valid[j] = true;
                if (pv[j] < 0 || pv[j] >= prm.offset || tv[j] < 0) {
                    set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                    valid[j] = false;
                }
                key[j] = (unsigned long long)tv[j] * (unsigned long long)prm.offset +
                         (unsigned long long)pv[j];
                if (CONFMAT) {
                    if (pv[j] != last_p) {  // pred // L, once per run of equal ids
                        last_p = pv[j];
                        const long long c =
                            prm.L_shift >= 0 ? (pv[j] >> prm.L_shift) : (pv[j] / prm.L);
                        last_c = (c >= 0 && c < n) ? (int)c : -1;
                    }
                    const int st = (sw >> (8 * j)) & 255;
                    if (last_c < 0 || st >= n) {
                        if (valid[j]) set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                        ckey[j] = -1;
                    } else {
                        ckey[j] = st * n + last_c;
                    }
                }
            }
        }

        // ---- pair counts: warp-aggregate equal keys, then one shared-memory insert each ----
        unsigned pending = 0u;
#pragma unroll
        for (int j = 0; j < VEC; ++j) pending |= (valid[j] ? 1u : 0u) << j;
        while (true) {
            const unsigned has = __ballot_sync(kFullMask, pending != 0u);
            if (!has) break;
            const int leader = __ffs(has) - 1;
            unsigned long long mine = 0;
#pragma unroll
            for (int j = VEC - 1; j >= 0; --j)
                if ((pending >> j) & 1u) mine = key[j];
            const unsigned long long cur = __shfl_sync(kFullMask, mine, leader);
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (((pending >> j) & 1u) && key[j] == cur) { ++cnt; pending &= ~(1u << j); }
            cnt = __reduce_add_sync(kFullMask, cnt);
            if (lane == 0) {
                if (!table_add(s_keys, s_cnts, kSmemSlots, 16, cur, (unsigned)cnt))
                    if (!table_add(fkeys, fcnts, kFrameSlots, kFrameSlots, cur, (unsigned)cnt))
                        set_status(prm.status + b, NPB_ERR_CAPACITY);
            }
        }

        // ---- confusion matrix: same aggregation on (target class, pred class) -------------
        if (CONFMAT) {
            unsigned cpend = 0u;
#pragma unroll
            for (int j = 0; j < VEC; ++j) cpend |= (ckey[j] >= 0 ? 1u : 0u) << j;
            while (true) {
                const unsigned has = __ballot_sync(kFullMask, cpend != 0u);
                if (!has) break;
                const int leader = __ffs(has) - 1;
                int mine = -1;
#pragma unroll
                for (int j = VEC - 1; j >= 0; --j)
                    if ((cpend >> j) & 1u) mine = ckey[j];
                const int cur = __shfl_sync(kFullMask, mine, leader);
                int cnt = 0;
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (((cpend >> j) & 1u) && ckey[j] == cur) { ++cnt; cpend &= ~(1u << j); }
                cnt = __reduce_add_sync(kFullMask, cnt);
                if (lane == 0) {
                    if (cm_smem) atomicAdd(s_cm + cur, (unsigned)cnt);
                    else atomicAdd(prm.confmat + cur, (unsigned long long)cnt);
                }
            }
        }
    }

    // ---- flush the CTA tables ----------------------------------------------------------
    __syncthreads();
    for (int i = tid; i < kSmemSlots; i += kPairThreads) {
        const unsigned long long k = s_keys[i];
        if (k != kEmptyKey && s_cnts[i])
            if (!table_add(fkeys, fcnts, kFrameSlots, kFrameSlots, k, s_cnts[i]))
                set_status(prm.status + b, NPB_ERR_CAPACITY);
    }
    if (cm_smem)
        for (int i = tid; i < n * n; i += kPairThreads)
            if (s_cm[i]) atomicAdd(prm.confmat + i, (unsigned long long)s_cm[i]);
}

// ---------------------------------------------------------------------------------------
struct MatchParams {
    const unsigned long long *frame_keys;
    const unsigned *frame_cnts;
    int num_categories;
    long long ignored_label, L, offset, void_segment_id;
    double *frame_stats;   // [B][4][num_categories]
    long long *matches;    // [B][match_cap][2] nullable
    int match_cap;
    int32_t *n_matches;    // [B] nullable
    int32_t *status;       // [B]
};

__global__ void __launch_bounds__(kMatchThreads) match_frames_kernel(const MatchParams prm)
{
    extern __shared__ unsigned char smem_raw[];
    long long *s_key = (long long *)smem_raw;                   // [kMaxPairs]
    long long *s_g = s_key + kMaxPairs;                         // gt segment id
    long long *s_p = s_g + kMaxPairs;                           // pred segment id
    double *s_iou = (double *)(s_p + kMaxPairs);                // IoU of matched pairs
    unsigned *s_cnt = (unsigned *)(s_iou + kMaxPairs);          // intersection area
    int *s_gcat = (int *)(s_cnt + kMaxPairs);                   // category of the gt segment
    unsigned char *s_flag = (unsigned char *)(s_gcat + kMaxPairs);  // bit0 = matched (TP)
    __shared__ int s_m, s_nm;
    __shared__ int s_tp[256], s_fn[256], s_fp[256];

    const int b = blockIdx.x, tid = threadIdx.x;
    const int NC = prm.num_categories;
    const unsigned long long *fkeys = prm.frame_keys + (size_t)b * kFrameSlots;
    const unsigned *fcnts = prm.frame_cnts + (size_t)b * kFrameSlots;
    if (tid == 0) { s_m = 0; s_nm = 0; }
    for (int c = tid; c < 256; c += kMatchThreads) { s_tp[c] = 0; s_fn[c] = 0; s_fp[c] = 0; }
    __syncthreads();
    for (int i = tid; i < kFrameSlots; i += kMatchThreads) {
        const unsigned long long k = fkeys[i];
        if (k != kEmptyKey) {
            const int slot = atomicAdd(&s_m, 1);
            if (slot < kMaxPairs) { s_key[slot] = (long long)k; s_cnt[slot] = fcnts[i]; }
        }
    }
    __syncthreads();
    int m = s_m;
    if (m > kMaxPairs) {
        if (tid == 0) set_status(prm.status + b, NPB_ERR_CAPACITY);
        m = kMaxPairs;
    }
    // bitonic sort by key ascending (= torch.unique order of target*offset + pred, pq.py:109)
    int npad = 1;
    while (npad < m) npad <<= 1;
    for (int i = m + tid; i < npad; i += kMatchThreads) { s_key[i] = 0x7fffffffffffffffll; s_cnt[i] = 0; }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += kMatchThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = ((i & k) == 0);
                    const long long a = s_key[i], c = s_key[ixj];
                    if ((a > c) == up) {
                        s_key[i] = c; s_key[ixj] = a;
                        const unsigned t = s_cnt[i]; s_cnt[i] = s_cnt[ixj]; s_cnt[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < m; t += kMatchThreads) {
        const long long key = s_key[t];
        const long long g = key / prm.offset;       // ids are validated non-negative
        s_g[t] = g;
        s_p[t] = key - g * prm.offset;
        const long long gc = g / prm.L;
        s_gcat[t] = (gc >= 0 && gc < 0x7fffffff) ? (int)gc : 0x7fffffff;
        s_flag[t] = 0;
        s_iou[t] = 0.0;
    }
    __syncthreads();

    // pass 1: areas, IoU, match decision                                    pq.py:119-152
    for (int t = tid; t < m; t += kMatchThreads) {
        const long long g = s_g[t], p = s_p[t], key = s_key[t];
        if (key == prm.void_segment_id) continue;                          // pq.py:120
        const long long gcat = s_gcat[t], pcat = p / prm.L;
        if (gcat != pcat) continue;                                        // pq.py:128
        const long long void_key = prm.void_segment_id * prm.offset + p;
        long long tsa = 0, psa = 0, r = 0;
        for (int u = 0; u < m; ++u) {
            const long long cu = s_cnt[u];
            if (s_g[u] == g) tsa += cu;
            if (s_p[u] == p) psa += cu;
            if (s_key[u] == void_key) r = cu;
        }
        const long long ia = s_cnt[t];
        const long long uni = tsa + psa - ia - r;                          // pq.py:143
        if (uni == 0) { set_status(prm.status + b, NPB_ERR_ZERO_DIVISION); continue; }
        const double iou = (double)ia / (double)uni;                       // pq.py:145
        if (iou > 0.5) {
            if (gcat >= NC || gcat >= 256) { set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE); continue; }
            s_flag[t] = 1;
            s_iou[t] = iou;
            atomicAdd(&s_tp[(int)gcat], 1);
            if (prm.matches) {
                const int slot = atomicAdd(&s_nm, 1);
                if (slot < prm.match_cap) {
                    long long *o = prm.matches + ((size_t)b * prm.match_cap + slot) * 2;
                    o[0] = g;
                    o[1] = p;
                } else {
                    set_status(prm.status + b, NPB_ERR_CAPACITY);
                }
            }
        }
    }
    __syncthreads();

    // pass 2: false negatives / false positives                            pq.py:155-177
    for (int t = tid; t < m; t += kMatchThreads) {
        const long long g = s_g[t], p = s_p[t];
        const bool first_g = (t == 0) || (s_g[t - 1] != g);
        bool first_p = true, g_matched = false, p_matched = false;
        long long psa = 0, pio = 0;
        for (int u = 0; u < m; ++u) {
            const bool same_p = (s_p[u] == p);
            if (same_p) {
                if (u < t) first_p = false;
                psa += s_cnt[u];
                if ((long long)s_gcat[u] == prm.ignored_label) pio += s_cnt[u];   // pq.py:47-57
                if (s_flag[u]) p_matched = true;
            }
            if (s_flag[u] && s_g[u] == g) g_matched = true;
        }
        if (first_g && !g_matched) {
            const long long cat = s_gcat[t];
            if (cat != prm.ignored_label) {                                 // pq.py:161
                if (cat >= NC || cat >= 256) set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                else atomicAdd(&s_fn[(int)cat], 1);
            }
        }
        if (first_p && !p_matched) {
            if (!((double)pio / (double)psa > 0.5)) {                       // pq.py:174
                const long long cat = p / prm.L;
                if (cat < 0 || cat >= NC || cat >= 256) set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                else atomicAdd(&s_fp[(int)cat], 1);
            }
        }
    }
    __syncthreads();

    // per category: IoU sum in ascending pair order (the reference's float64 add order)
    for (int c = tid; c < NC; c += kMatchThreads) {
        double acc = 0.0;
        if (c < 256 && s_tp[c] > 0)
            for (int t = 0; t < m; ++t)
                if (s_flag[t] && s_gcat[t] == c) acc += s_iou[t];
        double *fs = prm.frame_stats + (size_t)b * 4 * NC;
        fs[c] = acc;
        fs[NC + c] = c < 256 ? (double)s_tp[c] : 0.0;
        fs[2 * NC + c] = c < 256 ? (double)s_fn[c] : 0.0;
        fs[3 * NC + c] = c < 256 ? (double)s_fp[c] : 0.0;
    }
    if (tid == 0 && prm.n_matches) prm.n_matches[b] = s_nm < prm.match_cap ? s_nm : prm.match_cap;
}

// state += frame result, frames in order (PanopticQuality.update, pq.py:298-303)
__global__ void accumulate_frames_kernel(const double *__restrict__ frame_stats, int B, int NC,
                                         double *iou, double *tp, double *fn, double *fp)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= NC) return;
    double a0 = iou[c], a1 = tp[c], a2 = fn[c], a3 = fp[c];
    for (int b = 0; b < B; ++b) {
        const double *fs = frame_stats + (size_t)b * 4 * NC;
        a0 += fs[c];
        a1 += fs[NC + c];
        a2 += fs[2 * NC + c];
        a3 += fs[3 * NC + c];
    }
    iou[c] = a0; tp[c] = a1; fn[c] = a2; fp[c] = a3;
}

// ---- stand-alone confusion matrix over arbitrary integer dtypes ---------------------------
__device__ __forceinline__ long long load_int(const void *p, int dtype, size_t i)
{
    switch (dtype) {
        case NPB_U8: case NPB_BOOL: return ((const uint8_t *)p)[i];
        case NPB_I16: return ((const int16_t *)p)[i];
        case NPB_I32: return ((const int32_t *)p)[i];
        default: return ((const long long *)p)[i];
    }
}

__global__ void __launch_bounds__(256)
confmat_kernel(const void *__restrict__ preds, int pd, const void *__restrict__ target, int td,
               long long N, int n, unsigned long long *__restrict__ confmat,
               int32_t *__restrict__ status)
{
    extern __shared__ unsigned s_cm[];
    const bool cm_smem = n <= kSmemConfmatMaxN;
    const int lane = threadIdx.x & 31;
    if (cm_smem) {
        for (int i = threadIdx.x; i < n * n; i += 256) s_cm[i] = 0;
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * 256;
    const long long n_round = ((N + 31) / 32) * 32;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_round; i += stride) {
        int key = -1;
        if (i < N) {
            const long long p = load_int(preds, pd, (size_t)i);
            const long long t = load_int(target, td, (size_t)i);
            if (p < 0 || p >= n || t < 0 || t >= n) set_status(status, NPB_ERR_CATEGORY_RANGE);
            else key = (int)(t * n + p);
        }
        unsigned pending = __ballot_sync(kFullMask, key >= 0);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int cur = __shfl_sync(kFullMask, key, leader);
            const unsigned same = __ballot_sync(kFullMask, key == cur);
            if (lane == leader) {
                if (cm_smem) atomicAdd(s_cm + cur, (unsigned)__popc(same));
                else atomicAdd(confmat + cur, (unsigned long long)__popc(same));
            }
            pending &= ~same;
        }
    }
    if (cm_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n * n; i += 256)
            if (s_cm[i]) atomicAdd(confmat + i, (unsigned long long)s_cm[i]);
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t match_smem_bytes()
{
    return (size_t)kMaxPairs * (8 + 8 + 8 + 8 + 4 + 4 + 1) + 16;
}

}  // namespace npb

using namespace npb;

extern "C" int npb_confmat_update(const void *preds, int preds_dtype, const void *target,
                                  int target_dtype, int64_t N, int n_classes, int64_t *confmat,
                                  int32_t *status, void *stream)
{
    if (!preds || !target || !confmat || !status) return NPB_ERR_ARG;
    if (N < 0 || n_classes < 1 || n_classes > 46340) return NPB_ERR_ARG;
    if (preds_dtype < NPB_U8 || preds_dtype > NPB_BOOL || target_dtype < NPB_U8 ||
        target_dtype > NPB_BOOL)
        return NPB_ERR_ARG;
    if (N == 0) return NPB_OK;
    long long blocks = (N + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 16) blocks = 148 * 16;
    const size_t smem = n_classes <= kSmemConfmatMaxN ? (size_t)n_classes * n_classes * 4 : 0;
    confmat_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(
        preds, preds_dtype, target, target_dtype, (long long)N, n_classes,
        (unsigned long long *)confmat, status);
    return record_launch("npb_confmat_update");
}

// workspace: [frame_keys | frame_cnts | frame_stats]
extern "C" size_t npb_pq_update_workspace_bytes(int B, int num_categories)
{
    size_t bytes = align256((size_t)B * kFrameSlots * sizeof(unsigned long long));
    bytes += align256((size_t)B * kFrameSlots * sizeof(unsigned));
    bytes += align256((size_t)B * 4 * num_categories * sizeof(double));
    return bytes;
}

extern "C" int npb_pq_update(const int64_t *pred, const int64_t *target, const uint8_t *sem_target,
                             int B, int64_t P, int num_categories, int64_t ignored_label,
                             int64_t max_instances_per_category, int64_t offset,
                             int64_t void_segment_id, void *workspace, double *iou, double *tp,
                             double *fn, double *fp, int64_t *confmat, int confmat_n,
                             double *frame_stats, int64_t *matches, int match_cap,
                             int32_t *n_matches, int32_t *status, void *stream)
{
    if (!pred || !target || !workspace || !iou || !tp || !fn || !fp || !status) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || P < 1 || num_categories < 1 || num_categories > 256 ||
        max_instances_per_category < 1 || offset < 1)
        return NPB_ERR_ARG;
    if ((sem_target != nullptr) != (confmat != nullptr)) return NPB_ERR_ARG;
    if (confmat && (confmat_n < 1 || confmat_n > 256)) return NPB_ERR_ARG;
    if (matches && (match_cap < 1 || !n_matches)) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;

    char *ws = (char *)workspace;
    unsigned long long *fkeys = (unsigned long long *)ws;
    ws += align256((size_t)B * kFrameSlots * sizeof(unsigned long long));
    unsigned *fcnts = (unsigned *)ws;
    ws += align256((size_t)B * kFrameSlots * sizeof(unsigned));
    double *fstats = frame_stats ? frame_stats : (double *)ws;

    cudaMemsetAsync(fkeys, 0xff, (size_t)B * kFrameSlots * sizeof(unsigned long long), s);
    cudaMemsetAsync(fcnts, 0, (size_t)B * kFrameSlots * sizeof(unsigned), s);

    PairParams pp;
    pp.pred = (const long long *)pred; pp.target = (const long long *)target;
    pp.sem_target = sem_target; pp.P = P; pp.offset = offset; pp.L = max_instances_per_category;
    pp.L_shift = -1;
    for (int sh = 0; sh < 62; ++sh)
        if ((1ll << sh) == max_instances_per_category) pp.L_shift = sh;
    pp.n = confmat ? confmat_n : 0;
    pp.frame_keys = fkeys; pp.frame_cnts = fcnts;
    pp.confmat = (unsigned long long *)confmat; pp.status = status;

    const bool vec4 = (P % 4 == 0) && (((uintptr_t)pred | (uintptr_t)target) & 15u) == 0 &&
                      ((uintptr_t)sem_target & 3u) == 0;
    const int vec = vec4 ? 4 : 1;
    const long long n_chunks = (P + (long long)kPairThreads * vec - 1) / ((long long)kPairThreads * vec);
    long long bx = (148ll * 8 + B - 1) / B;   // ~8 CTAs per SM over the whole batch
    if (bx > n_chunks) bx = n_chunks;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, B);
    const size_t cm_smem = (confmat && confmat_n <= kSmemConfmatMaxN)
                               ? (size_t)confmat_n * confmat_n * sizeof(unsigned) : 0;
    if (confmat) {
        if (vec4) pair_count_kernel<4, true><<<grid, kPairThreads, cm_smem, s>>>(pp);
        else pair_count_kernel<1, true><<<grid, kPairThreads, cm_smem, s>>>(pp);
    } else {
        if (vec4) pair_count_kernel<4, false><<<grid, kPairThreads, 0, s>>>(pp);
        else pair_count_kernel<1, false><<<grid, kPairThreads, 0, s>>>(pp);
    }

    MatchParams mp;
    mp.frame_keys = fkeys; mp.frame_cnts = fcnts; mp.num_categories = num_categories;
    mp.ignored_label = ignored_label; mp.L = max_instances_per_category; mp.offset = offset;
    mp.void_segment_id = void_segment_id; mp.frame_stats = fstats;
    mp.matches = (long long *)matches; mp.match_cap = match_cap; mp.n_matches = n_matches;
    mp.status = status;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(match_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)match_smem_bytes());
        attr_set = true;
    }
    match_frames_kernel<<<B, kMatchThreads, match_smem_bytes(), s>>>(mp);
    accumulate_frames_kernel<<<(num_categories + 127) / 128, 128, 0, s>>>(fstats, B, num_categories,
                                                                         iou, tp, fn, fp);
    return record_launch("npb_pq_update");
}
