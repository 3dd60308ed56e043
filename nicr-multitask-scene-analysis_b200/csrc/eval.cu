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
// counts pairs in per-CTA shared-memory tables after warp-level aggregation (neighbouring pixels
// nearly always share the pair) and hands them to the matcher through a dense per-frame
// class-pair table (RED) and a per-frame entry list (plain stores).  match_frames_kernel (one
// CTA per frame) merges them into a shared-memory pair table, builds the per-segment tables
// (areas, void / ignored overlap) in shared-memory hash tables straight from the pairs (O(m)),
// does the matching, and orders only the MATCHED pairs by `target*offset + pred` -- the
// reference's visiting order, which fixes the float64 summation order;
// accumulate_frames_kernel adds the frames to the running state in frame order.  Both float64
// orders equal the reference's, so the states are bit-identical, not merely close.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "finalize.cuh"

namespace npb {

constexpr int kPairThreads = 256;
#ifndef NPB_PAIR_MIN_CTAS
#define NPB_PAIR_MIN_CTAS 3
#endif
#ifndef NPB_PAIR_SLOTS
#define NPB_PAIR_SLOTS 2048
#endif
constexpr int kSmemSlots = NPB_PAIR_SLOTS;       // per-CTA pair hash table (instance pairs; class pairs go dense)
constexpr int kMaxPairs = 4096;        // distinct pairs per frame handled by the matcher
constexpr int kMatchThreads = 1024;
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kSmemConfmatMaxN = 96;   // n*n*4 B <= 36 KB privatised in shared memory

// 64-bit pair key -> 32-bit hash.  Pair keys are highly structured (class << 16 | instance in
// both halves, most bits zero), so both words go through multiplicative hashing followed by
// an avalanche step; table indices are taken after the final mix.
__device__ __forceinline__ unsigned hash64(unsigned long long k)
{
    unsigned h = (unsigned)k * 0x9E3779B1u + (unsigned)(k >> 32) * 0x85EBCA6Bu;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    return h;
}

// insert (key, cnt) into an open-addressing table; returns false when no slot was found
__device__ __forceinline__ bool table_add(unsigned long long *keys, unsigned *cnts, int slots,
                                          int max_probe, unsigned long long key, unsigned cnt)
{
    unsigned h = hash64(key) & (unsigned)(slots - 1);
    NPB_ASSERT((slots & (slots - 1)) == 0 && cnt > 0u);
    for (int probe = 0; probe < max_probe; ++probe) {
        NPB_ASSERT(h < (unsigned)slots);
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

// Hand-over from the pixel pass to the matcher.
//   * pairs of two instance-free segments (class, class): a dense per-frame table [nd][nd] that
//     every CTA adds to with fire-and-forget reductions (RED, nothing to wait for);
//   * pairs that involve an instance: a per-frame LIST of (pair key, pixels) entries.  Every
//     CTA appends its hash table with plain stores after reserving a range with one atomic;
//     the same pair may appear once per CTA (and once more per overflow of a CTA table), the
//     matcher merges the duplicates in shared memory.
// No global hash table and no compare-and-swap round trips; the entry counters and the dense
// tables are cleared by one memset per launch.
struct FrameEntries {
    unsigned long long *keys;   // [cap]
    unsigned *cnts;             // [cap]
    unsigned *n;                // entries appended so far (may exceed cap: then the frame failed)
    unsigned cap;
    int32_t *status;
};

// single entry straight to the list (a CTA table overflowed); out of line: it is the rare path
// and the pixel loop should stay small in the instruction cache
__device__ __noinline__ void emit_entry(const FrameEntries fe, unsigned long long key, unsigned cnt)
{
    const unsigned idx = atomicAdd(fe.n, 1u);
    if (idx < fe.cap) {
        fe.keys[idx] = key;
        fe.cnts[idx] = cnt;
    } else {
        set_status(fe.status, NPB_ERR_CAPACITY);
    }
}

struct PairParams {
    const long long *pred;      // null in the fused variant
    // fused variant (panoptic ids produced on the fly, see write_panoptic_kernel in merge.cu)
    const uint8_t *sem_map;     // (B,P) network class per pixel
    const uint8_t *inst_map;    // (B,P) raw instance id per pixel
    const long long *inst_pan_id;   // [B][kMaxInst] panoptic id of every instance
    int fold_finalize;          // the instance tables are derived here from the vote histograms
    FinalizeParams fin;         //   (finalize.cuh), CTA 0 of every frame stores them
    ClassSet thing;
    long long *pan_out;         // (B,P)
    uint8_t *pan_sem_out;       // (B,P) nullable
    const long long *target;
    const uint8_t *sem_target;  // nullable
    long long P;
    int B;                      // frames
    long long offset, L;
    int L_shift;                // >= 0 when L is a power of two
    int O_shift;                // >= 0 when offset is a power of two
    int n;                      // confusion-matrix size (0 = no confmat)
    int nd;                     // side of the dense class-pair table (0 = disabled)
    unsigned long long *entry_keys;  // [B][entry_cap]   (see FrameEntries)
    unsigned *entry_cnts;            // [B][entry_cap]
    unsigned *entry_n;               // [B], zeroed before the launch
    unsigned entry_cap;
    unsigned *frame_dense;           // [B][nd][nd] class-pair pixels of the frame, zeroed before the launch
    unsigned long long *confmat;     // [n][n] int64, accumulated
    int32_t *status;                 // [B]
    NPB_TL_FIELD
};

// Shared-memory state of one CTA of the pixel pass.
//   keys/cnts : open-addressing table (pair key -> pixels) for pairs that involve an instance
//   dense     : [nd][nd] counters for (class, class) pairs of two instance-free segments
//               (stuff / void / misclassified single pixels): no hashing, no probing
//   cm        : privatised confusion matrix
//   q_*       : one work queue per warp (see pair_count_kernel)
__device__ __forceinline__ FrameEntries frame_entries(const PairParams &prm, int b)
{
    return FrameEntries{prm.entry_keys + (size_t)b * prm.entry_cap,
                        prm.entry_cnts + (size_t)b * prm.entry_cap, prm.entry_n + b, prm.entry_cap,
                        prm.status + b};
}

struct PairTables {
    unsigned long long *keys;
    unsigned *cnts;
    unsigned *dense;
    unsigned *cm;
};

constexpr int kQueueCap = 192;   // < 32 left over + at most 128 pixels + 32 group entries

// one queue entry = `cnt` pixels of pair `key` whose semantic target is `st`.
// STD = the standard id geometry of the reference's task helper (offset = 256^3, 65536
// instances per category, task_helper/panoptic.py:42,61): shifts and masks become immediates.
template <bool CONFMAT, bool STD>
__device__ __forceinline__ void pair_consume(const PairTables &t, const PairParams &prm, int b,
                                             unsigned long long key, unsigned cnt, int st,
                                             bool cm_smem)
{
    const int OS = STD ? 24 : prm.O_shift, LS = STD ? 16 : prm.L_shift;
    const long long offset = STD ? (1ll << 24) : prm.offset, L = STD ? (1ll << 16) : prm.L;
    long long tv, pv;
    if (STD || OS >= 0) {
        tv = (long long)(key >> OS);
        pv = (long long)(key & ((unsigned long long)offset - 1ull));
    } else {
        tv = (long long)(key / (unsigned long long)offset);
        pv = (long long)(key - (unsigned long long)tv * (unsigned long long)offset);
    }
    long long pc = -1;
    if (STD || LS >= 0) pc = pv >> LS;
    else if (CONFMAT) pc = pv / L;
    bool dense = false;
    if (prm.nd > 0 && ((pv | tv) & (L - 1)) == 0) {          // two instance-free segments
        const long long tc = tv >> LS;
        if (pc < prm.nd && tc < prm.nd) {
            atomicAdd(t.dense + (int)tc * prm.nd + (int)pc, cnt);
            dense = true;
        }
    }
    if (!dense && !table_add(t.keys, t.cnts, kSmemSlots, 8, key, cnt))
        emit_entry(frame_entries(prm, b), key, cnt);
    if (CONFMAT) {
        if (pc < 0 || pc >= prm.n || st >= prm.n) {
            set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
        } else {
            const int cell = st * prm.n + (int)pc;
            if (cm_smem) atomicAdd(t.cm + cell, cnt);
            else atomicAdd(prm.confmat + cell, (unsigned long long)cnt);
        }
    }
}

// The same for the standard geometry when both ids passed the range check of the pixel loop
// (0 <= pred, target < 2^24).  One queue entry is ONE 64-bit word:
//     lo = pred | (target & 0xff) << 24,   hi = target >> 8 | st << 16 | pixels << 24
// so `hi & 0xffff : lo` is the pair key `target << 24 | pred` and every field is a 32-bit
// shift / mask away (no 64-bit arithmetic in the hot loop).
template <bool CONFMAT>
__device__ __forceinline__ void pair_consume_std(const PairTables &t, const PairParams &prm, int b,
                                                 unsigned lo, unsigned hi, bool cm_smem)
{
    const unsigned cnt = hi >> 24;
    const unsigned st = (hi >> 16) & 255u;
    const unsigned pc = (lo >> 16) & 255u;            // predicted category  (pred >> 16)
    const unsigned tc = (hi >> 8) & 255u;             // target category     (target >> 16)
    // instance-free on both sides: pred & 0xffff == 0 and target & 0xffff == 0
    const bool inst_free = ((lo & 0xff00ffffu) | (hi & 255u)) == 0u;
    bool dense = false;
    NPB_ASSERT(cnt >= 1u && cnt <= 128u);      // a queue entry stands for 1..128 pixels of a warp
    if (inst_free && pc < (unsigned)prm.nd && tc < (unsigned)prm.nd) {
        NPB_ASSERT(tc * (unsigned)prm.nd + pc < (unsigned)(prm.nd * prm.nd));
        atomicAdd(t.dense + tc * (unsigned)prm.nd + pc, cnt);
        dense = true;
    }
    if (!dense) {
        const unsigned long long key = ((unsigned long long)(hi & 0xffffu) << 32) | lo;
        if (!table_add(t.keys, t.cnts, kSmemSlots, 8, key, cnt))
            emit_entry(frame_entries(prm, b), key, cnt);
    }
    if (CONFMAT) {
        if (pc >= (unsigned)prm.n || st >= (unsigned)prm.n) {
            set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
        } else {
            const unsigned cell = st * (unsigned)prm.n + pc;
            NPB_ASSERT(cell < (unsigned)(prm.n * prm.n));
            if (cm_smem) atomicAdd(t.cm + cell, cnt);
            else atomicAdd(prm.confmat + cell, (unsigned long long)cnt);
        }
    }
}

// queue accesses of that loop by 32-bit shared-memory address (no generic -> shared conversion
// per access)
__device__ __forceinline__ void sts_entry(unsigned addr, unsigned lo, unsigned hi)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ uint2 lds_entry(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

// Pixel pass.  A thread reads 4 consecutive pixels (128-bit loads).  The pixels of a lane that
// agree with its lead pixel (all 4 inside a segment, 3 next to an isolated pixel, 2 at a
// boundary) join the group of lanes with the same pair (MATCH.ANY); the lowest lane of a group
// emits ONE entry for all pixels of the group, the remaining pixels are emitted singly.  Entries
// go to a per-warp shared-memory queue and are consumed 32 at a time, one entry per lane, so
// the table updates (the expensive, divergent part) always run with a full warp.
template <int VEC, bool CONFMAT, bool STD, bool FUSED = false>
__global__ void __launch_bounds__(kPairThreads, NPB_PAIR_MIN_CTAS) pair_count_kernel(const PairParams prm)
{
    static_assert(!FUSED || (STD && VEC == 4), "the fused variant exists for the standard geometry only");
    // fused variant: panoptic id of an instance / of a stuff class (0 for thing classes: a thing
    // pixel without instance stays void, panoptic_merge.py:213-224)
    __shared__ unsigned s_pan32[FUSED ? kMaxInst : 1];
    __shared__ unsigned s_stuff[FUSED ? 256 : 1];
    __shared__ int s_fin_cls[FUSED ? kMaxInst : 1];
    static_assert(!FUSED || kPairThreads == kMaxInst, "one table entry per thread");
    NPB_TL(prm, 2, start);
    grid_launch_dependents();       // the matcher: its CTAs need a whole SM each, they only move in
                                    // where this grid has left
    extern __shared__ unsigned long long s_dyn[];
    PairTables t;
    t.keys = s_dyn;                                           // [kSmemSlots]
    unsigned long long *q_key_all = t.keys + kSmemSlots;      // [warps][kQueueCap]
    t.cnts = (unsigned *)(q_key_all + (kPairThreads / 32) * kQueueCap);   // [kSmemSlots]
    t.dense = t.cnts + kSmemSlots;                            // [nd*nd]
    t.cm = t.dense + prm.nd * prm.nd;                         // [n*n] when privatised

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = prm.n, nd = prm.nd;
    const bool cm_smem = CONFMAT && n <= kSmemConfmatMaxN;
    unsigned short *q_meta_all = (unsigned short *)(t.cm + (cm_smem ? n * n : 0));
    unsigned long long *q_key = q_key_all + warp * kQueueCap;
    unsigned short *q_meta = q_meta_all + warp * kQueueCap;
    const long long P = prm.P;
    const long long chunk = (long long)kPairThreads * VEC;
    const long long n_chunks = (P + chunk - 1) / chunk;
    // The grid is ONE wave of persistent CTAs over the chunks of the whole batch: CTA i takes
    // the i-th equal share of all B * n_chunks chunks, a contiguous range that may cross frame
    // boundaries -- every CTA has the same amount of work whatever the ratio of CTA slots to
    // frames is (2 CTAs per frame on 2.3 slots per frame left 14 % of the SMs' slots empty).
    // Within a frame a CTA's range is a band of rows: a segment then shows up in the tables of
    // the few CTAs whose band it crosses instead of in all of them -- fewer entries per CTA table
    // and far fewer duplicates for the matcher to merge.  The tables are flushed (and cleared)
    // at the end of every frame part.
    const long long g_total = n_chunks * prm.B;
    const long long g_begin = g_total * blockIdx.x / gridDim.x;
    const long long g_end = g_total * (blockIdx.x + 1) / gridDim.x;
    unsigned lt_mask = (1u << lane) - 1u;
#ifndef NPB_PAIR_NO_PIN
    // (an opaque definition: the compiler keeps the value in a register instead of rebuilding
    // it from %tid in every iteration of the pixel loop)
    asm volatile("" : "+r"(lt_mask));
#endif
    bool waited = false;
  for (long long g_at = g_begin; g_at < g_end;) {
    const int b = (int)(g_at / n_chunks);
    const long long ch_begin = g_at - (long long)b * n_chunks;
    const long long ch_end = (ch_begin + (g_end - g_at) < n_chunks) ? ch_begin + (g_end - g_at) : n_chunks;
    g_at += ch_end - ch_begin;
    for (int i = tid; i < kSmemSlots; i += kPairThreads) { t.keys[i] = kEmptyKey; t.cnts[i] = 0; }
    for (int i = tid; i < nd * nd; i += kPairThreads) t.dense[i] = 0;
    if (cm_smem)
        for (int i = tid; i < n * n; i += kPairThreads) t.cm[i] = 0;
    // everything above touched shared memory only: it overlaps the tail of the predecessor
    if (!waited) { grid_dependency_wait(); waited = true; }
    NPB_TL(prm, 2, wait);
    if (FUSED) {
        // panoptic id of every instance of the frame: from the table, or derived right here from
        // the vote histograms of the grouping kernel (every CTA of the frame repeats the few
        // hundred loads; CTA 0 stores the tables) -- no finalize launch in between
        const long long pan = prm.fold_finalize
                                  ? finalize_frame(prm.fin, b, tid, s_fin_cls, ch_begin == 0)
                                  : prm.inst_pan_id[(size_t)b * kMaxInst + tid];
        s_pan32[tid] = (unsigned)pan;
        s_stuff[tid] = prm.thing.has(tid) ? 0u : ((unsigned)tid + 1u) << 16;
    }
    __syncthreads();
    NPB_TL(prm, 10, wait);      // instance tables of the frame derived
    int q_len = 0;      // warp-uniform

    if constexpr (STD && VEC == 4) {
        // ---- standard geometry, 4 pixels per thread: everything in 32-bit words ----------------
        // entry of a pixel (see pair_consume_std): lo = pred | target << 24, hi = target >> 8 |
        // st << 16, one PRMT each; valid when 0 <= pred, target < 2^24, which one OR over the
        // lane's words checks.  Comparing / MATCHing the 64-bit entry decides "same pair and
        // same confusion cell" at once.
        const long long stride = chunk;
        const long long q_end = ch_end * chunk < P ? ch_end * chunk : P;
        long long q_next = ch_begin * chunk + (long long)tid * 4;   // next pixel to fetch
        const long long *pred_b = FUSED ? nullptr : prm.pred + (size_t)b * P;
        const long long *target_b = prm.target + (size_t)b * P;
        const uint8_t *sem_b = CONFMAT ? prm.sem_target + (size_t)b * P : nullptr;
        const uint8_t *csem_b = FUSED ? prm.sem_map + (size_t)b * P : nullptr;
        const uint8_t *cinst_b = FUSED ? prm.inst_map + (size_t)b * P : nullptr;
        unsigned q_base = (unsigned)__cvta_generic_to_shared(q_key);   // this warp's queue
#ifndef NPB_PAIR_NO_PIN
        asm volatile("" : "+r"(q_base));      // (kept, not rebuilt from the shared window base)
#endif
        // (zero once: a lane beyond the end of its range keeps its last words, which nothing
        // looks at -- every use below is gated by `act`)
        uint4 n_p0 = make_uint4(0u, 0u, 0u, 0u), n_p1 = n_p0, n_t0 = n_p0, n_t1 = n_p0;
        unsigned n_sw = 0u, n_cw = 0u, n_iw = 0u;
        // software pipeline: the loads of the next chunk are issued before the current chunk is
        // processed, so their DRAM latency hides behind the (instruction bound) table updates
        // running pointers of the lane (advanced by one chunk per fetch: two adds per stream
        // instead of rebuilding frame base + pixel offset from the parameters every time)
        const char *pp_at = FUSED ? nullptr : (const char *)(pred_b + q_next);
        const char *tp_at = (const char *)(target_b + q_next);
        const char *sp_at = CONFMAT ? (const char *)(sem_b + q_next) : nullptr;
        const char *cs_at = FUSED ? (const char *)(csem_b + q_next) : nullptr;
        const char *ci_at = FUSED ? (const char *)(cinst_b + q_next) : nullptr;
        auto fetch = [&]() {
            if (q_next < q_end) {
                if (FUSED) {
                    // class / instance maps were written by the grouping kernel just before: L2
                    n_cw = *(const unsigned *)cs_at;
                    n_iw = *(const unsigned *)ci_at;
                } else {
                    const uint4 *pp = (const uint4 *)pp_at;
                    n_p0 = __ldcs(pp); n_p1 = __ldcs(pp + 1);
                }
                const uint4 *tp = (const uint4 *)tp_at;
                n_t0 = __ldcs(tp); n_t1 = __ldcs(tp + 1);
                if (CONFMAT) n_sw = __ldcs((const unsigned *)sp_at);
            }
            q_next += stride;
            if (!FUSED) pp_at += stride * 8;
            tp_at += stride * 8;
            if (CONFMAT) sp_at += stride;
            if (FUSED) { cs_at += stride; ci_at += stride; }
        };
        fetch();
        for (long long ch = ch_begin; ch < ch_end; ++ch) {
            bool act = q_next - stride < P;
            const unsigned sw = n_sw;
            if (FUSED) {
                // the prediction of these 4 pixels: instance id -> its panoptic id, otherwise the
                // stuff id of the class; written out here and evaluated from registers
                unsigned v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned c = (n_cw >> (8 * j)) & 255u, ii = (n_iw >> (8 * j)) & 255u;
                    v[j] = ii ? s_pan32[ii] : s_stuff[c];
                }
                n_p0 = make_uint4(v[0], 0u, v[1], 0u);
                n_p1 = make_uint4(v[2], 0u, v[3], 0u);
                if (act) {
                    const size_t fq = (size_t)b * P + (size_t)(q_next - stride);
                    uint4 *o = (uint4 *)(prm.pan_out + fq);
                    __stcs(o, n_p0);
                    __stcs(o + 1, n_p1);
                    if (prm.pan_sem_out)        // pan // 65536, one byte per pixel
                        *(unsigned *)(prm.pan_sem_out + fq) =
                            __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
                }
            }
            const unsigned hi_or = n_p0.y | n_p0.w | n_p1.y | n_p1.w | n_t0.y | n_t0.w | n_t1.y | n_t1.w;
            const unsigned p_or = n_p0.x | n_p0.z | n_p1.x | n_p1.z;
            const unsigned t_or = n_t0.x | n_t0.z | n_t1.x | n_t1.z;
            if (act && (hi_or | ((p_or | t_or) >> 24)) != 0u) {
                // rare: the lane's ids fail the 24-bit range check -- an error (pred outside
                // [0, offset) or a negative target), or a target id >= 2^24 (a category >= 256:
                // only meaningful as the ignored label) whose pixels are counted one by one
                // with the full 64-bit key
                const unsigned p_bad = n_p0.y | n_p0.w | n_p1.y | n_p1.w | (p_or >> 24);
                const unsigned t_neg = (n_t0.y | n_t0.w | n_t1.y | n_t1.w) >> 31;
                if ((p_bad | t_neg) != 0u) {
                    set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                } else {
#pragma unroll 1
                    for (int j = 0; j < 4; ++j) {
                        const uint4 pw = j < 2 ? n_p0 : n_p1, tw = j < 2 ? n_t0 : n_t1;
                        const unsigned plo = (j & 1) ? pw.z : pw.x;
                        const unsigned long long tv =
                            ((unsigned long long)((j & 1) ? tw.w : tw.y) << 32) | ((j & 1) ? tw.z : tw.x);
                        pair_consume<CONFMAT, true>(t, prm, b, (tv << 24) | plo, 1u,
                                                    (int)((sw >> (8 * j)) & 255u), cm_smem);
                    }
                }
                act = false;
            }
            unsigned klo[4], khi[4];
            {
                const unsigned plo[4] = {n_p0.x, n_p0.z, n_p1.x, n_p1.z};
                const unsigned tlo[4] = {n_t0.x, n_t0.z, n_t1.x, n_t1.z};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    klo[j] = __byte_perm(plo[j], tlo[j], 0x4210);                    // p0 p1 p2 t0
                    khi[j] = __byte_perm(tlo[j], sw, 0x3021 | ((4 + j) << 8));       // t1 t2 st 0
                }
            }
            fetch();

            // Lead pixels of the lane: its first run of >= 2 equal NEIGHBOURS (the background
            // segment when an isolated pixel or a boundary cuts the lane), else pixel 0.  The run
            // joins the warp-wide group of that entry, the other pixels ("minors") are queued as
            // single pixels -- also a pixel that equals the lead but is not adjacent to its run
            // (A x A A): one more queue entry in a rare pattern buys three compares instead of
            // six and a table look-up instead of the select chains.  Branch free: nearly every
            // warp has lanes of each kind.
            const unsigned d01 = (klo[0] ^ klo[1]) | (khi[0] ^ khi[1]);
            const unsigned d12 = (klo[1] ^ klo[2]) | (khi[1] ^ khi[2]);
            const unsigned d23 = (klo[2] ^ klo[3]) | (khi[2] ^ khi[3]);
            // member mask by (e01, e12, e23): 000 -> 0001, 001 -> 0011, 010 -> 0110, 011 -> 0111,
            // 100 -> 1100, 101 -> 0011, 110 -> 1110, 111 -> 1111 (bit j: pixel j is in the run)
            const unsigned sh = (d01 ? 0u : 4u) + (d12 ? 0u : 8u) + (d23 ? 0u : 16u);
            const unsigned member = (0xFE3C7631u >> sh) & 15u;
            const unsigned lk_lo = (member & 1u) ? klo[0] : (member & 2u) ? klo[1] : klo[2];
            const unsigned lk_hi = (member & 1u) ? khi[0] : (member & 2u) ? khi[1] : khi[2];
            const unsigned actm = __ballot_sync(kFullMask, act);
            const unsigned peers =
                __match_any_sync(kFullMask, ((unsigned long long)lk_hi << 32) | lk_lo) & actm;
            const unsigned n_minor = 4u - (unsigned)__popc(member);     // 0..3
            // pixels of the group = 4 * lanes - minors of its lanes, the minors counted from two
            // ballots of the bits of n_minor (no REDUX: a reduction with per-group masks would
            // serialise over the groups)
            const unsigned n0 = __ballot_sync(kFullMask, (n_minor & 1u) != 0u) & actm;
            const unsigned n1 = __ballot_sync(kFullMask, (n_minor & 2u) != 0u) & actm;
            const unsigned total = 4 * __popc(peers) - __popc(peers & n0) - 2 * __popc(peers & n1);
            const bool leader = act && (peers & lt_mask) == 0u;
            const unsigned leader_mask = __ballot_sync(kFullMask, leader);
            const int total_minor = __popc(n0) + 2 * __popc(n1);

            // push: minors as single pixels, then one entry per group leader
            const unsigned q_tail = q_base + 8u * (unsigned)q_len;
            if (act && n_minor > 0) {
                // one address register per store: a store still in flight keeps reading its own
                const unsigned mm = ~member & 15u;              // bit j: pixel j is a minor
                const unsigned a0 = q_tail + 8u * (unsigned)(__popc(n0 & lt_mask) + 2 * __popc(n1 & lt_mask));
                const unsigned a1 = a0 + 8u * (mm & 1u);
                const unsigned a2 = a0 + 8u * (unsigned)__popc(mm & 3u);
                const unsigned a3 = a0 + 8u * (unsigned)__popc(mm & 7u);
                if (mm & 1u) sts_entry(a0, klo[0], khi[0] | (1u << 24));
                if (mm & 2u) sts_entry(a1, klo[1], khi[1] | (1u << 24));
                if (mm & 4u) sts_entry(a2, klo[2], khi[2] | (1u << 24));
                if (mm & 8u) sts_entry(a3, klo[3], khi[3] | (1u << 24));
            }
            if (leader)
                sts_entry(q_tail + 8u * (unsigned)(total_minor + __popc(leader_mask & lt_mask)), lk_lo,
                          lk_hi | (total << 24));
            q_len += total_minor + __popc(leader_mask);
            NPB_ASSERT(q_len <= kQueueCap);
            __syncwarp();
            while (q_len >= 32) {           // consume full warps of entries from the tail
                q_len -= 32;
                const uint2 e = lds_entry(q_base + 8u * (unsigned)(q_len + lane));
                pair_consume_std<CONFMAT>(t, prm, b, e.x, e.y, cm_smem);
                __syncwarp();
            }
        }
        if (lane < q_len) {
            const uint2 e = lds_entry(q_base + 8u * (unsigned)lane);
            pair_consume_std<CONFMAT>(t, prm, b, e.x, e.y, cm_smem);
        }
    } else {
        // software pipeline: the loads of the next chunk are issued before the current chunk is
        // processed, so their DRAM latency hides behind the (instruction bound) table updates
        long long n_pv[VEC], n_tv[VEC];
        unsigned n_sw = 0;
        auto fetch = [&](long long ch) {
            const long long q0 = ch * chunk + (long long)tid * VEC;
    #pragma unroll
            for (int j = 0; j < VEC; ++j) { n_pv[j] = 0; n_tv[j] = 0; }
            n_sw = 0;
            if (ch < ch_end && q0 < P) {
                const size_t fq = (size_t)b * P + q0;
                if (VEC == 4) {
                    const longlong2 a0 = __ldcs((const longlong2 *)(prm.pred + fq));
                    const longlong2 a1 = __ldcs((const longlong2 *)(prm.pred + fq) + 1);
                    const longlong2 t0 = __ldcs((const longlong2 *)(prm.target + fq));
                    const longlong2 t1 = __ldcs((const longlong2 *)(prm.target + fq) + 1);
                    n_pv[0] = a0.x; n_pv[1 % VEC] = a0.y; n_pv[2 % VEC] = a1.x; n_pv[3 % VEC] = a1.y;
                    n_tv[0] = t0.x; n_tv[1 % VEC] = t0.y; n_tv[2 % VEC] = t1.x; n_tv[3 % VEC] = t1.y;
                    if (CONFMAT) n_sw = *(const unsigned *)(prm.sem_target + fq);
                } else {
                    n_pv[0] = __ldcs(prm.pred + fq);
                    n_tv[0] = __ldcs(prm.target + fq);
                    if (CONFMAT) n_sw = prm.sem_target[fq];
                }
            }
        };
        fetch(ch_begin);

        for (long long ch = ch_begin; ch < ch_end; ++ch) {
            const long long p0 = ch * chunk + (long long)tid * VEC;
            const bool act = p0 < P;    // P % VEC == 0 guaranteed by the launcher
            unsigned long long key[VEC];
            unsigned sw = n_sw;
            long long pv[VEC], tv[VEC];
    #pragma unroll
            for (int j = 0; j < VEC; ++j) { key[j] = 0; pv[j] = n_pv[j]; tv[j] = n_tv[j]; }
            fetch(ch + 1);
            if (act) {
                // ids must satisfy 0 <= pred < offset, 0 <= target (checked on the OR of the lane)
                long long any_neg = 0;
                bool too_big = false;
    #pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    any_neg |= pv[j] | tv[j];
                    if (STD) {
                        too_big |= (pv[j] >> 24) != 0;
                        key[j] = ((unsigned long long)tv[j] << 24) | (unsigned long long)pv[j];
                    } else {
                        too_big |= pv[j] >= prm.offset;
                        key[j] = prm.O_shift >= 0
                                     ? (((unsigned long long)tv[j] << prm.O_shift) | (unsigned long long)pv[j])
                                     : (unsigned long long)tv[j] * (unsigned long long)prm.offset +
                                           (unsigned long long)pv[j];
                    }
                }
                if (any_neg < 0 || too_big) set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
            }

            // ---- grouping.  kk = pair key with the semantic target folded into the top byte, so one
            // 64-bit compare / MATCH decides "same pair AND same confusion cell" (ids that reach
            // into the top byte take the degenerate every-pixel-alone path below).
            unsigned long long kk[VEC];
            bool big = false;
    #pragma unroll
            for (int j = 0; j < VEC; ++j) {
                big |= (key[j] >> 56) != 0;
                kk[j] = key[j] ^ ((unsigned long long)((sw >> (8 * j)) & 255u) << 56);
            }
            // lead pixel of the lane: the first one that agrees with another pixel of the lane (the
            // background segment when an isolated pixel or a boundary cuts the lane); its `cnt0`
            // agreeing pixels join the warp-wide group of that key, the others ("minors") are
            // queued as single pixels
            int lj = 0;
            unsigned member = 1u;                     // bit j: pixel j agrees with the lead
            if (VEC == 4 && !big) {
                const bool e01 = kk[0] == kk[1 % VEC], e02 = kk[0] == kk[2 % VEC], e03 = kk[0] == kk[3 % VEC];
                const bool e12 = kk[1 % VEC] == kk[2 % VEC], e13 = kk[1 % VEC] == kk[3 % VEC];
                const bool e23 = kk[2 % VEC] == kk[3 % VEC];
                if (e01 | e02 | e03) { lj = 0; member = 1u | (e01 ? 2u : 0u) | (e02 ? 4u : 0u) | (e03 ? 8u : 0u); }
                else if (e12 | e13) { lj = 1; member = 2u | (e12 ? 4u : 0u) | (e13 ? 8u : 0u); }
                else if (e23) { lj = 2; member = 4u | 8u; }
            }
            const int cnt0 = act ? __popc(member) : 0;
            const unsigned long long lk = lj == 0 ? kk[0] : lj == 1 ? kk[1 % VEC] : kk[2 % VEC];
            const unsigned long long mk = (act && !big) ? lk : (kEmptyKey - 1ull - (unsigned)lane);
            const unsigned peers = __match_any_sync(kFullMask, mk);
            // pixels of the group = sum of cnt0 over its lanes = 4 * lanes - deficits (no REDUX:
            // a reduction with per-group masks would serialise over the groups)
            const unsigned d1 = __ballot_sync(kFullMask, cnt0 == VEC - 1);
            const unsigned d2 = __ballot_sync(kFullMask, cnt0 == VEC - 2);
            const unsigned d3 = __ballot_sync(kFullMask, cnt0 == VEC - 3);
            const int total = VEC * __popc(peers) - __popc(peers & d1) - 2 * __popc(peers & d2) -
                              3 * __popc(peers & d3);
            const bool leader = act && lane == __ffs(peers) - 1;
            const unsigned leader_mask = __ballot_sync(kFullMask, leader);
            const int n_minor = act ? VEC - cnt0 : 0;
            const unsigned b1 = __ballot_sync(kFullMask, n_minor >= 1);
            const unsigned b2 = __ballot_sync(kFullMask, n_minor >= 2);
            const unsigned b3 = __ballot_sync(kFullMask, n_minor >= 3);
            const int total_minor = __popc(b1) + __popc(b2) + __popc(b3);

            // push: minors as single pixels, then one entry per group leader
            if (n_minor > 0) {
                int pos = q_len + __popc(b1 & lt_mask) + __popc(b2 & lt_mask) + __popc(b3 & lt_mask);
    #pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (!((member >> j) & 1u)) {
                        q_key[pos] = key[j];
                        q_meta[pos] = (unsigned short)((1u << 8) | ((sw >> (8 * j)) & 255u));
                        ++pos;
                    }
                }
            }
            if (leader) {
                const int pos = q_len + total_minor + __popc(leader_mask & lt_mask);
                q_key[pos] = lj == 0 ? key[0] : lj == 1 ? key[1 % VEC] : key[2 % VEC];
                q_meta[pos] = (unsigned short)(((unsigned)total << 8) | ((sw >> (8 * lj)) & 255u));
            }
            q_len += total_minor + __popc(leader_mask);
            __syncwarp();
            while (q_len >= 32) {           // consume full warps of entries from the tail
                q_len -= 32;
                const unsigned long long k = q_key[q_len + lane];
                const unsigned m = q_meta[q_len + lane];
                pair_consume<CONFMAT, STD>(t, prm, b, k, m >> 8, (int)(m & 255u), cm_smem);
                __syncwarp();
            }
        }
        if (lane < q_len) {
            const unsigned long long k = q_key[lane];
            const unsigned m = q_meta[lane];
            pair_consume<CONFMAT, STD>(t, prm, b, k, m >> 8, (int)(m & 255u), cm_smem);
        }

    }

    NPB_TL(prm, 11, wait);      // pixel loop done
    // ---- flush: append the CTA tables to the frame's entry list, add the confusion matrix ----
    // count, reserve a range of the list with ONE global atomic, then store
    __shared__ unsigned s_total, s_base, s_cursor;
    if (tid == 0) { s_total = 0; s_cursor = 0; }
    __syncthreads();
    // class pairs first: reductions without a return value, their latency is never waited for
    if (nd > 0) {
        unsigned *fd = prm.frame_dense + (size_t)b * nd * nd;
        for (int i = tid; i < nd * nd; i += kPairThreads) {
            const unsigned c = t.dense[i];
            if (c) atomicAdd(fd + i, c);
        }
    }
    unsigned mine = 0;
    for (int i = tid; i < kSmemSlots; i += kPairThreads) mine += (t.keys[i] != kEmptyKey && t.cnts[i]) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFullMask, mine, o);
    if (lane == 0 && mine) atomicAdd(&s_total, mine);
    __syncthreads();
    const FrameEntries fe = frame_entries(prm, b);
    if (tid == 0) {
        s_base = s_total ? atomicAdd(fe.n, s_total) : 0u;
        if (s_total && s_base + s_total > fe.cap) set_status(fe.status, NPB_ERR_CAPACITY);
    }
    __syncthreads();
    if (s_total) {
        const unsigned base = s_base;
        for (int i = tid; i < kSmemSlots; i += kPairThreads) {
            const unsigned long long k = t.keys[i];
            const unsigned c = t.cnts[i];
            if (k != kEmptyKey && c) {
                const unsigned idx = base + atomicAdd(&s_cursor, 1u);
                if (idx < fe.cap) { fe.keys[idx] = k; fe.cnts[idx] = c; }
            }
        }
    }
    if (cm_smem)
        for (int i = tid; i < n * n; i += kPairThreads)
            if (t.cm[i]) atomicAdd(prm.confmat + i, (unsigned long long)t.cm[i]);
    __syncthreads();        // the tables are cleared for the next frame part
  }
    if (!waited) grid_dependency_wait();     // (a CTA without work: the chain rule still holds)
    NPB_TL(prm, 2, end);
}

// ---------------------------------------------------------------------------------------
struct MatchParams {
    const unsigned long long *entry_keys;   // see FrameEntries
    const unsigned *entry_cnts;
    unsigned *entry_n;                      // read, then put back to zero
    unsigned entry_cap;
    unsigned *frame_dense;                  // [B][nd][nd] or null; every cell read is put back to zero
    int nd;
    int num_categories;
    long long ignored_label, L, offset, void_segment_id;
    int L_shift, O_shift;  // >= 0 when L / offset are powers of two (shifts instead of 64-bit divisions)
    double *frame_stats;   // [B + 1][4][num_categories]; row B: the states before this update
    unsigned *done_cnt;    // frames matched so far (zero at rest); null: no accumulation
    double *iou, *tp, *fn, *fp;     // running state [num_categories]
    long long *matches;    // [B][match_cap][2] nullable
    int match_cap;
    int32_t *n_matches;    // [B] nullable
    int32_t *status;       // [B]
    NPB_TL_FIELD
};

// ---- matcher: one CTA per frame ------------------------------------------------------------
// The frame's hand-over tables are merged into a shared-memory pair table (pair -> pixels);
// segment tables (distinct gt ids / pred ids of the frame) are shared-memory hash tables filled
// from the pairs with atomics; only the matched pairs (one per matched gt segment at most) are
// ordered, because only their float64 IoU sum depends on the visiting order.
// The phases after the merge are shared with the fall-back for frames that do not fit
// (match_big_frame_kernel: the same tables in global memory, larger).
constexpr int kPairSlots = 2 * kMaxPairs;   // pair table of the matcher (load factor <= 0.5)
constexpr int kSegSlots = 2048;       // distinct gt (and pred) segments per frame: <= 1536
constexpr int kMaxMatched = 1024;

__host__ __device__ constexpr size_t match_smem_bytes()
{
    return (size_t)kPairSlots * (8 + 4 + 2 + 2) + (size_t)kMaxMatched * (8 + 4 + 4 + 2) +
           (size_t)kMaxPairs * 2 + (size_t)kSegSlots * (8 + 8 + 4 * 4 + 2) + 16;
}

// SlotT: index type of segment slots / pair slots (uint16 in shared memory, uint32 in the
// global-memory fall-back)
template <typename SlotT>
struct MatchTables {
    // pair table: key -> pixels (+ the slots of its two segments, filled by phase 1)
    unsigned long long *t_key;  // [pair_slots], kEmptyKey = free
    unsigned *t_cnt;
    SlotT *t_gslot, *t_pslot;
    int pair_slots;
    const SlotT *walk;          // the used pair slots ([n_walk]), or null: walk all pair_slots
    // segment tables
    unsigned long long *g_id, *p_id;    // [seg_slots], kEmptyKey = free
    unsigned *g_area, *p_area;          // pixels of the segment
    unsigned *p_void;                   // pred: pixels inside the gt void segment
    unsigned *p_pio;                    // pred: pixels inside ignored gt segments
    unsigned char *g_matched, *p_matched;
    int seg_slots;
    // matched pairs (unordered), then ordered IoUs / categories
    long long *m_key;
    unsigned *m_ia, *m_uni;
    unsigned short *m_cat, *s_cat;
    double *s_iou;
    int max_matched;
};

__device__ __forceinline__ long long div_pow2(long long v, long long d, int shift)
{
    return shift >= 0 ? (v >> shift) : (v / d);     // v is validated non-negative
}

// returns the slot of `id` (inserting it), or -1 when the table is full
__device__ __forceinline__ int seg_slot(unsigned long long *ids, int slots, unsigned long long id)
{
    unsigned h = hash64(id) & (unsigned)(slots - 1);
    for (int probe = 0; probe < slots; ++probe) {
        unsigned long long k = ids[h];
        if (k == kEmptyKey) k = atomicCAS(ids + h, kEmptyKey, id);
        if (k == kEmptyKey || k == id) return (int)h;
        h = (h + 1) & (unsigned)(slots - 1);
    }
    return -1;
}

// counters of one frame, in shared memory of the matching CTA
struct MatchCounters {
    int nm;                 // matched pairs
    int fail;               // a capacity was exceeded: the frame contributes nothing
    int tp[256], fn[256], fp[256];
};

// Phases 1-4 of the matcher on a filled pair table.  All threads of the CTA call it; the tables
// (other than the pair keys / counts) and the counters must be cleared.
template <typename SlotT>
__device__ void match_phases(const MatchParams &prm, int b, const MatchTables<SlotT> &T,
                             int n_walk, MatchCounters &C)
{
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int NC = prm.num_categories;

    // (1) segment tables: areas (pq.py:83-84), void overlap (pq.py:34-43), ignored overlap
    //     (pq.py:47-57); every pair remembers the slots of its two segments
    for (int i = tid; i < n_walk; i += nthreads) {
        const int t = T.walk ? (int)T.walk[i] : i;
        const long long key = (long long)T.t_key[t];
        if ((unsigned long long)key == kEmptyKey) continue;
        const unsigned cnt = T.t_cnt[t];
        const long long g = div_pow2(key, prm.offset, prm.O_shift);
        const long long p = key - g * prm.offset;
        const int gs = seg_slot(T.g_id, T.seg_slots, (unsigned long long)g);
        const int ps = seg_slot(T.p_id, T.seg_slots, (unsigned long long)p);
        NPB_ASSERT(gs < T.seg_slots && ps < T.seg_slots && t < T.pair_slots);
        if (gs < 0 || ps < 0) {
            C.fail = 1;
            T.t_key[t] = kEmptyKey;         // phase 2 skips the pair (its frame failed anyway)
            continue;
        }
        T.t_gslot[t] = (SlotT)gs;
        T.t_pslot[t] = (SlotT)ps;
        atomicAdd(T.g_area + gs, cnt);
        atomicAdd(T.p_area + ps, cnt);
        if (g == prm.void_segment_id) T.p_void[ps] = cnt;     // key == void*offset + p, unique
        if (div_pow2(g, prm.L, prm.L_shift) == prm.ignored_label) atomicAdd(T.p_pio + ps, cnt);
    }
    __syncthreads();
    NPB_TL(prm, 7, wait);

    // (2) IoU + match decision per intersecting pair                       pq.py:119-152
    for (int i = tid; i < n_walk; i += nthreads) {
        const int t = T.walk ? (int)T.walk[i] : i;
        const long long key = (long long)T.t_key[t];
        if ((unsigned long long)key == kEmptyKey) continue;
        if (key == prm.void_segment_id) continue;                          // pq.py:120
        const long long g = div_pow2(key, prm.offset, prm.O_shift), p = key - g * prm.offset;
        const long long gcat = div_pow2(g, prm.L, prm.L_shift), pcat = div_pow2(p, prm.L, prm.L_shift);
        if (gcat != pcat) continue;                                        // pq.py:128
        const int gs = T.t_gslot[t], ps = T.t_pslot[t];
        const long long ia = T.t_cnt[t];
        const long long uni = (long long)T.g_area[gs] + (long long)T.p_area[ps] - ia -
                              (long long)T.p_void[ps];                     // pq.py:143
        if (uni == 0) { set_status(prm.status + b, NPB_ERR_ZERO_DIVISION); continue; }
        // iou = ia / uni > 0.5 in float64 (pq.py:145-146) <=> 2 * ia > uni (integers < 2^33)
        if (uni > 0 && 2 * ia > uni) {
            if (gcat < 0 || gcat >= NC || gcat >= 256) { set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE); continue; }
            T.g_matched[gs] = 1;
            T.p_matched[ps] = 1;
            atomicAdd(&C.tp[(int)gcat], 1);
            const int slot = atomicAdd(&C.nm, 1);
            NPB_ASSERT(slot >= 0);
            if (slot < T.max_matched) {
                T.m_key[slot] = key;
                T.m_ia[slot] = (unsigned)ia;
                T.m_uni[slot] = (unsigned)uni;
                T.m_cat[slot] = (unsigned short)gcat;
            } else {
                C.fail = 1;
            }
            if (prm.matches) {
                if (slot < prm.match_cap) {
                    long long *o = prm.matches + ((size_t)b * prm.match_cap + slot) * 2;
                    o[0] = g;
                    o[1] = p;
                } else {
                    C.fail = 1;
                }
            }
        }
    }
    __syncthreads();
    NPB_TL(prm, 8, wait);

    // (3) false negatives: unmatched gt segments outside the ignored label    pq.py:155-163
    //     false positives: unmatched pred segments, unless more than half of their area lies
    //     in ignored gt segments                                            pq.py:165-177
    for (int i = tid; i < T.seg_slots; i += nthreads) {
        if (T.g_id[i] != kEmptyKey && !T.g_matched[i]) {
            const long long cat = div_pow2((long long)T.g_id[i], prm.L, prm.L_shift);
            if (cat != prm.ignored_label) {
                if (cat < 0 || cat >= NC || cat >= 256) set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                else atomicAdd(&C.fn[(int)cat], 1);
            }
        }
        if (T.p_id[i] != kEmptyKey && !T.p_matched[i]) {
            // pio / area > 0.5 in float64 (pq.py:172) <=> 2 * pio > area: the quotient of two
            // integers < 2^32 is never closer to 0.5 than 2^-33 unless it equals 0.5
            if (!(2ull * T.p_pio[i] > (unsigned long long)T.p_area[i])) {
                const long long cat = div_pow2((long long)T.p_id[i], prm.L, prm.L_shift);
                if (cat < 0 || cat >= NC || cat >= 256) set_status(prm.status + b, NPB_ERR_CATEGORY_RANGE);
                else atomicAdd(&C.fp[(int)cat], 1);
            }
        }
    }

    // (4) float64 IoU sums per category with the matched pairs in ascending key order (= the
    //     reference's visiting order, pq.py:109/119).  Keys are unique, so the rank of a pair is
    //     the number of smaller keys: every thread ranks its pairs against all others
    //     (broadcast reads, no barriers) and drops IoU + category at the ranked position.
    const int nm = C.nm < T.max_matched ? C.nm : T.max_matched;     // stable since the last barrier
    for (int i = tid; i < nm; i += nthreads) {
        const long long key = T.m_key[i];
        int rank = 0;
        for (int j = 0; j < nm; ++j) rank += T.m_key[j] < key;
        NPB_ASSERT(rank >= 0 && rank < nm);
        T.s_iou[rank] = (double)T.m_ia[i] / (double)T.m_uni[i];            // pq.py:145
        T.s_cat[rank] = T.m_cat[i];
    }
    __syncthreads();
    // a frame beyond the capacities reports it and contributes nothing: the caller may evaluate
    // it again with npb_pq_update_big_frame
    NPB_TL(prm, 9, wait);
    const bool failed = C.fail != 0 || prm.status[b] == NPB_ERR_CAPACITY;
    if (failed && tid == 0) set_status(prm.status + b, NPB_ERR_CAPACITY);
    for (int c = tid; c < NC; c += nthreads) {
        double acc = 0.0;
        if (c < 256 && C.tp[c] > 0)
            for (int i = 0; i < nm; ++i)
                if (T.s_cat[i] == c) acc += T.s_iou[i];
        double *fs = prm.frame_stats + (size_t)b * 4 * NC;
        fs[c] = failed ? 0.0 : acc;
        fs[NC + c] = (c < 256 && !failed) ? (double)C.tp[c] : 0.0;
        fs[2 * NC + c] = (c < 256 && !failed) ? (double)C.fn[c] : 0.0;
        fs[3 * NC + c] = (c < 256 && !failed) ? (double)C.fp[c] : 0.0;
    }
    if (tid == 0 && prm.n_matches)
        prm.n_matches[b] = failed ? 0 : (C.nm < prm.match_cap ? C.nm : prm.match_cap);
}

__global__ void __launch_bounds__(kMatchThreads, 1) match_frames_kernel(const MatchParams prm)
{
    extern __shared__ unsigned char smem_raw[];
    // pair table of the frame
    unsigned long long *t_key = (unsigned long long *)smem_raw;     // [kPairSlots]
    // matched pairs: key, intersection, union
    long long *s_mkey = (long long *)(t_key + kPairSlots);          // [kMaxMatched]
    unsigned long long *g_id = (unsigned long long *)(s_mkey + kMaxMatched);   // [kSegSlots]
    unsigned long long *p_id = g_id + kSegSlots;                    // [kSegSlots]
    unsigned *t_cnt = (unsigned *)(p_id + kSegSlots);               // [kPairSlots]
    unsigned *s_mia = t_cnt + kPairSlots;                           // [kMaxMatched]
    unsigned *s_muni = s_mia + kMaxMatched;                         // [kMaxMatched]
    unsigned *g_area = s_muni + kMaxMatched;                        // [kSegSlots] ...
    unsigned *p_area = g_area + kSegSlots;
    unsigned *p_void = p_area + kSegSlots;
    unsigned *p_pio = p_void + kSegSlots;
    unsigned short *t_gslot = (unsigned short *)(p_pio + kSegSlots);       // [kPairSlots]
    unsigned short *t_pslot = t_gslot + kPairSlots;                        // [kPairSlots]
    unsigned short *s_mcat = t_pslot + kPairSlots;                         // [kMaxMatched]
    unsigned short *s_idx = s_mcat + kMaxMatched;                          // [kMaxPairs] used slots
    unsigned char *g_matched = (unsigned char *)(s_idx + kMaxPairs);       // [kSegSlots]
    unsigned char *p_matched = g_matched + kSegSlots;                      // [kSegSlots]
    __shared__ int s_m;
    __shared__ MatchCounters C;

    const int b = blockIdx.x, tid = threadIdx.x;
    NPB_TL(prm, 3, start);
    if (tid == 0) { s_m = 0; C.nm = 0; C.fail = 0; }
    for (int c = tid; c < 256; c += kMatchThreads) { C.tp[c] = 0; C.fn[c] = 0; C.fp[c] = 0; }
    for (int i = tid; i < kPairSlots; i += kMatchThreads) { t_key[i] = kEmptyKey; t_cnt[i] = 0; }
    for (int i = tid; i < kSegSlots; i += kMatchThreads) {
        g_id[i] = kEmptyKey; p_id[i] = kEmptyKey;
        g_area[i] = 0; g_matched[i] = 0; p_area[i] = 0; p_void[i] = 0; p_pio[i] = 0; p_matched[i] = 0;
    }
    grid_dependency_wait();       // the pixel pass (launch_dependent: the set-up above overlaps its tail)
    // Dependents may be staged from here on (not earlier: everything before this grid has
    // completed now, which is what a successor that does not wait for THIS grid relies on -- the
    // NMS pass of the next pipelined call, api.cu)
    grid_launch_dependents();
    NPB_TL(prm, 3, wait);
    __syncthreads();

    // (0) the frame's pairs -> pair table.  Every used slot is remembered in s_idx, so the later
    //     phases walk the (few hundred) pairs, not the table.
    auto add_pair = [&](unsigned long long key, unsigned cnt) {
        unsigned h = hash64(key) & (unsigned)(kPairSlots - 1);
        for (int probe = 0; probe < kPairSlots; ++probe) {
            unsigned long long cur = t_key[h];
            if (cur == kEmptyKey) {
                cur = atomicCAS(t_key + h, kEmptyKey, key);
                if (cur == kEmptyKey) {                      // this thread claimed the slot
                    NPB_ASSERT(h < (unsigned)kPairSlots);
                    const int idx = atomicAdd(&s_m, 1);
                    if (idx < kMaxPairs) s_idx[idx] = (unsigned short)h;
                    else C.fail = 1;
                }
            }
            if (cur == kEmptyKey || cur == key) {
                atomicAdd(t_cnt + h, cnt);
                return;
            }
            h = (h + 1) & (unsigned)(kPairSlots - 1);
        }
        C.fail = 1;
    };
    // The hand-over tables are ZERO AT REST: what the matcher has consumed -- the entry counter of
    // the frame, the cells of the dense table, the frame counter -- it puts back to zero, so the
    // pixel pass of the next update finds them clean without a memset in between (and the matcher
    // of an update may run while the NEXT call's centre detection and grouping are already busy,
    // npb_panoptic_forward_eval_pipelined).
    // entries of the frame (loaded first: the dense part below hides the latency)
    const unsigned n_raw = prm.entry_n[b];
    const unsigned n_ent = n_raw < prm.entry_cap ? n_raw : prm.entry_cap;   // overflow: flagged by the writer
    // class pairs: dense per-frame table, every pair exactly once (independent loads first)
    if (prm.frame_dense) {
        const int nd = prm.nd, nd2 = nd * nd;
        unsigned *fd = prm.frame_dense + (size_t)b * nd2;
        constexpr int kBatch = 4;
        for (int i0 = tid; i0 < nd2; i0 += kMatchThreads * kBatch) {
            unsigned c[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int i = i0 + u * kMatchThreads;
                c[u] = i < nd2 ? fd[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (c[u] == 0u) continue;
                const int i = i0 + u * kMatchThreads;
                fd[i] = 0u;
                const unsigned long long tc = (unsigned)(i / nd), pc = (unsigned)(i % nd);
                add_pair((tc << prm.L_shift) * (unsigned long long)prm.offset + (pc << prm.L_shift), c[u]);
            }
        }
    }
    NPB_TL(prm, 5, wait);
    // instance pairs: the entry list (one entry per pair and CTA of the pixel pass)
    {
        const unsigned long long *ekeys = prm.entry_keys + (size_t)b * prm.entry_cap;
        const unsigned *ecnts = prm.entry_cnts + (size_t)b * prm.entry_cap;
        constexpr int kBatch = 4;
        for (unsigned e0 = tid; e0 < n_ent; e0 += kMatchThreads * kBatch) {
            unsigned long long k[kBatch];
            unsigned c[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const unsigned e = e0 + u * kMatchThreads;
                k[u] = e < n_ent ? ekeys[e] : kEmptyKey;
                c[u] = e < n_ent ? ecnts[e] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
                if (c[u] != 0u) add_pair(k[u], c[u]);
        }
    }
    __syncthreads();
    NPB_TL(prm, 6, wait);       // pair table merged (slot 5 = dense part done)
    const int m = s_m < kMaxPairs ? s_m : kMaxPairs;

    MatchTables<unsigned short> T;
    T.t_key = t_key; T.t_cnt = t_cnt; T.t_gslot = t_gslot; T.t_pslot = t_pslot;
    T.pair_slots = kPairSlots; T.walk = s_idx;
    T.g_id = g_id; T.p_id = p_id; T.g_area = g_area; T.p_area = p_area; T.p_void = p_void;
    T.p_pio = p_pio; T.g_matched = g_matched; T.p_matched = p_matched; T.seg_slots = kSegSlots;
    T.m_key = s_mkey; T.m_ia = s_mia; T.m_uni = s_muni; T.m_cat = s_mcat;
    // ordered IoUs / categories reuse the slot arrays of the pair table (free after phase 2)
    T.s_iou = (double *)t_gslot;
    T.s_cat = (unsigned short *)((double *)t_gslot + kMaxMatched);
    T.max_matched = kMaxMatched;
    match_phases(prm, b, T, m, C);
    if (tid == 0) prm.entry_n[b] = 0u;      // zero at rest (every thread has read it barriers ago)
    NPB_TL(prm, 4, start);      // slot 4: the accumulation tail (start = matching done)

    // state += frame results, frames in order (PanopticQuality.update, pq.py:298-303), by the CTA
    // that finishes last: the results of `chunk` frames are staged in shared memory (coalesced,
    // independent loads), then thread (statistic, category) adds them strictly in frame order --
    // the float64 order of the reference.
    if (!prm.done_cnt) return;
    __shared__ int s_last;
    __syncthreads();            // the frame results of every thread are issued ...
    if (tid == 0) {             // ... and made visible by one fence (cumulative over the barrier)
        __threadfence();
        s_last = (atomicAdd(prm.done_cnt, 1u) == gridDim.x - 1);
        if (s_last) __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    const int NC = prm.num_categories, row = 4 * NC, B = (int)gridDim.x;
    static_assert(kMatchThreads >= 4 * 256, "one thread per (statistic, category)");
    if (tid < row) {
        const int stat = tid / NC, c = tid - stat * NC;
        double *dst = (stat == 0 ? prm.iou : stat == 1 ? prm.tp : stat == 2 ? prm.fn : prm.fp) + c;
        double acc = *dst;
        // journal: the states as they were before this update (row B).  The host re-accumulates
        // from it, in frame order, when a frame of this update has to be redone by
        // npb_pq_update_big_frame (PanopticQuality._replay): the float64 sums then still equal
        // the reference's frame-by-frame order, pq.py:298-303
        prm.frame_stats[(size_t)B * row + tid] = acc;
        // batches of independent loads (adjacent threads read adjacent words), added strictly in
        // frame order
        constexpr int kFrames = 8;
        for (int b0 = 0; b0 < B; b0 += kFrames) {
            double v[kFrames];
#pragma unroll
            for (int j = 0; j < kFrames; ++j)
                v[j] = b0 + j < B ? __ldcg(prm.frame_stats + (size_t)(b0 + j) * row + tid) : 0.0;
#pragma unroll
            for (int j = 0; j < kFrames; ++j)
                if (b0 + j < B) acc += v[j];
        }
        *dst = acc;
    }
    if (tid == 0) *prm.done_cnt = 0u;       // zero at rest
    NPB_TL(prm, 4, end);
}

// ---- fall-back for frames beyond the shared-memory capacities ---------------------------------
// One frame per call, every table in global memory (npb_pq_update_big_frame): a plain pixel pass
// into one large hash table (pair -> pixels), then the phases of the matcher above on it.  Slow
// (one CTA matches), but it takes a frame of per-pixel random ids.
struct BigParams {
    const long long *pred, *target;
    long long P, offset;
    int O_shift;
    unsigned long long *t_key;
    unsigned *t_cnt;
    int pair_slots;
    int32_t *status;
};

constexpr int kBigRun = 8;      // consecutive pixels per thread (equal neighbours are merged)

__global__ void __launch_bounds__(256) big_pair_kernel(const BigParams prm)
{
    for (long long p0 = ((long long)blockIdx.x * 256 + threadIdx.x) * kBigRun; p0 < prm.P;
         p0 += (long long)gridDim.x * 256 * kBigRun) {
        unsigned long long run_key = kEmptyKey;
        unsigned run = 0;
        for (int j = 0; j < kBigRun && p0 + j < prm.P; ++j) {
            const long long pv = prm.pred[p0 + j], tv = prm.target[p0 + j];
            if (pv < 0 || tv < 0 || pv >= prm.offset) {
                set_status(prm.status, NPB_ERR_CATEGORY_RANGE);
                continue;
            }
            const unsigned long long key =
                prm.O_shift >= 0 ? (((unsigned long long)tv << prm.O_shift) | (unsigned long long)pv)
                                 : (unsigned long long)tv * (unsigned long long)prm.offset + (unsigned long long)pv;
            if (key == run_key) { ++run; continue; }
            if (run && !table_add(prm.t_key, prm.t_cnt, prm.pair_slots, prm.pair_slots, run_key, run))
                set_status(prm.status, NPB_ERR_CAPACITY);
            run_key = key;
            run = 1;
        }
        if (run && !table_add(prm.t_key, prm.t_cnt, prm.pair_slots, prm.pair_slots, run_key, run))
            set_status(prm.status, NPB_ERR_CAPACITY);
    }
}

__global__ void __launch_bounds__(kMatchThreads)
match_big_frame_kernel(const MatchParams prm, const MatchTables<unsigned> T)
{
    __shared__ MatchCounters C;
    const int tid = threadIdx.x;
    if (tid == 0) { C.nm = 0; C.fail = 0; }
    for (int c = tid; c < 256; c += kMatchThreads) { C.tp[c] = 0; C.fn[c] = 0; C.fp[c] = 0; }
    __syncthreads();
    match_phases(prm, 0, T, T.pair_slots, C);
}

// state += frame result, frames in order (PanopticQuality.update, pq.py:298-303).
// One warp per (category, statistic): the lanes fetch 32 frames at once (the loads are
// independent), lane 0 then adds them strictly in frame order -- the float64 order of the
// reference -- so the sequential part touches registers only.
__global__ void __launch_bounds__(128)
accumulate_frames_kernel(const double *frame_stats, int B, int NC,
                         double *iou, double *tp, double *fn, double *fp)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    grid_dependency_wait();       // the matcher (this kernel is launched with launch_dependent)
    if (warp >= 4 * NC) return;
    const int stat = warp / NC, c = warp - stat * NC;
    double *dst = stat == 0 ? iou : stat == 1 ? tp : stat == 2 ? fn : fp;
    double acc = dst[c];
    for (int b0 = 0; b0 < B; b0 += 32) {
        const int b = b0 + lane;
        const double v = b < B ? frame_stats[((size_t)b * 4 + stat) * NC + c] : 0.0;
        const int n = min(32, B - b0);
        for (int j = 0; j < n; ++j) acc += __shfl_sync(kFullMask, v, j);
    }
    if (lane == 0) dst[c] = acc;
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

// Generic form: one element per thread and round, any dtype / alignment, any n.
template <bool SKIP_VOID>
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
            long long t = load_int(target, td, (size_t)i);
            if (!SKIP_VOID || t != 0) {
                if (SKIP_VOID) t -= 1;
                if (p < 0 || p >= n || t < 0 || t >= n) set_status(status, NPB_ERR_CATEGORY_RANGE);
                else key = (int)(t * n + p);
            }
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

// Streaming form (n <= kSmemConfmatMaxN, 16-byte aligned maps).  A warp takes 512 consecutive
// elements per round.  Loads: 8 rows of 64 elements, lane l owns the elements 64 j + 2 l, + 1 of
// row j, so every load instruction of the warp covers one contiguous, fully used range (int64
// predictions: 512 B per LDG.128 = 4 lines); all 16 loads of a thread are issued before the
// first use (136 B in flight per thread for int64 + uint8).  Counting: the keys (16 bit, two
// per word) are transposed through a 1 KB shared-memory tile per warp so that every lane holds
// 16 CONSECUTIVE elements.  Class maps are piecewise constant: the elements of a lane that
// agree with its first one are counted together, adjacent lanes with the same first key are
// merged into runs (inclusive scan of the counts with shuffles, one ballot for the run heads)
// and each run costs ONE shared-memory atomic; the (few) other elements are added singly.
// Measured alternatives on B200 (256 frames 480x640, int64 + uint8, fraction of the HBM peak):
// 16 consecutive elements per lane loaded directly (32 lines per LDG.128) + MATCH.ANY / REDUX
// 0.65; row layout with one MATCH.ANY + per-group REDUX per row 0.40-0.54.
constexpr int kCmRows = 8;
constexpr int kCmChunk = 256 * 2 * kCmRows;      // elements per CTA and round

// two adjacent elements of a map as 32-bit words; `wide`: the value does not fit (a 64-bit
// element with a non-zero high word: negative or >= 2^32, never a valid class).  Negative
// 16 / 32-bit values become large unsigned numbers and fail the range check the same way.
template <typename T> struct CmPair;
template <> struct CmPair<uint8_t> {
    typedef unsigned short raw;
    static __device__ __forceinline__ void split(raw r, unsigned &a, unsigned &b, bool &wa, bool &wb)
    { a = r & 0xffu; b = r >> 8; wa = wb = false; }
};
template <> struct CmPair<int16_t> {
    typedef int raw;
    static __device__ __forceinline__ void split(raw r, unsigned &a, unsigned &b, bool &wa, bool &wb)
    { a = (unsigned)(int)(short)(r & 0xffff); b = (unsigned)(r >> 16); wa = wb = false; }
};
template <> struct CmPair<int32_t> {
    typedef int2 raw;
    static __device__ __forceinline__ void split(raw r, unsigned &a, unsigned &b, bool &wa, bool &wb)
    { a = (unsigned)r.x; b = (unsigned)r.y; wa = wb = false; }
};
template <> struct CmPair<long long> {
    typedef uint4 raw;
    static __device__ __forceinline__ void split(raw r, unsigned &a, unsigned &b, bool &wa, bool &wb)
    { a = r.x; wa = r.y != 0u; b = r.z; wb = r.w != 0u; }
};

// key = target * n + pred (16 bit), 0xffff for an element that is not counted
template <bool SKIP_VOID>
__device__ __forceinline__ unsigned confmat_key(unsigned p, bool p_wide, unsigned t, bool t_wide,
                                                unsigned n, bool &bad)
{
    const bool skip = SKIP_VOID && t == 0u && !t_wide;
    if (SKIP_VOID) t -= 1u;                                  // void wraps to 0xffffffff
    const bool ok = p < n && t < n && !p_wide && !t_wide;
    bad |= !ok && !skip;
    return ok ? t * n + p : 0xffffu;
}

template <bool SKIP_VOID, typename PT, typename TT>
__device__ __forceinline__ unsigned confmat_key_at(const PT *preds, const TT *target, long long i,
                                                   long long N, unsigned n, bool &bad)
{
    if (i >= N) return 0xffffu;
    const long long p = (long long)preds[i], t = (long long)target[i];
    return confmat_key<SKIP_VOID>((unsigned)p, (unsigned long long)p >> 32 != 0ull, (unsigned)t,
                                  (unsigned long long)t >> 32 != 0ull, n, bad);
}

template <typename PT, typename TT, bool SKIP_VOID>
__global__ void __launch_bounds__(256, 4)
confmat_stream_kernel(const PT *__restrict__ preds, const TT *__restrict__ target, long long N,
                      int n, unsigned long long *__restrict__ confmat,
                      int32_t *__restrict__ status)
{
    typedef typename CmPair<PT>::raw PRaw;
    typedef typename CmPair<TT>::raw TRaw;
    extern __shared__ unsigned s_cm[];
    __shared__ __align__(16) unsigned s_tile[8][32 * kCmRows];     // per warp: [row][lane]
    for (int i = threadIdx.x; i < n * n; i += 256) s_cm[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *tile = s_tile[warp];
    const long long n_chunks = (N + kCmChunk - 1) / kCmChunk;
    bool bad = false;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        // first element of this lane in row 0 of its warp's 512 elements
        const long long i0 = c * kCmChunk + warp * (64 * kCmRows) + 2 * lane;
        if (c * kCmChunk + kCmChunk <= N) {
            PRaw pr[kCmRows];
            TRaw tr[kCmRows];
#pragma unroll
            for (int j = 0; j < kCmRows; ++j) pr[j] = __ldcs((const PRaw *)(preds + i0 + 64 * j));
#pragma unroll
            for (int j = 0; j < kCmRows; ++j) tr[j] = __ldcs((const TRaw *)(target + i0 + 64 * j));
#pragma unroll
            for (int j = 0; j < kCmRows; ++j) {
                unsigned pa, pb, ta, tb;
                bool wpa, wpb, wta, wtb;
                CmPair<PT>::split(pr[j], pa, pb, wpa, wpb);
                CmPair<TT>::split(tr[j], ta, tb, wta, wtb);
                const unsigned ka = confmat_key<SKIP_VOID>(pa, wpa, ta, wta, (unsigned)n, bad);
                const unsigned kb = confmat_key<SKIP_VOID>(pb, wpb, tb, wtb, (unsigned)n, bad);
                tile[32 * j + lane] = ka | (kb << 16);
            }
        } else {
            // the last, partial chunk of the map
#pragma unroll 1
            for (int j = 0; j < kCmRows; ++j) {
                const long long i = i0 + 64 * j;
                const unsigned ka = confmat_key_at<SKIP_VOID>(preds, target, i, N, (unsigned)n, bad);
                const unsigned kb = confmat_key_at<SKIP_VOID>(preds, target, i + 1, N, (unsigned)n, bad);
                tile[32 * j + lane] = ka | (kb << 16);
            }
        }
        __syncwarp();
        // 16 consecutive elements of this lane: words 8 lane .. 8 lane + 7
        const uint4 w0 = *(const uint4 *)(tile + 8 * lane);
        const uint4 w1 = *(const uint4 *)(tile + 8 * lane + 4);
        __syncwarp();
        const unsigned w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const unsigned lead = w[0] & 0xffffu;           // 0xffff: not counted
        const unsigned lead2 = lead * 0x10001u;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned x = w[i] ^ lead2;
            if (x == 0u) { cnt += 2; continue; }
            const unsigned a = w[i] & 0xffffu, b = w[i] >> 16;
            if ((x & 0xffffu) == 0u) cnt += 1;
            else if (a != 0xffffu) atomicAdd(s_cm + a, 1u);
            if ((x >> 16) == 0u) cnt += 1;
            else if (b != 0xffffu) atomicAdd(s_cm + b, 1u);
        }
        if (lead == 0xffffu) cnt = 0;
        // runs of adjacent lanes with the same first key
        const unsigned prev = __shfl_up_sync(kFullMask, lead, 1);
        const unsigned heads = __ballot_sync(kFullMask, lane == 0 || prev != lead);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl += up;
        }
        const unsigned above = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);   // heads after this lane
        const int last = above ? __ffs(above) - 2 : 31;                          // last lane of my run
        const int run = __shfl_sync(kFullMask, incl, last) - (incl - cnt);
        if (((heads >> lane) & 1u) && run > 0) atomicAdd(s_cm + lead, (unsigned)run);
    }
    if (bad) set_status(status, NPB_ERR_CATEGORY_RANGE);
    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += 256)
        if (s_cm[i]) atomicAdd(confmat + i, (unsigned long long)s_cm[i]);
}


}  // namespace npb

using namespace npb;

// Entries per frame of the list between pixel pass and matcher: every CTA of the pixel pass
// contributes at most its distinct pairs; the fewer frames, the more CTAs work on one frame.
static int device_sm_count()
{
    static int n_sm_dev[64] = {0};
    static std::mutex guard;
    std::lock_guard<std::mutex> lock(guard);
    int dev = 0;
    cudaGetDevice(&dev);
    int &n = n_sm_dev[dev & 63];
    if (n == 0) {
        n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <typename PT, typename TT>
static void launch_confmat_stream(const void *preds, const void *target, long long N, int n,
                                  bool skip_void, int64_t *confmat, int32_t *status,
                                  cudaStream_t stream)
{
    long long blocks = (N + kCmChunk - 1) / kCmChunk;
    const size_t smem = (size_t)n * n * 4;
    // one wave of persistent CTAs (occupancy of the instance, normally 4 per SM)
    static std::atomic<int> ctas_per_sm_skip{0}, ctas_per_sm_all{0};
    std::atomic<int> &cached = skip_void ? ctas_per_sm_skip : ctas_per_sm_all;
    int per_sm = cached.load(std::memory_order_relaxed);
    if (per_sm == 0) {
        per_sm = 3;
        if (skip_void)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &per_sm, confmat_stream_kernel<PT, TT, true>, 256, kSmemConfmatMaxN * kSmemConfmatMaxN * 4);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &per_sm, confmat_stream_kernel<PT, TT, false>, 256, kSmemConfmatMaxN * kSmemConfmatMaxN * 4);
        if (per_sm < 1) per_sm = 1;
        cached.store(per_sm, std::memory_order_relaxed);
    }
    const long long wave = (long long)device_sm_count() * per_sm;
    if (blocks > wave) blocks = wave;
    if (skip_void)
        confmat_stream_kernel<PT, TT, true><<<(int)blocks, 256, smem, stream>>>(
            (const PT *)preds, (const TT *)target, N, n, (unsigned long long *)confmat, status);
    else
        confmat_stream_kernel<PT, TT, false><<<(int)blocks, 256, smem, stream>>>(
            (const PT *)preds, (const TT *)target, N, n, (unsigned long long *)confmat, status);
}

template <typename PT>
static void launch_confmat_stream_p(const void *preds, const void *target, int td, long long N,
                                    int n, bool skip_void, int64_t *confmat, int32_t *status,
                                    cudaStream_t stream)
{
    switch (td) {
        case NPB_U8: case NPB_BOOL:
            launch_confmat_stream<PT, uint8_t>(preds, target, N, n, skip_void, confmat, status, stream);
            break;
        case NPB_I16:
            launch_confmat_stream<PT, int16_t>(preds, target, N, n, skip_void, confmat, status, stream);
            break;
        case NPB_I32:
            launch_confmat_stream<PT, int32_t>(preds, target, N, n, skip_void, confmat, status, stream);
            break;
        default:
            launch_confmat_stream<PT, long long>(preds, target, N, n, skip_void, confmat, status, stream);
    }
}

static int confmat_update(const void *preds, int preds_dtype, const void *target, int target_dtype,
                          int64_t N, int n_classes, bool skip_void, int64_t *confmat,
                          int32_t *status, void *stream, const char *what)
{
    if (N < 0 || n_classes < 1 || n_classes > 46340) return NPB_ERR_ARG;
    if (preds_dtype < NPB_U8 || preds_dtype > NPB_BOOL || target_dtype < NPB_U8 ||
        target_dtype > NPB_BOOL)
        return NPB_ERR_ARG;
    if (N == 0) return NPB_OK;          // empty maps (their pointers may be null) add nothing
    if (!preds || !target || !confmat || !status) return NPB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)preds | (uintptr_t)target) & 15u) == 0;
    if (aligned && n_classes <= kSmemConfmatMaxN) {
        switch (preds_dtype) {
            case NPB_U8: case NPB_BOOL:
                launch_confmat_stream_p<uint8_t>(preds, target, target_dtype, N, n_classes,
                                                 skip_void, confmat, status, st);
                break;
            case NPB_I16:
                launch_confmat_stream_p<int16_t>(preds, target, target_dtype, N, n_classes,
                                                 skip_void, confmat, status, st);
                break;
            case NPB_I32:
                launch_confmat_stream_p<int32_t>(preds, target, target_dtype, N, n_classes,
                                                 skip_void, confmat, status, st);
                break;
            default:
                launch_confmat_stream_p<long long>(preds, target, target_dtype, N, n_classes,
                                                   skip_void, confmat, status, st);
        }
        return record_launch(what);
    }
    long long blocks = (N + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 16) blocks = 148 * 16;
    const size_t smem = n_classes <= kSmemConfmatMaxN ? (size_t)n_classes * n_classes * 4 : 0;
    if (skip_void)
        confmat_kernel<true><<<(int)blocks, 256, smem, st>>>(
            preds, preds_dtype, target, target_dtype, (long long)N, n_classes,
            (unsigned long long *)confmat, status);
    else
        confmat_kernel<false><<<(int)blocks, 256, smem, st>>>(
            preds, preds_dtype, target, target_dtype, (long long)N, n_classes,
            (unsigned long long *)confmat, status);
    return record_launch(what);
}

extern "C" int npb_confmat_update(const void *preds, int preds_dtype, const void *target,
                                  int target_dtype, int64_t N, int n_classes, int64_t *confmat,
                                  int32_t *status, void *stream)
{
    return confmat_update(preds, preds_dtype, target, target_dtype, N, n_classes, false, confmat,
                          status, stream, "npb_confmat_update");
}

extern "C" int npb_confmat_update_nonvoid(const void *preds, int preds_dtype, const void *target,
                                          int target_dtype, int64_t N, int n_classes,
                                          int64_t *confmat, int32_t *status, void *stream)
{
    return confmat_update(preds, preds_dtype, target, target_dtype, N, n_classes, true, confmat,
                          status, stream, "npb_confmat_update_nonvoid");
}

constexpr int kMaxPairCtasPerSm = 8;

static unsigned pq_entry_cap(int B)
{
    // every CTA of the pixel pass appends at most its hash table; the rest is head room for
    // single entries of overflowing CTA tables
    // (a frame is touched by at most ceil(slots / B) + 1 CTAs: the equal shares of the batch's
    // chunks may start and end inside a frame)
    const long long ctas_per_frame = ((long long)device_sm_count() * kMaxPairCtasPerSm + B - 1) / B + 1;
    return (unsigned)(4 * kMaxPairs + (long long)kSmemSlots * ctas_per_frame);
}

// side of the dense class-pair table: it needs `id >> shift` decoding and has to fit into the
// shared memory of the pixel pass next to the other tables
static int pq_dense_side(int num_categories, int64_t max_instances_per_category)
{
    const bool pow2 = (max_instances_per_category & (max_instances_per_category - 1)) == 0 &&
                      max_instances_per_category < (1ll << 62);
    return (pow2 && num_categories <= kSmemConfmatMaxN) ? num_categories : 0;
}

// workspace: [entry_keys | entry_cnts | (entry_n, frames done | frame_dense: one memset) | frame_stats (B + 1 rows)]
static size_t pq_counter_bytes(int B) { return align256((size_t)(B + 1) * sizeof(unsigned)); }
static size_t pq_cleared_bytes(int B)
{
    return pq_counter_bytes(B) +
           align256((size_t)B * kSmemConfmatMaxN * kSmemConfmatMaxN * sizeof(unsigned));
}

struct PqWorkspace {
    unsigned long long *ekeys;
    unsigned *ecnts;
    unsigned *en;           // [B] entries per frame, [B] = frames matched
    unsigned *fdense;
    double *fstats;
    unsigned entry_cap;
};

static PqWorkspace pq_workspace(void *workspace, int B)
{
    PqWorkspace w;
    w.entry_cap = pq_entry_cap(B);
    char *ws = (char *)workspace;
    w.ekeys = (unsigned long long *)ws;
    ws += align256((size_t)B * w.entry_cap * sizeof(unsigned long long));
    w.ecnts = (unsigned *)ws;
    ws += align256((size_t)B * w.entry_cap * sizeof(unsigned));
    w.en = (unsigned *)ws;
    w.fdense = (unsigned *)(ws + pq_counter_bytes(B));
    ws += pq_cleared_bytes(B);
    w.fstats = (double *)ws;
    return w;
}

// the one memset of an update (entry counters, frame counter, dense class-pair tables); the
// forward chain issues it at its head so that no memset separates its kernels
void npb::pq_cleared_range(void *workspace, int B, int num_categories,
                           int64_t max_instances_per_category, void **p, size_t *bytes)
{
    const PqWorkspace w = pq_workspace(workspace, B);
    const int nd = pq_dense_side(num_categories, max_instances_per_category);
    *p = w.en;
    *bytes = pq_counter_bytes(B) + (size_t)B * nd * nd * sizeof(unsigned);
}

void npb::pq_clear_workspace(void *workspace, int B, int num_categories,
                             int64_t max_instances_per_category, void *stream)
{
    void *p;
    size_t bytes;
    pq_cleared_range(workspace, B, num_categories, max_instances_per_category, &p, &bytes);
    cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream);
}

extern "C" size_t npb_pq_update_workspace_bytes(int B, int num_categories)
{
    const size_t cap = pq_entry_cap(B);
    size_t bytes = align256((size_t)B * cap * sizeof(unsigned long long));
    bytes += align256((size_t)B * cap * sizeof(unsigned));
    bytes += pq_cleared_bytes(B);
    bytes += align256((size_t)(B + 1) * 4 * num_categories * sizeof(double));
    return bytes;
}

// panoptic ids produced inside the pixel pass (npb_write_panoptic_eval) instead of read from `pred`
struct FusedWrite {
    const uint8_t *sem, *inst;
    const int64_t *inst_pan_id;
    ClassSet thing;
    int64_t *pan_out;
    uint8_t *pan_sem_out;
    const FinalizeParams *fold;     // non-null: the pixel pass derives the instance tables itself
};

// the fused pixel pass exists for the reference's id geometry and 4-pixel alignment only
static bool fused_write_supported(int64_t P, int64_t max_instances_per_category, int64_t offset,
                                  const FusedWrite &fw, const void *target, const void *sem_target)
{
    return max_instances_per_category == (1ll << 16) && offset == (1ll << 24) && P % 4 == 0 &&
           (((uintptr_t)fw.pan_out | (uintptr_t)target) & 15u) == 0 &&
           (((uintptr_t)fw.sem | (uintptr_t)fw.inst | (uintptr_t)fw.pan_sem_out |
             (uintptr_t)sem_target) & 3u) == 0;
}

// The matcher + frame accumulation of an update whose pixel pass has been issued.  `dependent`:
// launched as a programmatic dependent of the pixel pass (same stream, right behind it); otherwise
// a plain launch (any stream that is ordered after the pixel pass).
static int launch_match(void *workspace, int B, int num_categories, int64_t ignored_label,
                        int64_t max_instances_per_category, int64_t offset, int64_t void_segment_id,
                        double *iou, double *tp, double *fn, double *fp, double *frame_stats,
                        int64_t *matches, int match_cap, int32_t *n_matches, int32_t *status,
                        bool dependent, cudaStream_t s)
{
    const PqWorkspace w = pq_workspace(workspace, B);
    const int nd = pq_dense_side(num_categories, max_instances_per_category);
    {   // one process may drive several devices: function attributes are per device
        static std::mutex guard;
        static bool attr_set_dev[64] = {false};
        std::lock_guard<std::mutex> lock(guard);
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_set_dev[dev & 63]) {
            cudaFuncSetAttribute(match_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)match_smem_bytes());
            attr_set_dev[dev & 63] = true;
        }
    }
    MatchParams mp = {};
    mp.entry_keys = w.ekeys; mp.entry_cnts = w.ecnts; mp.entry_n = w.en; mp.entry_cap = w.entry_cap;
    mp.frame_dense = nd > 0 ? w.fdense : nullptr; mp.nd = nd;
    mp.num_categories = num_categories;
    mp.ignored_label = ignored_label; mp.L = max_instances_per_category; mp.offset = offset;
    mp.void_segment_id = void_segment_id; mp.frame_stats = frame_stats ? frame_stats : w.fstats;
    mp.L_shift = -1;
    mp.O_shift = -1;
    for (int sh = 0; sh < 62; ++sh) {
        if ((1ll << sh) == max_instances_per_category) mp.L_shift = sh;
        if ((1ll << sh) == offset) mp.O_shift = sh;
    }
    mp.matches = (long long *)matches; mp.match_cap = match_cap; mp.n_matches = n_matches;
    mp.status = status;
    mp.done_cnt = w.en + B; mp.iou = iou; mp.tp = tp; mp.fn = fn; mp.fp = fp;
    NPB_TL_SET(mp);
    if (dependent)
        launch_dependent(match_frames_kernel, dim3(B), dim3(kMatchThreads), match_smem_bytes(), s, mp);
    else
        match_frames_kernel<<<dim3(B), dim3(kMatchThreads), match_smem_bytes(), s>>>(mp);
    return record_launch("npb_pq_update (matcher)");
}

static int pq_update_impl(const int64_t *pred, const FusedWrite *fw, const int64_t *target,
                          const uint8_t *sem_target, int B, int64_t P, int num_categories,
                          int64_t ignored_label, int64_t max_instances_per_category, int64_t offset,
                          int64_t void_segment_id, void *workspace, double *iou, double *tp,
                          double *fn, double *fp, int64_t *confmat, int confmat_n,
                          double *frame_stats, int64_t *matches, int match_cap,
                          int32_t *n_matches, int32_t *status, bool cleared, bool skip_match,
                          void *stream)
{
    if ((!pred && !fw) || !target || !workspace || !iou || !tp || !fn || !fp || !status) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || P < 1 || num_categories < 1 || num_categories > 256 ||
        max_instances_per_category < 1 || offset < 1)
        return NPB_ERR_ARG;
    if ((sem_target != nullptr) != (confmat != nullptr)) return NPB_ERR_ARG;
    if (confmat && (confmat_n < 1 || confmat_n > 256)) return NPB_ERR_ARG;
    if (matches && (match_cap < 1 || !n_matches)) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;

    const PqWorkspace w = pq_workspace(workspace, B);
    const unsigned entry_cap = w.entry_cap;
    unsigned long long *ekeys = w.ekeys;
    unsigned *ecnts = w.ecnts, *en = w.en, *fdense = w.fdense;

    const int nd = pq_dense_side(num_categories, max_instances_per_category);
    if (!cleared) pq_clear_workspace(workspace, B, num_categories, max_instances_per_category, stream);

    PairParams pp = {};
    pp.pred = (const long long *)pred; pp.target = (const long long *)target;
    pp.sem_map = nullptr; pp.inst_map = nullptr; pp.inst_pan_id = nullptr;
    pp.pan_out = nullptr; pp.pan_sem_out = nullptr;
    for (int i = 0; i < 8; ++i) pp.thing.w[i] = 0;
    pp.fold_finalize = 0;
    memset(&pp.fin, 0, sizeof(pp.fin));
    if (fw) {
        pp.sem_map = fw->sem; pp.inst_map = fw->inst;
        pp.inst_pan_id = (const long long *)fw->inst_pan_id; pp.thing = fw->thing;
        pp.pan_out = (long long *)fw->pan_out; pp.pan_sem_out = fw->pan_sem_out;
        if (fw->fold) { pp.fold_finalize = 1; pp.fin = *fw->fold; }
    }
    pp.sem_target = sem_target; pp.P = P; pp.B = B; pp.offset = offset; pp.L = max_instances_per_category;
    pp.L_shift = -1;
    pp.O_shift = -1;
    for (int sh = 0; sh < 62; ++sh) {
        if ((1ll << sh) == max_instances_per_category) pp.L_shift = sh;
        if ((1ll << sh) == offset) pp.O_shift = sh;
    }
    pp.n = confmat ? confmat_n : 0;
    pp.nd = nd;
    pp.frame_dense = nd > 0 ? fdense : nullptr;
    pp.entry_keys = ekeys; pp.entry_cnts = ecnts; pp.entry_n = en; pp.entry_cap = entry_cap;
    pp.confmat = (unsigned long long *)confmat; pp.status = status;
    NPB_TL_SET(pp);

    const bool vec4 = fw != nullptr ||     // checked by fused_write_supported()
                      ((P % 4 == 0) && (((uintptr_t)pred | (uintptr_t)target) & 15u) == 0 &&
                       ((uintptr_t)sem_target & 3u) == 0);
    const int vec = vec4 ? 4 : 1;
    const long long n_chunks = (P + (long long)kPairThreads * vec - 1) / ((long long)kPairThreads * vec);
    // [hash keys | queue keys | hash counts | dense | confmat (if privatised) | queue meta]
    const size_t pc_smem = (size_t)kSmemSlots * 12 + (size_t)(kPairThreads / 32) * kQueueCap * 10 +
                           (size_t)pp.nd * pp.nd * sizeof(unsigned) +
                           ((confmat && confmat_n <= kSmemConfmatMaxN)
                                ? (size_t)confmat_n * confmat_n * sizeof(unsigned) : 0) + 16;
    // one process may drive several devices: function attributes and the SM count are per device
    // specialisation for the reference's id geometry (offset = 256^3, L = 65536)
    const bool std_geom = pp.O_shift == 24 && pp.L_shift == 16;
    typedef void (*PairKernel)(const PairParams);
    constexpr int kVariants = 10;
    static const PairKernel kernels[kVariants] = {
        pair_count_kernel<1, false, false>, pair_count_kernel<4, false, false>,
        pair_count_kernel<1, true, false>,  pair_count_kernel<4, true, false>,
        pair_count_kernel<1, false, true>,  pair_count_kernel<4, false, true>,
        pair_count_kernel<1, true, true>,   pair_count_kernel<4, true, true>,
        pair_count_kernel<4, false, true, true>, pair_count_kernel<4, true, true, true>};
    const int variant = fw ? 8 + (confmat ? 1 : 0)
                           : (std_geom ? 4 : 0) + (confmat ? 2 : 0) + (vec4 ? 1 : 0);
    const PairKernel kernel = kernels[variant];
    // one process may drive several devices: function attributes and the SM count are per device
    // (host threads may call concurrently: the per-device caches below are guarded)
    static std::mutex cache_guard;
    std::unique_lock<std::mutex> cache_lock(cache_guard);
    static bool pc_attr_set_dev[64] = {false};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const int dslot = cur_dev & 63;
    if (!pc_attr_set_dev[dslot]) {
        for (int v = 0; v < kVariants; ++v)
            cudaFuncSetAttribute(kernels[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        pc_attr_set_dev[dslot] = true;
    }
    const int n_sm = device_sm_count();
    // persistent CTAs: the whole batch in ONE wave (a second, partial wave would leave most SMs
    // idle for the length of a CTA), every CTA of a frame gets the same number of chunks
    static size_t occ_smem_dev[64][kVariants] = {{0}};
    static int occ_blocks_dev[64][kVariants] = {{0}};
    int per_sm = occ_blocks_dev[dslot][variant];
    if (per_sm == 0 || occ_smem_dev[dslot][variant] != pc_smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPairThreads, pc_smem);
        if (per_sm < 1) per_sm = 1;
        occ_smem_dev[dslot][variant] = pc_smem;
        occ_blocks_dev[dslot][variant] = per_sm;
    }
    // NPB_PAIR_CTAS_PER_SM=<n> caps the residency of the pixel pass, leaving registers / shared
    // memory for kernels of another stream (evaluation overlapped with the next batch)
    static int env_cap = -1;
    if (env_cap < 0) {
        const char *e = getenv("NPB_PAIR_CTAS_PER_SM");
        env_cap = e ? atoi(e) : 0;
    }
    cache_lock.unlock();
    const int natural_per_sm = per_sm;
    if (env_cap > 0 && per_sm > env_cap) per_sm = env_cap;
    if (per_sm > kMaxPairCtasPerSm) per_sm = kMaxPairCtasPerSm;     // pq_entry_cap() assumes it
    // few frames: every CTA only gets a handful of chunks, so its fixed cost (table set-up, flush)
    // and the duplicates it hands to the matcher weigh more than the extra residency
    static int small_cap = -1;     // NPB_PAIR_SMALL_CTAS=<n>: A/B of that rule
    if (small_cap < 0) {
        const char *e = getenv("NPB_PAIR_SMALL_CTAS");
        small_cap = e ? atoi(e) : 2;
        if (small_cap < 1) small_cap = 2;
    }
    if (B <= 32 && per_sm > small_cap) per_sm = small_cap;
    // one wave of CTAs, each with the same share of the B * n_chunks chunks of the batch
    long long n_ctas = (long long)n_sm * per_sm;
    if (n_ctas > n_chunks * B) n_ctas = n_chunks * B;
    if (n_ctas < 1) n_ctas = 1;
    dim3 grid((unsigned)n_ctas);
    // The grid is one wave of `per_sm` CTAs per SM.  As a programmatic dependent its CTAs are
    // placed while the predecessor drains, SM by SM: with room for more than `per_sm` of them the
    // first free SMs would take four and others none.  Padding the dynamic shared memory makes
    // `per_sm` the residency limit, so the wave spreads evenly whenever it is placed.
    size_t launch_smem = pc_smem;
    if (per_sm < natural_per_sm) {
        const size_t pad = (size_t)(228 * 1024) / (size_t)(per_sm + 1);
        if (pad > launch_smem) launch_smem = pad < (size_t)(200 * 1024) ? pad : (size_t)(200 * 1024);
    }
    launch_dependent(kernel, grid, dim3(kPairThreads), launch_smem, s, pp);

    if (skip_match) return record_launch("npb_pq_update (pixel pass)");
    // the attribute of the matcher is set above
    const int rc = launch_match(workspace, B, num_categories, ignored_label, max_instances_per_category,
                                offset, void_segment_id, iou, tp, fn, fp, frame_stats, matches,
                                match_cap, n_matches, status, true, s);
    if (rc != NPB_OK) return rc;
    return record_launch("npb_pq_update");
}

extern "C" int npb_pq_update(const int64_t *pred, const int64_t *target, const uint8_t *sem_target,
                             int B, int64_t P, int num_categories, int64_t ignored_label,
                             int64_t max_instances_per_category, int64_t offset,
                             int64_t void_segment_id, void *workspace, double *iou, double *tp,
                             double *fn, double *fp, int64_t *confmat, int confmat_n,
                             double *frame_stats, int64_t *matches, int match_cap,
                             int32_t *n_matches, int32_t *status, void *stream)
{
    if (!pred) return NPB_ERR_ARG;
    return pq_update_impl(pred, nullptr, target, sem_target, B, P, num_categories, ignored_label,
                          max_instances_per_category, offset, void_segment_id, workspace, iou, tp,
                          fn, fp, confmat, confmat_n, frame_stats, matches, match_cap, n_matches,
                          status, false, false, stream);
}

// `fold`: derive the instance tables from the vote histograms inside the fused pass (or, when
// the fused pass does not apply, with the finalize kernel first); `cleared`: the caller has
// issued pq_clear_workspace()
int npb::write_panoptic_eval_impl(const uint8_t *sem, const uint8_t *inst, int64_t *inst_pan_id,
                                  int32_t *inst_class, int B, int C, int H, int W,
                                  const uint8_t *h_thing_lut, int64_t max_instances_per_category,
                                  int64_t *pan_out, uint8_t *pan_sem_out, const npb_eval_args *ev,
                                  const FinalizeParams *fold, bool cleared, bool skip_match,
                                  void *stream)
{
    if (!sem || !inst || !inst_pan_id || !pan_out || !h_thing_lut || !ev) return NPB_ERR_ARG;
    if (C < 1 || C > 255 || B < 1 || H < 1 || W < 1) return NPB_ERR_ARG;
    const int64_t P = (int64_t)H * W;
    FusedWrite fw{sem, inst, inst_pan_id, make_class_set(h_thing_lut, C), pan_out, pan_sem_out, fold};
    if (fused_write_supported(P, max_instances_per_category, ev->offset, fw, ev->target,
                              ev->sem_target))
        return pq_update_impl(nullptr, &fw, ev->target, ev->sem_target, B, P, ev->num_categories,
                              ev->ignored_label, max_instances_per_category, ev->offset,
                              ev->void_segment_id, ev->workspace, ev->iou, ev->tp, ev->fn, ev->fp,
                              ev->confmat, ev->confmat_n, ev->frame_stats, ev->matches,
                              ev->match_cap, ev->n_matches, ev->status, cleared, skip_match, stream);
    // other id geometries / unaligned maps: the passes one after the other (same results)
    if (fold) {
        const int rc0 = launch_finalize(*fold, B, stream);
        if (rc0 != NPB_OK) return rc0;
    }
    const int rc = npb_write_panoptic(sem, inst, inst_pan_id, inst_class, B, C, H, W, h_thing_lut,
                                      max_instances_per_category, pan_out, pan_sem_out, stream);
    if (rc != NPB_OK) return rc;
    return pq_update_impl(pan_out, nullptr, ev->target, ev->sem_target, B, P, ev->num_categories,
                          ev->ignored_label, max_instances_per_category, ev->offset,
                          ev->void_segment_id, ev->workspace, ev->iou, ev->tp, ev->fn, ev->fp,
                          ev->confmat, ev->confmat_n, ev->frame_stats, ev->matches, ev->match_cap,
                          ev->n_matches, ev->status, cleared, skip_match, stream);
}

extern "C" int npb_write_panoptic_eval(const uint8_t *sem, const uint8_t *inst,
                                       const int64_t *inst_pan_id, const int32_t *inst_class, int B,
                                       int C, int H, int W, const uint8_t *h_thing_lut,
                                       int64_t max_instances_per_category, int64_t *pan_out,
                                       uint8_t *pan_sem_out, const npb_eval_args *ev, void *stream)
{
    return write_panoptic_eval_impl(sem, inst, (int64_t *)inst_pan_id, (int32_t *)inst_class, B, C,
                                    H, W, h_thing_lut, max_instances_per_category, pan_out,
                                    pan_sem_out, ev, nullptr, false, false, stream);
}

// matcher of an update whose pixel pass was issued with `skip_match` (api.cu: pipelined chain),
// as a programmatic dependent of whatever precedes it on the stream
int npb::pq_match_impl(const npb_eval_args *ev, int B, int64_t max_instances_per_category,
                       void *stream)
{
    if (!ev || !ev->workspace || !ev->iou || !ev->tp || !ev->fn || !ev->fp || !ev->status)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || ev->num_categories < 1 || ev->num_categories > 256 ||
        max_instances_per_category < 1 || ev->offset < 1)
        return NPB_ERR_ARG;
    if (ev->matches && (ev->match_cap < 1 || !ev->n_matches)) return NPB_ERR_ARG;
    return launch_match(ev->workspace, B, ev->num_categories, ev->ignored_label,
                        max_instances_per_category, ev->offset, ev->void_segment_id, ev->iou, ev->tp,
                        ev->fn, ev->fp, ev->frame_stats, ev->matches, ev->match_cap, ev->n_matches,
                        ev->status, true, (cudaStream_t)stream);
}

extern "C" int npb_pq_match_pending(const npb_eval_args *pending, int B,
                                    int64_t max_instances_per_category, void *stream)
{
    return pq_match_impl(pending, B, max_instances_per_category, stream);
}

// ---- fall-back entry point -----------------------------------------------------------------------
static int big_pow2_at_least(long long v, int lo, int hi)
{
    int sh = lo;
    while (sh < hi && (1ll << sh) < v) ++sh;
    return sh;
}

struct BigLayout {
    int pair_slots, seg_slots, max_matched;
    size_t ff_bytes;        // [t_key | g_id | p_id]: cleared to 0xff
    size_t zero_bytes;      // everything else that must start at zero
    size_t total;
};

static BigLayout big_layout(int64_t P, int num_categories)
{
    BigLayout l;
    l.pair_slots = 1 << big_pow2_at_least(2 * P, 12, 23);           // distinct pairs <= pixels
    l.seg_slots = 1 << big_pow2_at_least(2 * (P < 65536 ? P : 65536), 12, 17);
    l.max_matched = 65536;
    l.ff_bytes = align256((size_t)l.pair_slots * 8) + align256((size_t)l.seg_slots * 8) * 2;
    l.zero_bytes = align256((size_t)l.pair_slots * 4) * 3 + align256((size_t)l.seg_slots * 4) * 4 +
                   align256((size_t)l.seg_slots) * 2;
    l.total = l.ff_bytes + l.zero_bytes +
              align256((size_t)l.max_matched * 8) * 2 + align256((size_t)l.max_matched * 4) * 2 +
              align256((size_t)l.max_matched * 2) * 2 + align256((size_t)4 * num_categories * sizeof(double));
    return l;
}

extern "C" size_t npb_pq_update_big_frame_workspace_bytes(int64_t P, int num_categories)
{
    return big_layout(P < 1 ? 1 : P, num_categories < 1 ? 1 : num_categories).total;
}

extern "C" int npb_pq_update_big_frame(const int64_t *pred, const int64_t *target, int64_t P,
                                       int num_categories, int64_t ignored_label,
                                       int64_t max_instances_per_category, int64_t offset,
                                       int64_t void_segment_id, void *workspace, double *iou,
                                       double *tp, double *fn, double *fp, int64_t *matches,
                                       int match_cap, int32_t *n_matches, int32_t *status,
                                       void *stream)
{
    if (!pred || !target || !workspace || !iou || !tp || !fn || !fp || !status) return NPB_ERR_ARG;
    if (P < 1 || num_categories < 1 || num_categories > 256 || max_instances_per_category < 1 ||
        offset < 1)
        return NPB_ERR_ARG;
    if (matches && (match_cap < 1 || !n_matches)) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const BigLayout l = big_layout(P, num_categories);
    char *ws = (char *)workspace;
    auto take = [&](size_t bytes) { char *p = ws; ws += align256(bytes); return p; };

    MatchTables<unsigned> T;
    T.t_key = (unsigned long long *)take((size_t)l.pair_slots * 8);
    T.g_id = (unsigned long long *)take((size_t)l.seg_slots * 8);
    T.p_id = (unsigned long long *)take((size_t)l.seg_slots * 8);
    char *zero0 = ws;
    T.t_cnt = (unsigned *)take((size_t)l.pair_slots * 4);
    T.t_gslot = (unsigned *)take((size_t)l.pair_slots * 4);
    T.t_pslot = (unsigned *)take((size_t)l.pair_slots * 4);
    T.g_area = (unsigned *)take((size_t)l.seg_slots * 4);
    T.p_area = (unsigned *)take((size_t)l.seg_slots * 4);
    T.p_void = (unsigned *)take((size_t)l.seg_slots * 4);
    T.p_pio = (unsigned *)take((size_t)l.seg_slots * 4);
    T.g_matched = (unsigned char *)take((size_t)l.seg_slots);
    T.p_matched = (unsigned char *)take((size_t)l.seg_slots);
    T.m_key = (long long *)take((size_t)l.max_matched * 8);
    T.s_iou = (double *)take((size_t)l.max_matched * 8);
    T.m_ia = (unsigned *)take((size_t)l.max_matched * 4);
    T.m_uni = (unsigned *)take((size_t)l.max_matched * 4);
    T.m_cat = (unsigned short *)take((size_t)l.max_matched * 2);
    T.s_cat = (unsigned short *)take((size_t)l.max_matched * 2);
    double *fstats = (double *)take((size_t)4 * num_categories * sizeof(double));
    T.pair_slots = l.pair_slots; T.seg_slots = l.seg_slots; T.max_matched = l.max_matched;
    T.walk = nullptr;
    cudaMemsetAsync(workspace, 0xff, l.ff_bytes, s);
    cudaMemsetAsync(zero0, 0, l.zero_bytes, s);

    BigParams bp;
    bp.pred = (const long long *)pred; bp.target = (const long long *)target;
    bp.P = P; bp.offset = offset; bp.O_shift = -1;
    MatchParams mp = {};
    mp.L_shift = -1; mp.O_shift = -1;
    for (int sh = 0; sh < 62; ++sh) {
        if ((1ll << sh) == max_instances_per_category) mp.L_shift = sh;
        if ((1ll << sh) == offset) { mp.O_shift = sh; bp.O_shift = sh; }
    }
    bp.t_key = T.t_key; bp.t_cnt = T.t_cnt; bp.pair_slots = l.pair_slots; bp.status = status;
    long long blocks = (P + 256 * kBigRun - 1) / (256 * kBigRun);
    if (blocks > 4096) blocks = 4096;
    big_pair_kernel<<<(unsigned)blocks, 256, 0, s>>>(bp);

    mp.entry_keys = nullptr; mp.entry_cnts = nullptr; mp.entry_n = nullptr; mp.entry_cap = 0;
    mp.frame_dense = nullptr; mp.nd = 0;
    mp.num_categories = num_categories; mp.ignored_label = ignored_label;
    mp.L = max_instances_per_category; mp.offset = offset; mp.void_segment_id = void_segment_id;
    mp.frame_stats = fstats; mp.matches = (long long *)matches; mp.match_cap = match_cap;
    mp.n_matches = n_matches; mp.status = status;
    match_big_frame_kernel<<<1, kMatchThreads, 0, s>>>(mp, T);
    accumulate_frames_kernel<<<(4 * num_categories * 32 + 127) / 128, 128, 0, s>>>(fstats, 1, num_categories,
                                                                        iou, tp, fn, fp);
    return record_launch("npb_pq_update_big_frame");
}
