// centers.cu -- instance-centre detection (heat-map threshold + k x k NMS + top-k)
//
// Replaces InstancePostprocessing._get_instance_centers
// (reference: model/postprocessing/instance.py:78-168).  Semantics kept bit-exactly:
//   :86-88   h = (x > thr) ? x : -1                       (strict >)
//   :97-109  VALID k x k max-pool, ATen keeps the FIRST maximum in row-major window order
//   :123-129 a pixel survives iff it is that first maximum of its own window
//            => strictly greater than the window entries before it, >= those after it;
//            the border ring of width r = (k-1)/2 never survives, except that pixel (0,0)
//            survives when its thresholded value is exactly 0.0 (zero padding quirk)
//   :133,147-149  kth = max(k-th largest of the post-NMS map, 0)
//   :152-166 centres = pixels >= kth (ties can give more than k), raster order
// Only survivors with value >= 0 can ever be selected, so only those are materialised
// ("candidates"); everything else is equivalent to the -1 fill.
//
// NMS pass (nms_candidates_kernel for k >= 5: shared-memory halo tiles of 8 x 128 pixels + halo
//   per CTA, a warp per tile row, whole-tile early out, warp-aggregated append of the survivors;
//   nms_candidates_direct_kernel for k <= 3).
// Selection (select_centers_frame): exact radix select of the k-th largest value over the
//   candidate list of a frame, compaction, rank sort by pixel index.  It runs in the LAST CTA of
//   the NMS pass that finishes the frame (fence + per-frame counter): the candidate list is a
//   few dozen entries, a kernel of its own would cost a launch and one CTA per frame of latency.
#include "common.cuh"

namespace npb {

constexpr int kTileW = 128;   // one warp covers one tile row, 4 consecutive pixels per lane
constexpr int kTileH = 8;
constexpr int kNmsThreads = 256;

struct SelectParams {
    const uint2 *cand;          // [B][cap] (value bits, flat pixel index), written by the NMS pass
    int cap;
    int32_t *cand_cnt;          // [B]
    int32_t *done_cnt;          // [B] CTAs of the NMS pass that finished the frame
    const float *heat;
    const uint8_t *fg;          // nullable
    int H, W, top_k;
    int32_t *centers_yx;
    int32_t *n_centers;
    float *center_score;
    int32_t *status;
    int reset_status;
    // 1: the predecessor on the stream is the matcher of the previous pipelined call, which this
    // pass does not depend on: the dependency wait moves to the END of the first CTA (completion
    // of this grid then still implies completion of the matcher, see api.cu)
    int late_wait;
    // scratch of the LATER stages of the chain, zeroed here by all CTAs (nothing touches it
    // before this grid has completed): no memset node between the kernels of a step
    uint32_t *clear0, *clear1;
    size_t clear0_words, clear1_words;
    // a frame with more than 255 centres (NPB_ERR_TOO_MANY_CENTERS) leaves ALL of them here, as
    // flat pixel indices in raster order, for npb_overflow_centers (the reference's uint8 wrap)
    int32_t *wide_n;            // [B] centres of the frame, -1: more than NPB_MAX_WIDE_CENTERS
    unsigned *wide_idx;         // [B][2][NPB_MAX_WIDE_CENTERS]: unsorted | sorted
    NPB_TL_FIELD
};

constexpr int kWideCap = NPB_MAX_WIDE_CENTERS;

// every CTA of the NMS pass zeroes its share of the downstream scratch
__device__ __forceinline__ void clear_downstream_scratch(const SelectParams &sp)
{
    if (!sp.clear0 && !sp.clear1) return;
    const size_t total = (size_t)gridDim.x * gridDim.y * gridDim.z * blockDim.x;
    const size_t gtid = (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) *
                            blockDim.x + threadIdx.x;
    for (size_t i = gtid; i < sp.clear0_words; i += total) sp.clear0[i] = 0u;
    for (size_t i = gtid; i < sp.clear1_words; i += total) sp.clear1[i] = 0u;
}

// All NT threads of one CTA.  Candidates were written by other CTAs of the same grid: they are
// read through L2 (__ldcg), after the caller's fence.
template <int NT>
__device__ void select_centers_frame(const SelectParams &sp, int b)
{
    static_assert(NT >= kMaxInst, "one thread per centre in the rank sort");
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining;
    __shared__ int s_n;
    __shared__ unsigned s_idx[kMaxInst];

    const int tid = threadIdx.x;
    const int W = sp.W, cap = sp.cap, top_k = sp.top_k;
    const size_t P = (size_t)sp.H * W;
    const uint2 *cb = sp.cand + (size_t)b * cap;
    // the selection is the only writer of the frame's status word in the forward chain: it may
    // start it from NPB_OK itself (saves the chain a memset); thread 0 also does the first write
    if (sp.reset_status && tid == 0) sp.status[b] = NPB_OK;
    int S = __ldcg(sp.cand_cnt + b);
    if (S > cap) {  // cannot happen (cap is the independent-set bound); be loud if it does
        if (tid == 0) set_status(sp.status + b, NPB_ERR_CAPACITY);
        S = cap;
    }

    // exact k-th largest candidate value: 4 passes of an 8-bit radix select on the f32 bits
    // (all candidate values are >= +0.0, so the unsigned bit pattern is order preserving)
    unsigned kth_bits = 0u;
    if (S > top_k) {
        unsigned prefix = 0u, mask = 0u;
        if (tid == 0) s_remaining = (unsigned)top_k;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) hist[tid] = 0u;
            __syncthreads();
            for (int i = tid; i < S; i += NT) {
                const unsigned bits = __ldcg(cb + i).x;
                if ((bits & mask) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned rem = s_remaining;
                int d = 255;
                for (; d > 0; --d) {
                    if (hist[d] >= rem) break;
                    rem -= hist[d];
                }
                s_prefix = prefix | ((unsigned)d << shift);
                s_remaining = rem;
            }
            __syncthreads();
            prefix = s_prefix;
            mask |= 0xFFu << shift;
        }
        kth_bits = prefix;
    }

    if (tid == 0) s_n = 0;
    __syncthreads();
    const uint8_t *fgb = sp.fg ? sp.fg + (size_t)b * P : nullptr;
    for (int i = tid; i < S; i += NT) {
        const uint2 c = __ldcg(cb + i);
        if (c.x >= kth_bits && (!fgb || fgb[c.y])) {
            const int slot = atomicAdd(&s_n, 1);
            NPB_ASSERT(slot >= 0 && c.y < (unsigned)P);
            if (slot < kMaxInst) s_idx[slot] = c.y;
        }
    }
    __syncthreads();
    const int n = s_n;
    // the counters of the frame go back to zero: the next call on this workspace needs no memset
    if (tid == 0) { sp.cand_cnt[b] = 0; sp.done_cnt[b] = 0; }
    if (n > kMaxInst - 1) {
        // Beyond the uint8 instance ids (instance.py:236 wraps them silently): an error for this
        // call.  The complete centre list stays in the workspace so that a caller who wants the
        // reference's wrapped result can redo the frame (npb_overflow_centers +
        // npb_group_pixels_wide).  Rare path: second pass over the candidates, rank sort in
        // global memory.
        if (sp.wide_idx) {
            unsigned *tmp = sp.wide_idx + (size_t)b * 2 * kWideCap, *sorted = tmp + kWideCap;
            if (n <= kWideCap) {
                __syncthreads();
                if (tid == 0) s_n = 0;
                __syncthreads();
                for (int i = tid; i < S; i += NT) {
                    const uint2 c = __ldcg(cb + i);
                    if (c.x >= kth_bits && (!fgb || fgb[c.y])) tmp[atomicAdd(&s_n, 1)] = c.y;
                }
                __syncthreads();
                for (int i = tid; i < n; i += NT) {
                    const unsigned my = tmp[i];
                    int rank = 0;
                    for (int j = 0; j < n; ++j) rank += (tmp[j] < my);
                    sorted[rank] = my;
                }
            }
            if (tid == 0) sp.wide_n[b] = n <= kWideCap ? n : -1;
        }
        if (tid == 0) {
            set_status(sp.status + b, NPB_ERR_TOO_MANY_CENTERS);
            sp.n_centers[b] = 0;
        }
        return;
    }
    if (tid < n) {  // rank sort by flat pixel index = raster (y, x) order of nonzero()
        const unsigned my = s_idx[tid];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (s_idx[j] < my);
        const int y = (int)(my / (unsigned)W), x = (int)(my - (unsigned)y * (unsigned)W);
        int32_t *o = sp.centers_yx + ((size_t)b * kMaxInst + rank) * 2;
        o[0] = y;
        o[1] = x;
        sp.center_score[(size_t)b * kMaxInst + rank] = sp.heat[(size_t)b * P + my];
    }
    if (tid == 0) sp.n_centers[b] = n;
}

// Late dependency wait (SelectParams::late_wait): ONE thread of the grid waits for the predecessor
// before it exits -- enough for "this grid has completed => the predecessor has completed", and
// the other CTAs leave their slots to the CTAs behind them instead of idling in them.
__device__ __forceinline__ void late_dependency_wait(const SelectParams &sp)
{
    if (sp.late_wait && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
        grid_dependency_wait();
}

// End of an NMS CTA: publish its candidates, count the CTA; the CTA that completes the frame
// selects the centres.  Every thread of the CTA must call it (barriers inside).
template <int NT>
__device__ __forceinline__ void nms_frame_epilogue(const SelectParams &sp, int b, int ctas_per_frame)
{
    __shared__ int s_last;
    __threadfence();            // this thread's candidates are visible before the CTA is counted
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(sp.done_cnt + b, 1) == ctas_per_frame - 1);
    __syncthreads();
    NPB_TL(sp, 0, end);
    if (!s_last) return;
    __threadfence();
    select_centers_frame<NT>(sp, b);
    NPB_TL(sp, 0, end);
}

__global__ void __launch_bounds__(kNmsThreads)
nms_candidates_kernel(const float *__restrict__ heat, int H, int W, float thr, int r,
                      uint2 *cand, int cap, int32_t *cand_cnt, const SelectParams sp)
{
    extern __shared__ float tile[];  // (kTileH + 2r) x (kTileW + 2r), thresholded heat + halo
    NPB_TL(sp, 0, start);
    grid_launch_dependents();
    if (!sp.late_wait) grid_dependency_wait();
    NPB_TL(sp, 0, wait);
    clear_downstream_scratch(sp);
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
    const int pitch = kTileW + 2 * r, rows = kTileH + 2 * r;
    const size_t P = (size_t)H * W;
    const float *hb = heat + (size_t)b * P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // warp per tile row: coalesced row segments, no index division
    bool any = false;
    for (int ty = warp; ty < rows; ty += kNmsThreads / 32) {
        const int y = y0 - r + ty;
        const bool row_ok = (y >= 0 && y < H);
        for (int tx = lane; tx < pitch; tx += 32) {
            const int x = x0 - r + tx;
            float v = -1.0f;
            if (row_ok && x >= 0 && x < W) {
                v = __ldg(hb + (size_t)y * W + x);
                v = (v > thr) ? v : -1.0f;
            }
            tile[ty * pitch + tx] = v;
            any |= (v >= 0.0f);
        }
    }
    // most tiles of a real heat-map hold nothing above the threshold: nothing to do there
    const bool tile_hot = __syncthreads_or(any);

    const int ty = warp;                 // kTileH == number of warps
    const int y = y0 + ty;
    const float *center_row = tile + (ty + r) * pitch + r;
#pragma unroll
    for (int j = 0; j < 4 && tile_hot; ++j) {
        const int tx = lane * 4 + j;
        const int x = x0 + tx;
        bool surv = false;
        float v = -1.0f;
        if (y < H && x < W) {
            if (y >= r && y < H - r && x >= r && x < W - r) {
                v = center_row[tx];
                if (v >= 0.0f) {
                    surv = true;
                    for (int dy = -r; dy <= r; ++dy) {
                        const float *row = center_row + dy * pitch + tx;
                        for (int dx = -r; dx <= r; ++dx) {
                            const float w = row[dx];
                            const bool before = (dy < 0) || (dy == 0 && dx < 0);
                            // entries before self must be strictly smaller, the rest <= self
                            if (before ? !(v > w) : (w > v)) surv = false;
                        }
                    }
                }
            } else if (r > 0 && y == 0 && x == 0) {
                v = center_row[tx];
                surv = (v == 0.0f);
            }
        }
        const unsigned m = __ballot_sync(kFullMask, surv);
        if (m) {   // warp-aggregated append: one atomic per warp and pixel slot
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(cand_cnt + b, __popc(m));
            base = __shfl_sync(kFullMask, base, leader);
            if (surv) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                // v + 0.0f canonicalises -0.0 so that the bit pattern orders like the value
                if (slot < cap)
                    cand[(size_t)b * cap + slot] =
                        make_uint2(__float_as_uint(v + 0.0f), (unsigned)(y * W + x));
            }
        }
    }
    nms_frame_epilogue<kNmsThreads>(sp, b, (int)(gridDim.x * gridDim.y));
    late_dependency_wait(sp);
}

// Direct variant for small windows (k <= 3, the default): 4 consecutive pixels per thread from
// one 128-bit load; a thread whose pixels are all below the threshold is done (on a real
// heat-map ~95 % of the threads).  Only pixels above the threshold look at their window,
// through the read-only path (neighbouring rows are L1 / L2 hits), stopping at the first
// neighbour that beats them.  4 B/px of compulsory traffic and ~15 instructions per thread.
constexpr int kNmsGroups = 4;   // independent 128-bit loads per thread

template <int VEC>
__global__ void __launch_bounds__(256)
nms_candidates_direct_kernel(const float *__restrict__ heat, int H, int W, float thr, int r,
                             uint2 *cand, int cap, int32_t *cand_cnt, const SelectParams sp)
{
    NPB_TL(sp, 0, start);
    grid_launch_dependents();
    if (!sp.late_wait) grid_dependency_wait();
    NPB_TL(sp, 0, wait);
    clear_downstream_scratch(sp);
    const int b = blockIdx.y;
    const int P = H * W;
    const int base = blockIdx.x * (256 * VEC * kNmsGroups) + threadIdx.x * VEC;
    const float *hb = heat + (size_t)b * P;
    float v[kNmsGroups][VEC];
    bool any = false;
#pragma unroll
    for (int u = 0; u < kNmsGroups; ++u) {
        const int p0 = base + u * 256 * VEC;
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[u][j] = -1.0f;
        if (p0 < P) {
            if (VEC == 4) {
                const float4 t = __ldg((const float4 *)(hb + p0));
                v[u][0] = t.x; v[u][1 % VEC] = t.y; v[u][2 % VEC] = t.z; v[u][3 % VEC] = t.w;
            } else {
                v[u][0] = __ldg(hb + p0);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kNmsGroups; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            v[u][j] = (v[u][j] > thr) ? v[u][j] : -1.0f;
            any |= (v[u][j] >= 0.0f);
        }
#pragma unroll
    for (int u = 0; u < kNmsGroups && any; ++u) {
        const int p0 = base + u * 256 * VEC;
        if (p0 >= P) continue;
        bool hot = false;
#pragma unroll
        for (int j = 0; j < VEC; ++j) hot |= (v[u][j] >= 0.0f);
        if (!hot) continue;
        int y = p0 / W, x = p0 - y * W;
        if (VEC == 4 && r == 1 && y >= 1 && y < H - 1 && x >= 1 && x + 4 <= W - 1) {
            // all 4 pixels are interior and in one row: fetch the 3 x 6 neighbourhood with
            // independent loads (one latency round), then decide in registers
            float top[6], bot[6], mid[6];
            const float *rt = hb + (size_t)(y - 1) * W + x - 1;
            const float *rb = hb + (size_t)(y + 1) * W + x - 1;
#pragma unroll
            for (int q = 0; q < 6; ++q) { top[q] = __ldg(rt + q); bot[q] = __ldg(rb + q); }
            mid[0] = __ldg(hb + p0 - 1);
            mid[5] = __ldg(hb + p0 + 4);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                top[q] = (top[q] > thr) ? top[q] : -1.0f;
                bot[q] = (bot[q] > thr) ? bot[q] : -1.0f;
            }
            mid[0] = (mid[0] > thr) ? mid[0] : -1.0f;
            mid[5] = (mid[5] > thr) ? mid[5] : -1.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) mid[j + 1] = v[u][j % VEC];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float vj = mid[j + 1];
                // strictly greater than the 4 window entries before it, >= the 4 after it
                const bool surv = vj >= 0.0f && vj > top[j] && vj > top[j + 1] && vj > top[j + 2] &&
                                  vj > mid[j] && !(mid[j + 2] > vj) && !(bot[j] > vj) &&
                                  !(bot[j + 1] > vj) && !(bot[j + 2] > vj);
                if (surv) {
                    const int slot = atomicAdd(cand_cnt + b, 1);
                    if (slot < cap)
                        cand[(size_t)b * cap + slot] =
                            make_uint2(__float_as_uint(vj + 0.0f), (unsigned)(p0 + j));
                }
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float vj = v[u][j];
            bool surv = false;
            if (vj >= 0.0f) {
                if (y >= r && y < H - r && x >= r && x < W - r) {
                    surv = true;
                    for (int dy = -r; dy <= r && surv; ++dy) {
                        const float *row = hb + (size_t)(y + dy) * W + x;
                        for (int dx = -r; dx <= r; ++dx) {
                            if (dy == 0 && dx == 0) continue;
                            float w = __ldg(row + dx);
                            w = (w > thr) ? w : -1.0f;
                            const bool before = (dy < 0) || (dy == 0 && dx < 0);
                            if (before ? !(vj > w) : (w > vj)) { surv = false; break; }
                        }
                    }
                } else if (r > 0 && y == 0 && x == 0) {
                    surv = (vj == 0.0f);
                }
            }
            if (surv) {     // survivors are a handful per frame: plain atomics
                const int slot = atomicAdd(cand_cnt + b, 1);
                if (slot < cap)
                    cand[(size_t)b * cap + slot] =
                        make_uint2(__float_as_uint(vj + 0.0f), (unsigned)(y * W + x));
            }
            if (++x == W) { x = 0; ++y; }
        }
    }
    nms_frame_epilogue<256>(sp, b, (int)gridDim.x);
    late_dependency_wait(sp);
}

static int cand_capacity(int H, int W, int ks)
{
    if (ks <= 1) return H * W;
    return ((H + 1) / 2) * ((W + 1) / 2) + 1;  // survivors are pairwise non-adjacent
}

}  // namespace npb

using namespace npb;

extern "C" size_t npb_instance_centers_workspace_bytes(int B, int H, int W, int nms_kernel_size)
{
    const size_t cap = (size_t)cand_capacity(H, W, nms_kernel_size);
    size_t bytes = (size_t)B * cap * sizeof(uint2);
    bytes = (bytes + 255) & ~(size_t)255;
    bytes += npb::centers_counter_bytes(B);
    bytes += npb::centers_wide_bytes(B);
    return bytes;
}

// internal form: `cleared` = the counters (candidates per frame, finished CTAs per frame: the
// last centers_counter_bytes(B) of the workspace) are zero -- every call leaves them at zero, so
// a workspace that was zeroed once stays usable without a memset; `reset_status` = status starts
// at NPB_OK; `downstream` = scratch of later stages that the NMS pass zeroes on the way
int npb::instance_centers_impl(const float *heat, int B, int H, int W, float threshold,
                               int nms_kernel_size, int top_k, const uint8_t *fg, int apply_fg_mask,
                               void *workspace, int32_t *centers_yx, int32_t *n_centers,
                               float *center_score, int32_t *status, bool cleared,
                               bool reset_status, const ScratchToClear *downstream, bool late_wait,
                               void *stream)
{
    static_assert(kNmsThreads >= kMaxInst, "the selection needs one thread per centre");
    if (!heat || !workspace || !centers_yx || !n_centers || !center_score || !status)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || H < 1 || W < 1 || (nms_kernel_size & 1) == 0 || nms_kernel_size < 1 ||
        nms_kernel_size > 31)
        return NPB_ERR_ARG;
    if (top_k < 1 || (long long)top_k > (long long)H * W) return NPB_ERR_ARG;  // torch.topk raises
    if ((long long)H * W >= (1ll << 31)) return NPB_ERR_ARG;
    if (apply_fg_mask && !fg) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int r = (nms_kernel_size - 1) / 2;
    const int cap = cand_capacity(H, W, nms_kernel_size);
    uint2 *cand = (uint2 *)workspace;
    size_t off = ((size_t)B * cap * sizeof(uint2) + 255) & ~(size_t)255;
    int32_t *cand_cnt = (int32_t *)((char *)workspace + off);

    int32_t *done_cnt = cand_cnt + B;
    int32_t *wide_n = (int32_t *)((char *)cand_cnt + centers_counter_bytes(B));
    unsigned *wide_idx = (unsigned *)((char *)wide_n + align256((size_t)B * sizeof(int32_t)));
    if (!cleared) cudaMemsetAsync(cand_cnt, 0, 2 * (size_t)B * sizeof(int32_t), s);
    SelectParams sp;
    sp.cand = cand; sp.cap = cap; sp.cand_cnt = cand_cnt; sp.done_cnt = done_cnt; sp.heat = heat;
    sp.fg = apply_fg_mask ? fg : nullptr; sp.H = H; sp.W = W; sp.top_k = top_k;
    sp.centers_yx = centers_yx; sp.n_centers = n_centers; sp.center_score = center_score;
    sp.status = status; sp.reset_status = reset_status ? 1 : 0;
    sp.late_wait = late_wait ? 1 : 0;
    sp.clear0 = sp.clear1 = nullptr;
    sp.clear0_words = sp.clear1_words = 0;
    sp.wide_n = wide_n; sp.wide_idx = wide_idx;
    if (downstream) {
        sp.clear0 = (uint32_t *)downstream->p0; sp.clear0_words = downstream->bytes0 / 4;
        sp.clear1 = (uint32_t *)downstream->p1; sp.clear1_words = downstream->bytes1 / 4;
    }
    NPB_TL_SET(sp);
    if (nms_kernel_size <= 3) {
        // small window: per-pixel early out beats staging tiles (see kernel comment)
        const int P = H * W;
        if (P % 4 == 0 && W >= 4 && ((uintptr_t)heat & 15u) == 0) {
            dim3 grid((P / 4 + 256 * kNmsGroups - 1) / (256 * kNmsGroups), B);
            launch_dependent(nms_candidates_direct_kernel<4>, grid, dim3(256), 0, s, heat, H, W,
                             threshold, r, cand, cap, cand_cnt, sp);
        } else {
            dim3 grid((P + 256 * kNmsGroups - 1) / (256 * kNmsGroups), B);
            launch_dependent(nms_candidates_direct_kernel<1>, grid, dim3(256), 0, s, heat, H, W,
                             threshold, r, cand, cap, cand_cnt, sp);
        }
    } else {
        // large window: shared-memory halo tiles bound the cost per pixel
        dim3 grid((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH, B);
        const size_t smem = (size_t)(kTileH + 2 * r) * (kTileW + 2 * r) * sizeof(float);
        launch_dependent(nms_candidates_kernel, grid, dim3(kNmsThreads), smem, s, heat, H, W,
                         threshold, r, cand, cap, cand_cnt, sp);
    }
    return record_launch("npb_instance_centers");
}

extern "C" int npb_instance_centers(const float *heat, int B, int H, int W, float threshold,
                                    int nms_kernel_size, int top_k, const uint8_t *fg,
                                    int apply_fg_mask, void *workspace, int32_t *centers_yx,
                                    int32_t *n_centers, float *center_score, int32_t *status,
                                    void *stream)
{
    return instance_centers_impl(heat, B, H, W, threshold, nms_kernel_size, top_k, fg, apply_fg_mask,
                                 workspace, centers_yx, n_centers, center_score, status, false,
                                 false, nullptr, false, stream);
}

// ---- > 255 centres: the reference's uint8 wrap (instance.py:236) -------------------------------
__global__ void __launch_bounds__(256)
overflow_centers_kernel(const int32_t *wide_n, const unsigned *sorted, const float *heat_frame,
                        int W, int cap, int32_t *n_out, int32_t *centers_yx, float *score)
{
    const int n = *wide_n;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
    if (n < 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n && i < cap; i += gridDim.x * blockDim.x) {
        const unsigned idx = sorted[i];
        centers_yx[2 * i] = (int32_t)(idx / (unsigned)W);
        centers_yx[2 * i + 1] = (int32_t)(idx % (unsigned)W);
        score[i] = heat_frame[idx];
    }
}

extern "C" int npb_overflow_centers(const void *workspace, const float *heat, int B, int H, int W,
                                    int nms_kernel_size, int frame, int32_t *n_out,
                                    int32_t *centers_yx, float *center_score, int cap, void *stream)
{
    if (!workspace || !heat || !n_out || !centers_yx || !center_score) return NPB_ERR_ARG;
    if (B < 1 || frame < 0 || frame >= B || H < 1 || W < 1 || cap < 1) return NPB_ERR_ARG;
    const int ccap = cand_capacity(H, W, nms_kernel_size);
    const size_t off = ((size_t)B * ccap * sizeof(uint2) + 255) & ~(size_t)255;
    const char *counters = (const char *)workspace + off;
    const int32_t *wide_n = (const int32_t *)(counters + centers_counter_bytes(B));
    const unsigned *wide_idx = (const unsigned *)((const char *)wide_n + align256((size_t)B * sizeof(int32_t)));
    overflow_centers_kernel<<<16, 256, 0, (cudaStream_t)stream>>>(
        wide_n + frame, wide_idx + ((size_t)frame * 2 + 1) * kWideCap,
        heat + (size_t)frame * H * W, W, cap, n_out, centers_yx, center_score);
    return record_launch("npb_overflow_centers");
}
