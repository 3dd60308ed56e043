// targets.cu -- instance training / evaluation targets from ground-truth maps (batched)
//
// Replaces InstanceTargetGenerator._preprocess (reference: data/preprocessing/instance.py:151-286),
// the step BEFORE the hot path that produces `instance_center`, `instance_offset`,
// `instance_foreground`, `instance_center_mask` of a sample:
//   :191-212  for every instance id != 0: semantic class = bincount(sem[mask]).argmax() (smallest
//             class on ties); instances whose class is not a thing are skipped
//   :216-221  foreground |= mask; centre = (int(mean(y)), int(mean(x))) of the mask
//   :222-238  centre heat-map = max(heat-map, precomputed Gaussian stamped around the centre)
//   :241-245  offset[mask] = (centre_y - y, centre_x - x) as int16
//   :247-251  optional normalisation: float32(offset) / (H, W)
//   :269-275  centre mask = foreground | isin(sem, stuff classes without void)
// Instance ids are arbitrary uint16 values, so the instances and their (instance, class) parts
// of a frame are collected in per-frame global hash tables (pixel pass 1, warp aggregated),
// resolved by one CTA per frame, stamped, and looked up again by pixel pass 2.
#include "common.cuh"

namespace npb {

constexpr int kTgtPartSlots = 8192;   // (instance, class) parts per frame
constexpr int kTgtInstSlots = 4096;   // instances per frame (<= 3072)
constexpr unsigned kNoKey = 0xffffffffu;

__device__ __forceinline__ unsigned tgt_hash(unsigned h)
{
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ int tgt_slot(unsigned *keys, int slots, unsigned key)
{
    unsigned h = tgt_hash(key) & (unsigned)(slots - 1);
    for (int probe = 0; probe < slots; ++probe) {
        unsigned k = keys[h];
        if (k == kNoKey) k = atomicCAS(keys + h, kNoKey, key);
        if (k == kNoKey || k == key) return (int)h;
        h = (h + 1) & (unsigned)(slots - 1);
    }
    return -1;
}

__device__ __forceinline__ int tgt_find(const unsigned *keys, int slots, unsigned key)
{
    unsigned h = tgt_hash(key) & (unsigned)(slots - 1);
    for (int probe = 0; probe < slots; ++probe) {
        const unsigned k = keys[h];
        if (k == key) return (int)h;
        if (k == kNoKey) return -1;
        h = (h + 1) & (unsigned)(slots - 1);
    }
    return -1;
}

struct TargetTables {
    unsigned *part_key;               // [B][kTgtPartSlots]  instance << 16 | class
    unsigned *part_cnt;               // [B][kTgtPartSlots]
    unsigned *inst_key;               // [B][kTgtInstSlots]  instance id
    unsigned *inst_n;                 // [B][kTgtInstSlots]  pixels
    unsigned long long *inst_sy;      // [B][kTgtInstSlots]  sum of y
    unsigned long long *inst_sx;      // [B][kTgtInstSlots]  sum of x
    unsigned long long *inst_best;    // [B][kTgtInstSlots]  max over parts of count << 16 | ~class
    int *inst_center;                 // [B][kTgtInstSlots]  cy << 16 | cx, -1 = not encoded
};

// pixel pass 1: part counts, per-instance pixel count and coordinate sums
__global__ void __launch_bounds__(256)
target_stats_kernel(const uint8_t *__restrict__ sem, const int32_t *__restrict__ ins, int H, int W,
                    TargetTables t, int32_t *__restrict__ status)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const long long P = (long long)H * W;
    unsigned *pk = t.part_key + (size_t)b * kTgtPartSlots, *pc = t.part_cnt + (size_t)b * kTgtPartSlots;
    unsigned *ik = t.inst_key + (size_t)b * kTgtInstSlots, *in_ = t.inst_n + (size_t)b * kTgtInstSlots;
    unsigned long long *sy = t.inst_sy + (size_t)b * kTgtInstSlots, *sx = t.inst_sx + (size_t)b * kTgtInstSlots;
    const long long stride = (long long)gridDim.x * 256;
    const long long n_round = ((P + 31) / 32) * 32;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n_round; p += stride) {
        unsigned id = 0, c = 0;
        int y = 0, x = 0;
        if (p < P) {
            const size_t q = (size_t)b * P + p;
            const int v = ins[q];
            if (v < 0 || v > 65535) set_status(status, NPB_ERR_CATEGORY_RANGE);
            else id = (unsigned)v;
            c = sem[q];
            y = (int)(p / W);
            x = (int)(p - (long long)y * W);
        }
        const unsigned part = id ? ((id << 16) | c) : kNoKey;
        const unsigned peers_p = __match_any_sync(kFullMask, part);
        const unsigned peers_i = __match_any_sync(kFullMask, id);
        const int ysum = __reduce_add_sync(peers_i, y), xsum = __reduce_add_sync(peers_i, x);
        if (id) {
            if (lane == __ffs(peers_p) - 1) {
                const int s = tgt_slot(pk, kTgtPartSlots, part);
                if (s < 0) set_status(status, NPB_ERR_CAPACITY);
                else atomicAdd(pc + s, (unsigned)__popc(peers_p));
            }
            if (lane == __ffs(peers_i) - 1) {
                const int s = tgt_slot(ik, kTgtInstSlots, id);
                if (s < 0) {
                    set_status(status, NPB_ERR_CAPACITY);
                } else {
                    atomicAdd(in_ + s, (unsigned)__popc(peers_i));
                    atomicAdd(sy + s, (unsigned long long)ysum);
                    atomicAdd(sx + s, (unsigned long long)xsum);
                }
            }
        }
    }
}

// one CTA per frame: majority class per instance, thing test, centre
__global__ void __launch_bounds__(512)
target_instances_kernel(TargetTables t, ClassSet thing, int32_t *__restrict__ enc_ids,
                        int32_t *__restrict__ skip_ids, int32_t *__restrict__ n_enc,
                        int32_t *__restrict__ n_skip, int list_cap)
{
    const int b = blockIdx.x, tid = threadIdx.x;
    const unsigned *pk = t.part_key + (size_t)b * kTgtPartSlots, *pc = t.part_cnt + (size_t)b * kTgtPartSlots;
    const unsigned *ik = t.inst_key + (size_t)b * kTgtInstSlots;
    unsigned long long *best = t.inst_best + (size_t)b * kTgtInstSlots;
    __shared__ int s_enc, s_skip;
    if (tid == 0) { s_enc = 0; s_skip = 0; }
    __syncthreads();
    // most frequent class of every instance: max of (count << 16 | 0xffff - class), so that
    // equal counts resolve to the SMALLEST class like bincount(...).argmax()
    for (int i = tid; i < kTgtPartSlots; i += 512) {
        const unsigned key = pk[i];
        if (key != kNoKey) {
            const int s = tgt_find(ik, kTgtInstSlots, key >> 16);
            if (s >= 0)
                atomicMax(best + s, ((unsigned long long)pc[i] << 16) | (0xffffu - (key & 0xffffu)));
        }
    }
    __threadfence_block();
    __syncthreads();
    for (int s = tid; s < kTgtInstSlots; s += 512) {
        const unsigned id = ik[s];
        int info = -1;
        if (id != kNoKey) {
            const int cls = 0xffff - (int)(best[s] & 0xffffull);
            if (thing.has(cls & 255) && cls < 256) {
                const unsigned n = t.inst_n[(size_t)b * kTgtInstSlots + s];
                const int cy = (int)(t.inst_sy[(size_t)b * kTgtInstSlots + s] / n);   // int(mean)
                const int cx = (int)(t.inst_sx[(size_t)b * kTgtInstSlots + s] / n);
                info = (cy << 16) | cx;
                const int k = atomicAdd(&s_enc, 1);
                if (k < list_cap) enc_ids[(size_t)b * list_cap + k] = (int)id;
            } else {
                const int k = atomicAdd(&s_skip, 1);
                if (k < list_cap) skip_ids[(size_t)b * list_cap + k] = (int)id;
            }
        }
        t.inst_center[(size_t)b * kTgtInstSlots + s] = info;
    }
    __syncthreads();
    if (tid == 0) { n_enc[b] = min(s_enc, list_cap); n_skip[b] = min(s_skip, list_cap); }
}

// Gaussian stamps: one CTA per (slot chunk, frame); heat values are >= 0 so the unsigned bit
// pattern orders like the value and atomicMax on it is an exact, order independent maximum
__global__ void __launch_bounds__(256)
target_stamp_kernel(TargetTables t, const float *__restrict__ gauss, int sigma, int H, int W,
                    float *__restrict__ center)
{
    const int b = blockIdx.y;
    const int size = 6 * sigma + 3;
    for (int s = blockIdx.x; s < kTgtInstSlots; s += gridDim.x) {
        const int info = t.inst_center[(size_t)b * kTgtInstSlots + s];
        if (info < 0) continue;
        const int cy = info >> 16, cx = info & 0xffff;
        const int ul_x = cx - 3 * sigma - 1, ul_y = cy - 3 * sigma - 1;
        for (int i = threadIdx.x; i < size * size; i += 256) {
            const int gy = i / size, gx = i - gy * size;
            const int y = ul_y + gy, x = ul_x + gx;
            if (y >= 0 && y < H && x >= 0 && x < W)
                atomicMax((unsigned *)center + ((size_t)b * H + y) * W + x,
                          __float_as_uint(gauss[i]));
        }
    }
}

// pixel pass 2: offsets, foreground, centre mask
template <bool NORMALIZED>
__global__ void __launch_bounds__(256)
target_write_kernel(const uint8_t *__restrict__ sem, const int32_t *__restrict__ ins, int H, int W,
                    TargetTables t, ClassSet stuff, void *__restrict__ offset_out,
                    uint8_t *__restrict__ fg_out, uint8_t *__restrict__ mask_out,
                    int32_t *__restrict__ status)
{
    const int b = blockIdx.y;
    const long long P = (long long)H * W;
    const unsigned *ik = t.inst_key + (size_t)b * kTgtInstSlots;
    const int *ic = t.inst_center + (size_t)b * kTgtInstSlots;
    const float fH = (float)H, fW = (float)W;
    const long long stride = (long long)gridDim.x * 256;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < P; p += stride) {
        const size_t q = (size_t)b * P + p;
        const int id = ins[q];
        const int c = sem[q];
        int dy = 0, dx = 0;
        bool fg = false;
        if (id > 0 && id <= 65535) {
            const int s = tgt_find(ik, kTgtInstSlots, (unsigned)id);
            const int info = s >= 0 ? ic[s] : -1;
            if (info >= 0) {
                const int y = (int)(p / W), x = (int)(p - (long long)y * W);
                dy = (info >> 16) - y;
                dx = (info & 0xffff) - x;
                fg = true;
            } else {
                // a stuff pixel that still carries an instance id: the reference asserts
                // (instance.py:260) -- InstanceClearStuffIDs has to run first
                set_status(status, NPB_ERR_ARG);
            }
        }
        if (NORMALIZED) {
            float *o = (float *)offset_out + (size_t)b * 2 * P;
            o[p] = __fdiv_rn((float)dy, fH);
            o[P + p] = __fdiv_rn((float)dx, fW);
        } else {
            int16_t *o = (int16_t *)offset_out + (size_t)b * 2 * P;
            o[p] = (int16_t)dy;
            o[P + p] = (int16_t)dx;
        }
        fg_out[q] = fg;
        mask_out[q] = fg || stuff.has(c);
    }
}

static size_t tgt_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace npb

using namespace npb;

extern "C" size_t npb_instance_targets_workspace_bytes(int B)
{
    return tgt_align((size_t)B * kTgtPartSlots * 8) + tgt_align((size_t)B * kTgtInstSlots * 4) * 3 +
           tgt_align((size_t)B * kTgtInstSlots * 8) * 3;
}

extern "C" int npb_instance_targets(const uint8_t *sem, const int32_t *ins, int B, int H, int W,
                                    const uint8_t *h_thing_lut, int n_classes, int sigma,
                                    const float *gauss, int normalized_offset, void *workspace,
                                    float *center_out, void *offset_out, uint8_t *fg_out,
                                    uint8_t *center_mask_out, int32_t *encoded_ids,
                                    int32_t *skipped_ids, int32_t *n_encoded, int32_t *n_skipped,
                                    int list_cap, int32_t *status, void *stream)
{
    if (!sem || !ins || !h_thing_lut || !gauss || !workspace || !center_out || !offset_out ||
        !fg_out || !center_mask_out || !encoded_ids || !skipped_ids || !n_encoded || !n_skipped ||
        !status)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || H < 1 || W < 1 || H > 32767 || W > 32767 || n_classes < 1 ||
        n_classes > 256 || sigma < 1 || sigma > 64 || list_cap < 1)
        return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    TargetTables t;
    t.part_key = (unsigned *)ws;
    t.part_cnt = t.part_key + (size_t)B * kTgtPartSlots;
    ws += tgt_align((size_t)B * kTgtPartSlots * 8);
    t.inst_key = (unsigned *)ws; ws += tgt_align((size_t)B * kTgtInstSlots * 4);
    t.inst_n = (unsigned *)ws; ws += tgt_align((size_t)B * kTgtInstSlots * 4);
    t.inst_center = (int *)ws; ws += tgt_align((size_t)B * kTgtInstSlots * 4);
    t.inst_sy = (unsigned long long *)ws; ws += tgt_align((size_t)B * kTgtInstSlots * 8);
    t.inst_sx = (unsigned long long *)ws; ws += tgt_align((size_t)B * kTgtInstSlots * 8);
    t.inst_best = (unsigned long long *)ws;
    cudaMemsetAsync(t.part_key, 0xff, (size_t)B * kTgtPartSlots * 4, s);
    cudaMemsetAsync(t.part_cnt, 0, (size_t)B * kTgtPartSlots * 4, s);
    cudaMemsetAsync(t.inst_key, 0xff, (size_t)B * kTgtInstSlots * 4, s);
    cudaMemsetAsync(t.inst_n, 0, (size_t)B * kTgtInstSlots * 4, s);
    cudaMemsetAsync(t.inst_sy, 0, (size_t)B * kTgtInstSlots * 8, s);
    cudaMemsetAsync(t.inst_sx, 0, (size_t)B * kTgtInstSlots * 8, s);
    cudaMemsetAsync(t.inst_best, 0, (size_t)B * kTgtInstSlots * 8, s);
    cudaMemsetAsync(center_out, 0, (size_t)B * H * W * sizeof(float), s);

    // thing classes WITH void (index = semantic label); stuff = not thing, void removed
    uint8_t stuff_lut[256] = {0};
    for (int c = 1; c < n_classes; ++c) stuff_lut[c] = h_thing_lut[c] ? 0 : 1;
    const ClassSet thing = make_class_set(h_thing_lut, n_classes);
    const ClassSet stuff = make_class_set(stuff_lut, n_classes);

    const long long P = (long long)H * W;
    long long bx = (P + 256 * 8 - 1) / (256 * 8);
    if (bx > 148 * 8 / B + 1) bx = 148 * 8 / B + 1;
    dim3 grid((unsigned)bx, B);
    target_stats_kernel<<<grid, 256, 0, s>>>(sem, ins, H, W, t, status);
    target_instances_kernel<<<B, 512, 0, s>>>(t, thing, encoded_ids, skipped_ids, n_encoded,
                                              n_skipped, list_cap);
    target_stamp_kernel<<<dim3(64, B), 256, 0, s>>>(t, gauss, sigma, H, W, center_out);
    if (normalized_offset)
        target_write_kernel<true><<<grid, 256, 0, s>>>(sem, ins, H, W, t, stuff, offset_out, fg_out,
                                                       center_mask_out, status);
    else
        target_write_kernel<false><<<grid, 256, 0, s>>>(sem, ins, H, W, t, stuff, offset_out, fg_out,
                                                        center_mask_out, status);
    return record_launch("npb_instance_targets");
}
