// merge.cu -- per-frame instance table + panoptic id map (the "deeplab merge")
//
// Replaces deeplab_merge_semantic_and_instance (reference: utils/panoptic_merge.py:172-225):
//   :192-210  instances in ascending id: class = torch.mode(sem[mask]) (smallest value on
//             ties), skip class 0, running number per class, pan id = class * L + number
//   :213-223  every non-thing, non-void class: pan[(sem == c) & (ins == 0)] = c * L
// plus the meta areas (instance.py:253) and the orientation angle (instance.py:313,
// utils/_orientation.py:39-42).  The per-pixel histogram half lives in group.cu (fused path)
// or in merge_votes_kernel below (stand-alone API).
#include "finalize.cuh"

namespace npb {

// one CTA per frame, thread i <-> raw instance id i (finalize.cuh)
__global__ void __launch_bounds__(kMaxInst) finalize_instances_kernel(const FinalizeParams f)
{
    __shared__ int s_cls[kMaxInst];
    grid_dependency_wait();
    finalize_frame(f, blockIdx.x, threadIdx.x, s_cls, true);
}

// 4 pixels per group (32-bit loads of the two uint8 maps, two 128-bit streaming stores of
// int64 ids), kWriteGroups independent groups per thread so that enough bytes are in flight
constexpr int kWriteGroups = 4;

template <int VEC>
__global__ void __launch_bounds__(256)
write_panoptic_kernel(const uint8_t *sem, const uint8_t *inst, const int64_t *inst_pan_id,
                      const int32_t *inst_class, int P, long long L, ClassSet thing,
                      int64_t *pan_out, uint8_t *pan_sem_out)
{
    // NO `const __restrict__` on inputs that a predecessor of the chain produces: such loads
    // become LDG.CONSTANT (ld.global.nc), which the compiler is free to hoist above
    // grid_dependency_wait() -- the "memory" clobber does not order read-only loads (seen in
    // the SASS of this kernel: the table loads sat above ACQBULK and read stale rows).
    __shared__ long long s_pan[kMaxInst];
    __shared__ int s_cls[kMaxInst];
    const int b = blockIdx.y;
    grid_dependency_wait();
    s_pan[threadIdx.x] = inst_pan_id[(size_t)b * kMaxInst + threadIdx.x];
    if (pan_sem_out) {
        const int c = inst_class[(size_t)b * kMaxInst + threadIdx.x];
        s_cls[threadIdx.x] = c < 0 ? 0 : c;  // dropped instance -> void
    }
    __syncthreads();
    const int base = blockIdx.x * (256 * VEC * kWriteGroups) + threadIdx.x * VEC;
    uint32_t sw[kWriteGroups], iw[kWriteGroups];
#pragma unroll
    for (int u = 0; u < kWriteGroups; ++u) {
        const int p0 = base + u * 256 * VEC;
        sw[u] = 0; iw[u] = 0;
        if (p0 < P) {
            const size_t fb = (size_t)b * P + p0;
            if (VEC == 4) {
                sw[u] = *(const uint32_t *)(sem + fb);
                iw[u] = *(const uint32_t *)(inst + fb);
            } else {
                sw[u] = sem[fb];
                iw[u] = inst[fb];
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kWriteGroups; ++u) {
        const int p0 = base + u * 256 * VEC;
        if (p0 >= P) continue;
        const size_t fb = (size_t)b * P + p0;
        long long out[VEC];
        uint32_t psw = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int c = (sw[u] >> (8 * j)) & 255, ii = (iw[u] >> (8 * j)) & 255;
            long long v;
            int pc;  // panoptic class = v / L without the 64-bit division
            if (ii > 0) {
                v = s_pan[ii];
                pc = pan_sem_out ? s_cls[ii] : 0;
            } else {
                pc = thing.has(c) ? 0 : c + 1;
                v = (long long)pc * L;
            }
            out[j] = v;
            psw |= (uint32_t)pc << (8 * j);
        }
        if (VEC == 4) {
            longlong2 *o = (longlong2 *)(pan_out + fb);
            __stcs(o, make_longlong2(out[0], out[1 % VEC]));
            __stcs(o + 1, make_longlong2(out[2 % VEC], out[3 % VEC]));
            if (pan_sem_out) *(uint32_t *)(pan_sem_out + fb) = psw;
        } else {
            pan_out[fb] = out[0];
            if (pan_sem_out) pan_sem_out[fb] = (uint8_t)psw;
        }
    }
}

// ---- stand-alone merge: arbitrary semantic / instance / foreground maps -----------------
__global__ void __launch_bounds__(256)
merge_votes_kernel(const int64_t *__restrict__ sem, const uint8_t *__restrict__ ins,
                   const uint8_t *__restrict__ fg, long long P, int n_classes,
                   uint32_t *__restrict__ vote_hist, int32_t *__restrict__ status)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    uint32_t *hist = vote_hist + (size_t)b * kMaxInst * n_classes;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < ((P + 31) / 32) * 32;
         p += (long long)gridDim.x * 256) {
        int key = -1;
        if (p < P) {
            const size_t q = (size_t)b * P + p;
            const long long c = sem[q];
            if (c < 0 || c >= n_classes) {
                set_status(status, NPB_ERR_CATEGORY_RANGE);
            } else if (ins[q] > 0 && fg[q]) {
                key = ((int)ins[q]) * n_classes + (int)c;
            }
        }
        // warp aggregation over equal keys
        unsigned pending = __ballot_sync(kFullMask, key >= 0);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int cur = __shfl_sync(kFullMask, key, leader);
            const unsigned same = __ballot_sync(kFullMask, key == cur);
            if (lane == leader) atomicAdd(hist + cur, (uint32_t)__popc(same));
            pending &= ~same;
        }
    }
}

__global__ void __launch_bounds__(256)
merge_write_kernel(const int64_t *__restrict__ sem, const uint8_t *__restrict__ ins,
                   const uint8_t *__restrict__ fg, const int64_t *__restrict__ inst_pan_id,
                   long long P, int n_classes, long long L, long long void_label,
                   const uint8_t *__restrict__ thing_lut, int64_t *__restrict__ pan_out)
{
    __shared__ long long s_pan[kMaxInst];
    const int b = blockIdx.y;
    s_pan[threadIdx.x] = inst_pan_id[(size_t)b * kMaxInst + threadIdx.x];
    __syncthreads();
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < P;
         p += (long long)gridDim.x * 256) {
        const size_t q = (size_t)b * P + p;
        const long long c = sem[q];
        const int ii = ins[q];
        long long v = void_label;
        if (ii > 0) {
            if (fg[q]) v = s_pan[ii];
        } else if (c > 0 && c < n_classes && !thing_lut[c]) {
            v = c * L;
        }
        pan_out[q] = v;
    }
}

}  // namespace npb

using namespace npb;

int npb::launch_finalize(const FinalizeParams &f, int B, void *stream)
{
    launch_dependent(finalize_instances_kernel, dim3(B), dim3(kMaxInst), 0, (cudaStream_t)stream, f);
    return record_launch("npb_finalize_instances");
}

extern "C" int npb_finalize_instances(const uint32_t *vote_hist, const double *ori_sum,
                                      const int32_t *n_centers, int B, int C, int class_offset,
                                      int64_t max_instances_per_category, int64_t void_label,
                                      const uint8_t *h_orientation_lut, int32_t *inst_class,
                                      int64_t *inst_pan_id, int32_t *inst_area,
                                      float *inst_angle, void *stream)
{
    if (!vote_hist || !inst_class || !inst_pan_id || !inst_area) return NPB_ERR_ARG;
    if (B < 1 || C < 1 || C > 65536 || max_instances_per_category < 1) return NPB_ERR_ARG;
    FinalizeParams f;
    f.vote_hist = vote_hist; f.ori_sum = ori_sum; f.n_centers = n_centers; f.C = C;
    f.class_offset = class_offset; f.L = (long long)max_instances_per_category;
    f.void_label = (long long)void_label;
    f.orient = orientation_class_set(h_orientation_lut, C, class_offset);
    f.inst_class = inst_class; f.inst_pan_id = inst_pan_id; f.inst_area = inst_area;
    f.inst_angle = inst_angle;
    return launch_finalize(f, B, stream);
}

extern "C" int npb_write_panoptic(const uint8_t *sem, const uint8_t *inst,
                                  const int64_t *inst_pan_id, const int32_t *inst_class, int B,
                                  int C, int H, int W,
                                  const uint8_t *h_thing_lut, int64_t max_instances_per_category,
                                  int64_t *pan_out, uint8_t *pan_sem_out, void *stream)
{
    if (!sem || !inst || !inst_pan_id || !pan_out || !h_thing_lut) return NPB_ERR_ARG;
    if (pan_sem_out && !inst_class) return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || C < 1 || C > 255 || max_instances_per_category < 1)
        return NPB_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return NPB_ERR_ARG;
    const int P = H * W;
    const ClassSet thing = make_class_set(h_thing_lut, C);
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec4 = (P % 4 == 0) &&
                      (((uintptr_t)sem | (uintptr_t)inst | (uintptr_t)pan_sem_out) & 3u) == 0 &&
                      ((uintptr_t)pan_out & 15u) == 0;
    if (vec4) {
        dim3 grid((P / 4 + 256 * kWriteGroups - 1) / (256 * kWriteGroups), B);
        launch_dependent(write_panoptic_kernel<4>, grid, dim3(256), 0, s, sem, inst, inst_pan_id,
                         inst_class, P, (long long)max_instances_per_category, thing, pan_out,
                         pan_sem_out);
    } else {
        dim3 grid((P + 256 * kWriteGroups - 1) / (256 * kWriteGroups), B);
        launch_dependent(write_panoptic_kernel<1>, grid, dim3(256), 0, s, sem, inst, inst_pan_id,
                         inst_class, P, (long long)max_instances_per_category, thing, pan_out,
                         pan_sem_out);
    }
    return record_launch("npb_write_panoptic");
}

extern "C" size_t npb_deeplab_merge_workspace_bytes(int B, int n_classes)
{
    size_t bytes = (size_t)B * kMaxInst * n_classes * sizeof(uint32_t);
    bytes = (bytes + 255) & ~(size_t)255;
    bytes += ((size_t)n_classes + 255) & ~(size_t)255;  // device copy of the thing lut
    return bytes;
}

extern "C" int npb_deeplab_merge(const int64_t *sem, const uint8_t *ins, const uint8_t *fg, int B,
                                 int64_t P, int n_classes, int64_t max_instances_per_category,
                                 const uint8_t *h_thing_lut, int64_t void_label, void *workspace,
                                 int64_t *pan_out, int32_t *inst_class, int64_t *inst_pan_id,
                                 int32_t *inst_area, int32_t *status, void *stream)
{
    if (!sem || !ins || !fg || !workspace || !pan_out || !inst_class || !inst_pan_id ||
        !inst_area || !status || !h_thing_lut)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || P < 1 || n_classes < 1 || n_classes > 65536) return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *hist = (uint32_t *)workspace;
    const size_t hist_bytes = (size_t)B * kMaxInst * n_classes * sizeof(uint32_t);
    uint8_t *d_lut = (uint8_t *)workspace + ((hist_bytes + 255) & ~(size_t)255);
    cudaMemsetAsync(hist, 0, hist_bytes, s);
    // pageable host source: the copy is staged before the call returns
    cudaMemcpyAsync(d_lut, h_thing_lut, (size_t)n_classes, cudaMemcpyHostToDevice, s);
    const int bx = (int)((P + 256 * 8 - 1) / (256 * 8));
    dim3 grid(bx < 1 ? 1 : bx, B);
    merge_votes_kernel<<<grid, 256, 0, s>>>(sem, ins, fg, (long long)P, n_classes, hist, status);
    FinalizeParams f;
    f.vote_hist = hist; f.ori_sum = nullptr; f.n_centers = nullptr; f.C = n_classes;
    f.class_offset = 0; f.L = (long long)max_instances_per_category;
    f.void_label = (long long)void_label; f.orient = make_class_set(nullptr, 0);
    f.inst_class = inst_class; f.inst_pan_id = inst_pan_id; f.inst_area = inst_area;
    f.inst_angle = nullptr;
    finalize_instances_kernel<<<B, kMaxInst, 0, s>>>(f);
    merge_write_kernel<<<grid, 256, 0, s>>>(sem, ins, fg, inst_pan_id, (long long)P, n_classes,
                                            (long long)max_instances_per_category,
                                            (long long)void_label, d_lut, pan_out);
    return record_launch("npb_deeplab_merge");
}

// ---- naive merge: ground-truth panoptic targets ----------------------------------------------
// Replaces naive_merge_semantic_and_instance_np (reference: utils/panoptic_merge.py:43-107), the
// body of PanopticTargetGenerator (data/preprocessing/panoptic.py:16-85):
//   :64-90   for instance ids ascending, for the semantic classes present inside the instance
//            ascending (void skipped): number = ++counter[class]; pan id = class * L + number;
//            every (instance, class) part gets its own id; id_dict[pan id] = instance id
//   :93-105  every non-void, non-thing class c: pan[(sem == c) & (ins == 0)] = c * L
// Instance ids are arbitrary (uint16 in the datasets), so the (instance, class) parts of a frame
// are collected in a per-frame hash table (pixel pass 1), sorted and numbered by one CTA per
// frame, and looked up again by pixel pass 2.
namespace npb {

constexpr int kPartSlots = 8192;     // (instance, class) parts per frame: <= 4096
constexpr int kMaxParts = 4096;

__device__ __forceinline__ unsigned hash32(unsigned h)
{
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

// sem (B,P) u8 classes, ins (B,P) i32 ids in [0, 65535]; part key = ins << 16 | class
__global__ void __launch_bounds__(256)
naive_parts_kernel(const uint8_t *__restrict__ sem, const int32_t *__restrict__ ins, long long P,
                   unsigned *__restrict__ part_keys, int32_t *__restrict__ status)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    unsigned *keys = part_keys + (size_t)b * kPartSlots;
    const long long stride = (long long)gridDim.x * 256;
    const long long n_round = ((P + 31) / 32) * 32;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n_round; p += stride) {
        unsigned key = 0xffffffffu;
        if (p < P) {
            const size_t q = (size_t)b * P + p;
            const int id = ins[q];
            const unsigned c = sem[q];
            if (id < 0 || id > 65535) set_status(status, NPB_ERR_CATEGORY_RANGE);
            else if (id != 0 && c != 0) key = ((unsigned)id << 16) | c;
        }
        // one insert per distinct key of the warp
        const unsigned peers = __match_any_sync(kFullMask, key);
        if (key != 0xffffffffu && lane == __ffs(peers) - 1) {
            unsigned h = hash32(key) & (kPartSlots - 1);
            bool done = false;
            for (int probe = 0; probe < kPartSlots && !done; ++probe) {
                unsigned k = keys[h];
                if (k == 0xffffffffu) k = atomicCAS(keys + h, 0xffffffffu, key);
                done = (k == 0xffffffffu || k == key);
                h = (h + 1) & (kPartSlots - 1);
            }
            if (!done) set_status(status, NPB_ERR_CAPACITY);
        }
    }
}

// one CTA per frame: sort the parts (instance major, class minor), number them per class in
// that order; part_keys is rewritten as a sorted list, part_pan holds the ids, n_parts the count
__global__ void __launch_bounds__(512)
naive_number_kernel(unsigned *__restrict__ part_keys, long long *__restrict__ part_pan,
                    int32_t *__restrict__ n_parts, long long L, int32_t *__restrict__ status)
{
    __shared__ unsigned s_key[kMaxParts];
    __shared__ int s_m;
    const int b = blockIdx.x, tid = threadIdx.x;
    unsigned *keys = part_keys + (size_t)b * kPartSlots;
    if (tid == 0) s_m = 0;
    __syncthreads();
    for (int i = tid; i < kPartSlots; i += 512) {
        const unsigned k = keys[i];
        if (k != 0xffffffffu) {
            const int slot = atomicAdd(&s_m, 1);
            if (slot < kMaxParts) s_key[slot] = k;
        }
    }
    __syncthreads();
    int m = s_m;
    if (m > kMaxParts) {
        if (tid == 0) set_status(status, NPB_ERR_CAPACITY);
        m = kMaxParts;
    }
    int npad = 1;
    while (npad < m) npad <<= 1;
    for (int i = m + tid; i < npad; i += 512) s_key[i] = 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += 512) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = ((i & k) == 0);
                    const unsigned a = s_key[i], c = s_key[ixj];
                    if ((a > c) == up) { s_key[i] = c; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < kPartSlots; t += 512) keys[t] = t < m ? s_key[t] : 0xffffffffu;
    for (int t = tid; t < m; t += 512) {
        const unsigned cls = s_key[t] & 0xffffu;
        int number = 1;
        for (int u = 0; u < t; ++u) number += ((s_key[u] & 0xffffu) == cls);
        part_pan[(size_t)b * kMaxParts + t] = (long long)cls * L + number;
    }
    if (tid == 0) n_parts[b] = m;
}

__global__ void __launch_bounds__(256)
naive_write_kernel(const uint8_t *__restrict__ sem, const int32_t *__restrict__ ins, long long P,
                   const unsigned *__restrict__ part_keys, const long long *__restrict__ part_pan,
                   const int32_t *__restrict__ n_parts, long long L, long long void_label,
                   ClassSet thing, int64_t *__restrict__ pan_out)
{
    const int b = blockIdx.y;
    const unsigned *keys = part_keys + (size_t)b * kPartSlots;      // sorted, n_parts[b] entries
    const long long *pans = part_pan + (size_t)b * kMaxParts;
    const int m = n_parts[b];
    const long long stride = (long long)gridDim.x * 256;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < P; p += stride) {
        const size_t q = (size_t)b * P + p;
        const int id = ins[q];
        const unsigned c = sem[q];
        long long v = void_label;
        if (id != 0) {
            if (c != 0) {
                const unsigned key = ((unsigned)id << 16) | c;
                int lo = 0, hi = m - 1;
                while (lo < hi) {               // the part exists by construction
                    const int mid = (lo + hi) >> 1;
                    if (keys[mid] < key) lo = mid + 1; else hi = mid;
                }
                v = pans[lo];
            }
        } else if (c != 0 && !thing.has((int)c)) {
            v = (long long)c * L;
        }
        pan_out[q] = v;
    }
}

}  // namespace npb

extern "C" size_t npb_naive_merge_workspace_bytes(int B)
{
    return ((size_t)B * npb::kPartSlots * sizeof(unsigned) + 255) & ~(size_t)255;
}

extern "C" int npb_naive_merge(const uint8_t *sem, const int32_t *ins, int B, int64_t P,
                               int64_t max_instances_per_category, const uint8_t *h_thing_lut,
                               int n_classes, int64_t void_label, void *workspace, int64_t *pan_out,
                               uint32_t *part_keys_out, int64_t *part_pan_out, int32_t *n_parts,
                               int32_t *status, void *stream)
{
    using namespace npb;
    if (!sem || !ins || !workspace || !pan_out || !part_keys_out || !part_pan_out || !n_parts ||
        !status || !h_thing_lut)
        return NPB_ERR_ARG;
    if (B < 1 || B > 65535 || P < 1 || n_classes < 1 || n_classes > 256 ||
        max_instances_per_category < 1)
        return NPB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned *keys = (unsigned *)workspace;
    cudaMemsetAsync(keys, 0xff, (size_t)B * kPartSlots * sizeof(unsigned), s);
    long long bx = (P + 256 * 8 - 1) / (256 * 8);
    if (bx > 148 * 8 / B + 1) bx = 148 * 8 / B + 1;
    dim3 grid((unsigned)bx, B);
    naive_parts_kernel<<<grid, 256, 0, s>>>(sem, ins, (long long)P, keys, status);
    naive_number_kernel<<<B, 512, 0, s>>>(keys, (long long *)part_pan_out, n_parts,
                                          (long long)max_instances_per_category, status);
    naive_write_kernel<<<grid, 256, 0, s>>>(sem, ins, (long long)P, keys,
                                            (const long long *)part_pan_out, n_parts,
                                            (long long)max_instances_per_category,
                                            (long long)void_label,
                                            make_class_set(h_thing_lut, n_classes), pan_out);
    // sorted keys (instance << 16 | class) of the first kMaxParts slots are the id-dict source
    cudaMemcpy2DAsync(part_keys_out, (size_t)kMaxParts * sizeof(unsigned), keys,
                      (size_t)kPartSlots * sizeof(unsigned), (size_t)kMaxParts * sizeof(unsigned),
                      (size_t)B, cudaMemcpyDeviceToDevice, s);
    return record_launch("npb_naive_merge");
}
