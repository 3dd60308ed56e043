// finalize.cuh -- per-frame instance table from the vote histograms (shared by the stand-alone
// finalize_instances_kernel, merge.cu, and the fused id writer + evaluation pass, eval.cu)
//
// Replaces (reference): utils/panoptic_merge.py:192-210 -- instances in ascending id: class =
// torch.mode(sem[mask]) (smallest value on ties), skip class 0, running number per class,
// pan id = class * L + number -- plus the meta areas (instance.py:253) and the orientation angle
// (instance.py:310-313, utils/_orientation.py:39-42).
#pragma once
#include <math.h>

#include "common.cuh"

namespace npb {

struct FinalizeParams {
    const uint32_t *vote_hist;      // [B][kMaxInst][C]
    const double *ori_sum;          // [B][kMaxInst][2] nullable
    const int32_t *n_centers;       // [B] nullable (then every row is an instance)
    int C, class_offset;
    long long L, void_label;
    ClassSet orient;                // indexed by PANOPTIC class
    int32_t *inst_class;            // [B][kMaxInst]
    int64_t *inst_pan_id;
    int32_t *inst_area;
    float *inst_angle;              // nullable
};

// Thread i of a CTA with kMaxInst threads <-> raw instance id i of frame b.  Returns the
// panoptic id of the instance (void_label for a dropped / absent one); `s_cls` is a
// [kMaxInst] shared scratch; with `write` the tables of the frame are stored.
// Contains one __syncthreads(): all kMaxInst threads call it.
__device__ __forceinline__ long long finalize_frame(const FinalizeParams &f, int b, int i,
                                                    int *s_cls, bool write)
{
    const int C = f.C;
    const int n = f.n_centers ? f.n_centers[b] : kMaxInst - 1;
    const uint32_t *row = f.vote_hist + ((size_t)b * kMaxInst + i) * C;

    int cls = -1;
    uint32_t best = 0, area = 0;
    if (i >= 1 && i <= n) {
        // batches of 16 independent loads, then the ordered compares: a frame has a dozen
        // instances, so this is a handful of threads and every batch is one memory latency on
        // the critical path of the CTA
        constexpr int U = 16;
        for (int c0 = 0; c0 < C; c0 += U) {
            uint32_t h[U];
#pragma unroll
            for (int u = 0; u < U; ++u) h[u] = (c0 + u < C) ? row[c0 + u] : 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                area += h[u];
                if (h[u] > best) { best = h[u]; cls = c0 + u; }    // strict > : smallest class wins ties
            }
        }
    }
    int pcls = (cls >= 0) ? cls + f.class_offset : -1;
    if (pcls == 0) pcls = -1;  // majority is void -> instance dropped (panoptic_merge.py:201)
    s_cls[i] = pcls;
    __syncthreads();
    int number = 1;
    if (pcls >= 0)
        for (int j = 1; j < i; ++j) number += (s_cls[j] == pcls);
    const long long pan = (pcls >= 0) ? (long long)pcls * f.L + number : f.void_label;
    if (write) {
        const size_t o = (size_t)b * kMaxInst + i;
        f.inst_class[o] = pcls;
        f.inst_pan_id[o] = pan;
        f.inst_area[o] = (int32_t)area;
        if (f.inst_angle) {
            float ang = nanf("");
            if (f.ori_sum && pcls >= 0 && f.orient.has(pcls)) {
                const double sc = f.ori_sum[2 * o], ss = f.ori_sum[2 * o + 1];
                // the reference forms f32 sums and calls atan2 on them (instance.py:310-313)
                ang = (float)atan2((double)(float)ss, (double)(float)sc);
            }
            f.inst_angle[o] = ang;
        }
    }
    return pan;
}

// host side: the orientation set is indexed by PANOPTIC class (network class + class_offset)
inline ClassSet orientation_class_set(const uint8_t *h_orientation_lut, int C, int class_offset)
{
    uint8_t lut[256] = {0};
    if (h_orientation_lut)
        for (int c = 0; c < C && c + class_offset < 256; ++c)
            if (c + class_offset >= 0) lut[c + class_offset] = h_orientation_lut[c];
    return make_class_set(lut, 256);
}

}  // namespace npb
