/*
 * panoptic_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the dense panoptic post-processing + evaluation hot
 * path of TUI-NICR/nicr-multitask-scene-analysis v0.3.0.  It exists so that
 * the CUDA path can be checked bit-for-bit on machines where the (Python)
 * reference is not available.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product package never does.
 *
 * Parity is PINNED: the .npz fixtures under tests/golden/ were produced by importing the
 * unmodified reference (tests/golden/make_golden.py) and
 * tests/test_oracle_golden.py checks every function here against them.
 *
 * All file:line citations are relative to /root/reference/src/nicr_mt_scene_analysis/.
 * The arithmetic lives in PyTorch ATen (torch 2.11 CPU kernels); where the
 * rounding sequence matters it is spelled out (compile with -ffp-contract=off).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp (see oracle/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_ARG -1
#define ORC_ERR_TOO_MANY_CENTERS -2
#define ORC_ERR_ZERO_DIVISION -3
#define ORC_ERR_CATEGORY_RANGE -4
#define ORC_ERR_CAPACITY -5

/* ------------------------------------------------------------------------ */
/* a1  semantic arg-max                                                       */
/* model/postprocessing/semantic.py:52-53  (softmax(dim=1) then max(dim=1)).  */
/* torch.max returns the FIRST maximal index; arg-max of softmax(x) equals    */
/* first-index arg-max of x except when two logits differ by < 2^-23 (the     */
/* softmax normalisation can merge adjacent floats), a   */
/* case the parity inputs exclude by quantising logits.                       */
/* logits: [B][C][P] float32 (NCHW with P = H*W), out: [B][P] uint8           */
/* ------------------------------------------------------------------------ */
int orc_semantic_argmax(const float *logits, int B, int C, long P, uint8_t *out)
{
    if (C < 1 || C > 256) return ORC_ERR_ARG;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float *lb = logits + (size_t)b * C * P;
        uint8_t *ob = out + (size_t)b * P;
        for (long p = 0; p < P; ++p) {
            float best = lb[p];
            int arg = 0;
            /* non-finite logits (semantic.py:52-53): softmax is NaN in every class as soon as
               a logit is NaN or +Inf (inf - inf) or all are -Inf, and torch.max of an all-NaN
               row answers index 0; -Inf next to finite logits just has probability 0 */
            int poisoned = isnan(best) || (isinf(best) && best > 0);
            for (int c = 1; c < C; ++c) {
                float v = lb[(size_t)c * P + p];
                if (v > best) { best = v; arg = c; }
                poisoned |= isnan(v) || (isinf(v) && v > 0);
            }
            if (poisoned) arg = 0;
            ob[p] = (uint8_t)arg;
        }
    }
    return ORC_OK;
}

/* soft-max probability of the winning class (semantic.py:52-53 'score'),    */
/* tolerance-checked only (1e-5 rel): exp(x-max)/sum exp(x-max)              */
int orc_semantic_score(const float *logits, int B, int C, long P, float *score)
{
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float *lb = logits + (size_t)b * C * P;
        for (long p = 0; p < P; ++p) {
            float mx = lb[p];
            for (int c = 1; c < C; ++c) {
                float v = lb[(size_t)c * P + p];
                if (v > mx) mx = v;
            }
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += exp((double)lb[(size_t)c * P + p] - (double)mx);
            score[(size_t)b * P + p] = (float)(1.0 / s);
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* a2  centre heat-map NMS   model/postprocessing/instance.py:78-131          */
/*  :86-88   h = (x > thr) ? x : -1            (F.threshold, strict >)        */
/*  :97-109  ks x ks stride-1 VALID max-pool with indices, zero padded back;  */
/*           ATen keeps the FIRST maximum in row-major window order           */
/*  :123-127 pixel dies unless the pooled index is its own flat index         */
/*  :129     pixel dies unless value == pooled value                          */
/* Border ring of width (ks-1)/2: pooled value/index are the zero padding, so */
/* every border pixel dies -- except pixel (0,0), whose flat index 0 equals   */
/* the padded index; it survives :127 and survives :129 only if its           */
/* thresholded value is exactly 0.0 (possible only with thr < 0).             */
/* m: [H][W] float32 post-NMS map (non-survivors = -1).                       */
/* ------------------------------------------------------------------------ */
static void nms_map(const float *heat, int H, int W, float thr, int ks, float *m)
{
    const int r = (ks - 1) / 2;
    for (long i = 0; i < (long)H * W; ++i) m[i] = -1.0f;
    for (int y = r; y < H - r; ++y) {
        for (int x = r; x < W - r; ++x) {
            float maxv = -INFINITY;
            long arg = (long)(y - r) * W + (x - r);
            for (int dy = -r; dy <= r; ++dy) {
                for (int dx = -r; dx <= r; ++dx) {
                    long q = (long)(y + dy) * W + (x + dx);
                    float v = heat[q];
                    v = (v > thr) ? v : -1.0f;
                    if (v > maxv) { maxv = v; arg = q; }
                }
            }
            long self = (long)y * W + x;
            if (arg == self) m[self] = maxv;
        }
    }
    if (r > 0) {
        float v0 = (heat[0] > thr) ? heat[0] : -1.0f;
        if (v0 == 0.0f) m[0] = v0;
    }
}

/* k-th largest value of m[0..n) (torch.topk(...)[..., -1], instance.py:133) */
static int cmp_float_desc(const void *a, const void *b)
{
    float fa = *(const float *)a, fb = *(const float *)b;
    return (fa < fb) - (fa > fb);
}

static float kth_largest(const float *m, long n, int k)
{
    /* survivors are sparse: gather everything above the -1 fill first */
    long cnt = 0;
    for (long i = 0; i < n; ++i) cnt += (m[i] > -1.0f);
    long below = 0; /* values <= -1 that are not exactly the fill are handled by full sort */
    for (long i = 0; i < n; ++i) below += (m[i] < -1.0f);
    if (below == 0) {
        if (cnt < k) return -1.0f;
        float *tmp = (float *)malloc(sizeof(float) * (size_t)cnt);
        long j = 0;
        for (long i = 0; i < n; ++i) if (m[i] > -1.0f) tmp[j++] = m[i];
        qsort(tmp, (size_t)cnt, sizeof(float), cmp_float_desc);
        float v = tmp[k - 1];
        free(tmp);
        return v;
    }
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(tmp, m, sizeof(float) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(float), cmp_float_desc);
    float v = tmp[k - 1];
    free(tmp);
    return v;
}

/* instance.py:131-166 for one frame: top-k value, clamp(min=0), optional     */
/* foreground masking AFTER top-k (:142-143), '>=' (ties give > k centres),   */
/* nonzero() -> raster (y, x) order, int32.                                   */
/* returns number of centres, or a negative error                             */
static int frame_centers(const float *heat, int H, int W, float thr, int ks, int topk,
                         const uint8_t *fg_or_null, int32_t *centers_yx, int cap,
                         uint8_t *center_mask_or_null)
{
    long P = (long)H * W;
    if (topk > P) return ORC_ERR_ARG;
    float *m = (float *)malloc(sizeof(float) * (size_t)P);
    nms_map(heat, H, W, thr, ks, m);
    float kth = kth_largest(m, P, topk);
    if (kth < 0.0f) kth = 0.0f;                       /* clamp_(min=0)   :149 */
    if (fg_or_null)                                   /* :142-143            */
        for (long i = 0; i < P; ++i) if (!fg_or_null[i]) m[i] = -1.0f;
    int n = 0;
    for (long i = 0; i < P; ++i) {
        int is_c = (m[i] >= kth);                     /* :152-155            */
        if (center_mask_or_null) center_mask_or_null[i] = (uint8_t)is_c;
        if (is_c) {
            if (n < cap) {
                centers_yx[2 * n] = (int32_t)(i / W);
                centers_yx[2 * n + 1] = (int32_t)(i % W);
            }
            ++n;
        }
    }
    free(m);
    return n;
}

/* heat: [B][H][W]; fg (nullable, only used when apply_fg_mask): [B][H][W] u8 */
/* centers: [B][cap][2] int32 (y, x); n_out: [B]; center_mask nullable [B][P] */
int orc_instance_centers(const float *heat, int B, int H, int W, float thr, int ks,
                         int topk, const uint8_t *fg, int apply_fg_mask,
                         int32_t *centers, int cap, int32_t *n_out, uint8_t *center_mask)
{
    if ((ks & 1) == 0) return ORC_ERR_ARG;
    long P = (long)H * W;
    int status = ORC_OK;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        int n = frame_centers(heat + (size_t)b * P, H, W, thr, ks, topk,
                              (apply_fg_mask && fg) ? fg + (size_t)b * P : NULL,
                              centers + (size_t)b * cap * 2, cap,
                              center_mask ? center_mask + (size_t)b * P : NULL);
        if (n < 0 || n > cap) {
#pragma omp critical
            status = (n < 0) ? n : ORC_ERR_TOO_MANY_CENTERS;
            n_out[b] = (n < 0) ? 0 : n;
        } else {
            n_out[b] = n;
        }
    }
    return status;
}

/* ------------------------------------------------------------------------ */
/* a3  offset grouping   model/postprocessing/instance.py:170-268             */
/*  caller de-normalises: off_y*H, off_x*W   panoptic.py:105-111 (one f32 mul)*/
/*  :194  loc = grid + offset                 (one f32 add; grid exact in f32)*/
/*  :231  d = norm(centre - loc, dim=-1)      (f32 sub per component, then    */
/*        ATen's 2-element norm = sqrtf(fmaf(dx, dx, dy*dy)) -- probed,       */
/*        SURVEY.md section 7; dy = y-difference, dx = x-difference)          */
/*  :235  torch.min(dim=0) -> first index on ties, on the sqrt'ed values      */
/*  :236  id = uint8(argmin + 1)   (wraps for > 255 centres: rejected here)   */
/*  :246  optional: id = 0 where min distance > offset_distance_threshold     */
/*  :253  areas = bincount(id)                                                */
/*  :257-265 meta: centre (y, x), area, score = heat[y, x] (un-thresholded)   */
/* ------------------------------------------------------------------------ */
/* > 255 centres: rejected by default; with the wrap switch the loop goes on like   */
/* the reference does (:236 casts to uint8, so centre 256 becomes "no instance",   */
/* centre 257 joins instance 1, ...; :253 bincount of the uint8 ids, so the areas  */
/* of the meta entries beyond 255 are zero)                                        */
static int g_allow_wrap = 0;
void orc_set_allow_wrap(int on) { g_allow_wrap = on; }

static int frame_grouping(const float *heat, const float *off_y, const float *off_x,
                          const uint8_t *fg, int H, int W, int normalized,
                          int use_dist_thr, float dist_thr,
                          const int32_t *centers_yx, int n, uint8_t *inst,
                          int32_t *area /* [cap+1] */, float *score /* [cap] */)
{
    long P = (long)H * W;
    memset(inst, 0, (size_t)P);
    if (n == 0) return ORC_OK;                        /* :214-215            */
    if (n > 255 && !g_allow_wrap) return ORC_ERR_TOO_MANY_CENTERS;
    const float fh = (float)H, fw = (float)W;
    for (long p = 0; p < P; ++p) {
        if (!fg[p]) continue;
        int y = (int)(p / W), x = (int)(p % W);
        float oy = off_y[p], ox = off_x[p];
        if (normalized) { oy = oy * fh; ox = ox * fw; }
        float ly = (float)y + oy;
        float lx = (float)x + ox;
        float best = INFINITY;
        int arg = 0;
        for (int i = 0; i < n; ++i) {
            float d0 = (float)centers_yx[2 * i] - ly;
            float d1 = (float)centers_yx[2 * i + 1] - lx;
            float d = sqrtf(fmaf(d1, d1, d0 * d0));
            if (d < best || i == 0) { best = d; arg = i; }
        }
        uint8_t id = (uint8_t)(arg + 1);
        if (use_dist_thr && best > dist_thr) id = 0;
        inst[p] = id;
        area[id] += 1;
    }
    for (int i = 0; i < n; ++i)
        score[i] = heat[(long)centers_yx[2 * i] * W + centers_yx[2 * i + 1]];
    return ORC_OK;
}

/* heat [B][H][W], offset [B][2][H][W] (ch0 = y, ch1 = x), fg [B][H][W] u8    */
/* inst [B][H][W] u8, centers [B][cap][2], n_out [B], area [B][cap+1],        */
/* score [B][cap]                                                             */
int orc_instance_segmentation(const float *heat, const float *offset, const uint8_t *fg,
                              int B, int H, int W, float thr, int ks, int topk,
                              int apply_fg_mask, int normalized, int use_dist_thr,
                              float dist_thr, uint8_t *inst, int32_t *centers, int cap,
                              int32_t *n_out, int32_t *area, float *score)
{
    long P = (long)H * W;
    int status = orc_instance_centers(heat, B, H, W, thr, ks, topk, fg, apply_fg_mask,
                                      centers, cap, n_out, NULL);
    if (status != ORC_OK) return status;
    memset(area, 0, sizeof(int32_t) * (size_t)B * (cap + 1));
    memset(score, 0, sizeof(float) * (size_t)B * cap);
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        int s = frame_grouping(heat + (size_t)b * P, offset + (size_t)b * 2 * P,
                               offset + (size_t)b * 2 * P + P, fg + (size_t)b * P, H, W,
                               normalized, use_dist_thr, dist_thr,
                               centers + (size_t)b * cap * 2, n_out[b],
                               inst + (size_t)b * P, area + (size_t)b * (cap + 1),
                               score + (size_t)b * cap);
        if (s != ORC_OK) {
#pragma omp critical
            status = s;
        }
    }
    return status;
}

/* ------------------------------------------------------------------------ */
/* a5  deeplab merge   utils/panoptic_merge.py:172-225 (torch variant)        */
/*  :181 pan = void_label everywhere; :182 is_thing = (ins > 0) & fg          */
/*  :192-210 for instance ids ascending (torch.unique): mask = (ins==id) &    */
/*     is_thing; skip if empty; class = torch.mode(sem[mask]) (smallest value */
/*     on ties -- probed); skip if class == 0; n = ++counter[class];          */
/*     pan[mask] = class*L + n; id_dict[pan_id] = id                          */
/*  :213-223 every non-zero, non-thing class c present: pan[(sem==c) &        */
/*     (ins==0)] = c*L                                                        */
/* sem: int32 in [0, n_classes); ins: u8; fg: u8; thing_lut: [n_classes] u8   */
/* id_pairs: [256][2] int64 (pan_id, ins_id) in creation order                */
/* ------------------------------------------------------------------------ */
static int frame_merge(const int32_t *sem, const uint8_t *ins, const uint8_t *fg, long P,
                       int n_classes, int64_t L, const uint8_t *thing_lut,
                       int64_t void_label, int64_t *pan, int64_t *id_pairs, int32_t *n_pairs)
{
    int64_t *hist = (int64_t *)calloc((size_t)256 * n_classes, sizeof(int64_t));
    for (long p = 0; p < P; ++p) {
        if (sem[p] < 0 || sem[p] >= n_classes) { free(hist); return ORC_ERR_CATEGORY_RANGE; }
        if (ins[p] > 0 && fg[p]) hist[(size_t)ins[p] * n_classes + sem[p]] += 1;
    }
    int64_t pan_of_ins[256];
    int64_t *counter = (int64_t *)calloc((size_t)n_classes, sizeof(int64_t));
    int np = 0;
    for (int id = 0; id < 256; ++id) pan_of_ins[id] = void_label;
    for (int id = 1; id < 256; ++id) {
        const int64_t *h = hist + (size_t)id * n_classes;
        int64_t best = 0;
        int cls = -1;
        for (int c = 0; c < n_classes; ++c)
            if (h[c] > best) { best = h[c]; cls = c; }
        if (cls < 0) continue;          /* empty mask            :197-198 */
        if (cls == 0) continue;         /* majority is void      :201-202 */
        counter[cls] += 1;
        pan_of_ins[id] = (int64_t)cls * L + counter[cls];
        id_pairs[2 * np] = pan_of_ins[id];
        id_pairs[2 * np + 1] = id;
        ++np;
    }
    *n_pairs = np;
    for (long p = 0; p < P; ++p) {
        int64_t v = void_label;
        if (ins[p] > 0) {
            if (fg[p]) v = pan_of_ins[ins[p]];
        } else if (sem[p] != 0 && !thing_lut[sem[p]]) {
            v = (int64_t)sem[p] * L;
        }
        pan[p] = v;
    }
    free(hist);
    free(counter);
    return ORC_OK;
}

int orc_deeplab_merge_batch(const int32_t *sem, const uint8_t *ins, const uint8_t *fg,
                            int B, long P, int n_classes, int64_t L,
                            const uint8_t *thing_lut, int64_t void_label, int64_t *pan,
                            int64_t *id_pairs /* [B][256][2] */, int32_t *n_pairs /* [B] */)
{
    int status = ORC_OK;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        int s = frame_merge(sem + (size_t)b * P, ins + (size_t)b * P, fg + (size_t)b * P, P,
                            n_classes, L, thing_lut, void_label, pan + (size_t)b * P,
                            id_pairs + (size_t)b * 512, n_pairs + b);
        if (s != ORC_OK) {
#pragma omp critical
            status = s;
        }
    }
    return status;
}

/* ------------------------------------------------------------------------ */
/* naive merge (ground-truth targets)  utils/panoptic_merge.py:43-107           */
/*  instance ids ascending, inside each the present classes ascending (void     */
/*  skipped): number = ++counter[class]; pan = class*L + number on the part;    */
/*  then every non-void non-thing class c: pan[(sem==c)&(ins==0)] = c*L         */
/* sem [B][P] u8, ins [B][P] i32 in [0,65535]; id_pairs [B][cap][2] (pan, ins)  */
/* ------------------------------------------------------------------------ */
int orc_naive_merge_batch(const uint8_t *sem, const int32_t *ins, int B, long P, int64_t L,
                          const uint8_t *thing_lut /* [256] */, int64_t void_label, int64_t *pan,
                          int64_t *id_pairs, int cap, int32_t *n_pairs)
{
    int status = ORC_OK;
    for (int b = 0; b < B; ++b) {
        const uint8_t *sb = sem + (size_t)b * P;
        const int32_t *ib = ins + (size_t)b * P;
        int64_t *pb = pan + (size_t)b * P;
        /* presence of (instance, class) parts */
        uint8_t *present = (uint8_t *)calloc((size_t)65536 * 256, 1);
        for (long p = 0; p < P; ++p) {
            if (ib[p] < 0 || ib[p] > 65535) { free(present); return ORC_ERR_CATEGORY_RANGE; }
            present[(size_t)ib[p] * 256 + sb[p]] = 1;
        }
        int64_t *part_pan = (int64_t *)calloc((size_t)65536 * 256, sizeof(int64_t));
        int64_t counter[256] = {0};
        int np = 0;
        for (int id = 1; id < 65536; ++id) {
            for (int c = 1; c < 256; ++c) {
                if (!present[(size_t)id * 256 + c]) continue;
                counter[c] += 1;
                const int64_t v = (int64_t)c * L + counter[c];
                part_pan[(size_t)id * 256 + c] = v;
                if (np < cap) { id_pairs[((size_t)b * cap + np) * 2] = v; id_pairs[((size_t)b * cap + np) * 2 + 1] = id; }
                ++np;
            }
        }
        n_pairs[b] = np;
        if (np > cap) status = ORC_ERR_CAPACITY;
        for (long p = 0; p < P; ++p) {
            int64_t v = void_label;
            if (ib[p] != 0) {
                if (sb[p] != 0) v = part_pan[(size_t)ib[p] * 256 + sb[p]];
            } else if (sb[p] != 0 && !thing_lut[sb[p]]) {
                v = (int64_t)sb[p] * L;
            }
            pb[p] = v;
        }
        free(present);
        free(part_pan);
    }
    return status;
}

/* ------------------------------------------------------------------------ */
/* instance targets   data/preprocessing/instance.py:151-286                    */
/*  per instance id != 0 (ascending): class = bincount(sem[mask]).argmax();     */
/*  skipped unless a thing; fg |= mask; centre = (int(mean y), int(mean x));    */
/*  heat = max(heat, gauss stamp); offset[mask] = (cy - y, cx - x);             */
/*  normalised: float32(offset) / (H, W); centre mask = fg | stuff (no void)    */
/* gauss: [(6s+3)^2] f32 stamp supplied by the caller (numpy float64 -> f32)    */
/* offset: [B][2][P] f32 (normalised) or the same integers as f32 (pixels)      */
/* returns -1 if a non-foreground pixel carries an instance id (the assert)     */
/* ------------------------------------------------------------------------ */
int orc_instance_targets(const uint8_t *sem, const int32_t *ins, int B, int H, int W,
                         const uint8_t *thing_lut /* [256] with void */, int sigma,
                         const float *gauss, int normalized, float *center, float *offset,
                         uint8_t *fg, uint8_t *cmask)
{
    const long P = (long)H * W;
    const int size = 6 * sigma + 3;
    int status = ORC_OK;
    for (int b = 0; b < B; ++b) {
        const uint8_t *sb = sem + (size_t)b * P;
        const int32_t *ib = ins + (size_t)b * P;
        float *cb = center + (size_t)b * P, *oy = offset + (size_t)b * 2 * P, *ox = oy + P;
        uint8_t *fb = fg + (size_t)b * P, *mb = cmask + (size_t)b * P;
        for (long p = 0; p < P; ++p) { cb[p] = 0.0f; oy[p] = 0.0f; ox[p] = 0.0f; fb[p] = 0; }
        int64_t *hist = (int64_t *)calloc((size_t)65536 * 256, sizeof(int64_t));
        int64_t *sy = (int64_t *)calloc(65536, sizeof(int64_t));
        int64_t *sx = (int64_t *)calloc(65536, sizeof(int64_t));
        int64_t *n = (int64_t *)calloc(65536, sizeof(int64_t));
        for (long p = 0; p < P; ++p) {
            const int id = ib[p];
            if (id < 0 || id > 65535) { status = ORC_ERR_CATEGORY_RANGE; continue; }
            hist[(size_t)id * 256 + sb[p]] += 1;
            sy[id] += p / W; sx[id] += p % W; n[id] += 1;
        }
        int32_t *cyx = (int32_t *)malloc(sizeof(int32_t) * 65536 * 2);
        for (int id = 1; id < 65536; ++id) {
            cyx[2 * id] = -1;
            if (!n[id]) continue;
            int cls = 0; int64_t best = -1;
            for (int c = 0; c < 256; ++c)
                if (hist[(size_t)id * 256 + c] > best) { best = hist[(size_t)id * 256 + c]; cls = c; }
            if (!thing_lut[cls]) continue;
            const int cy = (int)(sy[id] / n[id]), cx = (int)(sx[id] / n[id]);
            cyx[2 * id] = cy; cyx[2 * id + 1] = cx;
            const int ul_x = cx - 3 * sigma - 1, ul_y = cy - 3 * sigma - 1;
            for (int gy = 0; gy < size; ++gy)
                for (int gx = 0; gx < size; ++gx) {
                    const int y = ul_y + gy, x = ul_x + gx;
                    if (y < 0 || y >= H || x < 0 || x >= W) continue;
                    const float v = gauss[gy * size + gx];
                    if (v > cb[(long)y * W + x]) cb[(long)y * W + x] = v;
                }
        }
        for (long p = 0; p < P; ++p) {
            const int id = ib[p];
            if (id > 0 && id <= 65535) {
                if (cyx[2 * id] >= 0) {
                    const float dy = (float)(cyx[2 * id] - (int)(p / W));
                    const float dx = (float)(cyx[2 * id + 1] - (int)(p % W));
                    oy[p] = normalized ? dy / (float)H : dy;
                    ox[p] = normalized ? dx / (float)W : dx;
                    fb[p] = 1;
                } else {
                    status = ORC_ERR_ARG;
                }
            }
            mb[p] = fb[p] || (sb[p] != 0 && !thing_lut[sb[p]]);
        }
        free(hist); free(sy); free(sx); free(n); free(cyx);
    }
    return status;
}

/* ------------------------------------------------------------------------ */
/* a6  per-instance orientation  model/postprocessing/instance.py:270-319     */
/*  :301-313 for every instance id != 0 present inside the mask:              */
/*     v = sum(orientation[:, mask & seg==id])  (f32 sum in ATen; order-      */
/*     dependent, hence tolerance 1e-5 rel) ; angle = atan2(v[1], v[0])       */
/*     (utils/_orientation.py:39-42: ch0 = cos, ch1 = sin)                    */
/* ori: [B][2][P]; seg: [B][P] int32; mask nullable [B][P] u8                 */
/* present/angle/sum: [B][max_id+1]([2])                                      */
/* ------------------------------------------------------------------------ */
int orc_instance_orientation(const float *ori, const int32_t *seg, const uint8_t *mask,
                             int B, long P, int max_id, uint8_t *present, float *angle,
                             double *sums)
{
    memset(present, 0, (size_t)B * (max_id + 1));
    memset(sums, 0, sizeof(double) * (size_t)B * (max_id + 1) * 2);
    int status = ORC_OK;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        const float *oc = ori + (size_t)b * 2 * P, *os = oc + P;
        const int32_t *sb = seg + (size_t)b * P;
        const uint8_t *mb = mask ? mask + (size_t)b * P : NULL;
        uint8_t *pb = present + (size_t)b * (max_id + 1);
        double *sm = sums + (size_t)b * (max_id + 1) * 2;
        for (long p = 0; p < P; ++p) {
            if (mb && !mb[p]) continue;
            int id = sb[p];
            if (id == 0) continue;
            if (id < 0 || id > max_id) { status = ORC_ERR_ARG; continue; }
            pb[id] = 1;
            sm[2 * id] += (double)oc[p];
            sm[2 * id + 1] += (double)os[p];
        }
        for (int id = 1; id <= max_id; ++id)
            angle[(size_t)b * (max_id + 1) + id] =
                pb[id] ? atan2f((float)sm[2 * id + 1], (float)sm[2 * id]) : NAN;
    }
    return status;
}

/* ------------------------------------------------------------------------ */
/* a8  mIoU confusion matrix   metric/miou.py:44-56                           */
/*  confmat[target][pred] += 1   (bincount(target*n + pred).reshape(n, n))    */
/* ------------------------------------------------------------------------ */
int orc_confmat(const int64_t *pred, const int64_t *target, long N, int n, int64_t *confmat)
{
    for (long i = 0; i < N; ++i) {
        int64_t t = target[i], p = pred[i];
        if (t < 0 || t >= n || p < 0 || p >= n) return ORC_ERR_CATEGORY_RANGE;
        confmat[t * n + p] += 1;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* a9  PQ segment matching for one frame   metric/pq.py:60-179                */
/* Python semantics kept: '//' and '%' are floor division / modulo; the IoU   */
/* is an int/int true division (correctly rounded float64); intersections are */
/* visited in ascending (target*offset + pred) order (torch.unique sorts) and */
/* the float64 IoU sum accumulates in that order.                             */
/* ------------------------------------------------------------------------ */
typedef struct { int64_t id; int64_t cnt; } id_count_t;

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

static long unique_counts(const int64_t *v, long n, id_count_t **out)
{
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    memcpy(tmp, v, sizeof(int64_t) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(int64_t), cmp_i64);
    id_count_t *u = (id_count_t *)malloc(sizeof(id_count_t) * (size_t)(n > 0 ? n : 1));
    long m = 0;
    for (long i = 0; i < n; ++i) {
        if (m > 0 && u[m - 1].id == tmp[i]) u[m - 1].cnt += 1;
        else { u[m].id = tmp[i]; u[m].cnt = 1; ++m; }
    }
    free(tmp);
    *out = u;
    return m;
}

static int64_t lookup(const id_count_t *u, long m, int64_t id, int *found)
{
    long lo = 0, hi = m - 1;
    while (lo <= hi) {
        long mid = (lo + hi) / 2;
        if (u[mid].id == id) { if (found) *found = 1; return u[mid].cnt; }
        if (u[mid].id < id) lo = mid + 1; else hi = mid - 1;
    }
    if (found) *found = 0;
    return 0;
}

static int64_t floordiv(int64_t a, int64_t b)
{
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

static int64_t floormod(int64_t a, int64_t b) { return a - floordiv(a, b) * b; }

/* iou/tp/fn/fp: [num_categories] float64 (zero-initialised by this call)     */
/* matches: [cap][2] int64 (gt_id, pred_id), n_matches out                    */
int orc_pq_compare(const int64_t *pred, const int64_t *target, long P, int num_categories,
                   int64_t ignored_label, int64_t L, int64_t offset, int64_t void_segment_id,
                   double *iou, double *tp, double *fn, double *fp, int64_t *matches,
                   int cap, int32_t *n_matches)
{
    for (int c = 0; c < num_categories; ++c) iou[c] = tp[c] = fn[c] = fp[c] = 0.0;
    id_count_t *ta, *pa, *ia;
    long nt = unique_counts(target, P, &ta);                       /* pq.py:83 */
    long npred = unique_counts(pred, P, &pa);                      /* pq.py:84 */
    int64_t *comb = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P > 0 ? P : 1));
    for (long i = 0; i < P; ++i)                                   /* pq.py:104 */
        comb[i] = (int64_t)((uint64_t)target[i] * (uint64_t)offset + (uint64_t)pred[i]);
    long ni = unique_counts(comb, P, &ia);                         /* pq.py:109 */
    free(comb);

    uint8_t *gt_matched = (uint8_t *)calloc((size_t)(nt > 0 ? nt : 1), 1);
    uint8_t *pred_matched = (uint8_t *)calloc((size_t)(npred > 0 ? npred : 1), 1);
    int status = ORC_OK;
    int nm = 0;

    for (long k = 0; k < ni && status == ORC_OK; ++k) {            /* pq.py:119 */
        int64_t iid = ia[k].id, iarea = ia[k].cnt;
        if (iid == void_segment_id) continue;                      /* pq.py:120 */
        int64_t g = floordiv(iid, offset), p = floormod(iid, offset);
        int64_t gcat = floordiv(g, L), pcat = floordiv(p, L);
        if (gcat != pcat) continue;                                /* pq.py:128 */
        int64_t r = lookup(ia, ni, void_segment_id * offset + p, NULL);   /* :134 */
        int fg_, fp_;
        int64_t tsa = lookup(ta, nt, g, &fg_);
        int64_t psa = lookup(pa, npred, p, &fp_);
        if (!fg_ || !fp_) { status = ORC_ERR_ARG; break; }         /* KeyError  */
        int64_t uni = tsa + psa - iarea - r;                       /* pq.py:143 */
        if (uni == 0) { status = ORC_ERR_ZERO_DIVISION; break; }
        double v = (double)iarea / (double)uni;                    /* pq.py:145 */
        if (v > 0.5) {
            if (gcat < 0 || gcat >= num_categories) { status = ORC_ERR_CATEGORY_RANGE; break; }
            tp[gcat] += 1.0;
            iou[gcat] += v;
            for (long j = 0; j < nt; ++j) if (ta[j].id == g) gt_matched[j] = 1;
            for (long j = 0; j < npred; ++j) if (pa[j].id == p) pred_matched[j] = 1;
            if (nm < cap) { matches[2 * nm] = g; matches[2 * nm + 1] = p; }
            ++nm;
        }
    }
    for (long j = 0; j < nt && status == ORC_OK; ++j) {            /* pq.py:155 */
        if (gt_matched[j]) continue;
        int64_t cat = floordiv(ta[j].id, L);
        if (cat == ignored_label) continue;
        if (cat < 0 || cat >= num_categories) { status = ORC_ERR_CATEGORY_RANGE; break; }
        fn[cat] += 1.0;
    }
    for (long j = 0; j < npred && status == ORC_OK; ++j) {         /* pq.py:165 */
        if (pred_matched[j]) continue;
        int64_t pio = 0;
        for (long t = 0; t < nt; ++t)                              /* pq.py:47-57 */
            if (floordiv(ta[t].id, L) == ignored_label)
                pio += lookup(ia, ni, ta[t].id * offset + pa[j].id, NULL);
        if ((double)pio / (double)pa[j].cnt > 0.5) continue;       /* pq.py:174 */
        int64_t cat = floordiv(pa[j].id, L);
        if (cat < 0 || cat >= num_categories) { status = ORC_ERR_CATEGORY_RANGE; break; }
        fp[cat] += 1.0;
    }
    *n_matches = nm;
    if (status == ORC_OK && nm > cap) status = ORC_ERR_CAPACITY;
    free(ta); free(pa); free(ia); free(gt_matched); free(pred_matched);
    return status;
}

/* ------------------------------------------------------------------------ */
/* Whole pipeline for a batch (a1 -> a2 -> a3 -> a5 -> a6), frames in         */
/* parallel with OpenMP.  Mirrors PanopticPostprocessing._postprocess_-       */
/* inference (model/postprocessing/panoptic.py:77-316) for the dense outputs. */
/* Used as the checker and as bench.py's CPU baseline ("port").               */
/*  thing_lut / orient_lut: [C] u8 indexed by network class (without void)    */
/* Outputs: sem [B][P] u8 (network class), inst [B][P] u8, pan [B][P] i64,    */
/*  centers [B][cap][2], n_out [B], area [B][cap+1], score [B][cap],          */
/*  id_pairs [B][256][2], n_pairs [B], ori_present [B][256] u8 (nullable),    */
/*  ori_angle [B][256] f32 (nullable)                                         */
/* ------------------------------------------------------------------------ */
int orc_panoptic_postprocess(const float *logits, const float *heat, const float *offset,
                             const float *orientation /* nullable */, int B, int C, int H,
                             int W, const uint8_t *thing_lut, const uint8_t *orient_lut,
                             float thr, int ks, int topk, int apply_fg_mask, int normalized,
                             int use_dist_thr, float dist_thr, int64_t L, uint8_t *sem,
                             uint8_t *inst, int64_t *pan, int32_t *centers, int cap,
                             int32_t *n_out, int32_t *area, float *score, int64_t *id_pairs,
                             int32_t *n_pairs, uint8_t *ori_present, float *ori_angle)
{
    long P = (long)H * W;
    int status = orc_semantic_argmax(logits, B, C, P, sem);
    if (status != ORC_OK) return status;
    uint8_t *fg = (uint8_t *)malloc((size_t)B * P);
    for (size_t i = 0; i < (size_t)B * P; ++i) fg[i] = thing_lut[sem[i]];   /* panoptic.py:123-127 */
    status = orc_instance_segmentation(heat, offset, fg, B, H, W, thr, ks, topk, apply_fg_mask,
                                       normalized, use_dist_thr, dist_thr, inst, centers, cap,
                                       n_out, area, score);
    if (status != ORC_OK) { free(fg); return status; }
    int32_t *sem1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * P);
    for (size_t i = 0; i < (size_t)B * P; ++i) sem1[i] = (int32_t)sem[i] + 1;  /* panoptic.py:146 */
    uint8_t *thing1 = (uint8_t *)calloc((size_t)C + 1, 1);
    for (int c = 0; c < C; ++c) thing1[c + 1] = thing_lut[c];
    status = orc_deeplab_merge_batch(sem1, inst, fg, B, P, C + 1, L, thing1, 0, pan, id_pairs,
                                     n_pairs);
    free(thing1);
    if (status == ORC_OK && orientation && ori_present && ori_angle) {
        /* panoptic.py:296-307: mask = isin(pan // L, orientation_ids)        */
        uint8_t *omask = (uint8_t *)malloc((size_t)B * P);
        int32_t *seg = sem1; /* reuse buffer for the int32 instance map */
        for (size_t i = 0; i < (size_t)B * P; ++i) {
            int64_t c = pan[i] / L;
            omask[i] = (c >= 1 && c <= C) ? orient_lut[c - 1] : 0;
            seg[i] = inst[i];
        }
        double *sums = (double *)malloc(sizeof(double) * (size_t)B * 256 * 2);
        status = orc_instance_orientation(orientation, seg, omask, B, P, 255, ori_present,
                                          ori_angle, sums);
        free(sums);
        free(omask);
    }
    free(sem1);
    free(fg);
    return status;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
