/*
 * oracle.c -- CPU restatement of the reference's dense per-anchor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ood_object_detection_b200/ may import, link or call
 * this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg can check and time the CUDA path against an independent statement of the
 * reference algorithm.  Parity is PINNED: tests/test_oracle_golden.py checks every function here
 * against the tests/golden npz files, which were produced by running the unmodified reference python
 * (DavidPetrus/ood_object_detection, effdet package) in the build container (tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it restates (paths relative to the reference root).
 * All arithmetic that feeds a threshold decision is done in fp32 in the reference's operation
 * order; compile with -ffp-contract=off so gcc never fuses a multiply-add.
 * Third-party arithmetic on the path (not in the reference tree): torchvision 0.26.0
 * ops.boxes._batched_nms_coordinate_trick + torchvision::nms (CPU kernel semantics: stable
 * descending score order, suppress iff inter/(area_i+area_j-inter) > thr, compared in double),
 * torch 2.11 topk/max/argmax (first index on ties), binary_cross_entropy_with_logits.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ A2: IoU
 * effdet/object_detection/region_similarity_calculator.py:24-73 (area, intersection, iou). */
static inline float area_yxyx(const float *b) { return (b[2] - b[0]) * (b[3] - b[1]); }

static inline float iou_yxyx(const float *g, float area_g, const float *a, float area_a) {
    float h = fminf(g[2], a[2]) - fmaxf(g[0], a[0]);
    if (h < 0.0f) h = 0.0f;
    float w = fminf(g[3], a[3]) - fmaxf(g[1], a[1]);
    if (w < 0.0f) w = 0.0f;
    float inter = h * w;
    if (inter == 0.0f) return 0.0f;
    float uni = (area_g + area_a) - inter;
    return inter / uni;
}

/* ------------------------------------------------------------------ A3-A7: target assignment
 * argmax_matcher.py:105-146 (match), matcher.py:151-179 (gather), box_list.py:152-164 +
 * box_coder.py:81-110 (encode), target_assigner.py:155-220, anchors.py:413-416,434.
 * One image.  gt rows are the already filtered (anchors.py:405-408) boxes, labels 1-based.
 * match: -1 unmatched, -2 ignored, >=0 gt row.  cls = label[match]-1 (0-1 = -1 for background).
 */
static void assign_image(const float *anchors, int64_t A, const float *gt, const int64_t *labels, int64_t M,
                         float matched_thr, float unmatched_thr, int negatives_lower, int force_match,
                         int64_t *match, int64_t *cls_t, float *box_t, float *num_pos) {
    if (M == 0) { /* argmax_matcher.py:105-114 */
        for (int64_t j = 0; j < A; ++j) match[j] = -1;
    } else {
        float *garea = (float *)malloc(sizeof(float) * M);
        for (int64_t i = 0; i < M; ++i) garea[i] = area_yxyx(gt + 4 * i);
        int nthreads = 1;
#ifdef _OPENMP
        nthreads = omp_get_max_threads();
#endif
        /* per-thread running argmax over anchors for every gt row (argmax_matcher.py:140) */
        float *tbest = (float *)malloc(sizeof(float) * M * nthreads);
        int64_t *tidx = (int64_t *)malloc(sizeof(int64_t) * M * nthreads);
        for (int64_t q = 0; q < M * nthreads; ++q) { tbest[q] = -2.0f; tidx[q] = 0; } /* idle threads never win */
#pragma omp parallel
        {
            int t = 0, nt = 1;
#ifdef _OPENMP
            t = omp_get_thread_num();
            nt = omp_get_num_threads();
#endif
            float *rb = tbest + (int64_t)t * M;
            int64_t *ri = tidx + (int64_t)t * M;
            for (int64_t i = 0; i < M; ++i) { rb[i] = -1.0f; ri[i] = 0; }
            int64_t lo = A * t / nt, hi = A * (t + 1) / nt;
            for (int64_t j = lo; j < hi; ++j) {
                const float *a = anchors + 4 * j;
                float aa = area_yxyx(a);
                /* column max over gt rows, first row on ties (argmax_matcher.py:123) */
                float best = 0.0f;
                int64_t bi = 0;
                for (int64_t i = 0; i < M; ++i) {
                    float v = iou_yxyx(gt + 4 * i, garea[i], a, aa);
                    if (i == 0 || v > best) { best = v; bi = i; }
                    if (v > rb[i]) { rb[i] = v; ri[i] = j; }
                }
                /* thresholds (argmax_matcher.py:126-137) */
                int below = unmatched_thr > best;
                int between = (best >= unmatched_thr) && (matched_thr > best);
                int64_t m = bi;
                if (negatives_lower) { if (below) m = -1; if (between) m = -2; }
                else { if (below) m = -2; if (between) m = -1; }
                match[j] = m;
            }
        }
        if (force_match) { /* argmax_matcher.py:139-144: lowest gt row wins a contested column */
            for (int64_t i = M - 1; i >= 0; --i) {
                float best = -1.0f;
                int64_t bj = 0;
                for (int t = 0; t < nthreads; ++t) { /* threads own ascending anchor ranges */
                    if (tbest[(int64_t)t * M + i] > best) { best = tbest[(int64_t)t * M + i]; bj = tidx[(int64_t)t * M + i]; }
                }
                match[bj] = i;
            }
        }
        free(garea); free(tbest); free(tidx);
    }
    int64_t npos = 0;
    const float eps = 1e-8f; /* box_coder.py:52 */
#pragma omp parallel for reduction(+ : npos) schedule(static)
    for (int64_t j = 0; j < A; ++j) {
        int64_t m = match[j];
        float *o = box_t + 4 * j;
        if (m >= 0) {
            npos += 1;
            const float *a = anchors + 4 * j;
            const float *g = gt + 4 * m;
            /* box_list.py:159-164: width/height first, centre = min + size/2 */
            float wa = a[3] - a[1], ha = a[2] - a[0];
            float yca = a[0] + ha / 2.0f, xca = a[1] + wa / 2.0f;
            float w = g[3] - g[1], h = g[2] - g[0];
            float yc = g[0] + h / 2.0f, xc = g[1] + w / 2.0f;
            ha += eps; wa += eps; h += eps; w += eps; /* box_coder.py:95-98 */
            o[0] = (yc - yca) / ha;                   /* ty */
            o[1] = (xc - xca) / wa;                   /* tx */
            o[2] = logf(h / ha);                      /* th */
            o[3] = logf(w / wa);                      /* tw */
            cls_t[j] = labels[m] - 1;
        } else {
            o[0] = o[1] = o[2] = o[3] = 0.0f;         /* target_assigner.py:181-184 */
            cls_t[j] = -1;                            /* unmatched_cls_target 0, minus 1 */
        }
    }
    *num_pos = (float)npos; /* anchors.py:434 */
}

ORC_API void orc_assign_batch(const float *anchors, int64_t A, const float *gt_boxes, const int64_t *gt_labels,
                              const int64_t *gt_count, int64_t B, int64_t Mmax, float matched_thr,
                              float unmatched_thr, int negatives_lower, int force_match, int64_t *match,
                              int64_t *cls_t, float *box_t, float *num_pos) {
    for (int64_t b = 0; b < B; ++b)
        assign_image(anchors, A, gt_boxes + b * Mmax * 4, gt_labels + b * Mmax, gt_count[b], matched_thr,
                     unmatched_thr, negatives_lower, force_match, match + b * A, cls_t + b * A,
                     box_t + b * A * 4, num_pos + b);
}

/* pairwise IoU matrix [N, M] (IouSimilarity.compare; used by the task_cls relabel, anchors.py:401) */
ORC_API void orc_iou_matrix(const float *b1, int64_t N, const float *b2, int64_t M, float *out) {
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < M; ++j)
            out[i * M + j] = iou_yxyx(b1 + 4 * i, area_yxyx(b1 + 4 * i), b2 + 4 * j, area_yxyx(b2 + 4 * j));
}

/* ------------------------------------------------------------------ A8: detection loss, one level
 * effdet/loss.py:182-186 (one_hot), :49-95 (new focal), :15-47 (legacy focal), :104-118 (huber),
 * :171-179 (_box_loss), :266-292 (per-level loop).  cls_out [B, 9C, H, W]; box_out [B, 36, H, W];
 * cls_t [B, H, W, 9] int64 (-1 background, -2 ignore); box_t [B, H, W, 36].
 * nps = sum(num_positives) + 1 (loss.py:261) as fp32.  Element values are formed in fp32 in the
 * reference's order and accumulated in double.  Optional gradients of
 *   total = cls_loss + box_w * box_loss     (loss.py:297)
 * w.r.t. cls_out / box_out are written when the pointers are non-NULL.
 */
static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

ORC_API void orc_loss_level(const float *cls_out, const float *box_out, const int64_t *cls_t, const float *box_t,
                            int64_t B, int64_t H, int64_t W, int64_t C, int64_t NA, float nps, float alpha,
                            float gamma, float delta, float smoothing, int legacy, float box_w, double *cls_sum,
                            double *box_sum, float *grad_cls, float *grad_box) {
    const int64_t HW = H * W;
    const float inv_n = 1.0f / nps;
    double csum = 0.0, bsum = 0.0;
#pragma omp parallel for reduction(+ : csum, bsum) schedule(static) collapse(2)
    for (int64_t b = 0; b < B; ++b) {
        for (int64_t a = 0; a < NA; ++a) {
            for (int64_t hw = 0; hw < HW; ++hw) {
                int64_t tc = cls_t[(b * HW + hw) * NA + a];
                float keep = (tc != -2) ? 1.0f : 0.0f; /* loss.py:285 */
                for (int64_t c = 0; c < C; ++c) {
                    int64_t off = ((b * NA * C) + a * C + c) * HW + hw;
                    float x = cls_out[off];
                    float t = (tc >= 0 && tc == c) ? 1.0f : 0.0f; /* loss.py:182-186 */
                    float v, g;
                    float sig = sigmoidf_(x);
                    if (!legacy) {
                        float af = t * alpha + (1.0f - t) * (1.0f - alpha); /* loss.py:78 */
                        float ts = t;
                        if (smoothing > 0.0f) ts = t * (1.0f - smoothing) + 0.5f * smoothing; /* :86 */
                        /* ATen binary_cross_entropy_with_logits: (1-t)*x - log_sigmoid(x) */
                        float bce = (1.0f - ts) * x - (fminf(x, 0.0f) - log1pf(expf(-fabsf(x))));
                        v = (inv_n * af) * bce; /* loss.py:93 */
                        g = inv_n * af * (sig - ts);
                    } else {
                        float bce = (1.0f - t) * x - (fminf(x, 0.0f) - log1pf(expf(-fabsf(x))));
                        float nx = -1.0f * x;
                        float mod = expf(gamma * t * nx - gamma * log1pf(expf(nx))); /* loss.py:43 */
                        float l = mod * bce;
                        float wl = (t == 1.0f) ? alpha * l : (1.0f - alpha) * l; /* loss.py:46 */
                        v = wl / nps;
                        float aw = (t == 1.0f) ? alpha : (1.0f - alpha);
                        g = aw * (mod * (sig - t) + bce * mod * gamma * (1.0f - sig - t)) / nps;
                    }
                    csum += (double)(v * keep);
                    if (grad_cls) grad_cls[off] = g * keep;
                }
                for (int64_t k = 0; k < 4; ++k) {
                    int64_t off = ((b * NA * 4) + a * 4 + k) * HW + hw;
                    float tg = box_t[(b * HW + hw) * NA * 4 + a * 4 + k];
                    float e = box_out[off] - tg;             /* loss.py:108 */
                    float ae = fabsf(e);
                    float q = ae < delta ? ae : delta;       /* clamp(max=delta) */
                    float lin = ae - q;
                    float l = 0.5f * (q * q) + delta * lin;  /* loss.py:112 */
                    float m = (tg != 0.0f) ? 1.0f : 0.0f;    /* loss.py:177 */
                    bsum += (double)(l * m);
                    if (grad_box) {
                        float d = (ae <= delta) ? e : (e > 0 ? delta : -delta);
                        grad_box[off] = box_w * m * d / (nps * 4.0f);
                    }
                }
            }
        }
    }
    *cls_sum = csum;
    *box_sum = bsum / ((double)nps * 4.0); /* loss.py:176,179 */
}

/* ------------------------------------------------------------------ A9: top-k over [A*C] per image
 * effdet/bench.py:36-54.  Level tensors are NCHW slices of ONE image: lvl[l] -> [9C, H_l, W_l].
 * Reference flat index = anchor*C + class with anchor = base_l + (h*W+w)*9 + a (bench.py:37).
 * Output sorted by (value desc, flat index asc) -- torch.topk's order on tie-free inputs.
 */
typedef struct { float v; int64_t i; } kv_t;
static inline int kv_better(kv_t a, kv_t b) { return a.v > b.v || (a.v == b.v && a.i < b.i); }

static void heap_sift_down(kv_t *h, int64_t n, int64_t p) { /* min-heap on "better": root = worst kept */
    for (;;) {
        int64_t l = 2 * p + 1, r = l + 1, w = p;
        if (l < n && kv_better(h[w], h[l])) w = l;
        if (r < n && kv_better(h[w], h[r])) w = r;
        if (w == p) return;
        kv_t t = h[p]; h[p] = h[w]; h[w] = t; p = w;
    }
}
static int kv_cmp_desc(const void *pa, const void *pb) {
    kv_t a = *(const kv_t *)pa, b = *(const kv_t *)pb;
    return kv_better(a, b) ? -1 : (kv_better(b, a) ? 1 : 0);
}

static void topk_image(const float *const *lvl, const int64_t *hw, int nlev, int64_t C, int64_t NA, int64_t K,
                       int64_t *out_idx, float *out_val) {
    kv_t *heap = (kv_t *)malloc(sizeof(kv_t) * K);
    int64_t n = 0, base = 0;
    for (int l = 0; l < nlev; ++l) {
        int64_t HW = hw[l];
        for (int64_t ch = 0; ch < NA * C; ++ch) {
            int64_t a = ch / C, c = ch % C;
            const float *p = lvl[l] + ch * HW;
            for (int64_t s = 0; s < HW; ++s) {
                kv_t e = {p[s], (base + s * NA + a) * C + c};
                if (n < K) {
                    heap[n++] = e;
                    if (n == K) for (int64_t q = K / 2 - 1; q >= 0; --q) heap_sift_down(heap, K, q);
                } else if (kv_better(e, heap[0])) {
                    heap[0] = e;
                    heap_sift_down(heap, K, 0);
                }
            }
        }
        base += HW * NA;
    }
    qsort(heap, n, sizeof(kv_t), kv_cmp_desc);
    for (int64_t q = 0; q < n; ++q) { out_idx[q] = heap[q].i; out_val[q] = heap[q].v; }
    free(heap);
}

/* cls_lvl[l] / box_lvl[l] point at [B, 9C, H_l, W_l] / [B, 36, H_l, W_l].  Outputs (bench.py:45-54):
 * cls_k [B,K] (the selected logit), box_k [B,K,4], anchor_idx [B,K], klass [B,K]. */
ORC_API void orc_post_process(const float *const *cls_lvl, const float *const *box_lvl, const int64_t *hw, int nlev,
                              int64_t B, int64_t C, int64_t NA, int64_t K, float *cls_k, float *box_k,
                              int64_t *anchor_idx, int64_t *klass) {
#pragma omp parallel for schedule(dynamic)
    for (int64_t b = 0; b < B; ++b) {
        const float *lv[16];
        for (int l = 0; l < nlev; ++l) lv[l] = cls_lvl[l] + b * NA * C * hw[l];
        int64_t *fi = (int64_t *)malloc(sizeof(int64_t) * K);
        topk_image(lv, hw, nlev, C, NA, K, fi, cls_k + b * K);
        for (int64_t q = 0; q < K; ++q) {
            int64_t anc = fi[q] / C; /* bench.py:45-46 */
            anchor_idx[b * K + q] = anc;
            klass[b * K + q] = fi[q] % C;
            int l = 0;
            int64_t base = 0;
            while (anc >= base + hw[l] * NA) { base += hw[l] * NA; ++l; }
            int64_t s = (anc - base) / NA, a = (anc - base) % NA;
            for (int k = 0; k < 4; ++k) /* bench.py:48-49 */
                box_k[(b * K + q) * 4 + k] = box_lvl[l][((b * NA * 4) + a * 4 + k) * hw[l] + s];
        }
        free(fi);
    }
}

/* ------------------------------------------------------------------ A11: soft-NMS
 * effdet/soft_nms.py:12-39 (pairwise_iou, xyxy), :42-112 (soft_nms loop).  max_rounds < 0: run
 * until nothing is left (the reference); the first D outputs do not depend on later rounds. */
static inline float iou_xyxy_soft(const float *p, const float *q) {
    float a1 = (p[2] - p[0]) * (p[3] - p[1]);
    float a2 = (q[2] - q[0]) * (q[3] - q[1]);
    float w = fminf(p[2], q[2]) - fmaxf(p[0], q[0]);
    float h = fminf(p[3], q[3]) - fmaxf(p[1], q[1]);
    if (w < 0.0f) w = 0.0f;
    if (h < 0.0f) h = 0.0f;
    float inter = w * h;
    return inter > 0.0f ? inter / ((a1 + a2) - inter) : 0.0f;
}

ORC_API int64_t orc_soft_nms(const float *boxes, const float *scores, int64_t n, int gaussian, float sigma,
                             float iou_thr, float score_thr, int64_t max_rounds, int64_t *idx_out,
                             float *score_out) {
    float *bx = (float *)malloc(sizeof(float) * 4 * (n ? n : 1));
    float *sc = (float *)malloc(sizeof(float) * (n ? n : 1));
    int64_t *id = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1));
    memcpy(bx, boxes, sizeof(float) * 4 * n);
    memcpy(sc, scores, sizeof(float) * n);
    for (int64_t i = 0; i < n; ++i) id[i] = i;
    int64_t m = n, count = 0;
    while (m > 0 && (max_rounds < 0 || count < max_rounds)) {
        int64_t top = 0;
        for (int64_t i = 1; i < m; ++i) if (sc[i] > sc[top]) top = i; /* argmax: first maximal */
        idx_out[count] = id[top];
        score_out[count] = sc[top];
        ++count;
        float tb[4] = {bx[4 * top], bx[4 * top + 1], bx[4 * top + 2], bx[4 * top + 3]};
        int64_t w = 0;
        for (int64_t i = 0; i < m; ++i) {
            float iou = iou_xyxy_soft(tb, bx + 4 * i);
            float decay;
            if (gaussian) decay = expf(-(iou * iou) / sigma); /* soft_nms.py:96 */
            else decay = (iou > iou_thr) ? 1.0f - iou : 1.0f;  /* :98-100 */
            float s = sc[i] * decay;
            if (s > score_thr && i != top) { /* :103-104 */
                memmove(bx + 4 * w, bx + 4 * i, sizeof(float) * 4);
                sc[w] = s; id[w] = id[i]; ++w;
            }
        }
        m = w;
    }
    free(bx); free(sc); free(id);
    return count;
}

/* ------------------------------------------------------------------ A10: generate_detections
 * effdet/anchors.py:51-85 (decode), :88-92 (clip), :95-172.  One image.  Inputs are the
 * _post_process outputs for the image.  soft=0: torchvision coordinate-trick batched_nms(0.3);
 * soft=1: batched_soft_nms(gaussian, sigma .5, iou .3, score thr .001) (anchors.py:146-148).
 * det [D,6] = x0,y0,x1,y1,score,class+1; src[q] = position (0..N-1) in the top-k list that
 * detection q came from.  Returns the number of rows.
 */
typedef struct { float s; int64_t i; } si_t;
static int si_cmp_desc_stable(const void *pa, const void *pb) {
    si_t a = *(const si_t *)pa, b = *(const si_t *)pb;
    if (a.s > b.s) return -1;
    if (a.s < b.s) return 1;
    return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
}

ORC_API int64_t orc_generate_detections(const float *cls, const float *box, const float *anchors,
                                        const int64_t *indices, const int64_t *classes, int64_t N, int has_scale,
                                        float img_scale, int has_size, const float *img_size, int64_t D, int soft,
                                        float score_min, double nms_thr, float soft_sigma, float soft_iou,
                                        float soft_score_thr, float *det, int64_t *src) {
    float *bx = (float *)malloc(sizeof(float) * 4 * (N ? N : 1));
    float *sc = (float *)malloc(sizeof(float) * (N ? N : 1));
    int64_t *kc = (int64_t *)malloc(sizeof(int64_t) * (N ? N : 1));
    int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (N ? N : 1));
    int64_t n = 0;
    float lim[4] = {0, 0, 0, 0};
    int clip = has_scale && has_size; /* anchors.py:137 */
    if (clip) { lim[0] = lim[2] = img_size[0] / img_scale; lim[1] = lim[3] = img_size[1] / img_scale; }
    for (int64_t q = 0; q < N; ++q) {
        const float *a = anchors + 4 * indices[q]; /* anchors.py:132 */
        const float *r = box + 4 * q;
        float yca = (a[0] + a[2]) / 2.0f, xca = (a[1] + a[3]) / 2.0f; /* anchors.py:66-69 */
        float ha = a[2] - a[0], wa = a[3] - a[1];
        float w = expf(r[3]) * wa, h = expf(r[2]) * ha;
        float yc = r[0] * ha + yca, xc = r[1] * wa + xca;
        float o[4] = {xc - w / 2.0f, yc - h / 2.0f, xc + w / 2.0f, yc + h / 2.0f}; /* xyxy */
        if (clip)
            for (int k = 0; k < 4; ++k) { /* anchors.py:88-92 */
                if (o[k] < 0.0f) o[k] = 0.0f;
                if (o[k] > lim[k]) o[k] = lim[k];
            }
        float s = 1.0f / (1.0f + expf(-cls[q])); /* anchors.py:140 */
        if (s > score_min) {                     /* anchors.py:141-144 */
            memcpy(bx + 4 * n, o, sizeof(o));
            sc[n] = s; kc[n] = classes[q]; pos[n] = q; ++n;
        }
    }
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1));
    int64_t nk = 0;
    if (n > 0) {
        /* coordinate trick (torchvision ops/boxes.py:94-111; soft_nms.py:163-165) */
        float mx = bx[0];
        for (int64_t i = 0; i < 4 * n; ++i) if (bx[i] > mx) mx = bx[i];
        float *ob = (float *)malloc(sizeof(float) * 4 * n);
        for (int64_t i = 0; i < n; ++i) {
            float off = (float)kc[i] * (mx + 1.0f);
            for (int k = 0; k < 4; ++k) ob[4 * i + k] = bx[4 * i + k] + off;
        }
        if (soft) {
            float *ss = (float *)malloc(sizeof(float) * n);
            nk = orc_soft_nms(ob, sc, n, 1, soft_sigma, soft_iou, soft_score_thr, -1, keep, ss);
            for (int64_t i = 0; i < nk; ++i) sc[keep[i]] = ss[i]; /* anchors.py:148 */
            free(ss);
        } else {
            /* torchvision::nms CPU kernel */
            si_t *ord = (si_t *)malloc(sizeof(si_t) * n);
            for (int64_t i = 0; i < n; ++i) { ord[i].s = sc[i]; ord[i].i = i; }
            qsort(ord, n, sizeof(si_t), si_cmp_desc_stable);
            unsigned char *sup = (unsigned char *)calloc(n, 1);
            float *ar = (float *)malloc(sizeof(float) * n);
            for (int64_t i = 0; i < n; ++i) ar[i] = (ob[4 * i + 2] - ob[4 * i]) * (ob[4 * i + 3] - ob[4 * i + 1]);
            for (int64_t _i = 0; _i < n; ++_i) {
                int64_t i = ord[_i].i;
                if (sup[i]) continue;
                keep[nk++] = i;
                for (int64_t _j = _i + 1; _j < n; ++_j) {
                    int64_t j = ord[_j].i;
                    if (sup[j]) continue;
                    float xx1 = fmaxf(ob[4 * i], ob[4 * j]), yy1 = fmaxf(ob[4 * i + 1], ob[4 * j + 1]);
                    float xx2 = fminf(ob[4 * i + 2], ob[4 * j + 2]), yy2 = fminf(ob[4 * i + 3], ob[4 * j + 3]);
                    float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
                    float inter = w * h;
                    float ovr = inter / (ar[i] + ar[j] - inter);
                    if ((double)ovr > nms_thr) sup[j] = 1;
                }
            }
            free(ord); free(sup); free(ar);
        }
        free(ob);
    }
    if (nk > D) nk = D; /* anchors.py:153 */
    for (int64_t q = 0; q < nk; ++q) {
        int64_t i = keep[q];
        for (int k = 0; k < 4; ++k) det[6 * q + k] = has_scale ? bx[4 * i + k] * img_scale : bx[4 * i + k];
        det[6 * q + 4] = sc[i];
        det[6 * q + 5] = (float)(kc[i] + 1); /* anchors.py:156 */
        src[q] = pos[i];
    }
    free(bx); free(sc); free(kc); free(pos); free(keep);
    return nk;
}

/* plain torchvision::nms (no class offsets) for the soft_nms-module goldens */
ORC_API int64_t orc_nms(const float *boxes, const float *scores, int64_t n, double thr, int64_t *keep) {
    si_t *ord = (si_t *)malloc(sizeof(si_t) * (n ? n : 1));
    for (int64_t i = 0; i < n; ++i) { ord[i].s = scores[i]; ord[i].i = i; }
    qsort(ord, n, sizeof(si_t), si_cmp_desc_stable);
    unsigned char *sup = (unsigned char *)calloc(n ? n : 1, 1);
    int64_t nk = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = ord[_i].i;
        if (sup[i]) continue;
        keep[nk++] = i;
        float ai = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = ord[_j].i;
            if (sup[j]) continue;
            float aj = (boxes[4 * j + 2] - boxes[4 * j]) * (boxes[4 * j + 3] - boxes[4 * j + 1]);
            float xx1 = fmaxf(boxes[4 * i], boxes[4 * j]), yy1 = fmaxf(boxes[4 * i + 1], boxes[4 * j + 1]);
            float xx2 = fminf(boxes[4 * i + 2], boxes[4 * j + 2]), yy2 = fminf(boxes[4 * i + 3], boxes[4 * j + 3]);
            float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
            float inter = w * h;
            float ovr = inter / (ai + aj - inter);
            if ((double)ovr > thr) sup[j] = 1;
        }
    }
    free(ord); free(sup);
    return nk;
}

/* decode_box_outputs (anchors.py:51-85) for [n,4] codes against [n,4] anchors */
ORC_API void orc_decode(const float *codes, const float *anchors, int64_t n, int xyxy, float *out) {
    for (int64_t q = 0; q < n; ++q) {
        const float *a = anchors + 4 * q, *r = codes + 4 * q;
        float yca = (a[0] + a[2]) / 2.0f, xca = (a[1] + a[3]) / 2.0f;
        float ha = a[2] - a[0], wa = a[3] - a[1];
        float w = expf(r[3]) * wa, h = expf(r[2]) * ha;
        float yc = r[0] * ha + yca, xc = r[1] * wa + xca;
        float ymin = yc - h / 2.0f, xmin = xc - w / 2.0f, ymax = yc + h / 2.0f, xmax = xc + w / 2.0f;
        float *o = out + 4 * q;
        if (xyxy) { o[0] = xmin; o[1] = ymin; o[2] = xmax; o[3] = ymax; }
        else { o[0] = ymin; o[1] = xmin; o[2] = ymax; o[3] = xmax; }
    }
}

/* ------------------------------------------------------------------ A12: OOD scores
 * Not in the reference (parity unpinned by the reference).  Definition:
 * energy = -T*logsumexp(row/T), max_logit = max(row) over the C raw logits of the anchor. */
ORC_API void orc_ood(const float *rows, int64_t n, int64_t C, float T, float *energy, float *max_logit) {
    for (int64_t q = 0; q < n; ++q) {
        const float *r = rows + q * C;
        float m = r[0];
        for (int64_t c = 1; c < C; ++c) if (r[c] > m) m = r[c];
        double s = 0.0;
        for (int64_t c = 0; c < C; ++c) s += exp(((double)r[c] - (double)m) / (double)T);
        energy[q] = (float)(-(double)T * ((double)m / (double)T + log(s)));
        max_logit[q] = m;
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the baseline leg asks for all host cores explicitly */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ F4: per-image TP/FP matching and CorLoc
 * effdet/evaluation/per_image_evaluation.py:29-92 (compute_object_detection_metrics), :93-175 (CorLoc),
 * :177-240 (_compute_tp_fp, per class), :276-303 (_get_overlaps_and_scores_box_mode), :305-470
 * (_compute_tp_fp_for_single_class: compute_match_iou / compute_match_ioa), :512-536 (_remove_invalid_boxes);
 * effdet/evaluation/np_box_list.py:128-205 (area / intersection / iou / ioa), :297-396 (sort_by_field,
 * non_max_suppression).  numpy arithmetic as the reference runs it on float32 boxes: areas and the coordinate
 * differences are rounded in fp32, but np.maximum(np.zeros(...), diff) promotes the intersection sides to
 * float64, so intersection, union (= (float)(area1 + area2) - inter) and the ratios are float64.
 * Classes are 0-based here (the evaluator has already removed its label offset).  Outputs per detection slot:
 * 1 true positive, 0 false positive, -1 ignored (matched a difficult or a group-of box), -2 removed (invalid box,
 * score filter, NMS); corloc[c] as _compute_cor_loc.  group_of_weight = 0 (the reference's default): detections
 * matched to group-of boxes are only ignored.  Ties between equal scores: numpy's argsort()[::-1] order is
 * unspecified; this restatement takes the later detection first (a stable ascending sort, reversed). */
static double eval_inter(const float *p, const float *q) {
    const float h32 = fminf(p[2], q[2]) - fmaxf(p[0], q[0]);
    const float w32 = fminf(p[3], q[3]) - fmaxf(p[1], q[1]);
    const double h = h32 > 0.0f ? (double)h32 : 0.0, w = w32 > 0.0f ? (double)w32 : 0.0;
    return h * w;
}
static double eval_iou(const float *p, const float *q) {
    const double inter = eval_inter(p, q);
    const float a1 = area_yxyx(p), a2 = area_yxyx(q);
    return inter / ((double)(a1 + a2) - inter);
}
typedef struct { float s; int64_t i; } si64_t;
static int si64_cmp_desc_later_first(const void *pa, const void *pb) {
    const si64_t *a = (const si64_t *)pa, *b = (const si64_t *)pb;
    if (a->s != b->s) return a->s > b->s ? -1 : 1;
    return a->i > b->i ? -1 : (a->i < b->i ? 1 : 0);
}

ORC_API void orc_match_detections(const float *det_boxes, const float *det_scores, const int64_t *det_classes, int64_t N,
                                  const float *gt_boxes, const int64_t *gt_classes, const uint8_t *gt_difficult,
                                  const uint8_t *gt_group_of, int64_t M, int64_t num_classes, double match_iou,
                                  double nms_iou, int64_t nms_max, int8_t *label, uint8_t *corloc) {
    si64_t *order = (si64_t *)malloc(sizeof(si64_t) * (size_t)(N > 0 ? N : 1));
    int64_t *sel = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    uint8_t *gt_done = (uint8_t *)malloc((size_t)(M > 0 ? M : 1));
    for (int64_t i = 0; i < N; ++i) label[i] = -2;
    for (int64_t c = 0; c < num_classes; ++c) {
        corloc[c] = 0;
        /* _remove_invalid_boxes + _get_ith_class_arrays */
        int64_t n = 0, best = -1;
        for (int64_t i = 0; i < N; ++i) {
            const float *b = det_boxes + 4 * i;
            if (det_classes[i] != c || !(b[0] < b[2] && b[1] < b[3])) continue;
            if (best < 0 || det_scores[i] > det_scores[best]) best = i;          /* np.argmax: first maximum */
            if (det_scores[i] > -10.0f) { order[n].s = det_scores[i]; order[n].i = i; ++n; }   /* filter_scores_greater_than */
        }
        int64_t mc = 0;
        for (int64_t m = 0; m < M; ++m) mc += gt_classes[m] == c;
        /* CorLoc (:143-175): the best-scoring detection against every gt box of the class */
        if (best >= 0 && mc > 0) {
            double mx = -1.0;
            for (int64_t m = 0; m < M; ++m)
                if (gt_classes[m] == c) { const double v = eval_iou(det_boxes + 4 * best, gt_boxes + 4 * m); if (v > mx) mx = v; }
            corloc[c] = mx >= match_iou;
        }
        if (n == 0) continue;
        qsort(order, (size_t)n, sizeof(si64_t), si64_cmp_desc_later_first);
        /* non_max_suppression (np_box_list.py:328-396) */
        int64_t k = 0;
        if (nms_iou >= 1.0) {
            for (int64_t j = 0; j < n && k < nms_max; ++j) sel[k++] = order[j].i;
        } else {
            for (int64_t j = 0; j < n && k < nms_max; ++j) {
                int keep = 1;
                for (int64_t q = 0; q < k && keep; ++q)
                    if (eval_iou(det_boxes + 4 * sel[q], det_boxes + 4 * order[j].i) > nms_iou) keep = 0;
                if (keep) sel[k++] = order[j].i;
            }
        }
        for (int64_t j = 0; j < k; ++j) label[sel[j]] = 0;
        if (mc == 0) continue;                      /* :365-366: no gt of the class, every detection is a false positive */
        memset(gt_done, 0, (size_t)(M > 0 ? M : 1));
        /* compute_match_iou (:379-407) over the non-group-of boxes, then compute_match_ioa (:409-441) */
        for (int64_t j = 0; j < k; ++j) {
            const float *d = det_boxes + 4 * sel[j];
            int64_t gid = -1;
            double gv = -1.0;
            for (int64_t m = 0; m < M; ++m) {
                if (gt_classes[m] != c || (gt_group_of && gt_group_of[m])) continue;
                const double v = eval_iou(d, gt_boxes + 4 * m);
                if (v > gv) { gv = v; gid = m; }                                  /* np.argmax: first maximum */
            }
            if (gid >= 0 && gv >= match_iou) {
                if (!(gt_difficult && gt_difficult[gid])) {
                    if (!gt_done[gid]) { label[sel[j]] = 1; gt_done[gid] = 1; }
                } else {
                    label[sel[j]] = -1;
                }
            }
        }
        if (gt_group_of) {
            for (int64_t j = 0; j < k; ++j) {
                if (label[sel[j]] != 0) continue;   /* not a true positive, not matched to a difficult box */
                const float *d = det_boxes + 4 * sel[j];
                int64_t gid = -1;
                double gv = -1.0;
                for (int64_t m = 0; m < M; ++m) {
                    if (gt_classes[m] != c || !gt_group_of[m]) continue;
                    const double v = eval_inter(gt_boxes + 4 * m, d) / (double)area_yxyx(d);   /* ioa(gt, det) */
                    if (v > gv) { gv = v; gid = m; }
                }
                if (gid >= 0 && gv >= match_iou) label[sel[j]] = -1;
            }
        }
    }
    free(order); free(sel); free(gt_done);
}
