// Detection generation (K4/K5/K6): box decode + score filter + class-aware NMS or Soft-NMS, and
// the per-detection OOD scores.  Replaces generate_detections / _batch_detection
// (reference effdet/anchors.py:95-172, effdet/bench.py:59-76), torchvision batched_nms
// (coordinate trick) and effdet/soft_nms.py.  See include/odk.h.
//
// One CTA (1024 threads) per image; all candidates of the image (<= 8192) live in shared
// memory.  Both suppressors are "frontier" algorithms that need at most max_det rounds instead
// of the O(n^2) mask of the classic bitmask NMS, because only the first max_det survivors are
// ever used (anchors.py:153):
//   hard NMS : round = take the first 16 still-alive candidates (score order), decide among them which are
//              kept (the greedy rule), kill every later candidate whose IoU with a newly kept one exceeds
//              the threshold; a warp owns 32 consecutive candidates = one word of the alive bitmask and
//              updates it with a ballot, so there are no atomics;
//   Soft-NMS : round = block-wide arg-max of the current scores (first index on ties), record it,
//              decay every alive score by exp(-iou^2/sigma), drop those at or below the score
//              threshold (soft_nms.py:88-110).
// All IoU arithmetic is in the reference's fp32 operation order on the class-offset boxes.
#include "odk_detect.cuh"

namespace odk {

struct DetArgs {
    const float *cls;       // [B,N]
    const float4 *box;      // [B,N]
    const long long *idx;   // [B,N] anchor index
    const long long *klass; // [B,N]
    const float4 *anchors;  // [A]
    const float *scale;     // [B] or null
    const float *size;      // [B,2] or null
    int N, cap;
    long long A;
    odk_detect_params p;
    float nms_thr_f;
    float *dets;            // [B,D,6]
    int *count;             // [B]
    int *src;               // [B,D]
    long long *det_anchor;  // [B,D] anchor index of every detection (-1 padded), or null
    const unsigned *only_flag;   // non-null: only images with only_flag[b] != 0 are processed
};

__global__ void __launch_bounds__(kDetThreads) detect_kernel(const __grid_constant__ DetArgs A) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_wcnt[kDetWarps];
    __shared__ float s_wmax[kDetWarps];
    __shared__ int s_kept[1024];
    __shared__ float s_keptscore[1024];
    __shared__ int s_n, s_unsorted;
    const DetSmem S = carve(s_raw, A.cap);
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_raw);   // aliases S.box until step 4
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (A.only_flag && __ldg(A.only_flag + b) == 0u) return;   // uniform
    const int N = A.N, D = A.p.max_det;
    const float *cls = A.cls + (size_t)b * N;
    const float4 *box = A.box + (size_t)b * N;
    const long long *idx = A.idx + (size_t)b * N;
    const long long *klass = A.klass + (size_t)b * N;
    const bool has_scale = A.scale != nullptr;
    const bool clip = has_scale && A.size != nullptr;                       // anchors.py:137
    const float scale = has_scale ? __ldg(A.scale + b) : 1.0f;
    const float lim_x = clip ? __fdiv_rn(__ldg(A.size + 2 * b), scale) : 0.f;
    const float lim_y = clip ? __fdiv_rn(__ldg(A.size + 2 * b + 1), scale) : 0.f;
    if (tid == 0) s_unsorted = 0;

    // 1. scores, score filter, order-preserving compaction (anchors.py:140-144)
    int total = 0;
    for (int base = 0; base < N; base += kDetThreads) {
        const int p = base + tid;
        float sc = 0.f;
        bool ok = false;
        if (p < N) { sc = sigmoid_ref(__ldg(cls + p)); ok = sc > A.p.score_min; }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = total, chunk = 0;
        for (int w = 0; w < kDetWarps; ++w) {
            const int c = s_wcnt[w];
            if (w < warp) off += c;
            chunk += c;
        }
        if (ok) {
            const int i = off + __popc(bal & ((1u << lane) - 1u));
            s_key[i] = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)p);
        }
        total += chunk;
        __syncthreads();
    }
    const int n = total;
    int kept_n = 0;
    if (n > 0) {
        // 2. torchvision::nms orders by descending score (stable); top-k output already is
        {
            int bad = 0;
            for (int i = tid; i + 1 < n; i += kDetThreads) bad |= (s_key[i] >> 32) < (s_key[i + 1] >> 32);
            if (bad) s_unsorted = 1;
            __syncthreads();
            if (s_unsorted && !A.p.soft_nms) {
                int P = 2;
                while (P < n) P <<= 1;
                for (int i = n + tid; i < P; i += kDetThreads) s_key[i] = 0ull;
                __syncthreads();
                bitonic_sort_desc_u64(s_key, P);
            }
        }
        // 3. keys -> registers, then decode boxes into the (aliased) box array
        unsigned long long mykeys[kDetMaxN / kDetThreads];
#pragma unroll
        for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
            const int i = tid + k * kDetThreads;
            mykeys[k] = i < n ? s_key[i] : 0ull;
        }
        __syncthreads();
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
            const int i = tid + k * kDetThreads;
            if (i < n) {
                const int p = (int)(0xFFFFFFFFu - (unsigned)(mykeys[k] & 0xFFFFFFFFull));
                const float4 o = decode_xyxy(__ldg(A.anchors + __ldg(idx + p)), __ldg(box + p), clip, lim_x, lim_y);
                S.box[i] = o;
                S.score[i] = __uint_as_float((unsigned)(mykeys[k] >> 32));
                S.src[i] = p;
                mx = fmaxf(mx, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
            }
        }
        // 4. coordinate trick: boxes + class * (max_coordinate + 1)  (torchvision boxes.py:105-108)
        mx = warp_max(mx);
        if (lane == 0) s_wmax[warp] = mx;
        init_alive(S.alive, n, A.cap);
        __syncthreads();
        mx = s_wmax[0];
        for (int w = 1; w < kDetWarps; ++w) mx = fmaxf(mx, s_wmax[w]);
        const float mul = __fadd_rn(mx, 1.0f);
        for (int i = tid; i < n; i += kDetThreads) {
            const float off = __fmul_rn((float)__ldg(klass + S.src[i]), mul);
            float4 o = S.box[i];
            o.x = __fadd_rn(o.x, off); o.y = __fadd_rn(o.y, off); o.z = __fadd_rn(o.z, off); o.w = __fadd_rn(o.w, off);
            S.box[i] = o;
        }
        __syncthreads();
        // 5. suppression, first D survivors
        if (A.p.soft_nms && s_unsorted)
            kept_n = soft_nms_rounds(S, n, true, A.p.soft_sigma, A.p.soft_iou, A.p.soft_score_thr, D, s_kept, n, n, kDetWarps,
                                     [&](int q, int i, float sc) { s_keptscore[q] = sc; });
        else if (A.p.soft_nms)
            // the lazy window and the batches need non-increasing scores: true for top-k output (checked above for API inputs)
            kept_n = soft_nms_batched(S, n, true, A.p.soft_sigma, A.p.soft_iou, A.p.soft_score_thr, D, s_kept,
                                      kDetFirstWindow, kDetThreads, [&](int q, int i, float sc) { s_keptscore[q] = sc; },
                                      s_raw + (size_t)A.cap * 24 + (size_t)(A.cap / 32) * 4 + 16);
        else
            kept_n = hard_nms_rounds(S, n, A.nms_thr_f, D, s_kept, reinterpret_cast<unsigned *>(s_raw + (size_t)A.cap * 24 + (size_t)(A.cap / 32) * 4 + 16));
        __syncthreads();
    }
    // 6. rows: boxes (re-decoded, unoffset) * img_scale, score, class + 1 (anchors.py:153-166)
    float *dets = A.dets + (size_t)b * D * 6;
    int *src = A.src + (size_t)b * D;
    for (int q = tid; q < D; q += kDetThreads) {
        float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int sp = -1;
        if (q < kept_n) {
            const int i = s_kept[q];
            sp = S.src[i];
            float4 o = decode_xyxy(__ldg(A.anchors + __ldg(idx + sp)), __ldg(box + sp), clip, lim_x, lim_y);
            if (has_scale) { o.x = __fmul_rn(o.x, scale); o.y = __fmul_rn(o.y, scale); o.z = __fmul_rn(o.z, scale); o.w = __fmul_rn(o.w, scale); }
            r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = o.w;
            r[4] = A.p.soft_nms ? s_keptscore[q] : S.score[i];
            r[5] = (float)(__ldg(klass + sp) + 1);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) dets[q * 6 + k] = r[k];
        src[q] = sp;
        if (A.det_anchor) A.det_anchor[(size_t)b * D + q] = sp >= 0 ? __ldg(idx + sp) : -1ll;
    }
    if (tid == 0) A.count[b] = kept_n;
}

// ---- stand-alone soft_nms / nms on one box set ---------------------------------------------------
__global__ void __launch_bounds__(kDetThreads)
soft_nms_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores, int n, int cap, int gaussian,
                float sigma, float iou_thr, float score_thr, int max_rounds, long long *__restrict__ idx_out,
                float *__restrict__ score_out, int *__restrict__ count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const DetSmem S = carve(s_raw, cap);
    for (int i = threadIdx.x; i < n; i += kDetThreads) { S.box[i] = __ldg(boxes + i); S.score[i] = __ldg(scores + i); }
    init_alive(S.alive, n, cap);
    __syncthreads();
    __shared__ int s_picked[kDetMaxN];   // 32 KB: ranks in pick order (needed to replay decays)
    const int c = soft_nms_rounds(S, n, gaussian != 0, sigma, iou_thr, score_thr, max_rounds, s_picked, n, n, kDetWarps,
                                  [&](int q, int i, float sc) { idx_out[q] = i; score_out[q] = sc; });
    if (threadIdx.x == 0) *count = c;
}

__global__ void __launch_bounds__(kDetThreads)
nms_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores, int n, int cap, float thr_f,
           long long *__restrict__ keep, int *__restrict__ count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_kept[kDetMaxN];   // 32 KB: every box may survive
    DetSmem S;                         // no score array here: boxes, source index, alive words, window masks
    S.box = reinterpret_cast<float4 *>(s_raw);
    S.score = nullptr;
    S.src = reinterpret_cast<int *>(s_raw + (size_t)cap * 16);
    S.alive = reinterpret_cast<unsigned *>(s_raw + (size_t)cap * 20);
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_raw);
    int P = 2;
    while (P < n) P <<= 1;
    // stable descending order on arbitrary (possibly negative) scores: order-preserving key
    for (int i = threadIdx.x; i < P; i += kDetThreads) {
        unsigned long long k = 0ull;
        if (i < n) {
            const unsigned u = __float_as_uint(__ldg(scores + i));
            const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
            k = ((unsigned long long)vk << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        }
        s_key[i] = k;
    }
    __syncthreads();
    bitonic_sort_desc_u64(s_key, P);
    int mine[kDetMaxN / kDetThreads];
#pragma unroll
    for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
        const int i = threadIdx.x + k * kDetThreads;
        mine[k] = i < n ? (int)(0xFFFFFFFFu - (unsigned)(s_key[i] & 0xFFFFFFFFull)) : -1;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
        const int i = threadIdx.x + k * kDetThreads;
        if (i < n) { S.box[i] = __ldg(boxes + mine[k]); S.src[i] = mine[k]; }
    }
    init_alive(S.alive, n, cap);
    __syncthreads();
    const int c = hard_nms_rounds(S, n, thr_f, n, s_kept, reinterpret_cast<unsigned *>(s_raw + (size_t)cap * 20 + (size_t)(cap / 32) * 4 + 16));
    __syncthreads();
    for (int q = threadIdx.x; q < c; q += kDetThreads) keep[q] = S.src[s_kept[q]];
    if (threadIdx.x == 0) *count = c;
}

// ---- OOD scores --------------------------------------------------------------------------------
// one warp per detection: energy = -T * logsumexp(row / T), max_logit = max(row)
struct LevelPtrs { const float *p[ODK_MAX_LEVELS]; unsigned char nhwc[ODK_MAX_LEVELS]; };

__global__ void __launch_bounds__(256)
ood_kernel(const __grid_constant__ Geo g, const __grid_constant__ LevelPtrs lv, int B, int C, const long long *__restrict__ anchor_idx, int D, float T,
           float *__restrict__ energy, float *__restrict__ max_logit, const unsigned *__restrict__ only_flag) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= B * D) return;
    const int b = wid / D;
    if (only_flag && __ldg(only_flag + b) == 0u) return;
    float e, m;
    ood_row(g, lv.p, lv.nhwc, b, C, __ldg(anchor_idx + wid), T, lane, e, m);
    if (lane == 0) { energy[wid] = e; max_logit[wid] = m; }
}

static inline size_t nms_smem_bytes(int cap) { return (size_t)cap * 20 + (size_t)(cap / 32) * 4 + 16 + kNmsMaskBytes; }

// the opt-in is per device and per launch size (cudaFuncSetAttribute is cheap): a process may use several GPUs
template <class Kernel>
static int smem_opt_in(Kernel k, size_t bytes, const char *what) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error((int)e, "%s: %zu bytes of shared memory: %s", what, bytes, cudaGetErrorString(e));
    return ODK_OK;
}

int launch_detect_flagged(const float *cls_topk, const float *box_topk, const int64_t *indices, const int64_t *classes, int B,
                          int N, const float *anchors, int64_t A, const float *img_scale, const float *img_size,
                          const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, int64_t *det_anchor,
                          const unsigned *only_flag, cudaStream_t st) {
    if (N > kDetMaxN) return set_error(ODK_EUNSUPPORTED, "odk_detect: more than %d candidates per image", kDetMaxN);
    if (((uintptr_t)box_topk | (uintptr_t)anchors) & 15) return set_error(ODK_EINVAL, "odk_detect: box_topk / anchors must be 16-byte aligned");
    DetArgs a;
    memset(&a, 0, sizeof(a));
    a.cls = cls_topk; a.box = (const float4 *)box_topk; a.idx = (const long long *)indices; a.klass = (const long long *)classes;
    a.anchors = (const float4 *)anchors; a.scale = img_scale; a.size = img_size; a.N = N; a.cap = det_cap(N); a.A = A;
    a.p = *params; a.nms_thr_f = float_at_or_below(params->nms_iou);
    a.dets = dets; a.count = count; a.src = src; a.det_anchor = (long long *)det_anchor; a.only_flag = only_flag;
    int rc = smem_opt_in(detect_kernel, det_smem_bytes(a.cap), "odk_detect");
    if (rc) return rc;
    detect_kernel<<<B, kDetThreads, det_smem_bytes(a.cap), st>>>(a);
    return check_launch("odk_detect");
}

int launch_ood_flagged(const Geo &g, const void *const *cls_levels, int layout, int B, int C, const int64_t *anchor_idx, int D, float T,
                       float *energy, float *max_logit, const unsigned *only_flag, cudaStream_t st) {
    LevelPtrs lv;
    memset(&lv, 0, sizeof(lv));
    for (int l = 0; l < g.nlev; ++l) {
        lv.p[l] = (const float *)cls_levels[l];
        lv.nhwc[l] = (layout >> l) & 1;
        if (!lv.p[l]) return set_error(ODK_EINVAL, "odk_ood: null level pointer (level %d)", l);
    }
    const long long warps = (long long)B * D;
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    ood_kernel<<<blocks, 256, 0, st>>>(g, lv, B, C, (const long long *)anchor_idx, D, T, energy, max_logit, only_flag);
    return check_launch("odk_ood");
}

}  // namespace odk

extern "C" {

int odk_detect(const float *cls_topk, const float *box_topk, const int64_t *indices, const int64_t *classes, int B,
               int N, const float *anchors, int64_t A, const float *img_scale, const float *img_size,
               const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, void *stream) {
    using namespace odk;
    if (!params) return set_error(ODK_EINVAL, "odk_detect: null params");
    if (B < 0 || N < 0) return set_error(ODK_EINVAL, "odk_detect: negative size");
    if (B == 0) return ODK_OK;
    if (!dets || !count || !src || !anchors || (N > 0 && (!cls_topk || !box_topk || !indices || !classes)))
        return set_error(ODK_EINVAL, "odk_detect: null pointer");
    if (params->max_det < 1 || params->max_det > 1024) return set_error(ODK_EUNSUPPORTED, "odk_detect: max_det must be in [1,1024]");
    if (((uintptr_t)box_topk | (uintptr_t)anchors) & 15) return set_error(ODK_EINVAL, "odk_detect: box_topk / anchors must be 16-byte aligned");
    return launch_detect_flagged(cls_topk, box_topk, indices, classes, B, N, anchors, A, img_scale, img_size, params, dets,
                                 count, src, nullptr, nullptr, (cudaStream_t)stream);
}

int odk_soft_nms(const float *boxes, const float *scores, int n, int method_gaussian, float sigma, float iou_thr,
                 float score_thr, int max_rounds, int64_t *idx_out, float *score_out, int32_t *count, void *stream) {
    using namespace odk;
    if (n < 0 || !count) return set_error(ODK_EINVAL, "odk_soft_nms: bad arguments");
    if (n > kDetMaxN) return set_error(ODK_EUNSUPPORTED, "odk_soft_nms: more than %d boxes", kDetMaxN);
    if (n > 0 && (!boxes || !scores || !idx_out || !score_out)) return set_error(ODK_EINVAL, "odk_soft_nms: null pointer");
    if ((uintptr_t)boxes & 15) return set_error(ODK_EINVAL, "odk_soft_nms: boxes must be 16-byte aligned");
    if (max_rounds < 0 || max_rounds > n) max_rounds = n;
    const int cap = det_cap(n);
    int rc = smem_opt_in(soft_nms_kernel, det_base_bytes(cap), "odk_soft_nms");
    if (rc) return rc;
    soft_nms_kernel<<<1, kDetThreads, det_base_bytes(cap), (cudaStream_t)stream>>>(
        (const float4 *)boxes, scores, n, cap, method_gaussian, sigma, iou_thr, score_thr, max_rounds,
        (long long *)idx_out, score_out, count);
    return check_launch("odk_soft_nms");
}

size_t odk_nms_workspace_bytes(int n) { (void)n; return 16; }

int odk_nms(const float *boxes, const float *scores, int n, double iou_thr, int64_t *keep, int32_t *count,
            void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    (void)workspace; (void)workspace_bytes;
    if (n < 0 || !count) return set_error(ODK_EINVAL, "odk_nms: bad arguments");
    if (n > kDetMaxN) return set_error(ODK_EUNSUPPORTED, "odk_nms: more than %d boxes", kDetMaxN);
    if (n > 0 && (!boxes || !scores || !keep)) return set_error(ODK_EINVAL, "odk_nms: null pointer");
    if ((uintptr_t)boxes & 15) return set_error(ODK_EINVAL, "odk_nms: boxes must be 16-byte aligned");
    const int cap = det_cap(n);
    int rc = smem_opt_in(nms_kernel, nms_smem_bytes(cap), "odk_nms");
    if (rc) return rc;
    nms_kernel<<<1, kDetThreads, nms_smem_bytes(cap), (cudaStream_t)stream>>>((const float4 *)boxes, scores, n, cap,
                                                                             float_at_or_below(iou_thr), (long long *)keep, count);
    return check_launch("odk_nms");
}

int odk_ood(const void *const *cls_levels, int B, int C, const int32_t *level_hw, int num_levels, int na, int layout,
            const int64_t *anchor_idx, int D, float temperature, float *energy, float *max_logit, void *stream) {
    using namespace odk;
    Geo g;
    int rc = make_geo(&g, level_hw, num_levels, na);
    if (rc) return rc;
    if (B < 0 || D < 0 || C < 1) return set_error(ODK_EINVAL, "odk_ood: bad sizes");
    if (B == 0 || D == 0) return ODK_OK;
    if (!cls_levels || !anchor_idx || !energy || !max_logit) return set_error(ODK_EINVAL, "odk_ood: null pointer");
    if (!(temperature > 0.0f)) return set_error(ODK_EINVAL, "odk_ood: temperature must be positive");
    return launch_ood_flagged(g, cls_levels, layout, B, C, anchor_idx, D, temperature, energy, max_logit, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
