// Per-image true/false-positive matching and CorLoc on the device: what the reference's numpy
// evaluator does image by image on the host after a .cpu().numpy() per image
// (effdet/evaluation/per_image_evaluation.py:29-92, :93-175, :177-240, :305-470; np_box_list.py:128-205,
// :297-396; caller effdet/evaluation/detection_evaluator.py:268-305).  See include/odk.h (odk_match_detections).
//
// One CTA per image, everything in shared memory.  The greedy rule is sequential in score order, but the only
// state it carries is one "already detected" bit per gt box, and which gt box a detection would claim does not
// depend on that state -- so the expensive part (every detection's arg-max IoU over the gt boxes of its class)
// is computed for all detections in parallel, and a single thread then walks the sorted list once.
// Arithmetic as numpy runs it on float32 boxes: areas and coordinate differences rounded in fp32, intersection,
// union and the ratios in float64 (np.maximum(np.zeros(...), diff) promotes), so the >= / > decisions are the
// reference's bit for bit.
#include "odk_detect.cuh"

namespace odk {

constexpr int kEvalThreads = 256;
constexpr int kEvalMaxD = 1024, kEvalMaxM = 1024, kEvalMaxC = 4096;

__device__ __forceinline__ float eval_area(float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }   // yxyx
__device__ __forceinline__ double eval_inter(float4 p, float4 q) {
    const float h32 = __fsub_rn(fminf(p.z, q.z), fmaxf(p.x, q.x));
    const float w32 = __fsub_rn(fminf(p.w, q.w), fmaxf(p.y, q.y));
    const double h = h32 > 0.0f ? (double)h32 : 0.0, w = w32 > 0.0f ? (double)w32 : 0.0;
    return h * w;
}
__device__ __forceinline__ double eval_iou(float4 p, float4 q) {
    const double inter = eval_inter(p, q);
    return inter / ((double)__fadd_rn(eval_area(p), eval_area(q)) - inter);
}

struct EvalArgs {
    const float *dets;          // [B,D,6] x0,y0,x1,y1,score,class (odk_detect layout; class in the gt labels' numbering)
    const int *count;           // [B] rows in use, or null (= D)
    const float4 *gt_boxes;     // [B,M] yxyx
    const int *gt_labels;       // [B,M]; < 0 = padding row
    const unsigned char *gt_difficult, *gt_group_of;   // [B,M] or null
    int B, D, M, C, label_offset;
    double match_iou, nms_iou;
    int nms_max;
    signed char *label;         // [B,D]
    unsigned char *corloc;      // [B,C]
};

__global__ void __launch_bounds__(kEvalThreads) match_kernel(const __grid_constant__ EvalArgs A) {
    __shared__ float4 s_det[kEvalMaxD];                 // yxyx
    __shared__ float s_score[kEvalMaxD];
    __shared__ int s_cls[kEvalMaxD];                    // 0-based class, -1 = not a detection / invalid box
    __shared__ unsigned long long s_key[kEvalMaxD];     // sort keys, then the sorted order
    __shared__ short s_gid[kEvalMaxD], s_goid[kEvalMaxD];   // arg-max gt over the non-group-of / group-of boxes of the class
    __shared__ unsigned char s_ghit[kEvalMaxD];         // bit 0: best IoU >= thr, bit 1: best IoA >= thr
    __shared__ signed char s_label[kEvalMaxD];
    __shared__ unsigned char s_gdone[kEvalMaxM];
    __shared__ int s_n;
    extern __shared__ __align__(16) unsigned char s_dyn[];   // gt boxes [M] float4, gt class [M] int, per-class NMS counters [C]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int D = A.D, M = A.M, C = A.C;
    float4 *s_gt = reinterpret_cast<float4 *>(s_dyn);
    int *s_gcls = reinterpret_cast<int *>(s_dyn + (size_t)M * 16);
    int *s_ccnt = s_gcls + M;
    const int n_det = A.count ? min(max(__ldg(A.count + b), 0), D) : D;
    for (int m = tid; m < M; m += kEvalThreads) {
        s_gt[m] = __ldg(A.gt_boxes + (size_t)b * M + m);
        const int l = __ldg(A.gt_labels + (size_t)b * M + m);
        s_gcls[m] = l < 0 ? -1 : l - A.label_offset;
        s_gdone[m] = 0;
    }
    for (int c = tid; c < C; c += kEvalThreads) s_ccnt[c] = 0;
    // 1. detections: xyxy -> yxyx, _remove_invalid_boxes (:512-536), the score filter of non_max_suppression
    int P = 2;
    while (P < D) P <<= 1;
    for (int i = tid; i < P; i += kEvalThreads) {
        unsigned long long key = 0ull;
        if (i < D) {
            int cls = -1;
            float sc = 0.f;
            float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_det) {
                const float *r = A.dets + ((size_t)b * D + i) * 6;
                box = make_float4(__ldg(r + 1), __ldg(r), __ldg(r + 3), __ldg(r + 2));
                sc = __ldg(r + 4);
                const int c = (int)__ldg(r + 5) - A.label_offset;
                if (box.x < box.z && box.y < box.w && c >= 0 && c < C) cls = c;
            }
            s_det[i] = box; s_score[i] = sc; s_cls[i] = cls;
            s_label[i] = -2;
            if (cls >= 0 && sc > -10.0f) {   // descending score; equal scores: the later detection first
                const unsigned u = __float_as_uint(sc);
                const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
                key = ((unsigned long long)vk << 32) | (unsigned long long)(unsigned)(i + 1);
            }
        }
        s_key[i] = key;
    }
    __syncthreads();
    bitonic_sort_desc_u64(s_key, P);
    if (tid == 0) {
        int n = 0;
        while (n < D && s_key[n] != 0ull) ++n;
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    // 2. per detection, independent of the matching state: the gt box it would claim (np.argmax: first maximum)
    for (int j = tid; j < n; j += kEvalThreads) {
        const int i = (int)(s_key[j] & 0xFFFFFFFFull) - 1;
        const float4 d = s_det[i];
        const int c = s_cls[i];
        int gid = -1, goid = -1;
        double gv = -1.0, gov = -1.0;
        for (int m = 0; m < M; ++m) {
            if (s_gcls[m] != c) continue;
            const bool go = A.gt_group_of && __ldg(A.gt_group_of + (size_t)b * M + m);
            if (!go) {
                const double v = eval_iou(d, s_gt[m]);
                if (v > gv) { gv = v; gid = m; }
            } else {
                const double v = eval_inter(s_gt[m], d) / (double)eval_area(d);   // ioa(gt, detection), :248-262
                if (v > gov) { gov = v; goid = m; }
            }
        }
        s_gid[i] = (short)gid; s_goid[i] = (short)goid;
        s_ghit[i] = (unsigned char)((gid >= 0 && gv >= A.match_iou ? 1 : 0) | (goid >= 0 && gov >= A.match_iou ? 2 : 0));
    }
    __syncthreads();
    // 3. one walk in score order: the class's own NMS (np_box_list.py:328-396; off at iou 1.0), compute_match_iou
    //    (:379-407) and compute_match_ioa (:409-441)
    if (tid < 32) {
        const int lane = tid;
        for (int j = 0; j < n; ++j) {
            const int i = (int)(s_key[j] & 0xFFFFFFFFull) - 1;
            const int c = s_cls[i];
            bool keep = s_ccnt[c] < A.nms_max;
            if (keep && A.nms_iou < 1.0) {
                bool hit = false;
                for (int q = lane; q < j && !hit; q += 32) {
                    const int k = (int)(s_key[q] & 0xFFFFFFFFull) - 1;
                    if (s_label[k] != -2 && s_cls[k] == c) hit = eval_iou(s_det[k], s_det[i]) > A.nms_iou;
                }
                keep = !__any_sync(0xffffffffu, hit);
            }
            if (lane == 0 && keep) {
                ++s_ccnt[c];
                signed char lab = 0;
                const int gid = s_gid[i];
                bool has_gt = false;
                for (int m = 0; m < M && !has_gt; ++m) has_gt = s_gcls[m] == c;   // (:365-366: no gt of the class -> all false positives)
                if (has_gt && (s_ghit[i] & 1)) {
                    if (!(A.gt_difficult && __ldg(A.gt_difficult + (size_t)b * M + gid))) {
                        if (!s_gdone[gid]) { lab = 1; s_gdone[gid] = 1; }
                    } else {
                        lab = -1;
                    }
                }
                if (has_gt && lab == 0 && (s_ghit[i] & 2)) lab = -1;   // matched to a group-of box: ignored (weight 0)
                s_label[i] = lab;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < D; i += kEvalThreads) A.label[(size_t)b * D + i] = s_label[i];
    // 4. CorLoc (:93-175): per class, the best-scoring valid detection (np.argmax: first) against every gt box of the class
    for (int c = tid; c < C; c += kEvalThreads) {
        int best = -1;
        for (int i = 0; i < n_det; ++i)
            if (s_cls[i] == c && (best < 0 || s_score[i] > s_score[best])) best = i;
        unsigned char ok = 0;
        if (best >= 0) {
            double mx = -1.0;
            bool any = false;
            for (int m = 0; m < M; ++m)
                if (s_gcls[m] == c) { any = true; mx = fmax(mx, eval_iou(s_det[best], s_gt[m])); }
            ok = any && mx >= A.match_iou;
        }
        A.corloc[(size_t)b * C + c] = ok;
    }
}

}  // namespace odk

extern "C" {

int odk_match_detections(const float *dets, const int32_t *count, int B, int D, const float *gt_boxes, const int32_t *gt_labels,
                         const uint8_t *gt_difficult, const uint8_t *gt_group_of, int M, int num_classes, int label_offset,
                         double match_iou, double nms_iou, int nms_max, int8_t *label, uint8_t *corloc, void *stream) {
    using namespace odk;
    if (B < 0 || D < 0 || M < 0 || num_classes < 1) return set_error(ODK_EINVAL, "odk_match_detections: bad sizes");
    if (B == 0 || D == 0) return ODK_OK;
    if (!dets || !label || !corloc || (M > 0 && (!gt_boxes || !gt_labels))) return set_error(ODK_EINVAL, "odk_match_detections: null pointer");
    if (D > kEvalMaxD || M > kEvalMaxM || num_classes > kEvalMaxC)
        return set_error(ODK_EUNSUPPORTED, "odk_match_detections: at most %d detections, %d gt boxes, %d classes per image", kEvalMaxD, kEvalMaxM, kEvalMaxC);
    if ((uintptr_t)gt_boxes & 15) return set_error(ODK_EINVAL, "odk_match_detections: gt_boxes must be 16-byte aligned");
    if (!(nms_iou >= 0.0 && nms_iou <= 1.0)) return set_error(ODK_EINVAL, "odk_match_detections: IOU threshold must be in [0, 1]");   // np_box_list.py:355
    if (nms_max < 0) return set_error(ODK_EINVAL, "odk_match_detections: max_output_size must be bigger than 0.");
    if (B > 65535) return set_error(ODK_EUNSUPPORTED, "odk_match_detections: batch > 65535");
    EvalArgs a;
    a.dets = dets; a.count = count; a.gt_boxes = (const float4 *)gt_boxes; a.gt_labels = gt_labels;
    a.gt_difficult = gt_difficult; a.gt_group_of = gt_group_of;
    a.B = B; a.D = D; a.M = M; a.C = num_classes; a.label_offset = label_offset;
    a.match_iou = match_iou; a.nms_iou = nms_iou; a.nms_max = nms_max; a.label = label; a.corloc = corloc;
    const size_t smem = (size_t)M * 20 + (size_t)num_classes * 4 + 16;
    cudaError_t e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, "odk_match_detections: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    match_kernel<<<B, kEvalThreads, smem, (cudaStream_t)stream>>>(a);
    return check_launch("odk_match_detections");
}

}  // extern "C"
