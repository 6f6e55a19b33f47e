// Target assignment on the GPU (K1): gt x anchor IoU matching with force-match, bit-exact
// against the reference's torch path.  See include/odk.h (odk_assign, odk_targets).
//
// Layout / mapping
//   * one thread per anchor, threads ordered in PLANAR order (level, anchor shape, y, x): a CTA's
//     256 anchors are spatially adjacent boxes of one shape, so its bounding box is small and
//     most gt boxes can be rejected for the whole CTA at once (IoU is exactly 0 when the boxes
//     do not overlap, so the cull never changes a result while match_thr > 0);
//   * the surviving gt boxes of an image are staged in shared memory (ordered compaction keeps
//     ascending gt order, which is what "first index wins" ties need);
//   * per-gt arg-max over anchors (force_match_for_each_row) is a 64-bit atomicMax on
//     (iou_bits << 32 | ~anchor_index): highest IoU, then LOWEST reference anchor index;
//     a shared-memory copy per CTA filters almost all candidates before the global atomic.
//   * a second tiny kernel applies the forced matches (lowest gt row wins a contested anchor,
//     argmax_matcher.py:141-143) and finishes num_positives.
#include "odk_common.cuh"

namespace odk {

constexpr int kAssignThreads = 256;
constexpr int kGtTile = 256;

__global__ void __launch_bounds__(kAssignThreads)
assign_kernel(const Geo g, const float4 *__restrict__ anchors, const float4 *__restrict__ gt_boxes,
              const int32_t *__restrict__ gt_labels, const int32_t *__restrict__ gt_count, int Mmax, float thr,
              int filter_valid, int cull, int32_t *__restrict__ match, unsigned long long *__restrict__ best,
              int32_t *__restrict__ pos_count) {
    __shared__ float s_red[4][kAssignThreads / 32];
    __shared__ float s_tile[4];
    __shared__ float4 s_box[kGtTile];
    __shared__ float s_area[kGtTile];
    __shared__ int s_idx[kGtTile];
    __shared__ unsigned long long s_best[kGtTile];
    __shared__ int s_wcnt[kAssignThreads / 32];
    __shared__ int s_pos;

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x * kAssignThreads + tid;
    const bool live = p < g.A;
    int r = 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float aarea = 0.f;
    if (live) {
        int l;
        r = planar_to_ref(g, p, l);
        a = __ldg(anchors + r);
        aarea = area_ref(a.x, a.y, a.z, a.w);
    }
    if (tid == 0) s_pos = 0;

    // this warp's own bounding box (lanes past the end hold +/-inf and never widen it)
    const float wy0 = warp_min(live ? a.x : INFINITY), wx0 = warp_min(live ? a.y : INFINITY);
    const float wy1 = warp_max(live ? a.z : -INFINITY), wx1 = warp_max(live ? a.w : -INFINITY);
    // bounding box of this CTA's anchors
    {
        if (lane == 0) { s_red[0][warp] = wy0; s_red[1][warp] = wx0; s_red[2][warp] = wy1; s_red[3][warp] = wx1; }
        __syncthreads();
        if (tid < 4) {
            float v = s_red[tid][0];
            for (int w = 1; w < kAssignThreads / 32; ++w) v = tid < 2 ? fminf(v, s_red[tid][w]) : fmaxf(v, s_red[tid][w]);
            s_tile[tid] = v;
        }
        __syncthreads();
    }
    const float ty0 = s_tile[0], tx0 = s_tile[1], ty1 = s_tile[2], tx1 = s_tile[3];

    int M = Mmax;
    if (gt_count) M = min(max(__ldg(gt_count + b), 0), Mmax);
    const float4 *gtb = gt_boxes + (size_t)b * Mmax;
    const int32_t *gtl = gt_labels + (size_t)b * Mmax;

    float best_iou = -1.0f;
    int best_g = -1;

    for (int base = 0; base < M; base += kGtTile) {
        // ---- stage the gt boxes that can overlap this CTA (ordered compaction) ----
        const int gi = base + tid;
        float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
        bool take = false;
        if (gi < M) {
            take = !filter_valid || __ldg(gtl + gi) >= 0;
            if (take) {
                gb = __ldg(gtb + gi);
                if (cull) take = (gb.z > ty0) && (gb.x < ty1) && (gb.w > tx0) && (gb.y < tx1);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int slot = __popc(bal & ((1u << lane) - 1u));
        int n = 0;
#pragma unroll
        for (int w = 0; w < kAssignThreads / 32; ++w) {
            const int c = s_wcnt[w];
            if (w < warp) slot += c;
            n += c;
        }
        if (take) {
            s_box[slot] = gb;
            s_area[slot] = area_ref(gb.x, gb.y, gb.z, gb.w);
            s_idx[slot] = gi;
            s_best[slot] = 0ull;
        }
        __syncthreads();

        // ---- every anchor against the staged gts, ascending gt order ----
        // A warp's 32 anchors are neighbours in one feature-map row, so most staged gts miss the
        // whole warp: reject those with one warp-uniform test on the warp's bounding box.
        for (int j = 0; j < n; ++j) {
            const float4 q = s_box[j];
            if (cull && !((q.z > wy0) && (q.x < wy1) && (q.w > wx0) && (q.y < wx1))) continue;   // warp-uniform
            float v = 0.0f;
            if (live) {
                // same arithmetic as iou_ref, with the (exact) early-outs for an empty intersection
                const float h = __fsub_rn(fminf(q.z, a.z), fmaxf(q.x, a.x));
                const float w = __fsub_rn(fminf(q.w, a.w), fmaxf(q.y, a.y));
                if (h > 0.0f && w > 0.0f) {
                    const float inter = __fmul_rn(h, w);
                    if (inter != 0.0f) v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(s_area[j], aarea), inter));
                }
                if (v > best_iou) { best_iou = v; best_g = s_idx[j]; }
            }
            // per-gt arg-max over anchors: reduce inside the warp, one shared atomic per warp at most
            unsigned long long key = 0ull;
            if (v > 0.0f)
                key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r);
            const unsigned long long cur = *(volatile unsigned long long *)&s_best[j];
            if (__any_sync(0xffffffffu, key > cur)) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                if (lane == 0) atomicMax(&s_best[j], key);
            }
        }
        __syncthreads();
        if (tid < n) {
            const unsigned long long k = s_best[tid];
            if (k != 0ull) atomicMax(best + (size_t)b * Mmax + s_idx[tid], k);
        }
        __syncthreads();
    }

    // thresholds (argmax_matcher.py:126-137 with matched == unmatched threshold)
    int m = -1;
    if (live) {
        if (best_g >= 0 && !(thr > best_iou)) m = best_g;
        match[(size_t)b * g.Apad + p] = m;
    }
    const unsigned posb = __ballot_sync(0xffffffffu, m >= 0);
    if (lane == 0 && posb) atomicAdd(&s_pos, __popc(posb));
    __syncthreads();
    if (tid == 0 && s_pos) atomicAdd(pos_count + b, s_pos);
}

// Forced matches: gt row i claims its arg-max anchor (anchor 0 if its IoU is 0 everywhere);
// the lowest gt row wins a contested anchor.  One CTA per image.
__global__ void __launch_bounds__(128)
assign_force_kernel(const Geo g, const int32_t *__restrict__ gt_labels, const int32_t *__restrict__ gt_count, int Mmax,
                    int filter_valid, const unsigned long long *__restrict__ best,
                    const int32_t *__restrict__ pos_count, int32_t *__restrict__ match, float *__restrict__ num_pos) {
    extern __shared__ int s_p[];
    __shared__ int s_extra;
    const int b = blockIdx.x;
    int M = Mmax;
    if (gt_count) M = min(max(__ldg(gt_count + b), 0), Mmax);
    if (threadIdx.x == 0) s_extra = 0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        int p = -1;
        if (!filter_valid || __ldg(gt_labels + (size_t)b * Mmax + i) >= 0) {
            const unsigned long long k = best[(size_t)b * Mmax + i];
            const int r = k ? (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)) : 0;
            int l;
            p = ref_to_planar(g, r, l);
        }
        s_p[i] = p;
    }
    __syncthreads();
    int extra = 0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const int p = s_p[i];
        if (p < 0) continue;
        bool win = true;
        for (int j = 0; j < i; ++j)
            if (s_p[j] == p) { win = false; break; }
        if (win) {
            int32_t *mp = match + (size_t)b * g.Apad + p;
            if (*mp < 0) ++extra;
            *mp = i;
        }
    }
    if (extra) atomicAdd(&s_extra, extra);
    __syncthreads();
    if (threadIdx.x == 0) num_pos[b] = (float)(pos_count[b] + s_extra);
}

// Reference-layout targets from `match` (one thread per anchor in REFERENCE order so the
// int64 / float4 stores are coalesced).
__global__ void __launch_bounds__(256)
targets_kernel(const Geo g, int B, const float4 *__restrict__ anchors, const float4 *__restrict__ gt_boxes,
               const int32_t *__restrict__ gt_labels, int Mmax, const int32_t *__restrict__ match,
               int64_t *__restrict__ cls_targets, float4 *__restrict__ box_targets) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.A) return;
    int l;
    const int p = ref_to_planar(g, r, l);
    const int m = __ldg(match + (size_t)b * g.Apad + p);
    const size_t o = (size_t)B * g.off[l] + (size_t)b * g.hw[l] * g.na + (r - g.off[l]);
    int64_t c = -1;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m >= 0) {
        c = (int64_t)__ldg(gt_labels + (size_t)b * Mmax + m) - 1;  // anchors.py:416
        t = encode_ref(__ldg(gt_boxes + (size_t)b * Mmax + m), __ldg(anchors + r));
    }
    cls_targets[o] = c;
    box_targets[o] = t;
}

__global__ void iou_matrix_kernel(const float4 *__restrict__ b1, int n, const float4 *__restrict__ b2, int m,
                                  float *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= m || i >= n) return;
    const float4 p = __ldg(b1 + i), q = __ldg(b2 + j);
    out[(size_t)i * m + j] = iou_ref(p.x, p.y, p.z, p.w, area_ref(p.x, p.y, p.z, p.w), q.x, q.y, q.z, q.w,
                                     area_ref(q.x, q.y, q.z, q.w));
}

}  // namespace odk

extern "C" {

int odk_iou_matrix(const float *boxes1, int n, const float *boxes2, int m, float *out, void *stream) {
    using namespace odk;
    if (n < 0 || m < 0) return set_error(ODK_EINVAL, "odk_iou_matrix: negative size");
    if (n == 0 || m == 0) return ODK_OK;
    if (!boxes1 || !boxes2 || !out) return set_error(ODK_EINVAL, "odk_iou_matrix: null pointer");
    if (((uintptr_t)boxes1 | (uintptr_t)boxes2) & 15) return set_error(ODK_EINVAL, "odk_iou_matrix: boxes must be 16-byte aligned");
    if (n > 65535) return set_error(ODK_EUNSUPPORTED, "odk_iou_matrix: more than 65535 rows");
    dim3 grid((m + 255) / 256, n);
    iou_matrix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4 *)boxes1, n, (const float4 *)boxes2, m, out);
    return check_launch("odk_iou_matrix");
}

size_t odk_assign_workspace_bytes(int B, int Mmax) {
    if (B < 0 || Mmax < 0) return 0;
    size_t best = (size_t)B * (size_t)(Mmax > 0 ? Mmax : 1) * sizeof(unsigned long long);
    size_t pos = (((size_t)B * sizeof(int32_t)) + 15) & ~(size_t)15;
    return best + pos + 16;
}

int odk_assign(const float *anchors, const float *gt_boxes, const int32_t *gt_labels, const int32_t *gt_count, int B,
               int Mmax, const int32_t *level_hw, int num_levels, int na, float match_thr, int filter_valid,
               int32_t *match, float *num_pos, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    Geo g;
    int rc = make_geo(&g, level_hw, num_levels, na);
    if (rc) return rc;
    if (B < 0 || Mmax < 0) return set_error(ODK_EINVAL, "odk_assign: negative size");
    if (B == 0) return ODK_OK;
    if (!anchors || !match || !num_pos || (Mmax > 0 && (!gt_boxes || !gt_labels)))
        return set_error(ODK_EINVAL, "odk_assign: null pointer");
    if (B > 65535) return set_error(ODK_EUNSUPPORTED, "odk_assign: batch > 65535");
    if (Mmax > 8192) return set_error(ODK_EUNSUPPORTED, "odk_assign: more than 8192 gt rows per image");
    if (workspace_bytes < odk_assign_workspace_bytes(B, Mmax) || !workspace)
        return set_error(ODK_EWORKSPACE, "odk_assign: workspace too small (%zu < %zu)", workspace_bytes,
                         odk_assign_workspace_bytes(B, Mmax));
    if (((uintptr_t)anchors | (uintptr_t)gt_boxes | (uintptr_t)workspace) & 15)
        return set_error(ODK_EINVAL, "odk_assign: anchors / gt_boxes / workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int Mw = Mmax > 0 ? Mmax : 1;
    unsigned long long *best = (unsigned long long *)workspace;
    int32_t *pos = (int32_t *)((char *)workspace + (size_t)B * Mw * sizeof(unsigned long long));
    cudaError_t e = cudaMemsetAsync(workspace, 0, odk_assign_workspace_bytes(B, Mmax), st);
    if (e != cudaSuccess) return set_error((int)e, "odk_assign memset: %s", cudaGetErrorString(e));
    dim3 grid((g.A + kAssignThreads - 1) / kAssignThreads, B);
    assign_kernel<<<grid, kAssignThreads, 0, st>>>(g, (const float4 *)anchors, (const float4 *)gt_boxes, gt_labels,
                                                   gt_count, Mmax, match_thr, filter_valid, match_thr > 0.0f ? 1 : 0,
                                                   match, best, pos);
    rc = check_launch("odk_assign/assign_kernel");
    if (rc) return rc;
    assign_force_kernel<<<B, 128, (size_t)Mw * sizeof(int), st>>>(g, gt_labels, gt_count, Mmax, filter_valid, best,
                                                                   pos, match, num_pos);
    return check_launch("odk_assign/assign_force_kernel");
}

int odk_targets(const float *anchors, const float *gt_boxes, const int32_t *gt_labels, int B, int Mmax,
                const int32_t *level_hw, int num_levels, int na, const int32_t *match, int64_t *cls_targets,
                float *box_targets, void *stream) {
    using namespace odk;
    Geo g;
    int rc = make_geo(&g, level_hw, num_levels, na);
    if (rc) return rc;
    if (B < 0 || Mmax < 0) return set_error(ODK_EINVAL, "odk_targets: negative size");
    if (B == 0) return ODK_OK;
    if (!anchors || !match || !cls_targets || !box_targets || (Mmax > 0 && (!gt_boxes || !gt_labels)))
        return set_error(ODK_EINVAL, "odk_targets: null pointer");
    if (B > 65535) return set_error(ODK_EUNSUPPORTED, "odk_targets: batch > 65535");
    if (((uintptr_t)anchors | (uintptr_t)gt_boxes | (uintptr_t)box_targets) & 15)
        return set_error(ODK_EINVAL, "odk_targets: anchors / gt_boxes / box_targets must be 16-byte aligned");
    dim3 grid((g.A + 255) / 256, B);
    targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, B, (const float4 *)anchors, (const float4 *)gt_boxes,
                                                           gt_labels, Mmax, match, cls_targets, (float4 *)box_targets);
    return check_launch("odk_targets");
}

}  // extern "C"
