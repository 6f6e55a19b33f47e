// Target assignment on the GPU (K1): gt x anchor IoU matching with force-match, bit-exact
// against the reference's torch path.  See include/odk.h (odk_assign, odk_targets).
//
// Layout / mapping
//   * four consecutive anchors per thread, in PLANAR order (level, anchor shape, y, x): a CTA's
//     1024 anchors are spatially adjacent boxes of one shape, so its bounding box is small and
//     most gt boxes can be rejected for the whole CTA at once (IoU is exactly 0 when the boxes
//     do not overlap, so the cull never changes a result while match_thr > 0);
//   * the surviving gt boxes of an image are staged in shared memory (ordered compaction keeps
//     ascending gt order, which is what "first index wins" ties need);
//   * per-gt arg-max over anchors (force_match_for_each_row) is a 64-bit atomicMax on
//     (iou_bits << 32 | ~anchor_index): highest IoU, then LOWEST reference anchor index;
//     a shared-memory copy per CTA filters almost all candidates before the global atomic.
//   * the last CTA of each image (threadfence + counter) applies the forced matches (lowest gt row
//     wins a contested anchor, argmax_matcher.py:141-143) and finishes num_positives: one launch.
#include "odk_common.cuh"

namespace odk {

constexpr int kAssignThreads = 256;
constexpr int kPerThread = 4;                    // consecutive planar anchors per thread (one int4 store)
constexpr int kGtTile = 256;
constexpr int kDirectMax = 48;
// gt rows per image up to which the direct path is used

// Forced matches for one image, run by the LAST CTA of that image: gt row i claims its arg-max
// anchor (anchor 0 if its IoU is 0 everywhere); the lowest gt row wins a contested anchor
// (argmax_matcher.py:139-144).  s_p: Mmax ints of dynamic shared memory.
__device__ void force_matches(const Geo &g, int b, const int32_t *__restrict__ gt_labels, int M, int Mmax, int filter_valid,
                              const unsigned long long *best, const int32_t *pos_count, int32_t *match,
                              float *__restrict__ num_pos, int *s_p) {
    __shared__ int s_extra;
    if (threadIdx.x == 0) s_extra = 0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        int p = -1;
        if (!filter_valid || __ldg(gt_labels + (size_t)b * Mmax + i) >= 0) {
            const unsigned long long k = __ldcg(best + (size_t)b * Mmax + i);
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
            if (__ldcg(mp) < 0) ++extra;
            *mp = i;
        }
    }
    if (extra) atomicAdd(&s_extra, extra);
    __syncthreads();
    if (threadIdx.x == 0) num_pos[b] = (float)(__ldcg(pos_count + (size_t)b * kCtrStride) + s_extra);
}

// STAGED = true : gt boxes that can touch the CTA are compacted into shared memory first (many gts);
// STAGED = false: few gts (<= kDirectMax): every warp walks the image's gt list straight from global
//                 memory (uniform, L1-resident loads) -- no staging phase and no block barriers in the
//                 matching loop, which is what dominates when there are only ~10 boxes per image.
template <bool STAGED>
__global__ void __launch_bounds__(kAssignThreads, 4)
assign_kernel(const Geo g, const float4 *__restrict__ anchors, const float4 *__restrict__ gt_boxes,
              const int32_t *__restrict__ gt_labels, const int32_t *__restrict__ gt_count, int Mmax, float thr,
              int filter_valid, int cull, int32_t *match, unsigned long long *best, int32_t *pos_count,
              unsigned *done, float *__restrict__ num_pos) {
    extern __shared__ int s_dyn[];   // Mmax ints for the forced-match step
    __shared__ float s_red[4][kAssignThreads / 32];
    __shared__ float s_tile[4];
    __shared__ float4 s_box[kGtTile];
    __shared__ float s_area[kGtTile];
    __shared__ int s_idx[kGtTile];
    __shared__ unsigned long long s_best[kGtTile];
    __shared__ int s_wcnt[kAssignThreads / 32];
    __shared__ int s_pos;
    __shared__ bool s_last;

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_pos = 0;
    int npos_thread = 0;
    // A CTA walks several 1024-anchor tiles of its image, so the fixed costs (gather latency, the
    // closing fence + counter) are paid once per CTA, not once per tile.
    const int ntiles = (g.Apad + kAssignThreads * kPerThread - 1) / (kAssignThreads * kPerThread);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = (tile * kAssignThreads + tid) * kPerThread;
    float4 a[kPerThread];
    float aarea[kPerThread];
    int r[kPerThread];
    bool live[kPerThread];
    {
        // decode the first anchor; its neighbours share (level, shape) unless the plane ends in between
        int l0 = 0, r0 = 0, left = 0;
        if (p0 < g.A) {
            r0 = planar_to_ref(g, p0, l0);
            const int loc = p0 - g.off[l0];
            left = g.hw[l0] - (loc % g.hw[l0]);   // positions left in this (level, shape) plane
        }
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            live[i] = p0 + i < g.A;
            r[i] = 0;
            a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            aarea[i] = 0.f;
            if (live[i]) {
                int l;
                r[i] = i < left ? r0 + i * g.na : planar_to_ref(g, p0 + i, l);
                a[i] = __ldg(anchors + r[i]);
                aarea[i] = area_ref(a[i].x, a[i].y, a[i].z, a[i].w);
            }
        }
    }
    // this warp's own bounding box (dead slots hold +/-inf and never widen it)
    float y0 = INFINITY, x0 = INFINITY, y1 = -INFINITY, x1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < kPerThread; ++i)
        if (live[i]) { y0 = fminf(y0, a[i].x); x0 = fminf(x0, a[i].y); y1 = fmaxf(y1, a[i].z); x1 = fmaxf(x1, a[i].w); }
    const float wy0 = warp_min(y0), wx0 = warp_min(x0), wy1 = warp_max(y1), wx1 = warp_max(x1);
    float best_iou[kPerThread];
    int best_g[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) { best_iou[i] = -1.0f; best_g[i] = -1; }
    if (STAGED) {
    // bounding box of this CTA's anchors
    if (lane == 0) { s_red[0][warp] = wy0; s_red[1][warp] = wx0; s_red[2][warp] = wy1; s_red[3][warp] = wx1; }
    __syncthreads();
    if (tid < 4) {
        float v = s_red[tid][0];
        for (int w = 1; w < kAssignThreads / 32; ++w) v = tid < 2 ? fminf(v, s_red[tid][w]) : fmaxf(v, s_red[tid][w]);
        s_tile[tid] = v;
    }
    __syncthreads();
    const float ty0 = s_tile[0], tx0 = s_tile[1], ty1 = s_tile[2], tx1 = s_tile[3];

    int M = Mmax;
    if (gt_count) M = min(max(__ldg(gt_count + b), 0), Mmax);
    const float4 *gtb = gt_boxes + (size_t)b * Mmax;
    const int32_t *gtl = gt_labels + (size_t)b * Mmax;


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
        // A warp's 128 anchors are neighbours in a few feature-map rows, so most staged gts miss the
        // whole warp: reject those with one warp-uniform test on the warp's bounding box.
        for (int j = 0; j < n; ++j) {
            const float4 q = s_box[j];
            if (cull && !((q.z > wy0) && (q.x < wy1) && (q.w > wx0) && (q.y < wx1))) continue;   // warp-uniform
            const float qa = s_area[j];
            const int qi = s_idx[j];
            unsigned long long key = 0ull;
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                float v = 0.0f;
                if (live[i]) {
                    // same arithmetic as iou_ref, with the (exact) early-outs for an empty intersection
                    const float h = __fsub_rn(fminf(q.z, a[i].z), fmaxf(q.x, a[i].x));
                    const float w = __fsub_rn(fminf(q.w, a[i].w), fmaxf(q.y, a[i].y));
                    if (h > 0.0f && w > 0.0f) {
                        const float inter = __fmul_rn(h, w);
                        if (inter != 0.0f) v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(qa, aarea[i]), inter));
                    }
                    if (v > best_iou[i]) { best_iou[i] = v; best_g[i] = qi; }
                    if (v > 0.0f) {
                        const unsigned long long k =
                            ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r[i]);
                        key = k > key ? k : key;
                    }
                }
            }
            // per-gt arg-max over anchors: reduce inside the warp, one shared atomic per warp at most
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

    }
    if (!STAGED) {
        int M = Mmax;
        if (gt_count) M = min(max(__ldg(gt_count + b), 0), Mmax);
        const float4 *gtb = gt_boxes + (size_t)b * Mmax;
        const int32_t *gtl = gt_labels + (size_t)b * Mmax;
        for (int j = 0; j < M; ++j) {
            const float4 q = __ldg(gtb + j);                     // warp-uniform address
            if (filter_valid && __ldg(gtl + j) < 0) continue;    // anchors.py:405-408
            if (cull && !((q.z > wy0) && (q.x < wy1) && (q.w > wx0) && (q.y < wx1))) continue;   // warp-uniform
            const float qa = area_ref(q.x, q.y, q.z, q.w);
            unsigned long long key = 0ull;
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                float v = 0.0f;
                if (live[i]) {
                    const float h = __fsub_rn(fminf(q.z, a[i].z), fmaxf(q.x, a[i].x));
                    const float w = __fsub_rn(fminf(q.w, a[i].w), fmaxf(q.y, a[i].y));
                    if (h > 0.0f && w > 0.0f) {
                        const float inter = __fmul_rn(h, w);
                        if (inter != 0.0f) v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(qa, aarea[i]), inter));
                    }
                    if (v > best_iou[i]) { best_iou[i] = v; best_g[i] = j; }
                    if (v > 0.0f) {
                        const unsigned long long k =
                            ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r[i]);
                        key = k > key ? k : key;
                    }
                }
            }
            if (__any_sync(0xffffffffu, key != 0ull)) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                if (lane == 0) atomicMax(best + (size_t)b * Mmax + j, key);   // result unused: a RED, no stall
            }
        }
    }
    // thresholds (argmax_matcher.py:126-137 with matched == unmatched threshold)
    int m[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        m[i] = (live[i] && best_g[i] >= 0 && !(thr > best_iou[i])) ? best_g[i] : -1;
        npos_thread += m[i] >= 0;
    }
    if (p0 < g.Apad)   // Apad is a multiple of 4: whole int4s, the padding slots get -1
        *reinterpret_cast<int4 *>(match + (size_t)b * g.Apad + p0) = make_int4(m[0], m[1], m[2], m[3]);
    }   // tiles
    int npos = (int)warp_sum((float)npos_thread);
    if (lane == 0 && npos) atomicAdd(&s_pos, npos);
    __syncthreads();
    if (tid == 0) {
        if (s_pos) atomicAdd(pos_count + (size_t)b * kCtrStride, s_pos);
        __threadfence();
        s_last = atomicAdd(done + (size_t)b * kCtrStride, 1u) == gridDim.x - 1;   // all CTAs of this image are through
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        int Mf = Mmax;
        if (gt_count) Mf = min(max(__ldg(gt_count + b), 0), Mmax);
        force_matches(g, b, gt_labels, Mf, Mmax, filter_valid, best, pos_count, match, num_pos, s_dyn);
    }
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

// ---- gt-centric assignment over regular anchor grids ----------------------------------------------
// The pyramid's anchors are regular grids: plane k = (level, shape) has centres (cy0 + y*sy, cx0 + x*sx)
// and one half-size.  A gt box can only have a positive IoU with the cells whose anchor overlaps it, and
// IoU <= min(area)/max(area), so instead of testing every (gt, anchor) pair a CTA per gt enumerates
//   pass 0: the planes whose area bound can reach the match threshold (these hold every match), and
//   pass 1: only if its best IoU is still below the threshold, the planes whose bound exceeds that best
//           (the forced match needs the true arg-max).
// The enumeration is a superset of the pairs that matter and every IoU is still computed from the fp32
// anchor table in the reference's operation order, so results are bit-identical to the dense kernel.
// One 64-bit atomicMax array keys[B, Apad] resolves everything: an IoU>=thr pair writes
// (iou_bits<<32 | ~gt) (highest IoU, then lowest gt row: torch.max's first index), a forced match writes
// (0xFFFFFFFF<<32 | ~gt) (beats any IoU; lowest gt row wins, argmax_matcher.py:141-143).  The first
// writer of an anchor (old key 0) counts it into num_positives.
constexpr int kGtcThreads = 256;
constexpr int kMaxPlanes = ODK_MAX_LEVELS * 16;
constexpr int kCellsInFlight = 4;
constexpr int kDescFloats = 12;   // cy0, cx0, sy, sx, hy, hx, area, W, H, off, a, level

// `gen` (nullable): the grids' generator in float64 -- cy0, cx0, sy, sx, half_y, half_x per plane, the very numbers
// the anchor table was built from (anchors.py:264-299).  With it a cell's anchor is RECOMPUTED, bit for bit what the
// table holds: the centre cy0 + y*sy is exact in double (small integers and halves), centre -/+ half size is one
// IEEE double operation each, like numpy's, and the cast rounds to nearest like `.float()`
// (odk_anchor_table + test_anchor_table_from_generator compare every anchor of every model shape).  That takes the
// dependent anchor gather -- an L2 round trip per group of cells -- out of the enumeration loop.
constexpr int kGenDoubles = 6;
__device__ __forceinline__ float4 anchor_from_gen(const double *gk, int yy, int xx) {
    const double cy = fma((double)yy, gk[2], gk[0]), cx = fma((double)xx, gk[3], gk[1]);   // exact
    return make_float4(__double2float_rn(__dsub_rn(cy, gk[4])), __double2float_rn(__dsub_rn(cx, gk[5])),
                       __double2float_rn(__dadd_rn(cy, gk[4])), __double2float_rn(__dadd_rn(cx, gk[5])));
}

__device__ __forceinline__ void assign_one_gt(const Geo &g, const float4 *__restrict__ anchors, const float *__restrict__ desc,
                                              const double *__restrict__ gen, int nplanes, const float4 *__restrict__ gt_boxes,
                                              const int32_t *__restrict__ gt_labels, const int32_t *__restrict__ gt_count,
                                              int Mmax, float thr, int filter_valid, unsigned long long *keys,
                                              int32_t *pos_count, unsigned *touched, int touched_cap) {
    __shared__ int s_off[kMaxPlanes + 1], s_x0[kMaxPlanes], s_nx[kMaxPlanes], s_y0[kMaxPlanes];
    __shared__ int s_w[kMaxPlanes], s_base[kMaxPlanes], s_sh[kMaxPlanes], s_hw[kMaxPlanes];
    __shared__ double s_gen[kMaxPlanes][kGenDoubles];
    __shared__ unsigned long long s_red[kGtcThreads / 32];
    __shared__ unsigned long long s_best;
    __shared__ unsigned s_bmax;
    __shared__ unsigned char s_done[kMaxPlanes];
    __shared__ float s_desc[kMaxPlanes][kDescFloats];
    const int b = blockIdx.y, i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // everything the CTA needs from global memory is requested at once (one exposed latency, not four dependent ones)
    const int cnt_raw = gt_count ? __ldg(gt_count + b) : Mmax;
    const int label = __ldg(gt_labels + (size_t)b * Mmax + i);
    const float4 q = __ldg(gt_boxes + (size_t)b * Mmax + i);
    for (int e = tid; e < nplanes * kDescFloats; e += kGtcThreads) (&s_desc[0][0])[e] = __ldg(desc + e);
    const int M = min(max(cnt_raw, 0), Mmax);
    if (i >= M) return;
    if (filter_valid && label < 0) return;
    const float qa = area_ref(q.x, q.y, q.z, q.w);
    if (tid == 0) s_best = 0ull;
    __syncthreads();
    for (int k = tid; k < nplanes; k += kGtcThreads) {
        const float *d = s_desc[k];
        s_w[k] = (int)d[7]; s_base[k] = (int)d[9]; s_sh[k] = (int)d[10]; s_hw[k] = g.hw[(int)d[11]];
    }
    if (gen)
        for (int e = tid; e < nplanes * kGenDoubles; e += kGtcThreads) (&s_gen[0][0])[e] = __ldg(gen + e);
    __syncthreads();
    unsigned long long *krow = keys + (size_t)b * g.Apad;

    // Passes: 0 = the planes whose area bound reaches the threshold, cells that can reach it (these hold every match).
    // If the gt matched nothing the forced match needs the TRUE arg-max over all anchors: pass 1 looks at the planes
    // whose bound is at least half the best bound of any plane (the anchors most like the gt), cells that can beat the
    // best IoU so far; pass 2 at whatever plane can still beat the best after that -- for a small gt, whose only
    // similar anchors are the smallest ones, usually none (one pass over every overlapping cell of every plane cost
    // such a gt 3-4 k cells and made its CTA the kernel's critical path: 16 us against a median of 7.6).
    // A plane enumerated in pass 1 is complete for pass 2 (the best only grows); pass-0 planes are not (need = thr).
    if (tid == 0) s_bmax = 0u;
    for (int k = tid; k < nplanes; k += kGtcThreads) s_done[k] = 0;
    __syncthreads();
    for (int pass = 0; pass < 3; ++pass) {
        // ---- which planes, which cells ----
        const float best_so_far = __uint_as_float((unsigned)(s_best >> 32));
        if (pass > 0 && !(best_so_far < thr)) break;   // uniform: the gt has its match(es), the arg-max is among them
        const float bmax = __uint_as_float(s_bmax);     // (pass 0 writes it below; read after that pass's barriers)
        for (int k = tid; k < nplanes; k += kGtcThreads) {
            const float *d = s_desc[k];
            int nx = 0, ny = 0, x0 = 0, y0 = 0;
            if (qa > 0.0f) {
                const float pa = d[6];
                // Two upper bounds on any IoU in this plane: min/max of the areas, and the tighter one from the dimensions (the
                // intersection is at most min(heights) * min(widths); for a long thin gt it rules out the square anchors of the
                // same area).  Pass 0 selects by the AREA bound on purpose: for a gt that ends up unmatched its planes are the
                // cheap first look (cells that could reach thr only) that gives passes 1-2 a best IoU to prune and shrink with
                // (measured: selecting pass 0 by the tight bound costs 1 us at D0 M=10 and 6 % at D7 M=100).
                const float bound = fminf(pa, qa) / fmaxf(pa, qa) * 1.00001f;
                const float imax = fminf(2.0f * d[4], q.z - q.x) * fminf(2.0f * d[5], q.w - q.y);
                const float tight = imax / fmaxf(pa + qa - imax, 1e-30f) * 1.0001f;
                if (pass == 0) atomicMax(&s_bmax, __float_as_uint(bound));      // positive floats order like their bits
                const bool want = pass == 0 ? (bound >= thr)
                                            : (!s_done[k] && tight > best_so_far && (pass == 2 || bound >= 0.5f * bmax));
                if (want) {
                    if (pass > 0) s_done[k] = 1;
                    const int W = (int)d[7], H = (int)d[8];
                    // Centres strictly inside (g0 - h, g1 + h) can overlap.  Pass 0 only needs the cells that can reach
                    // IoU >= thr: inter >= thr * union >= thr * max(areas), and inter = ih * iw with iw <= min(widths),
                    // so ih >= thr * max(areas) / min(widths) =: ih_min, and ih <= (ha + hg) / 2 - |dy|: the centre range
                    // shrinks by ih_min on both sides (likewise in x) -- about 4x fewer cells at thr = 0.5.  A 0.1 %
                    // slack on the bound and the one-cell widening below absorb the rounding.
                    // (pass 1 the same with the best IoU so far in place of thr: only a cell that can beat it matters)
                    const float need = pass == 0 ? thr : best_so_far;
                    const float amax = fmaxf(pa, qa) * need * 0.999f;
                    const float shy = amax / fminf(2.0f * d[5], q.w - q.y);
                    const float shx = amax / fminf(2.0f * d[4], q.z - q.x);
                    const float ylo = (q.x - d[4] + shy - d[0]) / d[2], yhi = (q.z + d[4] - shy - d[0]) / d[2];
                    const float xlo = (q.y - d[5] + shx - d[1]) / d[3], xhi = (q.w + d[5] - shx - d[1]) / d[3];
                    const int ya = max((int)floorf(ylo) - 1, 0), yb = min((int)ceilf(yhi) + 1, H - 1);
                    const int xa = max((int)floorf(xlo) - 1, 0), xb = min((int)ceilf(xhi) + 1, W - 1);
                    if (yb >= ya && xb >= xa && yhi >= ylo && xhi >= xlo) { ny = yb - ya + 1; nx = xb - xa + 1; y0 = ya; x0 = xa; }
                }
            }
            s_x0[k] = x0; s_y0[k] = y0; s_nx[k] = nx;
            s_off[k + 1] = nx * ny;   // turned into a prefix sum below
        }
        __syncthreads();
        if (tid < 32) {   // inclusive scan of up to 128 plane counts by one warp (4 per lane)
            int v[kMaxPlanes / 32], run = 0;
#pragma unroll
            for (int j = 0; j < kMaxPlanes / 32; ++j) {
                const int k = tid * (kMaxPlanes / 32) + j;
                run += k < nplanes ? s_off[k + 1] : 0;
                v[j] = run;
            }
            int inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += t;
            }
            const int excl = inc - run;
#pragma unroll
            for (int j = 0; j < kMaxPlanes / 32; ++j) {
                const int k = tid * (kMaxPlanes / 32) + j;
                if (k < nplanes) s_off[k + 1] = excl + v[j];
            }
            if (tid == 0) s_off[0] = 0;
        }
        __syncthreads();
        const int total = s_off[nplanes];
        if (total == 0) continue;   // uniform (pass 0 may select nothing: pass 1 still has to look)

        // ---- enumerate the cells ----
        // a thread's cells c = tid, tid+256, ... ascend, so its plane index only moves forward;
        // kCellsInFlight cells per iteration (independent anchor gathers)
        unsigned long long best = 0ull;
        int kcur = 0;
        auto locate = [&](int c, int &k, int &r, int &p, bool &ok, float4 &a) {
            ok = c < total;
            r = 0; p = 0;
            if (!ok) return;
            while (s_off[k + 1] <= c) ++k;
            const int loc = c - s_off[k];
            const int nx = s_nx[k];
            const int yy = s_y0[k] + loc / nx, xx = s_x0[k] + loc % nx;
            const int cell = yy * s_w[k] + xx;
            r = s_base[k] + cell * g.na + s_sh[k];
            p = s_base[k] + s_sh[k] * s_hw[k] + cell;
            if (gen) a = anchor_from_gen(s_gen[k], yy, xx);
        };
        auto visit = [&](float4 a, int r, int p) {
            const float h = __fsub_rn(fminf(q.z, a.z), fmaxf(q.x, a.x));
            const float w = __fsub_rn(fminf(q.w, a.w), fmaxf(q.y, a.y));
            if (h > 0.0f && w > 0.0f) {
                const float inter = __fmul_rn(h, w);
                if (inter != 0.0f) {
                    const float v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(qa, area_ref(a.x, a.y, a.z, a.w)), inter));
                    if (v > 0.0f) {
                        const unsigned long long kr =
                            ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r);
                        best = kr > best ? kr : best;
                        if (!(thr > v)) {   // a candidate match for this anchor (argmax_matcher.py:126-137)
                            const unsigned long long kg =
                                ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                            if (atomicMax(krow + p, kg) == 0ull) {   // first key of this anchor: count it, list it
                                const int slot = atomicAdd(pos_count + (size_t)b * kCtrStride, 1);
                                if (slot < touched_cap) touched[(size_t)b * touched_cap + slot] = (unsigned)p;
                            }
                        }
                    }
                }
            }
        };
        for (int c = tid; c < total; c += kCellsInFlight * kGtcThreads) {
            int r[kCellsInFlight], p[kCellsInFlight];
            bool ok[kCellsInFlight];
            float4 a[kCellsInFlight];
#pragma unroll
            for (int u = 0; u < kCellsInFlight; ++u) locate(c + u * kGtcThreads, kcur, r[u], p[u], ok[u], a[u]);
            if (!gen) {
#pragma unroll
                for (int u = 0; u < kCellsInFlight; ++u) a[u] = __ldg(anchors + r[u]);   // r = 0 when out of range: a harmless, cached read
            }
#pragma unroll
            for (int u = 0; u < kCellsInFlight; ++u)
                if (ok[u]) visit(a[u], r[u], p[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) s_red[warp] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long v = s_best;
            for (int w = 0; w < kGtcThreads / 32; ++w) v = s_red[w] > v ? s_red[w] : v;
            s_best = v;
        }
        __syncthreads();
    }
    // ---- forced match: the arg-max anchor (anchor 0 when the gt overlaps nothing) ----
    if (tid == 0) {
        const unsigned long long kb = s_best;
        const int r = kb ? (int)(0xFFFFFFFFu - (unsigned)(kb & 0xFFFFFFFFull)) : 0;
        int l;
        const int p = ref_to_planar(g, r, l);
        const unsigned long long kf = (0xFFFFFFFFull << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        if (atomicMax(krow + p, kf) == 0ull) {
            const int slot = atomicAdd(pos_count + (size_t)b * kCtrStride, 1);
            if (slot < touched_cap) touched[(size_t)b * touched_cap + slot] = (unsigned)p;
        }
    }
}

// num_positives per image and the loss normaliser sum + 1 (loss.py:261), by one CTA
template <bool COHERENT>
__device__ __forceinline__ void finish_counts(int B, const int32_t *pos_count, float *num_pos, float *normalizer) {
    __shared__ float s_w[kGtcThreads / 32];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += kGtcThreads) {
        const int32_t *pc = pos_count + (size_t)b * kCtrStride;
        const float v = (float)(COHERENT ? __ldcg(pc) : __ldg(pc));
        num_pos[b] = v;
        acc += v;   // integer-valued, exact in fp32 below 2^24 like the reference's fp32 sum
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && normalizer) {
        float t = 0.f;
        for (int w = 0; w < kGtcThreads / 32; ++w) t += s_w[w];
        normalizer[0] = t + 1.0f;
    }
}

// One CTA per (gt row, image); the CTA that finishes last turns the counters into num_positives and
// the normaliser, so the loss kernel can follow without another launch.
__global__ void __launch_bounds__(kGtcThreads, 5)   // 5 CTAs/SM: B*M = 640 CTAs at D0 B=64 M=10 are one resident wave
assign_gt_kernel(const Geo g, const float4 *__restrict__ anchors, const float *__restrict__ desc, const double *__restrict__ gen, int nplanes,
                 const float4 *__restrict__ gt_boxes, const int32_t *__restrict__ gt_labels,
                 const int32_t *__restrict__ gt_count, int Mmax, float thr, int filter_valid,
                 unsigned long long *keys, int32_t *pos_count, unsigned *touched, int touched_cap, unsigned *done, int B,
                 float *num_pos, float *normalizer) {
    __shared__ bool s_last;
    assign_one_gt(g, anchors, desc, gen, nplanes, gt_boxes, gt_labels, gt_count, Mmax, thr, filter_valid, keys, pos_count, touched,
                  touched_cap);
    __threadfence();   // this thread's counter updates are visible before the CTA reports in
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        finish_counts<true>(B, pos_count, num_pos, normalizer);
    }
}

__global__ void __launch_bounds__(kGtcThreads)
finish_counts_kernel(int B, const int32_t *__restrict__ pos_count, float *__restrict__ num_pos, float *__restrict__ normalizer) {
    finish_counts<false>(B, pos_count, num_pos, normalizer);
}

// the anchor table from its generator (reference order r = off_l + (y*W + x)*na + a): one thread per anchor
__global__ void __launch_bounds__(256)
anchor_table_kernel(const Geo g, const float *__restrict__ desc, const double *__restrict__ gen, float4 *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.A) return;
    const int l = geo_level(g, r);
    const int loc = r - g.off[l];
    const int cell = loc / g.na, a = loc - cell * g.na;
    const int k = l * g.na + a;
    const int W = (int)__ldg(desc + (size_t)k * kDescFloats + 7);
    const int yy = cell / W, xx = cell - yy * W;
    out[r] = anchor_from_gen(gen + (size_t)k * kGenDoubles, yy, xx);
}

__global__ void __launch_bounds__(256)
keys_to_match_kernel(int B, int Apad, const unsigned long long *__restrict__ keys, int32_t *__restrict__ match) {
    const size_t n4 = (size_t)B * Apad / 4;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    const ulonglong2 k01 = __ldcg(reinterpret_cast<const ulonglong2 *>(keys) + 2 * t);
    const ulonglong2 k23 = __ldcg(reinterpret_cast<const ulonglong2 *>(keys) + 2 * t + 1);
    int4 m;
    m.x = k01.x ? (int)(0xFFFFFFFFu - (unsigned)(k01.x & 0xFFFFFFFFull)) : -1;
    m.y = k01.y ? (int)(0xFFFFFFFFu - (unsigned)(k01.y & 0xFFFFFFFFull)) : -1;
    m.z = k23.x ? (int)(0xFFFFFFFFu - (unsigned)(k23.x & 0xFFFFFFFFull)) : -1;
    m.w = k23.y ? (int)(0xFFFFFFFFu - (unsigned)(k23.y & 0xFFFFFFFFull)) : -1;
    reinterpret_cast<int4 *>(match)[t] = m;
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
    size_t pos = (size_t)B * odk::kCtrStride * sizeof(int32_t);
    return best + 2 * pos + 16;   // best keys, positive counts, per-image CTA counters (one 128 B line each)
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
    unsigned *done = (unsigned *)((char *)pos + (size_t)B * kCtrStride * sizeof(int32_t));
    cudaError_t e = cudaMemsetAsync(workspace, 0, odk_assign_workspace_bytes(B, Mmax), st);
    if (e != cudaSuccess) return set_error((int)e, "odk_assign memset: %s", cudaGetErrorString(e));
    const int per_cta = kAssignThreads * kPerThread;
    const int ntiles = (g.Apad + per_cta - 1) / per_cta;
    const int sms = device_sm_count();
    // One tile per CTA measured faster than a persistent walk (70 vs 88 us at D0, B=64): the kernel is
    // bound by dependent latency per warp, so more independent CTAs in flight win.  The tile loop in
    // the kernel stays for very large grids.
    int per_image = ntiles;
    if ((long long)per_image * B > 65535ll * 16) per_image = (sms * 4 + B - 1) / B;
    if (per_image > ntiles) per_image = ntiles;
    if (per_image < 1) per_image = 1;
    dim3 grid(per_image, B);
    if (Mmax <= kDirectMax)
        assign_kernel<false><<<grid, kAssignThreads, (size_t)Mw * sizeof(int), st>>>(
            g, (const float4 *)anchors, (const float4 *)gt_boxes, gt_labels, gt_count, Mmax, match_thr, filter_valid,
            match_thr > 0.0f ? 1 : 0, match, best, pos, done, num_pos);
    else
        assign_kernel<true><<<grid, kAssignThreads, (size_t)Mw * sizeof(int), st>>>(
            g, (const float4 *)anchors, (const float4 *)gt_boxes, gt_labels, gt_count, Mmax, match_thr, filter_valid,
            match_thr > 0.0f ? 1 : 0, match, best, pos, done, num_pos);
    return check_launch("odk_assign/assign_kernel");
}


size_t odk_assign_grid_workspace_bytes(int B, int64_t A) {
    if (B < 0 || A < 0) return 0;
    return odk::assign_grid_layout(B, odk_planar_stride(A)).total;
}

int odk_keys_to_match(const void *keys, int B, int64_t A, int32_t *match, void *stream) {
    using namespace odk;
    if (B < 0 || A < 0) return set_error(ODK_EINVAL, "odk_keys_to_match: negative size");
    if (B == 0 || A == 0) return ODK_OK;
    if (!keys || !match || (((uintptr_t)keys | (uintptr_t)match) & 15))
        return set_error(ODK_EINVAL, "odk_keys_to_match: keys / match must be non-null and 16-byte aligned");
    const int Apad = (int)odk_planar_stride(A);
    const size_t n4 = (size_t)B * Apad / 4;
    keys_to_match_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        B, Apad, (const unsigned long long *)keys, match);
    return check_launch("odk_keys_to_match");
}

int odk_anchor_table(const float *plane_desc, const double *plane_gen, int num_planes, const int32_t *level_hw, int num_levels,
                     int na, float *anchors_out, void *stream) {
    using namespace odk;
    Geo g;
    int rc = make_geo(&g, level_hw, num_levels, na);
    if (rc) return rc;
    if (!plane_desc || !plane_gen || !anchors_out || ((uintptr_t)anchors_out & 15))
        return set_error(ODK_EINVAL, "odk_anchor_table: null pointer / anchors_out not 16-byte aligned");
    if (num_planes != num_levels * na) return set_error(ODK_EINVAL, "odk_anchor_table: need one descriptor per (level, shape)");
    if (g.A == 0) return ODK_OK;
    anchor_table_kernel<<<(unsigned)((g.A + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, plane_desc, plane_gen, (float4 *)anchors_out);
    return check_launch("odk_anchor_table");
}

int odk_assign_grid(const float *anchors, const float *plane_desc, const double *plane_gen, int num_planes, const float *gt_boxes,
                    const int32_t *gt_labels, const int32_t *gt_count, int B, int Mmax, const int32_t *level_hw,
                    int num_levels, int na, float match_thr, int filter_valid, int32_t *match, float *num_pos,
                    float *normalizer, int flags, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    Geo g;
    int rc = make_geo(&g, level_hw, num_levels, na);
    if (rc) return rc;
    if (B < 0 || Mmax < 0) return set_error(ODK_EINVAL, "odk_assign_grid: negative size");
    if (B == 0) return ODK_OK;
    if (!anchors || !plane_desc || !num_pos || (Mmax > 0 && (!gt_boxes || !gt_labels)))
        return set_error(ODK_EINVAL, "odk_assign_grid: null pointer");
    if (num_planes != num_levels * na || num_planes > kMaxPlanes)
        return set_error(ODK_EINVAL, "odk_assign_grid: need one descriptor per (level, shape), at most %d", kMaxPlanes);
    if (!(match_thr > 0.0f)) return set_error(ODK_EUNSUPPORTED, "odk_assign_grid: match_thr must be > 0 (use odk_assign)");
    if (B > 65535) return set_error(ODK_EUNSUPPORTED, "odk_assign_grid: batch > 65535");
    if (workspace_bytes < odk_assign_grid_workspace_bytes(B, g.A) || !workspace)
        return set_error(ODK_EWORKSPACE, "odk_assign_grid: workspace too small (%zu < %zu)", workspace_bytes,
                         odk_assign_grid_workspace_bytes(B, g.A));
    if (((uintptr_t)anchors | (uintptr_t)gt_boxes | (uintptr_t)workspace | (uintptr_t)match) & 15)
        return set_error(ODK_EINVAL, "odk_assign_grid: anchors / gt_boxes / match / workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const AssignGridWs w = assign_grid_layout(B, g.Apad);
    unsigned long long *keys = (unsigned long long *)workspace;
    int32_t *pos = (int32_t *)((char *)workspace + w.pos);
    unsigned *touched = (unsigned *)((char *)workspace + w.touched);
    unsigned *done = (unsigned *)((char *)workspace + w.done);
    if (!(flags & ODK_ASSIGN_WS_CLEAN)) {   // else: keys, counters and `done` are zero already (odk_loss cleared them)
        cudaError_t e = cudaMemsetAsync(workspace, 0, w.total, st);
        if (e != cudaSuccess) return set_error((int)e, "odk_assign_grid memset: %s", cudaGetErrorString(e));
    }
    if (Mmax > 0) {
        dim3 grid(Mmax, B);
        assign_gt_kernel<<<grid, kGtcThreads, 0, st>>>(g, (const float4 *)anchors, plane_desc, plane_gen, num_planes,
                                                       (const float4 *)gt_boxes, gt_labels, gt_count, Mmax, match_thr,
                                                       filter_valid, keys, pos, touched, w.touched_cap, done, B, num_pos,
                                                       normalizer);
        rc = check_launch("odk_assign_grid/assign_gt_kernel");
    } else {
        finish_counts_kernel<<<1, kGtcThreads, 0, st>>>(B, pos, num_pos, normalizer);
        rc = check_launch("odk_assign_grid/finish_counts_kernel");
    }
    if (rc || !match) return rc;
    return odk_keys_to_match(keys, B, g.A, match, stream);
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
