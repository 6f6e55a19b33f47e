// Global top-k over [B, A*C] class logits read in place from the NCHW pyramid levels (K3).
// Replaces _post_process (reference effdet/bench.py:12-56).  See include/odk.h (odk_topk).
//
// HBM-bound design: the logits are streamed ONCE.
//   P0  sample   : ~1/64 of each image (pseudo-random 512-byte units of every channel plane); a lane keeps
//                  the maximum of what it sampled in one of 8192 slots, and the image's last CTA turns
//                  the slot maxima into a threshold (12-bit key bin) that keeps between K and kCap
//                  elements with overwhelming probability;
//   P1  collect  : every CTA streams its share of the image with 128-bit loads and appends (key, flat
//                  index) of the few elements at or above the threshold (~1.7*K per image) to a
//                  candidate list;
//   P2  select   : one CTA per image cuts the candidates to just over K with a 1024-bin histogram, uses
//                  the same histogram as a counting sort (in-bin ranking by direct comparison; bitonic
//                  network for heavily tied inputs) on the 64-bit composite key (value desc, flat index
//                  asc) and emits the top K.  If the candidate count is not in [K, kCap] it raises a flag;
//   P3  cluster  : an 8-CTA thread-block cluster per image gathers the box regressions of the selected
//                  anchors; for flagged images (never for i.i.d. logits; e.g. constant inputs) it runs an
//                  exact MSD radix select on the composite key instead (histograms merged through
//                  distributed shared memory), collects, sorts and emits everything itself.
// Composite key: hi32 = order-preserving map of the fp32 logit, lo32 = ~flat_index, so keys are
// unique and "largest key first" is torch.topk's order with ties broken by ascending index.
#include "odk_stream.cuh"

namespace cg = cooperative_groups;

namespace odk {

// ---- P1: single streaming pass, keep elements at or above the threshold bin ---------------
constexpr int kStage = 1024;   // per-CTA staging slots for hits (~200 expected)


__global__ void __launch_bounds__(kTopkThreads) topk_collect_kernel(const __grid_constant__ TopkArgs A) {
    __shared__ unsigned long long s_stage[kStage];
    __shared__ unsigned s_nstage, s_base;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) s_nstage = 0u;
    __syncthreads();
    const unsigned thr = __ldcg(A.thr + b);
    const float thr_f = thr_float(thr);
    const int lane = threadIdx.x & 31;
    const int W = gridDim.x * (kTopkThreads / 32);
    const int ntasks = A.task_off[A.g.nlev];
    unsigned *cnt = A.cnt + b;
    unsigned long long *cand = A.cand + (size_t)b * kCap;
    for (int t = blockIdx.x * (kTopkThreads / 32) + (threadIdx.x >> 5); t < ntasks; t += W) {
        const Task k = decode_task(A, b, t);
        auto hit = [&](float x, int s) {
            if (x >= thr_f) {
                const unsigned flat = k.fbase + (unsigned)s * k.fstride;
                const unsigned long long key = ((unsigned long long)vkey_of(x) << 32) | (unsigned long long)(~flat);
                const unsigned slot = atomicAdd(&s_nstage, 1u);
                if (slot < (unsigned)kStage) {
                    s_stage[slot] = key;
                } else {   // staging full (threshold far too low): straight to the global list
                    const unsigned pos = atomicAdd(cnt, 1u);
                    if (pos < (unsigned)kCap) {
                        cand[pos] = key;
                        if (A.cand_box) A.cand_box[(size_t)b * kCap + pos] = gather_box(A, b, (int)(flat / (unsigned)A.C));
                    }
                }
            }
        };
        visit_task(k, lane,
                   [&](float4 v, int s) {
                       // one compare per float4 on the common path; hits are ~1 in 2000 elements
                       if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) >= thr_f) {
                           hit(v.x, s); hit(v.y, s + 1); hit(v.z, s + 2); hit(v.w, s + 3);
                       }
                   },
                   hit);
    }
    __syncthreads();
    const unsigned n = min(s_nstage, (unsigned)kStage);
    if (threadIdx.x == 0) s_base = n ? atomicAdd(cnt, n) : 0u;
    __syncthreads();
    const unsigned base = s_base;
    for (unsigned i = threadIdx.x; i < n; i += kTopkThreads) {
        if (base + i < (unsigned)kCap) {
            const unsigned long long key = s_stage[i];
            cand[base + i] = key;
            // the detection stage needs the candidate's box regression (4 scattered sectors): gather it here, spread
            // over all SMs and hidden behind the stream, and leave it next to the key (one sector per row later)
            if (A.cand_box)
                A.cand_box[(size_t)b * kCap + base + i] = gather_box(A, b, (int)(~(unsigned)(key & 0xFFFFFFFFull) / (unsigned)A.C));
        }
    }
}

__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const __grid_constant__ TopkArgs A) {
    extern __shared__ __align__(16) unsigned long long s_keys[];
    const int b = blockIdx.x;
    const unsigned n = A.cnt[b];
    if (n < (unsigned)A.K || n > (unsigned)kCap) {
        if (threadIdx.x == 0) A.flag[b] = 1u;
        return;
    }
    sort_and_emit<false>(A, b, (int)n, s_keys);   // out_box: by the cluster kernel launched next
}

// ---- P3: exact radix select for flagged images (8-CTA cluster per image) ---------------------
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSelThreads)
topk_exact_kernel(const __grid_constant__ TopkArgs A) {
    extern __shared__ __align__(16) unsigned long long s_keys[];   // rank 0: sort buffer
    __shared__ unsigned s_hist[1 << kRadixBits];
    __shared__ unsigned long long s_prefix;   // decided high bits, already shifted into place
    __shared__ unsigned s_need;               // rank still to find inside the prefix
    __shared__ int s_done;
    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.x / kClusterSize;
    const unsigned rank = cluster.block_rank();
    if (A.fused == 2) {
        // staged odk_postprocess: this kernel runs BESIDE the tail kernel (forked stream), so it decides by itself which
        // images the sampled path cannot take -- the same test on the final candidate count as run_tail -- and publishes
        // the flag for the flagged-image kernels that follow it
        const unsigned n = __ldcg(A.cnt + b);
        const bool flagged = n < (unsigned)A.K || n > (unsigned)kCap || A.force_exact;
        if (!flagged) return;
        if (rank == 0 && threadIdx.x == 0) A.flag[b] = 1u;
    } else if (A.flag[b] == 0u) {   // uniform across the cluster: the select kernel did this image, its box rows are left
        if (!A.fused) gather_selected_boxes(A, b, rank);   // (behind odk_postprocess the image is complete already)
        return;
    }

    const int lane = threadIdx.x & 31;
    const int W = kClusterSize * (kSelThreads / 32);
    const int wid = rank * (kSelThreads / 32) + (threadIdx.x >> 5);
    const int ntasks = A.task_off[A.g.nlev];
    if (threadIdx.x == 0) { s_prefix = 0ull; s_need = (unsigned)A.K; s_done = 0; }
    unsigned long long lower = 0ull;   // collect every key >= lower
    int shift = 64;
    for (int level = 0; level < 6; ++level) {
        const int bits = shift >= kRadixBits ? kRadixBits : shift;
        shift -= bits;
        for (int i = threadIdx.x; i < (1 << kRadixBits); i += kSelThreads) s_hist[i] = 0;
        cluster.sync();   // also orders the state written by rank 0 in the previous round
        const unsigned long long prefix = *cluster.map_shared_rank(&s_prefix, 0);
        const int pshift = shift + bits;
        for (int t = wid; t < ntasks; t += W) {
            const Task k = decode_task(A, b, t);
            auto one = [&](float x, int s) {
                const unsigned flat = k.fbase + (unsigned)s * k.fstride;
                const unsigned long long key = ((unsigned long long)vkey_of(x) << 32) | (unsigned long long)(~flat);
                const bool in = pshift >= 64 ? true : ((key >> pshift) == (prefix >> pshift));
                if (in) atomicAdd(&s_hist[(unsigned)(key >> shift) & ((1u << bits) - 1u)], 1u);
            };
            visit_task(k, lane, [&](float4 v, int s) { one(v.x, s); one(v.y, s + 1); one(v.z, s + 2); one(v.w, s + 3); }, one);
        }
        cluster.sync();
        if (rank == 0) {
            // merge the 8 histograms into mine through distributed shared memory
            for (int i = threadIdx.x; i < (1 << kRadixBits); i += kSelThreads) {
                unsigned v = s_hist[i];
                for (unsigned r = 1; r < kClusterSize; ++r) v += cluster.map_shared_rank(s_hist, r)[i];
                s_hist[i] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned need = s_need, run = 0;
                int bin = (1 << bits) - 1;
                for (; bin > 0; --bin) {
                    if (run + s_hist[bin] >= need) break;
                    run += s_hist[bin];
                }
                const unsigned in_bin = s_hist[bin];
                s_prefix = prefix | ((unsigned long long)bin << shift);
                s_need = need - run;
                // everything above the chosen bin plus the bin itself fits -> stop refining
                const unsigned taken = (unsigned)A.K - need + run;
                s_done = (taken + in_bin <= (unsigned)kCap) || shift == 0;
            }
        }
        cluster.sync();
        lower = *cluster.map_shared_rank(&s_prefix, 0);
        if (*cluster.map_shared_rank(&s_done, 0)) break;
    }
    // collect
    if (rank == 0 && threadIdx.x == 0) A.cnt[b] = 0u;
    cluster.sync();
    unsigned long long *cand = A.cand + (size_t)b * kCap;
    for (int t = wid; t < ntasks; t += W) {
        const Task k = decode_task(A, b, t);
        auto one = [&](float x, int s) {
            const unsigned flat = k.fbase + (unsigned)s * k.fstride;
            const unsigned long long key = ((unsigned long long)vkey_of(x) << 32) | (unsigned long long)(~flat);
            if (key >= lower) {
                const unsigned pos = atomicAdd(A.cnt + b, 1u);
                if (pos < (unsigned)kCap) cand[pos] = key;
            }
        };
        visit_task(k, lane, [&](float4 v, int s) { one(v.x, s); one(v.y, s + 1); one(v.z, s + 2); one(v.w, s + 3); }, one);
    }
    __threadfence();
    cluster.sync();
    if (rank == 0) {
        const unsigned n = *(volatile unsigned *)(A.cnt + b);
        if (threadIdx.x == 0) A.thr[b] = (unsigned)(lower >> 32);   // every collected key is >= this
        __syncthreads();
        sort_and_emit<true>(A, b, (int)min(n, (unsigned)kCap), s_keys);
    }
    cluster.sync();   // keep peers' shared memory alive until rank 0 is done with DSMEM
}

struct TopkWs { size_t slots, thr, thr_hi, cnt, flag, cand, total; int slot_stride, tps; };

static TopkWs topk_ws_layout(const StreamGeo &G, int B) {
    TopkWs w;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.slot_stride = (int)sample_slot_stride(G, &w.tps);
    size_t off = 0;
    w.slots = off; off = al(off + (size_t)B * w.slot_stride * 4);
    w.thr = off; off = al(off + (size_t)B * 4);
    w.thr_hi = off; off = al(off + (size_t)B * 4);
    w.cnt = off; off = al(off + (size_t)B * 4);
    w.flag = off; off = al(off + (size_t)B * 4);
    w.cand = off; off = al(off + (size_t)B * kCap * sizeof(unsigned long long));
    w.total = off;
    return w;
}

// per-level task model of the collect / exact kernels: NCHW levels are na * C planes of hw elements, channels_last
// levels ([B, H, W, channels], bit l of `layout`; bit 8 + l for the box levels) one run of hw * na * C elements
int fill_topk_levels(TopkArgs *a, const void *const *cls_levels, const void *const *box_levels, int na, int layout, int *ntasks,
                     const char *who) {
    a->planes = na * a->C;
    a->div_C = make_fastdiv((unsigned)a->C);
    int toff = 0;
    for (int l = 0; l < a->g.nlev; ++l) {
        a->cls[l] = (const float *)cls_levels[l];
        a->box[l] = (const float *)box_levels[l];
        if (!a->cls[l] || !a->box[l]) return set_error(ODK_EINVAL, "%s: null level pointer (level %d)", who, l);
        a->cls_nhwc[l] = (layout >> l) & 1;
        a->box_nhwc[l] = (layout >> (8 + l)) & 1;
        if (a->box_nhwc[l] && ((uintptr_t)a->box[l] & 15)) return set_error(ODK_EINVAL, "%s: channels_last box level %d must be 16-byte aligned", who, l);
        a->nplanes[l] = a->cls_nhwc[l] ? 1 : a->planes;
        const long long plen = a->cls_nhwc[l] ? (long long)a->g.hw[l] * a->planes : a->g.hw[l];
        if (plen > 0x7fffffffll) return set_error(ODK_EUNSUPPORTED, "%s: level %d too large", who, l);
        a->plane_len[l] = (int)plen;
        a->vec[l] = (plen % 4 == 0 && ((uintptr_t)a->cls[l] & 15) == 0) ? 4 : 1;
        a->nvec[l] = (int)(plen / a->vec[l]);
        a->nseg[l] = (a->nvec[l] + kSegVec - 1) / kSegVec;
        a->div_nseg[l] = make_fastdiv((unsigned)a->nseg[l]);
        a->task_off[l] = toff;
        toff += a->nplanes[l] * a->nseg[l];
    }
    for (int l = a->g.nlev; l <= ODK_MAX_LEVELS; ++l) a->task_off[l] = toff;
    *ntasks = toff;
    return ODK_OK;
}

// one co-resident wave of collect CTAs over the whole batch (per-plane task model)
int launch_topk_collect(const TopkArgs &a, int ntasks, cudaStream_t st) {
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topk_collect_kernel, kTopkThreads, 0);
    if (occ < 1) occ = 1;
    const int warps_per_cta = kTopkThreads / 32;
    int per_image = (sms * occ) / a.B;                           // all CTAs co-resident: exactly one wave
    const int max_useful = (ntasks + warps_per_cta - 1) / warps_per_cta;
    if (per_image > max_useful) per_image = max_useful;
    if (per_image < 1) per_image = 1;
    dim3 grid(per_image, a.B);
    topk_collect_kernel<<<grid, kTopkThreads, 0, st>>>(a);
    return check_launch("odk_topk/collect");
}

int launch_topk_exact_flagged(const TopkArgs &a, cudaStream_t st) {
    cudaFuncSetAttribute(topk_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmemBytes);
    topk_exact_kernel<<<a.B * kClusterSize, kSelThreads, kSelSmemBytes, st>>>(a);
    return check_launch("odk_topk/exact");
}

}  // namespace odk

extern "C" {

size_t odk_topk_workspace_bytes(int B, int C, const int32_t *level_hw, int num_levels, int na, int K) {
    using namespace odk;
    (void)K;
    Geo g;
    if (B < 1 || C < 1 || make_geo(&g, level_hw, num_levels, na)) return 0;
    StreamGeo G;
    memset(&G, 0, sizeof(G));
    for (int l = 0; l < g.nlev; ++l)
        G.ntask_img += (int)((((long long)g.na * C * g.hw[l] + 6) / 4 + kGroupsPerTask - 1) / kGroupsPerTask);
    return topk_ws_layout(G, B).total;
}

int odk_topk(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
             int num_levels, int na, int K, int layout, float *cls_topk, float *box_topk, int64_t *indices, int64_t *classes,
             void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    TopkArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_geo(&a.g, level_hw, num_levels, na);
    if (rc) return rc;
    if (!cls_levels || !box_levels || !cls_topk || !box_topk || !indices || !classes)
        return set_error(ODK_EINVAL, "odk_topk: null pointer");
    if (B < 1 || C < 1 || K < 1) return set_error(ODK_EINVAL, "odk_topk: B, C, K must be positive");
    if (B > 65535) return set_error(ODK_EUNSUPPORTED, "odk_topk: batch > 65535");
    a.N = (long long)a.g.A * C;
    if (a.N > 0xFFFFFFFFll) return set_error(ODK_EUNSUPPORTED, "odk_topk: A*C does not fit 32 bits");
    if ((long long)K > a.N) return set_error(ODK_EINVAL, "odk_topk: K=%d exceeds A*C=%lld (selected index k out of range)", K, a.N);
    if (K > kCap / 2) return set_error(ODK_EUNSUPPORTED, "odk_topk: K > %d not supported", kCap / 2);
    if ((size_t)B * na * (size_t)C > 0x7fffffffull) return set_error(ODK_EUNSUPPORTED, "odk_topk: B*na*C overflows int");
    if (((uintptr_t)box_topk | (uintptr_t)workspace) & 15) return set_error(ODK_EINVAL, "odk_topk: box_topk / workspace must be 16-byte aligned");
    if (!workspace) return set_error(ODK_EWORKSPACE, "odk_topk: null workspace");
    a.B = B; a.C = C; a.K = K;
    int toff = 0;
    rc = fill_topk_levels(&a, cls_levels, box_levels, na, layout, &toff, "odk_topk");
    if (rc) return rc;
    StreamGeo G;
    rc = make_stream_geo(&G, a.g, cls_levels, C, layout);
    if (rc) return rc;
    const TopkWs w = topk_ws_layout(G, B);
    if (workspace_bytes < w.total)
        return set_error(ODK_EWORKSPACE, "odk_topk: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    char *ws = (char *)workspace;
    a.slots = (unsigned *)(ws + w.slots); a.thr = (unsigned *)(ws + w.thr); a.cnt = (unsigned *)(ws + w.cnt);
    a.flag = (unsigned *)(ws + w.flag); a.cand = (unsigned long long *)(ws + w.cand);
    a.thr_hi = (unsigned *)(ws + w.thr_hi); a.cand_box = nullptr;
    a.out_val = cls_topk; a.out_box = box_topk; a.out_idx = (long long *)indices; a.out_cls = (long long *)classes;

    cudaStream_t st = (cudaStream_t)stream;
    // sample + threshold per image; the same launch zeroes the candidate counters and flags (no memset)
    SampleLaunch sl;
    memset(&sl, 0, sizeof(sl));
    sl.G = G; sl.B = B; sl.K = K; sl.N = a.N; sl.slots = a.slots; sl.slot_stride = w.slot_stride; sl.tps = w.tps; sl.thr = a.thr; sl.thr_hi = (unsigned *)(ws + w.thr_hi);
    sl.zero0 = a.cnt; sl.zero1 = a.flag;
    rc = launch_sample(sl, st);
    if (rc) return rc;

    rc = launch_topk_collect(a, toff, st);
    if (rc) return rc;
    cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmemBytes);
    topk_select_kernel<<<B, kSelThreads, kSelSmemBytes, st>>>(a);
    rc = check_launch("odk_topk/select");
    if (rc) return rc;
    return launch_topk_exact_flagged(a, st);
}

}  // extern "C"