// Fused post-process (K3 + K4/K5 + K6 in one pipeline): top-k over the class logits, box decode, score
// filter, class-aware NMS or Soft-NMS and the per-detection OOD scores.  Replaces the whole
// DetBenchPredict chain _post_process -> _batch_detection (reference effdet/bench.py:12-76,
// effdet/anchors.py:95-172).  See include/odk.h (odk_postprocess).
//
// The logits are streamed ONCE, image-major, by a persistent grid (one 1024-thread CTA per SM); everything
// that is not streaming -- selecting the top K of an image's candidates, gathering and decoding its boxes,
// suppression, OOD scores: per-image chains of dependent latencies -- runs WHILE the later images are still
// streaming, so only the last image's chain is exposed after the stream ends:
//   sample   : (odk::launch_sample) an 8-CTA cluster per image reads ~1/64 of it, keeps per-lane maxima in
//              slots, and its leader turns them into the collect threshold (no global counters, no memset:
//              it also zeroes the counters of the kernels that follow);
//   stream   : every WARP of the persistent grid takes 16 KB tasks from one global queue (image-major order),
//              compares 8 x 128-bit loads in flight per lane against the image's threshold and appends the
//              few hits (~1 in 1600) to the image's candidate list through a warp-private staging buffer;
//   tail     : the warp that completes an image's last task hands the image to its CTA: at the next task
//              boundary the CTA's 32 warps meet and run select (counting sort of the candidates, odk_topk.cuh)
//              -> decode -> suppression (odk_detect.cuh) -> rows -> OOD scores for it out of shared memory,
//              then return to the queue.  Nothing ever waits on another CTA.
// Images whose candidate count leaves [K, kCap] (constant / heavily tied logits) are only flagged here; the
// exact cluster radix select of odk_topk.cu and the stand-alone detect kernel pick them up afterwards.
#include "odk_stream.cuh"
#include "odk_detect.cuh"

namespace cg = cooperative_groups;

namespace odk {

constexpr int kPostThreads = 1024;
constexpr int kPostWarps = kPostThreads / 32;
constexpr int kWarpStage = 48;        // staged hits per warp and task (~2.6 expected)
constexpr int kPostKeys = 6;          // sorted keys per thread: K <= 6144
constexpr int kPostMaxK = kPostKeys * kPostThreads;
constexpr int kSampleThreads = 256;
constexpr int kSampleWarps = kClusterSize * (kSampleThreads / 32);

// ---- sample ---------------------------------------------------------------------------------------------
// Task t of image b contributes its 512-byte unit j = hash(b, t) >> 26 when j < 32 (and the unit exists), so
// every unit is taken with probability 1/64.  Lane maxima over `tps` consecutive tasks form one slot group.
// The slots live in the shared memory of the cluster's leader CTA: the other CTAs merge their maxima into
// them with distributed-shared-memory atomics, so the threshold statistics never touch global memory.
constexpr int kSampleBatch = 8;   // 128-bit loads per lane in flight

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSampleThreads)
sample_kernel(const __grid_constant__ SampleLaunch S) {
    extern __shared__ unsigned s_slots[];   // leader only: [nslots]
    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.x / kClusterSize;
    const int rank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31;
    const int w = rank * (kSampleThreads / 32) + (threadIdx.x >> 5);
    const int ntask = S.G.ntask_img;
    const int tps = S.tps;
    const int nslots = ((ntask + tps - 1) / tps) * 32;
    if (rank == 0)
        for (int i = threadIdx.x; i < nslots; i += kSampleThreads) s_slots[i] = 0u;
    cluster.sync();
    unsigned *slots = cluster.map_shared_rank(s_slots, 0);
    const int nsuper = (ntask + 31) / 32;
    for (int sp = w; sp < nsuper; sp += kSampleWarps) {
        // lane i decodes task t0 + i ONCE (the decode is ~100 instructions); the warp then walks the 32 tasks
        // with the unit's first address and valid lane range broadcast by shuffles
        const int t0 = sp * 32;
        const float *uptr = nullptr;
        int range = 0;   // lo | hi << 8: lanes lo <= L < hi of the unit exist
        {
            const int t = t0 + lane;
            if (t < ntask) {
                const unsigned j = stream_hash((unsigned)b, (unsigned)t) >> 26;
                if (j < 32u) {
                    const STask k = stream_task(S.G, b, t);
                    const int g = k.g0 + (int)j * 32;
                    const int lo = max(k.f0 - g, 0), hi = min(k.f1 - g, 32);
                    if (hi > lo) {
                        uptr = k.blk + ((ptrdiff_t)g * 4 - k.mis);
                        range = lo | (hi << 8);
                    }
                }
            }
        }
        float acc = -INFINITY;
        bool acc_any = false;
#pragma unroll 1
        for (int i8 = 0; i8 < 32; i8 += kSampleBatch) {
            float4 v[kSampleBatch];
            bool on[kSampleBatch];
#pragma unroll
            for (int i = 0; i < kSampleBatch; ++i) {
                const float *ptr = reinterpret_cast<const float *>(__shfl_sync(0xffffffffu, (unsigned long long)uptr, i8 + i));
                const int rg = __shfl_sync(0xffffffffu, range, i8 + i);
                on[i] = lane >= (rg & 0xFF) && lane < (rg >> 8);
                v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                if (on[i]) v[i] = ld_stream4(ptr + lane * 4);
            }
            if (tps >= kSampleBatch) {
#pragma unroll
                for (int i = 0; i < kSampleBatch; ++i)
                    if (on[i]) { acc = fmaxf(acc, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w))); acc_any = true; }
                const int tend = t0 + i8 + kSampleBatch;   // a slot group ends here, or the warp's 32 tasks do
                if ((tend & (tps - 1)) == 0 || i8 + kSampleBatch == 32) {
                    if (acc_any) atomicMax(slots + (size_t)((tend - 1) / tps) * 32 + lane, vkey_of(acc));
                    acc = -INFINITY; acc_any = false;
                }
            } else {
                for (int i0 = 0; i0 < kSampleBatch; i0 += tps) {
                    float m = -INFINITY;
                    bool any = false;
#pragma unroll
                    for (int i = 0; i < kSampleBatch; ++i)
                        if (i >= i0 && i < i0 + tps && on[i]) { m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w))); any = true; }
                    if (any) atomicMax(slots + (size_t)((t0 + i8 + i0) / tps) * 32 + lane, vkey_of(m));
                }
            }
        }
    }
    cluster.sync();   // every maximum has landed in the leader's slots
    if (rank == 0) {
        const unsigned thr = threshold_from_slots<kSampleThreads>(s_slots, nslots, S.N, S.K);
        if (threadIdx.x == 0) {
            S.thr[b] = (S.N <= kCap) ? 0u : thr;
            if (S.zero0) S.zero0[b] = 0u;
            if (S.zero1) S.zero1[b] = 0u;
            if (S.zero2) S.zero2[b] = 0u;
            if (b == 0 && S.zero_scalar) *S.zero_scalar = 0u;
        }
    }
}

// ---- fused stream + tails ---------------------------------------------------------------------------------
struct PostArgs {
    StreamGeo G;
    TopkArgs T;               // geometry, box levels, K, thr / cnt / flag / cand, top-k outputs (always valid)
    int emit_topk;            // also write the top-k tensors of every image (the caller asked for them)
    unsigned *queue, *done;   // next task; completed tasks per image
    unsigned total_tasks;
    const float4 *anchors;
    const float *scale, *size;
    odk_detect_params p;
    float nms_thr_f;
    int cap;                  // candidate slots carved out of shared memory (multiple of 1024, >= K)
    float *dets;
    int *count, *src;
    long long *det_anchor;    // nullable
    float *energy, *max_logit;   // nullable
    float ood_T;
};

static inline size_t post_det_bytes(int cap) { return (size_t)cap * 26 + (size_t)(cap / 32) * 4 + 16 + kNmsMaskBytes; }

// rows of one image out of the sorted keys each thread holds (rank = tid + k * 1024)
static __device__ void detect_sorted(const PostArgs &P, int b, const unsigned long long (&keys)[kPostKeys], unsigned char *raw) {
    __shared__ int s_cnt[kPostKeys * kPostWarps + 1];
    __shared__ float s_wmax[kPostWarps];
    __shared__ int s_kept[1024];
    __shared__ float s_keptscore[1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.T.K, D = P.p.max_det, cap = P.cap;
    float4 *sbox = reinterpret_cast<float4 *>(raw);
    float *sscore = reinterpret_cast<float *>(raw + (size_t)cap * 16);
    unsigned *sflat = reinterpret_cast<unsigned *>(raw + (size_t)cap * 20);
    unsigned short *srank = reinterpret_cast<unsigned short *>(raw + (size_t)cap * 24);
    unsigned *alive = reinterpret_cast<unsigned *>(raw + (size_t)cap * 26);
    const bool has_scale = P.scale != nullptr;
    const bool clip = has_scale && P.size != nullptr;                       // anchors.py:137
    const float scale = has_scale ? __ldg(P.scale + b) : 1.0f;
    const float lim_x = clip ? __fdiv_rn(__ldg(P.size + 2 * b), scale) : 0.f;
    const float lim_y = clip ? __fdiv_rn(__ldg(P.size + 2 * b + 1), scale) : 0.f;

    // 1. scores, score filter, order-preserving compaction (anchors.py:140-144): one scan for all rounds
    float sc[kPostKeys];
    unsigned bal[kPostKeys];
#pragma unroll
    for (int k = 0; k < kPostKeys; ++k) {
        const int i = tid + k * kPostThreads;
        bool ok = i < K;
        sc[k] = 0.f;
        if (ok) { sc[k] = sigmoid_ref(val_of((unsigned)(keys[k] >> 32))); ok = sc[k] > P.p.score_min; }
        bal[k] = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_cnt[k * kPostWarps + warp] = __popc(bal[k]);
    }
    __syncthreads();
    if (warp == 0) {
        int v[kPostKeys], run = 0;
#pragma unroll
        for (int j = 0; j < kPostKeys; ++j) { v[j] = s_cnt[lane * kPostKeys + j]; run += v[j]; }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        int excl = inc - run;
#pragma unroll
        for (int j = 0; j < kPostKeys; ++j) { s_cnt[lane * kPostKeys + j] = excl; excl += v[j]; }
        if (lane == 31) s_cnt[kPostKeys * kPostWarps] = inc;
    }
    __syncthreads();
    const int n = s_cnt[kPostKeys * kPostWarps];
    int kept_n = 0;
    if (n > 0) {   // uniform
        // 2. anchor / regression gathers and decode, three rows per thread in flight
        int slot[kPostKeys];
        float mx = -INFINITY;
#pragma unroll
        for (int k0 = 0; k0 < kPostKeys; k0 += 3) {
            float4 an[3], rg[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int k = k0 + j;
                slot[k] = -1;
                an[j] = rg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((bal[k] >> lane) & 1u) {
                    slot[k] = s_cnt[k * kPostWarps + warp] + __popc(bal[k] & ((1u << lane) - 1u));
                    const unsigned flat = ~(unsigned)(keys[k] & 0xFFFFFFFFull);
                    const int anchor = (int)fd_div(flat, P.T.div_C);
                    an[j] = __ldg(P.anchors + anchor);
                    rg[j] = gather_box(P.T, b, anchor);
                }
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int k = k0 + j;
                if (slot[k] >= 0) {
                    const float4 o = decode_xyxy(an[j], rg[j], clip, lim_x, lim_y);
                    sbox[slot[k]] = o;
                    sscore[slot[k]] = sc[k];
                    sflat[slot[k]] = ~(unsigned)(keys[k] & 0xFFFFFFFFull);
                    srank[slot[k]] = (unsigned short)(tid + k * kPostThreads);
                    mx = fmaxf(mx, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
                }
            }
        }
        // 3. coordinate trick: boxes + class * (max_coordinate + 1)  (torchvision boxes.py:105-108)
        mx = warp_max(mx);
        if (lane == 0) s_wmax[warp] = mx;
        init_alive(alive, n, cap);
        __syncthreads();
        mx = s_wmax[0];
        for (int w = 1; w < kPostWarps; ++w) mx = fmaxf(mx, s_wmax[w]);
        const float mul = __fadd_rn(mx, 1.0f);
#pragma unroll
        for (int k = 0; k < kPostKeys; ++k) {
            if (slot[k] >= 0) {
                const unsigned flat = ~(unsigned)(keys[k] & 0xFFFFFFFFull);
                const unsigned anchor = fd_div(flat, P.T.div_C);
                const float off = __fmul_rn((float)(flat - anchor * (unsigned)P.T.C), mul);
                float4 o = sbox[slot[k]];
                o.x = __fadd_rn(o.x, off); o.y = __fadd_rn(o.y, off); o.z = __fadd_rn(o.z, off); o.w = __fadd_rn(o.w, off);
                sbox[slot[k]] = o;
            }
        }
        __syncthreads();
        // 4. suppression, first D survivors (the candidates are in descending score order by construction)
        DetSmem S;
        S.box = sbox; S.score = sscore; S.src = nullptr; S.alive = alive;
        if (P.p.soft_nms)
            kept_n = soft_nms_rounds(S, n, true, P.p.soft_sigma, P.p.soft_iou, P.p.soft_score_thr, D, s_kept,
                                     kDetFirstWindow, kDetThreads, kSoftGroup, [&](int q, int i, float s) { s_keptscore[q] = s; });
        else
            kept_n = hard_nms_rounds(S, n, P.nms_thr_f, D, s_kept, alive + cap / 32 + 4);
        __syncthreads();
    }
    // 5. rows: boxes (re-decoded, unoffset) * img_scale, score, class + 1 (anchors.py:153-166)
    float *dets = P.dets + (size_t)b * D * 6;
    int *src = P.src + (size_t)b * D;
    for (int q = tid; q < D; q += kPostThreads) {
        float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int sp = -1;
        long long anc = -1;
        if (q < kept_n) {
            const int i = s_kept[q];
            const unsigned flat = sflat[i];
            const unsigned anchor = fd_div(flat, P.T.div_C);
            sp = (int)srank[i];
            anc = (long long)anchor;
            float4 o = decode_xyxy(__ldg(P.anchors + anchor), gather_box(P.T, b, (int)anchor), clip, lim_x, lim_y);
            if (has_scale) { o.x = __fmul_rn(o.x, scale); o.y = __fmul_rn(o.y, scale); o.z = __fmul_rn(o.z, scale); o.w = __fmul_rn(o.w, scale); }
            r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = o.w;
            r[4] = P.p.soft_nms ? s_keptscore[q] : sscore[i];
            r[5] = (float)((flat - anchor * (unsigned)P.T.C) + 1u);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) dets[q * 6 + k] = r[k];
        src[q] = sp;
        if (P.det_anchor) P.det_anchor[(size_t)b * D + q] = anc;
    }
    if (tid == 0) P.count[b] = kept_n;
    // 6. OOD scores over the C raw logits of each detection's source anchor (one warp per detection)
    if (P.energy) {
        for (int q = warp; q < D; q += kPostWarps) {
            float e = 0.f, m = 0.f;
            if (q < kept_n) ood_row(P.T.g, P.G.cls, b, P.T.C, (long long)fd_div(sflat[s_kept[q]], P.T.div_C), P.ood_T, lane, e, m);
            if (lane == 0) { P.energy[(size_t)b * D + q] = e; P.max_logit[(size_t)b * D + q] = m; }
        }
    }
    __syncthreads();   // shared memory is free again
}

template <int E>
__device__ __forceinline__ void load_sorted(const unsigned long long *s, int K, unsigned long long (&keys)[kPostKeys]) {
#pragma unroll
    for (int k = 0; k < kPostKeys; ++k) {
        const int i = threadIdx.x + k * kPostThreads;
        keys[k] = i < K ? sorted_at<E>(s, i) : 0ull;
    }
}

// everything of image b that is not streaming, by the whole CTA
static __device__ void run_tail(const PostArgs &P, int b, unsigned long long *s) {
    __threadfence();   // acquire side of the done-counter hand-off: the candidates of every other SM are visible
    const unsigned n = __ldcg(P.T.cnt + b);
    if (n < (unsigned)P.T.K || n > (unsigned)kCap) {   // uniform: the exact path (odk_topk.cu) takes this image
        if (threadIdx.x == 0) P.T.flag[b] = 1u;
        return;
    }
    unsigned long long keys[kPostKeys];
    const Refined R = refine_candidates(P.T, b, (int)n, s);
    if (R.ranked) {
        if (P.emit_topk) emit_topk<0, true>(P.T, b, s + kSortSlots);
        load_sorted<0>(s + kSortSlots, P.T.K, keys);
    } else if (R.m <= kSortSlots) {
        for (int i = R.m + threadIdx.x; i < kSortSlots; i += blockDim.x) s[i] = 0ull;
        __syncthreads();
        block_sort_desc<8>(s);
        if (P.emit_topk) emit_topk<8, true>(P.T, b, s);
        load_sorted<8>(s, P.T.K, keys);
    } else {   // more than 8192 keys tie inside one sub-bin: sort everything
        const unsigned long long *cand = P.T.cand + (size_t)b * kCap;
        __syncthreads();
        for (int i = threadIdx.x; i < 16 * kSelThreads; i += blockDim.x) s[i] = i < (int)n ? __ldcg(cand + i) : 0ull;
        __syncthreads();
        block_sort_desc<16>(s);
        if (P.emit_topk) emit_topk<16, true>(P.T, b, s);
        load_sorted<16>(s, P.T.K, keys);
    }
    __syncthreads();   // the keys are in registers: the buffer becomes the detection arrays
    detect_sorted(P, b, keys, reinterpret_cast<unsigned char *>(s));
}

__global__ void __launch_bounds__(kPostThreads, 1) post_fused_kernel(const __grid_constant__ PostArgs P) {
    extern __shared__ __align__(16) unsigned long long s_dyn[];
    __shared__ unsigned long long s_wst[kPostWarps][kWarpStage];
    __shared__ unsigned s_wn[kPostWarps];
    __shared__ int s_tailq[64];
    __shared__ unsigned s_tail_n, s_tail_done;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_tail_n = 0u; s_tail_done = 0u; }
    if (tid < kPostWarps) s_wn[tid] = 0u;
    __syncthreads();

    const unsigned total = P.total_tasks;
    const int ntask = P.G.ntask_img;
    auto grab = [&]() {
        unsigned t = 0u;
        if (lane == 0) t = atomicAdd(P.queue, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    // A task a warp holds but has not streamed yet keeps its image open, so no warp may sit on one while its
    // CTA runs a tail (the image would wait for the tail, and the CTA would then finish -- and inherit the tail
    // of -- every following image too): the next task is only taken ahead of time while no hand-off is pending.
    constexpr unsigned kNoTask = 0xFFFFFFFFu;
    auto tail_pending = [&]() {
        unsigned p = 0u;
        if (lane == 0) p = *(volatile unsigned *)&s_tail_n != *(volatile unsigned *)&s_tail_done;
        return __shfl_sync(0xffffffffu, p, 0) != 0u;
    };
    bool drained = false;
    unsigned next = grab();
    for (;;) {
        if (!drained) {
            if (next == kNoTask) next = grab();
            const unsigned cur = next;
            if (cur < total) {
                next = tail_pending() ? kNoTask : grab();   // its latency hides behind this task's loads
                const unsigned b = fd_div(cur, P.G.div_ntask);
                const STask k = stream_task(P.G, (int)b, (int)(cur - b * (unsigned)ntask));
                const float thr_f = thr_float(__ldcg(P.T.thr + b));
                unsigned *cnt = P.T.cnt + b;
                unsigned long long *cand = P.T.cand + (size_t)b * kCap;
                auto hit = [&](float x, int e) {
                    if (x >= thr_f) {
                        const unsigned long long key = ((unsigned long long)vkey_of(x) << 32) |
                                                       (unsigned long long)(~stream_flat(P.G, k.l, (unsigned)e));
                        const unsigned slot = atomicAdd(&s_wn[warp], 1u);
                        if (slot < (unsigned)kWarpStage) {
                            s_wst[warp][slot] = key;
                        } else {   // staging full (threshold far too low): straight to the global list
                            const unsigned pos = atomicAdd(cnt, 1u);
                            if (pos < (unsigned)kCap) cand[pos] = key;
                        }
                    }
                };
                for (int base = k.f0; base < k.f1; base += 256) {
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int g = base + lane + 32 * j;
                        v[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                        if (g < k.f1) v[j] = ld_stream4(k.blk + ((ptrdiff_t)g * 4 - k.mis));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // one compare per float4 on the common path; hits are ~1 in 1600 elements
                        if (fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)) >= thr_f) {
                            const int g = base + lane + 32 * j;
                            if (g < k.f1) {
                                const int e = g * 4 - k.mis;
                                hit(v[j].x, e); hit(v[j].y, e + 1); hit(v[j].z, e + 2); hit(v[j].w, e + 3);
                            }
                        }
                    }
                }
                // partial groups at the two ends of a block whose length or start is not a multiple of 16 bytes
                if (k.g0 < k.g1) {
                    const bool first = k.g0 < k.f0;
                    const bool last = k.f1 < k.g1 && !(first && k.g1 - 1 == k.g0);
                    if (first && lane < 4) {
                        const int e = k.g0 * 4 - k.mis + lane;
                        if (e >= 0 && (unsigned)e < k.lb) hit(ld_stream1(k.blk + e), e);
                    }
                    if (last && lane >= 4 && lane < 8) {
                        const int e = (k.g1 - 1) * 4 - k.mis + (lane - 4);
                        if (e >= 0 && (unsigned)e < k.lb) hit(ld_stream1(k.blk + e), e);
                    }
                }
                __syncwarp();
                const unsigned nst = min(*(volatile unsigned *)&s_wn[warp], (unsigned)kWarpStage);
                if (nst) {   // warp-uniform
                    unsigned pos0 = 0u;
                    if (lane == 0) pos0 = atomicAdd(cnt, nst);
                    pos0 = __shfl_sync(0xffffffffu, pos0, 0);
                    for (unsigned i = lane; i < nst; i += 32)
                        if (pos0 + i < (unsigned)kCap) cand[pos0 + i] = s_wst[warp][i];
                    __syncwarp();
                    if (lane == 0) s_wn[warp] = 0u;
                }
                // release: every lane's candidate stores are visible before the task is reported done
                __threadfence();
                __syncwarp();
                unsigned d = 0u;
                if (lane == 0) d = atomicAdd(P.done + b, 1u);
                d = __shfl_sync(0xffffffffu, d, 0);
                if (d == (unsigned)ntask - 1u && lane == 0) {   // the image is complete: hand it to this CTA
                    const unsigned slot = atomicAdd(&s_tail_n, 1u);
                    s_tailq[slot & 63u] = (int)b;
                }
                __syncwarp();
            } else {
                drained = true;
            }
        }
        if (drained || tail_pending()) {
            if (!drained && next != kNoTask) continue;   // stream the task in hand first (nothing new is taken)
            __syncthreads();   // all 32 warps are here: nobody is streaming, the tail queue is stable
            const unsigned n1 = s_tail_n, d0 = s_tail_done;
            for (unsigned i = d0; i != n1; ++i) run_tail(P, s_tailq[i & 63u], s_dyn);
            __syncthreads();
            if (tid == 0) s_tail_done = n1;
            if (__syncthreads_and(drained ? 1 : 0)) break;   // every warp is out of tasks, every hand-off served
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------
int make_stream_geo(StreamGeo *G, const Geo &g, const void *const *cls_levels, int C) {
    memset(G, 0, sizeof(*G));
    G->nlev = g.nlev; G->na = g.na; G->C = C;
    G->div_C = make_fastdiv((unsigned)C);
    int toff = 0;
    for (int l = 0; l < g.nlev; ++l) {
        G->cls[l] = (const float *)cls_levels[l];
        if (!G->cls[l]) return set_error(ODK_EINVAL, "null class-logit level pointer (level %d)", l);
        if ((uintptr_t)G->cls[l] & 3) return set_error(ODK_EINVAL, "class-logit level %d is not 4-byte aligned", l);
        const long long lb = (long long)g.na * C * g.hw[l];
        if (lb > 0x7fffffffll / 2) return set_error(ODK_EUNSUPPORTED, "level %d has more than 2^30 logits per image", l);
        G->lb[l] = (unsigned)lb;
        G->hw[l] = g.hw[l];
        G->off[l] = g.off[l];
        G->div_hw[l] = make_fastdiv((unsigned)g.hw[l]);
        G->task_off[l] = toff;
        const long long groups_max = (lb + 6) / 4;   // whatever the block's alignment is
        toff += (int)((groups_max + kGroupsPerTask - 1) / kGroupsPerTask);
    }
    for (int l = g.nlev; l <= ODK_MAX_LEVELS; ++l) G->task_off[l] = toff;
    G->ntask_img = toff;
    G->div_ntask = make_fastdiv((unsigned)toff);
    return ODK_OK;
}

size_t sample_slot_stride(const StreamGeo &G, int *tps_out) {
    // aim at 4096..8192 slots per image: enough resolution for the rank statistics, little to histogram
    int tps = 1;
    while ((long long)G.ntask_img * 32 / (tps * 2) >= 4096) tps *= 2;
    if (tps_out) *tps_out = tps;
    return (size_t)((G.ntask_img + tps - 1) / tps) * 32;
}

int launch_sample(const SampleLaunch &s, cudaStream_t st) {
    const size_t smem = (size_t)((s.G.ntask_img + s.tps - 1) / s.tps) * 32 * sizeof(unsigned);
    if (smem > 64 * 1024) return set_error(ODK_EUNSUPPORTED, "odk sample: %zu bytes of slots", smem);
    if (smem > 32 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "odk sample: %s", cudaGetErrorString(e));
    }
    sample_kernel<<<s.B * kClusterSize, kSampleThreads, smem, st>>>(s);
    return check_launch("odk sample_kernel");
}

struct PostWs {
    size_t slots, thr, cnt, flag, done, queue, cand, tk_val, tk_box, tk_idx, tk_cls, total;
    int slot_stride, tps;
};

static PostWs post_ws_layout(const StreamGeo &G, int B, int K) {
    PostWs w;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.slot_stride = (int)sample_slot_stride(G, &w.tps);
    size_t off = 0;
    w.slots = off; off = al(off + (size_t)B * w.slot_stride * 4);
    w.thr = off; off = al(off + (size_t)B * 4);
    w.cnt = off; off = al(off + (size_t)B * 4);
    w.flag = off; off = al(off + (size_t)B * 4);
    w.done = off; off = al(off + (size_t)B * 4);
    w.queue = off; off = al(off + 4);
    w.cand = off; off = al(off + (size_t)B * kCap * 8);
    w.tk_val = off; off = al(off + (size_t)B * K * 4);
    w.tk_box = off; off = al(off + (size_t)B * K * 16);
    w.tk_idx = off; off = al(off + (size_t)B * K * 8);
    w.tk_cls = off; off = al(off + (size_t)B * K * 8);
    w.total = off;
    return w;
}

// defined in odk_topk.cu / odk_detect.cu: the exact path for the images the fused kernel flagged
int launch_topk_exact_flagged(const TopkArgs &a, cudaStream_t st);
int launch_detect_flagged(const float *cls_topk, const float *box_topk, const int64_t *indices, const int64_t *classes, int B,
                          int N, const float *anchors, int64_t A, const float *img_scale, const float *img_size,
                          const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, int64_t *det_anchor,
                          const unsigned *only_flag, cudaStream_t st);
int launch_ood_flagged(const Geo &g, const void *const *cls_levels, int B, int C, const int64_t *anchor_idx, int D, float T,
                       float *energy, float *max_logit, const unsigned *only_flag, cudaStream_t st);

}  // namespace odk

extern "C" {

static int post_ws_for(int B, int C, const int32_t *level_hw, int num_levels, int na, int K, odk::PostWs *w) {
    using namespace odk;
    Geo g;
    if (B < 1 || C < 1 || K < 1 || make_geo(&g, level_hw, num_levels, na)) return 1;
    StreamGeo G;
    memset(&G, 0, sizeof(G));
    for (int l = 0; l < g.nlev; ++l) {
        const long long lb = (long long)g.na * C * g.hw[l];
        G.ntask_img += (int)(((lb + 6) / 4 + kGroupsPerTask - 1) / kGroupsPerTask);
    }
    *w = post_ws_layout(G, B, K);
    return 0;
}

size_t odk_postprocess_workspace_bytes(int B, int C, const int32_t *level_hw, int num_levels, int na, int K) {
    odk::PostWs w;
    return post_ws_for(B, C, level_hw, num_levels, na, K, &w) ? 0 : w.total;
}

size_t odk_postprocess_flags_offset(int B, int C, const int32_t *level_hw, int num_levels, int na, int K) {
    odk::PostWs w;
    return post_ws_for(B, C, level_hw, num_levels, na, K, &w) ? 0 : w.flag;
}

int odk_postprocess(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
                    int num_levels, int na, int K, const float *anchors, const float *img_scale, const float *img_size,
                    const odk_detect_params *params, float temperature, float *dets, int32_t *count, int32_t *src,
                    int64_t *det_anchor, float *energy, float *max_logit, float *cls_topk, float *box_topk,
                    int64_t *indices, int64_t *classes, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    PostArgs P;
    memset(&P, 0, sizeof(P));
    TopkArgs &a = P.T;
    int rc = make_geo(&a.g, level_hw, num_levels, na);
    if (rc) return rc;
    if (!cls_levels || !box_levels || !anchors || !params || !dets || !count || !src)
        return set_error(ODK_EINVAL, "odk_postprocess: null pointer");
    if (B < 1 || C < 1 || K < 1) return set_error(ODK_EINVAL, "odk_postprocess: B, C, K must be positive");
    if (B > 65535) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: batch > 65535");
    a.N = (long long)a.g.A * C;
    if (a.N > 0xFFFFFFFFll) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: A*C does not fit 32 bits");
    if ((long long)K > a.N) return set_error(ODK_EINVAL, "odk_postprocess: K=%d exceeds A*C=%lld (selected index k out of range)", K, a.N);
    if (K > kPostMaxK) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: K > %d (use odk_topk + odk_detect)", kPostMaxK);
    if ((size_t)B * na * (size_t)C > 0x7fffffffull) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: B*na*C overflows int");
    if (params->max_det < 1 || params->max_det > 1024) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: max_det must be in [1,1024]");
    if ((energy == nullptr) != (max_logit == nullptr)) return set_error(ODK_EINVAL, "odk_postprocess: energy and max_logit go together");
    if (energy && !(temperature > 0.0f)) return set_error(ODK_EINVAL, "odk_postprocess: temperature must be positive");
    if (energy && !det_anchor) return set_error(ODK_EINVAL, "odk_postprocess: OOD scores need det_anchor");
    const bool want_topk = cls_topk || box_topk || indices || classes;
    if (want_topk && !(cls_topk && box_topk && indices && classes))
        return set_error(ODK_EINVAL, "odk_postprocess: the four top-k outputs go together");
    if (((uintptr_t)anchors | (uintptr_t)workspace | (uintptr_t)box_topk) & 15)
        return set_error(ODK_EINVAL, "odk_postprocess: anchors / box_topk / workspace must be 16-byte aligned");
    rc = make_stream_geo(&P.G, a.g, cls_levels, C);
    if (rc) return rc;
    const PostWs w = post_ws_layout(P.G, B, K);
    if (!workspace || workspace_bytes < w.total)
        return set_error(ODK_EWORKSPACE, "odk_postprocess: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    a.B = B; a.C = C; a.K = K; a.planes = na * C;
    a.div_C = make_fastdiv((unsigned)C);
    int toff = 0;
    for (int l = 0; l < num_levels; ++l) {   // per-plane task model of odk_topk.cu (the exact fallback uses it)
        a.cls[l] = (const float *)cls_levels[l];
        a.box[l] = (const float *)box_levels[l];
        if (!a.box[l]) return set_error(ODK_EINVAL, "odk_postprocess: null box level pointer (level %d)", l);
        a.vec[l] = (a.g.hw[l] % 4 == 0 && ((uintptr_t)a.cls[l] & 15) == 0) ? 4 : 1;
        a.nvec[l] = a.g.hw[l] / a.vec[l];
        a.nseg[l] = (a.nvec[l] + kSegVec - 1) / kSegVec;
        a.div_nseg[l] = make_fastdiv((unsigned)a.nseg[l]);
        a.task_off[l] = toff;
        toff += a.planes * a.nseg[l];
    }
    for (int l = num_levels; l <= ODK_MAX_LEVELS; ++l) a.task_off[l] = toff;
    char *ws = (char *)workspace;
    a.slots = (unsigned *)(ws + w.slots); a.thr = (unsigned *)(ws + w.thr); a.cnt = (unsigned *)(ws + w.cnt);
    a.flag = (unsigned *)(ws + w.flag); a.cand = (unsigned long long *)(ws + w.cand);
    a.out_val = want_topk ? cls_topk : (float *)(ws + w.tk_val);
    a.out_box = want_topk ? box_topk : (float *)(ws + w.tk_box);
    a.out_idx = want_topk ? (long long *)indices : (long long *)(ws + w.tk_idx);
    a.out_cls = want_topk ? (long long *)classes : (long long *)(ws + w.tk_cls);
    a.fused = 1;
    P.emit_topk = want_topk ? 1 : 0;
    P.queue = (unsigned *)(ws + w.queue); P.done = (unsigned *)(ws + w.done);
    P.total_tasks = (unsigned)B * (unsigned)P.G.ntask_img;
    if ((long long)B * P.G.ntask_img > 0x7fffffffll) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: too many tasks");
    P.anchors = (const float4 *)anchors; P.scale = img_scale; P.size = img_size;
    P.p = *params; P.nms_thr_f = float_at_or_below(params->nms_iou);
    P.cap = det_cap(K);
    P.dets = dets; P.count = count; P.src = src; P.det_anchor = (long long *)det_anchor;
    P.energy = energy; P.max_logit = max_logit; P.ood_T = temperature;

    cudaStream_t st = (cudaStream_t)stream;
    SampleLaunch s;
    memset(&s, 0, sizeof(s));
    s.G = P.G; s.B = B; s.K = K; s.N = a.N; s.slots = a.slots; s.slot_stride = w.slot_stride; s.tps = w.tps; s.thr = a.thr;
    s.zero0 = a.cnt; s.zero1 = a.flag; s.zero2 = P.done; s.zero_scalar = P.queue;
    rc = launch_sample(s, st);
    if (rc) return rc;

    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
    size_t smem = (size_t)kCap * 8;
    if (post_det_bytes(P.cap) > smem) smem = post_det_bytes(P.cap);
    cudaError_t e = cudaFuncSetAttribute(post_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, "odk_postprocess: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    unsigned grid = (unsigned)sms;
    const unsigned useful = (P.total_tasks + kPostWarps - 1) / kPostWarps;
    if (grid > useful) grid = useful;
    if (grid < 1) grid = 1;
    post_fused_kernel<<<grid, kPostThreads, smem, st>>>(P);
    rc = check_launch("odk_postprocess/post_fused_kernel");
    if (rc) return rc;
    // flagged images only (none for real score distributions): exact select, then their detections / OOD scores
    rc = launch_topk_exact_flagged(a, st);
    if (rc) return rc;
    rc = launch_detect_flagged(a.out_val, a.out_box, (const int64_t *)a.out_idx, (const int64_t *)a.out_cls, B, K, anchors,
                               a.g.A, img_scale, img_size, params, dets, count, src, det_anchor, a.flag, st);
    if (rc) return rc;
    if (energy) {
        rc = launch_ood_flagged(a.g, cls_levels, B, C, det_anchor, params->max_det, temperature, energy, max_logit, a.flag, st);
    }
    return rc;
}

}  // extern "C"
