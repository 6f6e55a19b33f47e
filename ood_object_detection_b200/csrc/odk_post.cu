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
#include <stdlib.h>

#include "odk_stream.cuh"
#include "odk_detect.cuh"

namespace cg = cooperative_groups;

namespace odk {

__device__ __forceinline__ unsigned long long global_ns() {   // (stamp() of odk_topk.cuh for the per-image marks)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

constexpr int kPostThreads = 1024;
constexpr int kPostWarps = kPostThreads / 32;
constexpr int kWarpStage = 24;        // staged hits per warp, task and buffer (~2.6 expected)
constexpr int kPostKeys = 6;          // sorted keys per thread: K <= 6144
constexpr int kPostMaxK = kPostKeys * kPostThreads;
constexpr int kSampleThreads = 512;
constexpr int kSampleWarps = kClusterSize * (kSampleThreads / 32);

// ---- sample ---------------------------------------------------------------------------------------------
// Task t of image b contributes its 512-byte unit j = hash(b, t) >> (32 - shift) when j < 64 (and the task has that
// many), so every unit is taken with probability 2^-shift.  Lane maxima over `tps` consecutive tasks form one slot group.
// The slots live in the shared memory of the cluster's leader CTA: the other CTAs merge their maxima into
// them with distributed-shared-memory atomics, so the threshold statistics never touch global memory.
static_assert(kUnitsPerTask <= (1 << kSampleShift), "one sampled unit per task at most");
constexpr int kSampleBatch = 8;   // 128-bit loads per lane in flight

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSampleThreads, 2)
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
                const unsigned j = stream_hash((unsigned)b, (unsigned)t) >> (32 - S.shift);   // 1 unit of 2^shift
                if (j < (unsigned)kUnitsPerTask) {
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
        const uint2 thr = threshold_from_slots<kSampleThreads>(s_slots, nslots, S.N, S.K, S.shift);
        if (threadIdx.x == 0) {
            S.thr[b] = (S.N <= kCap) ? 0u : thr.x;
            S.thr_hi[b] = thr.y;
            if (S.zero0) S.zero0[b] = 0u;
            if (S.zero1) S.zero1[b] = 0u;
            if (S.zero2) S.zero2[b] = 0u;
            if (S.zero3) S.zero3[b] = 0u;
            if (b == 0 && S.zero_scalar)
                for (int i = 0; i < S.zero_scalars; ++i) S.zero_scalar[i] = 0u;
            if (b == 0 && S.zero_u64) *S.zero_u64 = 0ull;
        }
    }
}

// ---- fused stream + tails ---------------------------------------------------------------------------------
struct PostArgs {
    StreamGeo G;
    TopkArgs T;               // geometry, box levels, K, thr / cnt / flag / cand, top-k outputs (always valid)
    int emit_topk;            // also write the top-k tensors of every image (the caller asked for them)
    unsigned *queue, *done;   // next task; completed tasks per image
    unsigned *tq_img;         // [B] completed images in completion order (image + 1; 0 = not yet published)
    unsigned *tq_pushed, *tq_claimed;   // completed images; images some CTA has taken the tail of
    int chunk, ahead;                   // tasks per global queue access (power of two); chunks the bookkeeper stays ahead
    int debug_skip_tails;               // diagnostics (ODK_POST_SKIP_TAILS=1): stream only, every image is left to the exact path
    unsigned long long *timeline;       // [B][kStampSlots] globaltimer ns: 0 stream complete, 1 tail claimed, 2 tail start, 3 tail end,
                                        // 4 select done, 5 filter + compaction, 6 gathers + decode, 7 class offsets, 8 suppression,
                                        // 9-12 inside select: histogram, scans, scatter, ranks; then kernel start, kernel end
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

static inline size_t post_det_bytes(int cap) { return (size_t)cap * 26 + (size_t)(cap / 32) * 4 + 16 + kSoftScratchBytes; }   // (>= the hard-NMS masks)

// rows of one image out of the sorted keys each thread holds (rank = tid + k * 1024)
template <bool SOFT>
static __device__ void detect_sorted(const PostArgs &P, int b, const unsigned long long (&keys)[kPostKeys],
                                     const int (&cpos)[kPostKeys], unsigned char *raw) {
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
    stamp(P.timeline, b, 5);
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
                    // the collect stage left every candidate's regression row next to its key: one sector instead of four
                    rg[j] = cpos[k] >= 0 ? __ldcg(P.T.cand_box + (size_t)b * kCap + cpos[k]) : gather_box(P.T, b, anchor);
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
        stamp(P.timeline, b, 6);
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
        stamp(P.timeline, b, 7);
        // 4. suppression, first D survivors (the candidates are in descending score order by construction)
        DetSmem S;
        S.box = sbox; S.score = sscore; S.src = nullptr; S.alive = alive;
        if (SOFT)
            kept_n = soft_nms_batched(S, n, true, P.p.soft_sigma, P.p.soft_iou, P.p.soft_score_thr, D, s_kept,
                                      kDetFirstWindow, kDetThreads, [&](int q, int i, float s) { s_keptscore[q] = s; },
                                      reinterpret_cast<unsigned char *>(alive + cap / 32 + 4));
        else
            kept_n = hard_nms_rounds(S, n, P.nms_thr_f, D, s_kept, alive + cap / 32 + 4);
        __syncthreads();
        stamp(P.timeline, b, 8);
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
            r[4] = SOFT ? s_keptscore[q] : sscore[i];
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
        if (P.T.C <= 128) {
            // the scattered loads of four detections per warp are issued together (one exposed latency per 128 rows)
            for (int q0 = warp; q0 < D; q0 += 4 * kPostWarps) {
                float v[4][4];
                bool ok[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = q0 + j * kPostWarps;
                    OodRow r;
                    r.row = nullptr; r.cstride = 0;
                    if (q < kept_n) r = ood_row_of(P.T.g, P.G.cls, P.G.nhwc, b, P.T.C, (long long)fd_div(sflat[s_kept[q]], P.T.div_C));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        v[j][k] = (r.row && lane + 32 * k < P.T.C) ? __ldg(r.row + (size_t)(lane + 32 * k) * r.cstride) : 0.f;
                    ok[j] = r.row != nullptr;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = q0 + j * kPostWarps;
                    if (q >= D) continue;
                    float e = 0.f, m = 0.f;
                    if (ok[j]) ood_reduce(v[j], P.T.C, P.ood_T, lane, e, m);
                    if (lane == 0) { P.energy[(size_t)b * D + q] = e; P.max_logit[(size_t)b * D + q] = m; }
                }
            }
        } else {
            for (int q = warp; q < D; q += kPostWarps) {
                float e = 0.f, m = 0.f;
                if (q < kept_n) ood_row(P.T.g, P.G.cls, P.G.nhwc, b, P.T.C, (long long)fd_div(sflat[s_kept[q]], P.T.div_C), P.ood_T, lane, e, m);
                if (lane == 0) { P.energy[(size_t)b * D + q] = e; P.max_logit[(size_t)b * D + q] = m; }
            }
        }
    }
    __syncthreads();   // shared memory is free again
}

template <int E>
__device__ __forceinline__ void load_sorted(const unsigned long long *s, const unsigned short *srcpos, int K,
                                            unsigned long long (&keys)[kPostKeys], int (&cpos)[kPostKeys]) {
#pragma unroll
    for (int k = 0; k < kPostKeys; ++k) {
        const int i = threadIdx.x + k * kPostThreads;
        keys[k] = i < K ? sorted_at<E>(s, i) : 0ull;
        cpos[k] = (i < K && srcpos) ? (int)srcpos[i] : -1;   // -1: position unknown (bitonic paths), gather from the level
    }
}

// everything of image b that is not streaming, by the whole CTA
template <bool SOFT>
static __device__ void run_tail(const PostArgs &P, int b, unsigned long long *s) {
    stamp(P.timeline, b, 2);
    __threadfence();   // acquire side of the done-counter hand-off: the candidates of every other SM are visible
    const unsigned n = __ldcg(P.T.cnt + b);
    if (n < (unsigned)P.T.K || n > (unsigned)kCap || P.debug_skip_tails) {   // uniform: the exact path (odk_topk.cu) takes this image
        if (threadIdx.x == 0) P.T.flag[b] = 1u;
        return;
    }
    unsigned long long keys[kPostKeys];
    int cpos[kPostKeys];
    const Refined R = refine_candidates(P.T, b, (int)n, s);
    if (threadIdx.x == 0 && P.timeline) {
        P.timeline[(size_t)b * kStampSlots + 13] = n;
        P.timeline[(size_t)b * kStampSlots + 14] = (unsigned long long)R.m | (R.ranked ? 1ull << 32 : 0ull);
    }
    if (R.ranked) {
        if (P.emit_topk) emit_topk<0, true>(P.T, b, s + kSortSlots);
        load_sorted<0>(s + kSortSlots, P.T.cand_box ? R.srcpos : nullptr, P.T.K, keys, cpos);
    } else if (R.m <= kSortSlots) {
        for (int i = R.m + threadIdx.x; i < kSortSlots; i += blockDim.x) s[i] = 0ull;
        __syncthreads();
        block_sort_desc<8>(s);
        if (P.emit_topk) emit_topk<8, true>(P.T, b, s);
        load_sorted<8>(s, nullptr, P.T.K, keys, cpos);
    } else {   // more than 8192 keys tie inside one sub-bin: sort everything
        const unsigned long long *cand = P.T.cand + (size_t)b * kCap;
        __syncthreads();
        for (int i = threadIdx.x; i < 16 * kSelThreads; i += blockDim.x) s[i] = i < (int)n ? __ldcg(cand + i) : 0ull;
        __syncthreads();
        block_sort_desc<16>(s);
        if (P.emit_topk) emit_topk<16, true>(P.T, b, s);
        load_sorted<16>(s, nullptr, P.T.K, keys, cpos);
    }
    __syncthreads();   // the keys are in registers: the buffer becomes the detection arrays
    stamp(P.timeline, b, 4);
    detect_sorted<SOFT>(P, b, keys, cpos, reinterpret_cast<unsigned char *>(s));
    stamp(P.timeline, b, 3);
}

// Staged pipeline: the tails of all images as one launch after the collect kernel of odk_topk.cu
// (two instantiations: the suppressors are large, and each kernel's register allocation is better without the other)
template <bool SOFT>
__global__ void __launch_bounds__(kPostThreads, 1) post_tail_kernel(const __grid_constant__ PostArgs P) {
    extern __shared__ __align__(16) unsigned long long s_dyn[];
    run_tail<SOFT>(P, (int)blockIdx.x, s_dyn);
}

// Streaming warps only ever touch shared memory between tasks; everything that needs a global round trip is
// done by ONE warp per CTA, the bookkeeper, and aggregated, because 4.6 k warps hammering single addresses
// (a task counter, an image's candidate and completion counters) is what limits a naive scheme (measured: one
// atomicAdd per task on one queue word caps the kernel at ~5.7 ns per 16 KB task = 2.9 TB/s):
//   tasks          : the bookkeeper takes CHUNKS of consecutive tasks from the global queue (one atomic per
//                    256 KB) and keeps two chunks ahead; streaming warps draw tickets from a shared counter;
//   streaming warp : stages its hits in one of its private buffers, posts a record {image, hits, warp, buffer}
//                    into the CTA's ring and goes on with the next task in the next buffer;
//   bookkeeper     : lane i takes record i; lanes of the same image combine: one atomicAdd on the image's
//                    candidate counter, the warp copies the staged hits to the candidate list, frees the
//                    buffers, fences once, one atomicAdd on the image's completed-task counter -- the lane that
//                    completes an image queues the image's tail for this CTA.
constexpr int kBookWarp = kPostWarps - 1;
constexpr int kStreamWarps = kPostWarps - 1;
constexpr int kStageBufs = 4;         // staging buffers per streaming warp: a buffer is free again ~2 bookkeeper rounds after its task
constexpr int kRecCap = 256;          // ring of task records (at most kStageBufs per streaming warp are outstanding)
constexpr int kChunkRing = 16;        // published chunk bases (the bookkeeper is at most `ahead` + 1 chunks ahead)

template <bool SOFT>
__global__ void __launch_bounds__(kPostThreads, 1) post_fused_kernel(const __grid_constant__ PostArgs P) {
    extern __shared__ __align__(16) unsigned long long s_dyn[];
    // the staging buffers alias the head of the dynamic buffer: the CTA only meets for a tail once every record is
    // flushed, i.e. while all of them are empty
    unsigned long long (*s_wst)[kStageBufs][kWarpStage] = reinterpret_cast<unsigned long long (*)[kStageBufs][kWarpStage]>(s_dyn);
    __shared__ unsigned s_wn[kStreamWarps];
    __shared__ unsigned s_busy[kStreamWarps][kStageBufs];
    __shared__ unsigned s_rec[kRecCap], s_rec_seq[kRecCap];
    __shared__ unsigned s_rec_head, s_arrived, s_drained;
    __shared__ unsigned s_ticket, s_chunks, s_chunk_base[kChunkRing];
    __shared__ int s_tailq[64];
    __shared__ unsigned s_tail_n, s_tail_done;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_tail_n = 0u; s_tail_done = 0u; s_rec_head = 0u; s_arrived = 0u; s_drained = 0u; s_ticket = 0u; s_chunks = 0u; }
    if (tid < kStreamWarps) {
        s_wn[tid] = 0u;
        for (int i = 0; i < kStageBufs; ++i) s_busy[tid][i] = 0u;
    }
    for (int i = tid; i < kRecCap; i += kPostThreads) s_rec_seq[i] = 0u;
    __syncthreads();

    const unsigned total = P.total_tasks;
    const int ntask = P.G.ntask_img;
    if (blockIdx.x == 0 && tid == 0) P.timeline[(size_t)P.T.B * kStampSlots] = global_ns();
    auto tail_pending = [&]() {
        unsigned p = 0u;
        if (lane == 0) p = *(volatile unsigned *)&s_tail_n != *(volatile unsigned *)&s_tail_done;
        return __shfl_sync(0xffffffffu, p, 0) != 0u;
    };
    // meeting point of the CTA: every image it has taken is processed by all 32 warps; returns true when every
    // streaming warp is out of tasks and the bookkeeper has seen every image of the batch taken by some CTA
    auto meet = [&](bool drained) {
        __syncthreads();   // all 32 warps are here: nobody is streaming, every record is flushed, the queue is stable
        const unsigned n1 = s_tail_n, d0 = s_tail_done;
        for (unsigned i = d0; i != n1; ++i) run_tail<SOFT>(P, s_tailq[i & 63u], s_dyn);
        __syncthreads();
        if (tid == 0) { s_tail_done = n1; s_arrived = 0u; }
        return __syncthreads_and(drained ? 1 : 0) != 0;
    };

    if (warp == kBookWarp) {
        // ---- bookkeeper ----
        unsigned tail = 0u, chunks = 0u;
        bool exhausted = false;
        for (;;) {
            // keep the CTA kChunksAhead chunks of tasks ahead of its ticket counter (past the end of the queue the
            // chunks are still published: their tasks are >= total, which is how the streaming warps learn to stop)
            while (chunks < *(volatile unsigned *)&s_ticket / (unsigned)P.chunk + (unsigned)P.ahead) {
                unsigned base = total;
                if (!exhausted) {
                    if (lane == 0) base = atomicAdd(P.queue, (unsigned)P.chunk);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    exhausted = base >= total;
                }
                if (lane == 0) {
                    s_chunk_base[chunks % kChunkRing] = base;
                    __threadfence_block();
                    *(volatile unsigned *)&s_chunks = chunks + 1u;
                }
                ++chunks;
            }
            const unsigned head = *(volatile unsigned *)&s_rec_head;
            const unsigned take = min(head - tail, 32u);
            // Any CTA may run the tail of any completed image (if the CTA that reports an image's last task always
            // ran it, the CTA that is last once would be last for every later image too and run all their tails in
            // series): take one from the global queue unless this CTA already has one.
            bool all_taken = false;
            if (*(volatile unsigned *)&s_tail_n == *(volatile unsigned *)&s_tail_done) {
                unsigned img1 = 0u;
                if (lane == 0) {
                    const unsigned claimed = __ldcg(P.tq_claimed), pushed = __ldcg(P.tq_pushed);
                    all_taken = claimed >= (unsigned)P.T.B;
                    if (claimed < pushed && atomicCAS(P.tq_claimed, claimed, claimed + 1u) == claimed) {
                        while ((img1 = __ldcg(P.tq_img + claimed)) == 0u) __nanosleep(50);   // published right after the count
                        __threadfence();   // acquire: the image's candidates
                        P.timeline[(size_t)(img1 - 1u) * kStampSlots + 1] = global_ns();
                        const unsigned q = atomicAdd(&s_tail_n, 1u);
                        s_tailq[q & 63u] = (int)(img1 - 1u);
                    }
                }
                all_taken = __shfl_sync(0xffffffffu, all_taken ? 1 : 0, 0) != 0;
            }
            if (take == 0u) {
                // nothing to flush: join the streaming warps once all of them wait at the meeting point -- for a tail,
                // or at the very end (all drained, every image taken by some CTA)
                if (*(volatile unsigned *)&s_arrived == (unsigned)kStreamWarps) {
                    if (*(volatile unsigned *)&s_rec_head != tail) continue;   // a record slipped in before the last arrival
                    const bool pending = *(volatile unsigned *)&s_tail_n != *(volatile unsigned *)&s_tail_done;
                    const bool finished = *(volatile unsigned *)&s_drained == (unsigned)kStreamWarps && all_taken;
                    if (pending || finished) {
                        if (meet(finished)) {
                            if (lane == 0) atomicMax(P.timeline + (size_t)P.T.B * kStampSlots + 1, global_ns());
                            break;
                        }
                    } else {
                        __nanosleep(200);   // drained, but some image is still open elsewhere: keep polling the queue
                    }
                } else {
                    __nanosleep(100);
                }
                continue;
            }
            const unsigned active = take >= 32u ? 0xFFFFFFFFu : ((1u << take) - 1u);
            unsigned rec = 0xFFFF0000u;   // idle lanes: an image number no record can carry (B <= 65535)
            if ((unsigned)lane < take) {
                const unsigned slot = (tail + (unsigned)lane) % kRecCap;
                while (*(volatile unsigned *)&s_rec_seq[slot] != tail + (unsigned)lane + 1u) { }   // written right after the claim
                rec = s_rec[slot];
            }
            const unsigned rb = rec >> 16, rn = (rec >> 8) & 0xFFu, rw = (rec >> 3) & 0x1Fu, rbuf = rec & 7u;
            // lanes of the same image combine: hits before mine, hits and tasks of the group
            unsigned before = 0u, hits = 0u, tasks = 0u, leader = 32u;
            for (unsigned j = 0; j < take; ++j) {
                const unsigned jb = __shfl_sync(0xffffffffu, rb, j), jn = __shfl_sync(0xffffffffu, rn, j);
                if (jb == rb) {
                    if (leader == 32u) leader = j;
                    if (j < (unsigned)lane) before += jn;
                    hits += jn;
                    ++tasks;
                }
            }
            const bool lead = (unsigned)lane < take && leader == (unsigned)lane;
            unsigned pos = 0u;
            if (lead && hits) pos = atomicAdd(P.T.cnt + rb, hits);
            pos = __shfl_sync(0xffffffffu, pos, leader & 31u) + before;
            for (unsigned j = 0; j < take; ++j) {
                const unsigned jn = __shfl_sync(0xffffffffu, rn, j);
                if (jn == 0u) continue;
                const unsigned jb = __shfl_sync(0xffffffffu, rb, j), jpos = __shfl_sync(0xffffffffu, pos, j);
                const unsigned jw = __shfl_sync(0xffffffffu, rw, j), jbuf = __shfl_sync(0xffffffffu, rbuf, j);
                if ((unsigned)lane < jn && jpos + (unsigned)lane < (unsigned)kCap)
                    P.T.cand[(size_t)jb * kCap + jpos + lane] = s_wst[jw][jbuf][lane];
            }
            __syncwarp();
            if ((unsigned)lane < take) *(volatile unsigned *)&s_busy[rw][rbuf] = 0u;   // the staging buffer may be reused
            // release: the candidate stores of the whole batch are visible before any of its tasks is reported done
            __threadfence();
            __syncwarp();
            if (lead) {
                const unsigned d = atomicAdd(P.done + rb, tasks);
                if (d + tasks == (unsigned)ntask) {   // the image is complete: publish it for whichever CTA polls next
                    const unsigned q = atomicAdd(P.tq_pushed, 1u);
                    P.timeline[(size_t)rb * kStampSlots] = global_ns();
                    __threadfence();
                    atomicExch(P.tq_img + q, rb + 1u);
                }
            }
            __syncwarp();
            (void)active;
            tail += take;
        }
        return;
    }

    // ---- streaming warps ----
    // No task is taken while a hand-off is pending: a task a warp holds but has not streamed keeps its image open,
    // the image would wait for this CTA's tail, and the CTA would then finish -- and inherit the tail of -- every
    // following image too.
    auto grab = [&]() {
        unsigned t = 0u;
        if (lane == 0) {
            const unsigned k = atomicAdd(&s_ticket, 1u);
            const unsigned c = k / (unsigned)P.chunk;
            while (*(volatile unsigned *)&s_chunks <= c) __nanosleep(20);   // the bookkeeper is two chunks ahead
            t = *(volatile unsigned *)&s_chunk_base[c % kChunkRing];
            t = t >= total ? total : t + k % (unsigned)P.chunk;
        }
        return __shfl_sync(0xffffffffu, t, 0);
    };
    static_assert(kWarpStage <= 32, "one staged hit per bookkeeper lane");
    bool drained = false;
    unsigned buf = 0u;
    for (;;) {
        if (!drained && !tail_pending()) {
            const unsigned cur = grab();
            if (cur < total) {
                const unsigned b = fd_div(cur, P.G.div_ntask);
                const STask k = stream_task(P.G, (int)b, (int)(cur - b * (unsigned)ntask));
                const float thr_f = thr_float(__ldcg(P.T.thr + b));
                // the buffer this task stages into must have been flushed (it was posted kStageBufs tasks ago)
                if (lane == 0) while (*(volatile unsigned *)&s_busy[warp][buf]) __nanosleep(50);
                __syncwarp();
                bool overflowed = false;
                auto hit = [&](float x, int e) {
                    if (x >= thr_f) {
                        const unsigned long long key = ((unsigned long long)vkey_of(x) << 32) |
                                                       (unsigned long long)(~stream_flat(P.G, k.l, (unsigned)e));
                        const unsigned slot = atomicAdd(&s_wn[warp], 1u);
                        if (slot < (unsigned)kWarpStage) {
                            s_wst[warp][buf][slot] = key;
                        } else {   // staging full (threshold far too low): straight to the global list
                            const unsigned pos = atomicAdd(P.T.cnt + b, 1u);
                            if (pos < (unsigned)kCap) P.T.cand[(size_t)b * kCap + pos] = key;
                            overflowed = true;
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
                if (__any_sync(0xffffffffu, overflowed)) __threadfence();   // direct stores before the record is posted
                __syncwarp();
                if (lane == 0) {   // post the task: the bookkeeper does everything that needs a global round trip
                    const unsigned nst = min(*(volatile unsigned *)&s_wn[warp], (unsigned)kWarpStage);
                    s_wn[warp] = 0u;
                    *(volatile unsigned *)&s_busy[warp][buf] = 1u;
                    const unsigned slot = atomicAdd(&s_rec_head, 1u);
                    s_rec[slot % kRecCap] = (b << 16) | (nst << 8) | ((unsigned)warp << 3) | buf;
                    __threadfence_block();
                    *(volatile unsigned *)&s_rec_seq[slot % kRecCap] = slot + 1u;
                }
                buf = (buf + 1u) % kStageBufs;
                __syncwarp();
            } else {
                drained = true;
                if (lane == 0) atomicAdd(&s_drained, 1u);
            }
        }
        if (drained || tail_pending()) {
            if (lane == 0) { __threadfence_block(); atomicAdd(&s_arrived, 1u); }   // after this warp's last record
            if (meet(drained)) break;
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------
int make_stream_geo(StreamGeo *G, const Geo &g, const void *const *cls_levels, int C, int layout) {
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
        G->nhwc[l] = (layout >> l) & 1;
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

int launch_sample(const SampleLaunch &s0, cudaStream_t st) {
    // The exposed cost of the sample is its scattered 512-byte reads (~1 TB/s): large images are sampled at 1/128
    // (the rank statistics still see ~100 k values and ~40 exceedances; the 5-sigma margin then keeps ~2.2 K
    // candidates instead of ~1.7 K), small ones at 1/64.
    SampleLaunch s = s0;
    s.shift = s.N >= (1ll << 23) ? kSampleShift + 1 : kSampleShift;
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
    size_t slots, thr, thr_hi, cnt, flag, done, tq_img, queue, timeline, cand, cand_box, tk_val, tk_box, tk_idx, tk_cls, total;
    int slot_stride, tps;
};

static PostWs post_ws_layout(const StreamGeo &G, int B, int K) {
    PostWs w;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.slot_stride = (int)sample_slot_stride(G, &w.tps);
    size_t off = 0;
    w.slots = off; off = al(off + (size_t)B * w.slot_stride * 4);
    w.thr = off; off = al(off + (size_t)B * 4);
    w.thr_hi = off; off = al(off + (size_t)B * 4);
    w.cnt = off; off = al(off + (size_t)B * 4);
    w.flag = off; off = al(off + (size_t)B * 4);
    w.done = off; off = al(off + (size_t)B * 4);
    w.tq_img = off; off = al(off + (size_t)B * 4);
    w.queue = off; off = al(off + 16);   // queue, tq_pushed, tq_claimed
    w.timeline = off; off = al(off + ((size_t)B * kStampSlots + 2) * 8);
    w.cand = off; off = al(off + (size_t)B * kCap * 8);
    w.cand_box = off; off = al(off + (size_t)B * kCap * 16);
    w.tk_val = off; off = al(off + (size_t)B * K * 4);
    w.tk_box = off; off = al(off + (size_t)B * K * 16);
    w.tk_idx = off; off = al(off + (size_t)B * K * 8);
    w.tk_cls = off; off = al(off + (size_t)B * K * 8);
    w.total = off;
    return w;
}

// defined in odk_topk.cu / odk_detect.cu: the exact path for the images the fused kernel flagged
int launch_topk_exact_flagged(const TopkArgs &a, cudaStream_t st);
int launch_topk_collect(const TopkArgs &a, int ntasks, cudaStream_t st);
int launch_detect_flagged(const float *cls_topk, const float *box_topk, const int64_t *indices, const int64_t *classes, int B,
                          int N, const float *anchors, int64_t A, const float *img_scale, const float *img_size,
                          const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, int64_t *det_anchor,
                          const unsigned *only_flag, cudaStream_t st);
int launch_ood_flagged(const Geo &g, const void *const *cls_levels, int layout, int B, int C, const int64_t *anchor_idx, int D, float T,
                       float *energy, float *max_logit, const unsigned *only_flag, cudaStream_t st);

}  // namespace odk

extern "C" {

// flagged images only (none for real score distributions): exact select, then their detections / OOD scores
static int flagged_path(const odk::TopkArgs &a, int B, int K, const float *anchors, const float *img_scale, const float *img_size,
                        const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, int64_t *det_anchor,
                        const void *const *cls_levels, int layout, int C, float temperature, float *energy, float *max_logit,
                        cudaStream_t st) {
    using namespace odk;
    int rc = launch_topk_exact_flagged(a, st);
    if (rc) return rc;
    rc = launch_detect_flagged(a.out_val, a.out_box, (const int64_t *)a.out_idx, (const int64_t *)a.out_cls, B, K, anchors, a.g.A,
                               img_scale, img_size, params, dets, count, src, det_anchor, a.flag, st);
    if (rc) return rc;
    if (energy) rc = launch_ood_flagged(a.g, cls_levels, layout, B, C, det_anchor, params->max_det, temperature, energy, max_logit, a.flag, st);
    return rc;
}

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

size_t odk_postprocess_timeline_offset(int B, int C, const int32_t *level_hw, int num_levels, int na, int K) {
    odk::PostWs w;
    return post_ws_for(B, C, level_hw, num_levels, na, K, &w) ? 0 : w.timeline;
}

int odk_postprocess(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
                    int num_levels, int na, int K, int layout, const float *anchors, const float *img_scale, const float *img_size,
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
    rc = make_stream_geo(&P.G, a.g, cls_levels, C, layout);
    if (rc) return rc;
    const PostWs w = post_ws_layout(P.G, B, K);
    if (!workspace || workspace_bytes < w.total)
        return set_error(ODK_EWORKSPACE, "odk_postprocess: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    a.B = B; a.C = C; a.K = K;
    int toff = 0;
    rc = fill_topk_levels(&a, cls_levels, box_levels, na, layout, &toff, "odk_postprocess");   // (the collect / exact kernels' task model)
    if (rc) return rc;
    char *ws = (char *)workspace;
    a.slots = (unsigned *)(ws + w.slots); a.thr = (unsigned *)(ws + w.thr); a.cnt = (unsigned *)(ws + w.cnt);
    a.flag = (unsigned *)(ws + w.flag); a.cand = (unsigned long long *)(ws + w.cand);
    a.thr_hi = (unsigned *)(ws + w.thr_hi); a.cand_box = (float4 *)(ws + w.cand_box);
    a.out_val = want_topk ? cls_topk : (float *)(ws + w.tk_val);
    a.out_box = want_topk ? box_topk : (float *)(ws + w.tk_box);
    a.out_idx = want_topk ? (long long *)indices : (long long *)(ws + w.tk_idx);
    a.out_cls = want_topk ? (long long *)classes : (long long *)(ws + w.tk_cls);
    a.fused = 1;
    P.emit_topk = want_topk ? 1 : 0;
    P.queue = (unsigned *)(ws + w.queue); P.done = (unsigned *)(ws + w.done);
    P.timeline = (unsigned long long *)(ws + w.timeline);
    a.stamp = P.timeline;
    {
        const char *dbg = getenv("ODK_POST_SKIP_TAILS");
        P.debug_skip_tails = (dbg && dbg[0] == '1') ? 1 : 0;
        // a chunk is what one queue access takes; the bookkeeper keeps `ahead` chunks published beyond the one in use:
        // enough to hide the atomic's latency, little enough that images still complete in order
        P.chunk = 8; P.ahead = 2;
        const char *ec = getenv("ODK_POST_CHUNK"), *ea = getenv("ODK_POST_AHEAD");
        if (ec && atoi(ec) >= 1 && atoi(ec) <= 64 && (atoi(ec) & (atoi(ec) - 1)) == 0) P.chunk = atoi(ec);
        if (ea && atoi(ea) >= 1 && atoi(ea) <= 8) P.ahead = atoi(ea);
    }
    P.tq_img = (unsigned *)(ws + w.tq_img); P.tq_pushed = P.queue + 1; P.tq_claimed = P.queue + 2;
    P.total_tasks = (unsigned)B * (unsigned)P.G.ntask_img;
    if ((long long)B * P.G.ntask_img > 0x7fffffffll) return set_error(ODK_EUNSUPPORTED, "odk_postprocess: too many tasks");
    P.anchors = (const float4 *)anchors; P.scale = img_scale; P.size = img_size;
    P.p = *params; P.nms_thr_f = float_at_or_below(params->nms_iou);
    P.cap = det_cap(K);
    P.dets = dets; P.count = count; P.src = src; P.det_anchor = (long long *)det_anchor;
    P.energy = energy; P.max_logit = max_logit; P.ood_T = temperature;

    cudaStream_t st = (cudaStream_t)stream;
    const bool persistent = params->pipeline == ODK_PIPELINE_PERSISTENT;
    if (persistent) a.cand_box = nullptr;   // (its bookkeeper does not gather the rows: the tail reads them from the levels)
    SampleLaunch s;
    memset(&s, 0, sizeof(s));
    s.G = P.G; s.B = B; s.K = K; s.N = a.N; s.slots = a.slots; s.slot_stride = w.slot_stride; s.tps = w.tps; s.thr = a.thr;
    s.thr_hi = (unsigned *)(ws + w.thr_hi);
    s.zero0 = a.cnt; s.zero1 = a.flag; s.zero2 = P.done; s.zero3 = P.tq_img; s.zero_scalar = P.queue; s.zero_scalars = 3;
    s.zero_u64 = P.timeline + (size_t)B * kStampSlots + 1;
    rc = launch_sample(s, st);
    if (rc) return rc;

    size_t smem = kSelSmemBytes;
    if (post_det_bytes(P.cap) > smem) smem = post_det_bytes(P.cap);
    if (persistent) {
        // one persistent kernel: image-major stream + tails of the earlier images (see the top of this file)
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms < 1) sms = 148;
        auto kernel = params->soft_nms ? post_fused_kernel<true> : post_fused_kernel<false>;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "odk_postprocess: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
        unsigned grid = (unsigned)sms;
        const unsigned useful = (P.total_tasks + kStreamWarps - 1) / kStreamWarps;
        if (grid > useful) grid = useful;
        if (grid < 1) grid = 1;
        kernel<<<grid, kPostThreads, smem, st>>>(P);
        rc = check_launch("odk_postprocess/post_fused_kernel");
    } else {
        // staged: the one-wave collect kernel of odk_topk.cu streams the whole batch, then one CTA per image runs
        // the image's tail (with B <= #SMs both pipelines expose exactly one tail after the last logit is read)
        rc = launch_topk_collect(a, toff, st);
        if (rc) return rc;
        auto kernel = params->soft_nms ? post_tail_kernel<true> : post_tail_kernel<false>;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, "odk_postprocess: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
        // Which images the sampled path cannot take is known from the candidate counts once the collect kernel is done,
        // so the flagged-image kernels (exact select -> detections -> OOD scores; they exit at once when nothing is flagged,
        // which is every real score distribution) run on a forked stream BESIDE the tail kernel instead of behind it:
        // three dependent launches (~9 us) leave the critical path.
        SideLane lane;
        rc = fork_side(st, &lane);
        if (rc) return rc;
        kernel<<<B, kPostThreads, smem, st>>>(P);
        rc = check_launch("odk_postprocess/post_tail_kernel");
        if (rc) return rc;
        a.fused = 2;
        a.force_exact = P.debug_skip_tails;
        rc = flagged_path(a, B, K, anchors, img_scale, img_size, params, dets, count, src, det_anchor, cls_levels, layout, C, temperature,
                          energy, max_logit, lane.side);
        const int rj = join_side(st, lane);
        return rc ? rc : rj;
    }
    if (rc) return rc;
    return flagged_path(a, B, K, anchors, img_scale, img_size, params, dets, count, src, det_anchor, cls_levels, layout, C, temperature,
                        energy, max_logit, st);
}

}  // extern "C"
