// Device-side building blocks of the top-k path (shared by odk_topk.cu and odk_post.cu).
// See odk_topk.cu for the design notes.
#pragma once
#include <cooperative_groups.h>
#include <string.h>

#include "odk_common.cuh"

namespace odk {

constexpr int kTopkThreads = 256;
constexpr int kSelThreads = 1024;
constexpr int kCap = 16384;            // candidate capacity per image (128 KB of 64-bit keys)
constexpr int kHistBins = 4096;        // top 12 bits of the value key
constexpr int kSlots = 8192;           // per-image sample slots (lane maxima)
constexpr int kSampleShift = 6;        // sample 1 / 64 of the 512-byte units (1 / 128 of large images, odk_stream.cuh)
constexpr int kSegVec = 1024;          // vec4 units per task segment
constexpr int kClusterSize = 8;
constexpr int kRadixBits = 11;

struct TopkArgs {
    Geo g;
    const float *cls[ODK_MAX_LEVELS];
    const float *box[ODK_MAX_LEVELS];
    int nplanes[ODK_MAX_LEVELS];  // contiguous runs per image: na * C channel planes (NCHW) or 1 (channels_last: the whole block)
    int plane_len[ODK_MAX_LEVELS];   // elements per run: hw (NCHW) or hw * na * C (channels_last)
    unsigned char cls_nhwc[ODK_MAX_LEVELS], box_nhwc[ODK_MAX_LEVELS];   // the level is stored channels_last ([B, H, W, channels])
    int vec[ODK_MAX_LEVELS];      // 4 or 1
    int nvec[ODK_MAX_LEVELS];     // vector units per plane
    int nseg[ODK_MAX_LEVELS];     // task segments per plane
    FastDiv div_nseg[ODK_MAX_LEVELS], div_C;   // the task decode runs once per 16 KB segment, and per SAMPLED unit
    int task_off[ODK_MAX_LEVELS + 1];
    int B, C, K, planes;          // planes = na * C channel planes per level
    long long N;                  // elements per image = A * C
    unsigned *slots;              // [B][kSlots] value keys: maxima of the sampled units of one lane
    unsigned *thr;                // [B] threshold key of the collect pass
    const unsigned *thr_hi;       // [B] sampled estimate of the key ~K/16 elements exceed (bin edge hint), or null
    float4 *cand_box;             // [B][kCap] box regression of every candidate (same order as cand), or null
    int nslots;                   // slots actually used per image
    unsigned *cnt;                // [B]
    unsigned *flag;               // [B]
    unsigned long long *cand;     // [B][kCap]
    float *out_val;               // [B][K]
    float *out_box;               // [B][K][4]
    long long *out_idx;           // [B][K]
    long long *out_cls;           // [B][K]
    int fused;                    // called behind odk_postprocess: unflagged images are already complete (2: beside the tail kernel)
    int force_exact;              // diagnostics (ODK_POST_SKIP_TAILS=1): every image goes the exact way
    unsigned long long *stamp;    // diagnostics: [B][kStampSlots] %globaltimer marks (odk_postprocess), or null
};

constexpr int kStampSlots = 16;
__device__ __forceinline__ void stamp(unsigned long long *base, int b, int slot) {
    if (base && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        base[(size_t)b * kStampSlots + slot] = t;
    }
}

__device__ __forceinline__ unsigned vkey_of(float x) {
    const unsigned u = __float_as_uint(x + 0.0f);   // -0.0 -> +0.0: they compare equal in torch.topk
    return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float val_of(unsigned k) {
    const unsigned u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(u);
}

struct Task {
    const float *base;   // first element of the plane
    int u0, u1;          // vector-unit range of this segment
    int vec;
    unsigned fbase;      // flat index of element 0 of this run: (off_l + a) * C + c (NCHW plane) or off_l * C (channels_last)
    unsigned fstride;    // flat-index step per element: na * C (NCHW: next position) or 1 (channels_last: memory order IS flat order)
};

__device__ __forceinline__ int task_level(const TopkArgs &A, int t) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < ODK_MAX_LEVELS; ++i)
        if (i < A.g.nlev && t >= A.task_off[i]) l = i;
    return l;
}

__device__ __forceinline__ Task decode_task(const TopkArgs &A, int b, int t) {
    const int l = task_level(A, t);
    const int local = t - A.task_off[l];
    const int ch = (int)fd_div((unsigned)local, A.div_nseg[l]);
    const int sg = local - ch * A.nseg[l];
    Task k;
    k.base = A.cls[l] + ((size_t)b * A.nplanes[l] + ch) * A.plane_len[l];
    k.vec = A.vec[l];
    k.u0 = sg * kSegVec;
    k.u1 = min(A.nvec[l], k.u0 + kSegVec);
    if (A.cls_nhwc[l]) {   // [B, H, W, na*C]: element e of the image's block is flat index off_l * C + e (bench.py:37)
        k.fbase = (unsigned)A.g.off[l] * (unsigned)A.C;
        k.fstride = 1u;
    } else {
        const int a = (int)fd_div((unsigned)ch, A.div_C), c = ch - a * A.C;
        k.fbase = (unsigned)(A.g.off[l] + a) * (unsigned)A.C + (unsigned)c;
        k.fstride = (unsigned)A.planes;
    }
    return k;
}

// Visit every element of a task segment with the warp: f4(v0..v3 of one lane's float4, first
// position) / f1(value, position).  Four independent 128-bit loads are in flight per lane.
template <class F4, class F1>
__device__ __forceinline__ void visit_task(const Task &k, int lane, F4 f4, F1 f1) {
    if (k.vec == 4) {
        int u = k.u0 + lane;
        for (; u + 96 < k.u1; u += 128) {
            const float4 v0 = ld_stream4(k.base + (size_t)u * 4);
            const float4 v1 = ld_stream4(k.base + (size_t)(u + 32) * 4);
            const float4 v2 = ld_stream4(k.base + (size_t)(u + 64) * 4);
            const float4 v3 = ld_stream4(k.base + (size_t)(u + 96) * 4);
            f4(v0, u * 4); f4(v1, u * 4 + 128); f4(v2, u * 4 + 256); f4(v3, u * 4 + 384);
        }
        for (; u < k.u1; u += 32) f4(ld_stream4(k.base + (size_t)u * 4), u * 4);
    } else {
        for (int u = k.u0 + lane; u < k.u1; u += 32) f1(ld_stream1(k.base + u), u);
    }
}

// Threshold per image from the slot maxima (one CTA of NT threads; `sl` may be shared memory).  With L slots of
// n samples each, a threshold exceeded by r slots is exceeded by about -ln(1 - r/L)/n of all elements; r is the
// smallest rank whose 5-sigma lower bound on that fraction still covers K of the N elements.  The collect pass
// keeps everything at or above the r-th largest slot maximum rounded down to 20 key bits: the rank is first
// located in a 4096-bin histogram of the top 12 key bits, then inside that bin on key bits 19..12 (a 12-bit
// bin alone is 2^-3 of the value wide and can hold most of a steep score distribution).
template <int NT>
static __device__ uint2 threshold_from_slots(const unsigned *sl, int nslots, long long N, int K, int sample_shift) {
    static_assert(kHistBins % NT == 0 && NT % 32 == 0 && NT >= 256, "bins per thread");
    constexpr int kPer = kHistBins / NT, kWarps = NT / 32;
    __shared__ unsigned s_hist[kHistBins];
    __shared__ unsigned s_wsum[kWarps], s_wtot[kWarps], s_sub[2][256];
    __shared__ unsigned s_coarse[2], s_above[2], s_fine[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kHistBins; i += NT) s_hist[i] = 0u;
    if (tid < 256) { s_sub[0][tid] = 0u; s_sub[1][tid] = 0u; }
    if (tid < 2) { s_coarse[tid] = 0u; s_above[tid] = 0u; s_fine[tid] = 0u; }
    __syncthreads();
    unsigned used = 0u;
    for (int i = tid; i < nslots; i += NT) {
        const unsigned k = sl[i];
        if (k) { atomicAdd(&s_hist[k >> 20], 1u); ++used; }
    }
    used = __reduce_add_sync(0xffffffffu, used);
    if (lane == 0) s_wsum[warp] = used;
    __syncthreads();   // histogram and per-warp counts complete
    unsigned Lu = 0u;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) Lu += s_wsum[w];
    const float L = (float)Lu;
    const float n_per = fmaxf((float)N / (float)(1 << sample_shift) / fmaxf(L, 1.0f), 1.0f);
    const float need = (float)K / (float)N;
    unsigned r = 0u;   // every thread runs the same search: no broadcast needed
    if (L >= 16.0f) {
        auto covers = [&](unsigned q) {
            return -logf(1.0f - (float)q / L) / n_per * (1.0f - 5.0f * rsqrtf((float)q)) >= need;
        };
        unsigned lo = 26, hi = (unsigned)(L * 0.75f);   // r > 25 keeps the 5-sigma factor positive
        if (hi > lo && covers(hi)) {
            while (lo < hi) {
                const unsigned mid = (lo + hi) >> 1;
                if (covers(mid)) hi = mid; else lo = mid + 1;
            }
            r = lo;
        }
    }
    if (r == 0u) return make_uint2(0u, 0u);   // uniform: keep everything (the select step then flags the image)
    // second rank: the slot maxima expected above the value that K/16 elements exceed (no margin: only a hint)
    unsigned rk[2];
    rk[0] = r;
    rk[1] = max(4u, (unsigned)(L * (1.0f - expf(-n_per * need * (1.0f / 16.0f)))));
    // ranks from the top bin down: thread t owns the kPer bins below hi = kHistBins - 1 - t * kPer
    const int hi = kHistBins - 1 - tid * kPer;
    unsigned mine = 0u;
#pragma unroll
    for (int i = 0; i < kPer; ++i) mine += s_hist[hi - i];
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    unsigned before = inc - mine;
    for (int w = 0; w < warp; ++w) before += s_wtot[w];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (before < rk[q] && before + mine >= rk[q]) {   // exactly one thread per rank
            unsigned run = before;
            for (int i = 0; i < kPer; ++i) {
                const unsigned c = s_hist[hi - i];
                if (run + c >= rk[q]) { s_coarse[q] = (unsigned)(hi - i) << 20; s_above[q] = run; break; }
                run += c;
            }
        }
    }
    __syncthreads();
    const unsigned c0 = s_coarse[0], c1 = s_coarse[1];
    for (int i = tid; i < nslots; i += NT) {
        const unsigned k = sl[i];
        if (k && (k >> 20) == (c0 >> 20)) atomicAdd(&s_sub[0][(k >> 12) & 255u], 1u);
        if (k && (k >> 20) == (c1 >> 20)) atomicAdd(&s_sub[1][(k >> 12) & 255u], 1u);
    }
    __syncthreads();
    // rank inside the coarse bin: warps 0 and 1 scan the 256 sub-bins of rank 0 and 1 from the top (8 per lane)
    if (warp < 2) {
        const unsigned *sub = s_sub[warp];
        unsigned v[8], tot = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = sub[255 - (lane * 8 + j)]; tot += v[j]; }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        unsigned run = s_above[warp] + incl - tot;
        const unsigned want = rk[warp];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (run < want && run + v[j] >= want) s_fine[warp] = (unsigned)(255 - (lane * 8 + j));
            run += v[j];
        }
    }
    __syncthreads();
    return make_uint2(c0 | (s_fine[0] << 12), c1 | (s_fine[1] << 12));
}

__device__ __forceinline__ float thr_float(unsigned thr_key) {
    // x >= thr_float  <=>  vkey_of(x) >= thr_key for every non-NaN x (-0.0 counts as +0.0)
    return thr_key == 0u ? -INFINITY : val_of(thr_key);
}

// ---- P2: sort candidates, emit top K + gathers ----------------------------------------------
// Block-wide bitonic sort (descending) of P = E * 1024 64-bit keys held E per thread in registers.
// Element e = t * E + r lives in register r of thread t, so compare-exchange distance j is
//   j <  E        : inside the thread (no communication),
//   E <= j < 32 E : a lane exchange inside the warp (shuffles),
//   j >= 32 E     : between warps, through shared memory in a transposed (conflict-free) layout.
// Only 15 of the 91 stages of an 8192-key sort need a block barrier this way.
// In: s[r * 1024 + t] (any assignment of keys to slots).  Out: rank q is at s[(q % E) * 1024 + q / E].
__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned long long u64min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

template <int E>
static __device__ void block_sort_desc(unsigned long long *s) {
    const int t = threadIdx.x;
    constexpr int P = E * kSelThreads;
    unsigned long long v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = s[r * kSelThreads + t];
    __syncthreads();
#pragma unroll 1
    for (int k = 2; k <= P; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j >= E; j >>= 1) {
            const int m = j / E;   // partner thread distance
            if (m >= 32) {
#pragma unroll
                for (int r = 0; r < E; ++r) s[r * kSelThreads + t] = v[r];
                __syncthreads();
                const int pt = t ^ m;
                const bool keep_max = ((t & m) == 0) == (((t * E) & k) == 0);
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const unsigned long long o = s[r * kSelThreads + pt];
                    v[r] = keep_max ? u64max(v[r], o) : u64min(v[r], o);
                }
                __syncthreads();
            } else {
                const bool keep_max = ((t & m) == 0) == (((t * E) & k) == 0);
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], m);
                    v[r] = keep_max ? u64max(v[r], o) : u64min(v[r], o);
                }
            }
        }
        // in-thread stages: j = min(k/2, E/2) ... 1
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1) {
            if (j < k) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & j) == 0) {
                        const bool desc = (((t * E + r) & k) == 0);
                        const unsigned long long x = v[r], y = v[r | j];
                        if ((x < y) == desc) { v[r] = y; v[r | j] = x; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < E; ++r) s[r * kSelThreads + t] = v[r];
    __syncthreads();
}

template <int E>
__device__ __forceinline__ unsigned long long sorted_at(const unsigned long long *s, int q) {
    if (E == 0) return s[q];   // linear (bucket-rank path)
    return s[(q % (E ? E : 1)) * kSelThreads + q / (E ? E : 1)];
}

// box regression rows of the selected anchors (bench.py:48-49): 4 scattered 4-byte loads per row
__device__ __forceinline__ float4 gather_box(const TopkArgs &A, int b, int anchor) {
    const Geo &g = A.g;
    const int l = geo_level(g, anchor);
    const int loc = anchor - g.off[l];
    const int sp = loc / g.na, a = loc - sp * g.na;
    if (A.box_nhwc[l])   // [B, H, W, na*4]: the four codes of (position, shape) are one aligned 16-byte load
        return __ldg(reinterpret_cast<const float4 *>(A.box[l] + ((size_t)b * g.hw[l] + sp) * (size_t)(g.na * 4)) + a);
    const float *bp = A.box[l] + ((size_t)(b * g.na + a) * 4) * g.hw[l] + sp;
    float4 r;
    r.x = __ldg(bp); r.y = __ldg(bp + g.hw[l]); r.z = __ldg(bp + 2 * (size_t)g.hw[l]); r.w = __ldg(bp + 3 * (size_t)g.hw[l]);
    return r;
}

// BOXES = false leaves out_box to the cluster kernel that follows (gather_selected_boxes): the scattered
// gather is LSU-bound on one SM, spread over the 8 CTAs of the image's cluster it is not.
template <int E, bool BOXES>
static __device__ void emit_topk(const TopkArgs &A, int b, const unsigned long long *s) {
    constexpr int kEmitRows = 4;   // rows per thread whose gathers are issued before any store
    for (int q0 = threadIdx.x; q0 < A.K; q0 += kEmitRows * blockDim.x) {
        float4 r[kEmitRows];
        unsigned long long key[kEmitRows];
#pragma unroll
        for (int u = 0; u < kEmitRows; ++u) {
            const int q = q0 + u * blockDim.x;
            key[u] = q < A.K ? sorted_at<E>(s, q) : ~0ull;   // ~0: flat index 0, never stored
            if (BOXES) {
                const unsigned flat = ~(unsigned)(key[u] & 0xFFFFFFFFull);
                r[u] = gather_box(A, b, (int)(flat / (unsigned)A.C));
            }
        }
#pragma unroll
        for (int u = 0; u < kEmitRows; ++u) {
            const int q = q0 + u * blockDim.x;
            if (q >= A.K) continue;
            const unsigned flat = ~(unsigned)(key[u] & 0xFFFFFFFFull);
            const int anchor = (int)(flat / (unsigned)A.C);
            const size_t o = (size_t)b * A.K + q;
            A.out_val[o] = val_of((unsigned)(key[u] >> 32));
            A.out_idx[o] = anchor;                                          // bench.py:45
            A.out_cls[o] = (int)(flat - (unsigned)anchor * (unsigned)A.C);  // bench.py:46
            if (BOXES) reinterpret_cast<float4 *>(A.out_box)[o] = r[u];    // bench.py:48-49
        }
    }
}

// rows of an image the select kernel finished: one row per thread of the 8-CTA cluster
static __device__ void gather_selected_boxes(const TopkArgs &A, int b, unsigned cluster_rank) {
    for (int q = (int)cluster_rank * kSelThreads + threadIdx.x; q < A.K; q += kClusterSize * kSelThreads) {
        const size_t o = (size_t)b * A.K + q;
        reinterpret_cast<float4 *>(A.out_box)[o] = gather_box(A, b, (int)A.out_idx[o]);
    }
}

// Cut n <= kCap candidates down to just over K and order them: a 1024-bin histogram on the value-key
// bits below the collect threshold finds the finest edge T with count(key >= T) >= K; everything below T
// cannot be in the top K.  The histogram is also a counting sort: the suffix sums give every sub-bin its
// first rank, survivors are scattered to their sub-bin's range and ranked inside it by direct comparison
// (a handful of mates per sub-bin on real score distributions), which replaces a 91-stage bitonic sort of
// 8192 keys.  Heavily tied inputs (a sub-bin with more than kBucketMax keys, or more than 8192 survivors)
// take the bitonic path instead.
constexpr int kRefineBins = 1024;
constexpr int kBucketMax = 512;    // largest sub-bin the direct ranking accepts
constexpr int kSortSlots = 8 * kSelThreads;

// survivors; ranked: s[kSortSlots + q] is the q-th largest key and srcpos[q] its position in the candidate list
struct Refined { int m; bool ranked; const unsigned short *srcpos; };
// dynamic shared memory of every kernel that selects: kCap keys + two position arrays
constexpr size_t kSelSmemBytes = (size_t)kCap * 8 + 2 * (size_t)(8 * kSelThreads) * 2;

// whole block (kSelThreads == kRefineBins: one bin per thread, top bin first): per-bin first ranks (start), the
// first bin from the top at which the running count reaches `need` (edge; 0 if it never does) with the count
// down to and including it (total), and the largest bin at or above the edge other than skip_bin (biggest).
// Two block barriers; the one-warp version of this scan was 15-25 us of a 30 us select.
static __device__ void block_suffix_scan(const unsigned *hist, unsigned *start, unsigned need, unsigned &edge, unsigned &total,
                                         unsigned &biggest, int skip_bin) {
    static_assert(kRefineBins == kSelThreads, "one bin per thread");
    __shared__ unsigned s_wt[kSelThreads / 32];
    __shared__ unsigned s_r_edge, s_r_total, s_r_big;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int bin = kRefineBins - 1 - t;
    const unsigned v = hist[bin];
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_wt[warp] = inc;
    if (t == 0) { s_r_edge = 0u; s_r_total = 0u; s_r_big = 0u; }
    __syncthreads();
    unsigned incl = inc;
    for (int w = 0; w < warp; ++w) incl += s_wt[w];
    const unsigned excl = incl - v;
    start[bin] = excl;
    if (incl >= need && excl < need) { s_r_edge = (unsigned)bin; s_r_total = incl; }
    if (bin == 0 && incl < need) s_r_total = incl;   // never reached: everything is kept
    unsigned big = (excl < need && bin != skip_bin) ? v : 0u;   // bins at or above the edge
    big = __reduce_max_sync(0xffffffffu, big);
    if (lane == 0 && big) atomicMax(&s_r_big, big);
    __syncthreads();
    edge = s_r_edge; total = s_r_total; biggest = s_r_big;
    __syncthreads();   // the result words may be reused by the next call
}

static __device__ Refined refine_candidates(const TopkArgs &A, int b, int n, unsigned long long *s) {
    __shared__ unsigned s_rh[kRefineBins], s_start[kRefineBins], s_fill[kRefineBins];
    __shared__ unsigned s_rh2[kRefineBins], s_start2[kRefineBins], s_fill2[kRefineBins];
    __shared__ unsigned s_cnt, s_maxbin, s_wmax[kSelThreads / 32], s_wmin[kSelThreads / 32];
    const int tid = threadIdx.x;
    const unsigned base = __ldcg(A.thr + b);   // every candidate key is >= base
    const unsigned hint = A.thr_hi ? __ldcg(A.thr_hi + b) : 0u;
    const unsigned long long *cand = A.cand + (size_t)b * kCap;
    unsigned short *pos1 = reinterpret_cast<unsigned short *>(s + 2 * kSortSlots);   // candidate-list position of s[p]
    unsigned short *pos2 = pos1 + kSortSlots;                                       // ... of the q-th largest key
    for (int i = tid; i < kRefineBins; i += blockDim.x) { s_rh[i] = 0; s_fill[i] = 0; s_rh2[i] = 0; s_fill2[i] = 0; }
    if (tid == 0) { s_cnt = 0u; s_maxbin = 0u; }
    unsigned long long mine[kCap / kSelThreads];
    unsigned vmax = 0u, vmin = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < kCap / kSelThreads; ++k) {
        const int i = tid + k * kSelThreads;
        mine[k] = i < n ? __ldcg(cand + i) : 0ull;
        if (i < n) {
            vmax = max(vmax, (unsigned)(mine[k] >> 32));
            vmin = min(vmin, (unsigned)(mine[k] >> 32));
        }
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    if ((tid & 31) == 0) { s_wmax[tid >> 5] = vmax; s_wmin[tid >> 5] = vmin; }
    __syncthreads();
#pragma unroll 4
    for (int w = 0; w < kSelThreads / 32; ++w) { vmax = max(vmax, s_wmax[w]); vmin = min(vmin, s_wmin[w]); }
    // Bins are LINEAR IN THE VALUE between the smallest candidate and an upper edge: the sampled estimate of the
    // value ~K/16 elements exceed (A.thr_hi) when it lies inside the candidates' range, else the largest
    // candidate.  A score distribution's tail decays roughly exponentially, so between those two edges the
    // densest of 1023 bins holds ~n*ln(16n/K)/1023 keys -- a few dozen -- wherever the values sit (bins on the
    // key BITS degenerate near zero, where a float's exponent field eats the range).  Everything at or above the
    // upper edge -- the strongest scores, possibly with far outliers -- shares the top bin, which gets a second
    // histogram, linear between that edge and the largest candidate.
    constexpr unsigned kTopBin = kRefineBins - 1;
    const unsigned key_lo = max(base, vmin);
    const bool has_top = hint > key_lo && hint < vmax;
    const unsigned top_key = has_top ? hint : 0xFFFFFFFFu;
    const float x_lo = val_of(key_lo);
    const float x_top = val_of(has_top ? hint : vmax);
    const float scale1 = x_top > x_lo ? (float)(kRefineBins - 2) / (x_top - x_lo) : 0.0f;
    auto bin_of = [&](unsigned long long key) {
        const unsigned kv = (unsigned)(key >> 32);
        if (has_top && kv >= top_key) return kTopBin;
        const float t = (val_of(kv) - x_lo) * scale1;   // monotone in the key; NaN -> 0
        return (unsigned)min(max((int)t, 0), kRefineBins - 2);
    };
#pragma unroll
    for (int k = 0; k < kCap / kSelThreads; ++k)
        if (tid + k * kSelThreads < n) atomicAdd(&s_rh[bin_of(mine[k])], 1u);
    __syncthreads();
    stamp(A.stamp, b, 9);
    unsigned edge, total0, biggest0;   // keep bins >= edge (edge 0: keep everything)
    block_suffix_scan(s_rh, s_start, (unsigned)A.K, edge, total0, biggest0, (int)kTopBin);
    if (tid == 0) s_maxbin = biggest0;
    const int m = (int)total0;       // survivors (>= K: the select step only runs with n >= K)
    if (m > kSortSlots) return {m, false, nullptr};
    const float x_hi = val_of(vmax);
    const float scale2 = x_hi > x_top ? (float)(kRefineBins - 1) / (x_hi - x_top) : 0.0f;
    auto bin2_of = [&](unsigned long long key) {
        const float t = (val_of((unsigned)(key >> 32)) - x_top) * scale2;   // monotone in the key; NaN/inf -> 0
        return (unsigned)min(max((int)t, 0), kRefineBins - 1);
    };
    const bool two_level = s_rh[kTopBin] > 32u;
    if (two_level) {
#pragma unroll
        for (int k = 0; k < kCap / kSelThreads; ++k) {
            const int i = tid + k * kSelThreads;
            if (i < n && bin_of(mine[k]) == kTopBin) atomicAdd(&s_rh2[bin2_of(mine[k])], 1u);
        }
        __syncthreads();
        unsigned e2, t2, big2;
        block_suffix_scan(s_rh2, s_start2, 0xFFFFFFFFu, e2, t2, big2, -1);   // the top bin starts at rank 0
        if (tid == 0) s_maxbin = max(s_maxbin, big2);
    } else if (tid == 0) {
        s_maxbin = max(s_maxbin, s_rh[kTopBin]);
    }
    __syncthreads();
    const bool ranked = s_maxbin <= (unsigned)kBucketMax;
    if (!ranked) {
        // compact in any order, the caller sorts
#pragma unroll
        for (int k = 0; k < kCap / kSelThreads; ++k) {
            const int i = tid + k * kSelThreads;
            if (i < n && bin_of(mine[k]) >= edge) s[atomicAdd(&s_cnt, 1u)] = mine[k];
        }
        __syncthreads();
        return {m, false, nullptr};
    }
    stamp(A.stamp, b, 10);
    // counting sort: scatter to the (sub-)bin's rank range ...
#pragma unroll
    for (int k = 0; k < kCap / kSelThreads; ++k) {
        const int i = tid + k * kSelThreads;
        if (i < n) {
            const unsigned d = bin_of(mine[k]);
            if (d >= edge) {
                unsigned pos;
                if (two_level && d == kTopBin) {
                    const unsigned d2 = bin2_of(mine[k]);
                    pos = s_start2[d2] + atomicAdd(&s_fill2[d2], 1u);
                } else {
                    pos = s_start[d] + atomicAdd(&s_fill[d], 1u);
                }
                s[pos] = mine[k];
                pos1[pos] = (unsigned short)i;
            }
        }
    }
    __syncthreads();
    stamp(A.stamp, b, 11);
    // ... then the exact rank inside the bin: keys are unique, so counting the larger mates is a permutation
    unsigned long long *sorted = s + kSortSlots;
    for (int p = tid; p < m; p += kSelThreads) {
        const unsigned long long key = s[p];
        const unsigned d = bin_of(key);
        unsigned lo, hi;
        if (two_level && d == kTopBin) {
            const unsigned d2 = bin2_of(key);
            lo = s_start2[d2]; hi = lo + s_rh2[d2];
        } else {
            lo = s_start[d]; hi = lo + s_rh[d];
        }
        unsigned r = lo;
        for (unsigned j = lo; j < hi; ++j) r += s[j] > key;
        sorted[r] = key;
        pos2[r] = pos1[p];
    }
    __syncthreads();
    stamp(A.stamp, b, 12);
    return {m, true, pos2};
}

template <bool BOXES>
static __device__ void sort_and_emit(const TopkArgs &A, int b, int n, unsigned long long *s) {
    const Refined R = refine_candidates(A, b, n, s);
    if (R.ranked) {
        emit_topk<0, BOXES>(A, b, s + kSortSlots);
    } else if (R.m <= kSortSlots) {
        for (int i = R.m + threadIdx.x; i < kSortSlots; i += blockDim.x) s[i] = 0ull;
        __syncthreads();
        block_sort_desc<8>(s);
        emit_topk<8, BOXES>(A, b, s);
    } else {   // more than 8192 keys tie inside one sub-bin: sort everything
        const unsigned long long *cand = A.cand + (size_t)b * kCap;
        __syncthreads();
        for (int i = threadIdx.x; i < 16 * kSelThreads; i += blockDim.x) s[i] = i < n ? __ldcg(cand + i) : 0ull;
        __syncthreads();
        block_sort_desc<16>(s);
        emit_topk<16, BOXES>(A, b, s);
    }
}

}  // namespace odk
