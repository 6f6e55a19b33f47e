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
constexpr int kSampleShift = 6;        // sample 1 / 64 of the 512-byte units
constexpr int kSegVec = 1024;          // vec4 units per task segment
constexpr int kClusterSize = 8;
constexpr int kRadixBits = 11;

struct TopkArgs {
    Geo g;
    const float *cls[ODK_MAX_LEVELS];
    const float *box[ODK_MAX_LEVELS];
    int vec[ODK_MAX_LEVELS];      // 4 or 1
    int nvec[ODK_MAX_LEVELS];     // vector units per plane
    int nseg[ODK_MAX_LEVELS];     // task segments per plane
    FastDiv div_nseg[ODK_MAX_LEVELS], div_C;   // the task decode runs once per 16 KB segment, and per SAMPLED unit
    int task_off[ODK_MAX_LEVELS + 1];
    int B, C, K, planes;          // planes = na * C channel planes per level
    long long N;                  // elements per image = A * C
    unsigned *slots;              // [B][kSlots] value keys: maxima of the sampled units of one lane
    unsigned *thr;                // [B] threshold key of the collect pass
    int nslots;                   // slots actually used per image
    unsigned *cnt;                // [B]
    unsigned *flag;               // [B]
    unsigned long long *cand;     // [B][kCap]
    float *out_val;               // [B][K]
    float *out_box;               // [B][K][4]
    long long *out_idx;           // [B][K]
    long long *out_cls;           // [B][K]
    int fused;                    // called behind odk_postprocess: unflagged images are already complete
};

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
    unsigned fbase;      // flat index of position 0 of this plane: (off_l + a) * C + c
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
    const int a = (int)fd_div((unsigned)ch, A.div_C), c = ch - a * A.C;
    Task k;
    k.base = A.cls[l] + ((size_t)b * A.planes + ch) * A.g.hw[l];
    k.vec = A.vec[l];
    k.u0 = sg * kSegVec;
    k.u1 = min(A.nvec[l], k.u0 + kSegVec);
    k.fbase = (unsigned)(A.g.off[l] + a) * (unsigned)A.C + (unsigned)c;
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
static __device__ unsigned threshold_from_slots(const unsigned *sl, int nslots, long long N, int K) {
    static_assert(kHistBins % NT == 0 && NT % 32 == 0, "bins per thread");
    constexpr int kPer = kHistBins / NT, kWarps = NT / 32;
    __shared__ unsigned s_hist[kHistBins];
    __shared__ unsigned s_wsum[kWarps], s_wtot[kWarps], s_sub[256];
    __shared__ unsigned s_coarse, s_above;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kHistBins; i += NT) s_hist[i] = 0u;
    if (tid < 256) s_sub[tid] = 0u;
    if (tid == 0) { s_coarse = 0u; s_above = 0u; }
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
    const float n_per = fmaxf((float)N / (float)(1 << kSampleShift) / fmaxf(L, 1.0f), 1.0f);
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
    if (r == 0u) return 0u;   // uniform: keep everything (the select step then flags the image)
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
    if (before < r && before + mine >= r) {   // exactly one thread
        unsigned run = before;
        for (int i = 0; i < kPer; ++i) {
            const unsigned c = s_hist[hi - i];
            if (run + c >= r) { s_coarse = (unsigned)(hi - i) << 20; s_above = run; break; }
            run += c;
        }
    }
    __syncthreads();
    const unsigned coarse = s_coarse;
    for (int i = tid; i < nslots; i += NT) {
        const unsigned k = sl[i];
        if (k && (k >> 20) == (coarse >> 20)) atomicAdd(&s_sub[(k >> 12) & 255u], 1u);
    }
    __syncthreads();
    unsigned run = s_above;
    int sub = 255;
    for (; sub > 0; --sub) {   // every thread walks the same 256 counters (broadcast reads)
        run += s_sub[sub];
        if (run >= r) break;
    }
    return coarse | ((unsigned)sub << 12);
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
constexpr int kRefineShift = 14;   // sub-bin = 2^14 key units: 64 sub-bins per 12-bit threshold bin
constexpr int kBucketMax = 512;    // largest sub-bin the direct ranking accepts
constexpr int kSortSlots = 8 * kSelThreads;

struct Refined { int m; bool ranked; };   // survivors; ranked: s[kSortSlots + q] is the q-th largest key

// one warp: per-bin first ranks from the top bin down (s_start), stops at the first bin where the running
// count reaches `need` (returns that bin and the count through lane-uniform values); tracks the largest bin
__device__ __forceinline__ void suffix_scan(const unsigned *hist, unsigned *start, unsigned need, unsigned first_rank,
                                            unsigned &edge, unsigned &total, unsigned &biggest, int skip_bin) {
    const int lane = threadIdx.x & 31;
    unsigned run = first_rank, big = 0;
    edge = 0u; total = 0u;
    for (int c = kRefineBins / 32 - 1; c >= 0; --c) {
        const int bin = c * 32 + (31 - lane);   // lane 0 holds the highest bin of the chunk
        const unsigned v = hist[bin];
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        start[bin] = run + inc - v;   // keys in higher bins = first rank of this one
        const unsigned hit = __ballot_sync(0xffffffffu, run + inc - first_rank >= need);
        if (hit) {
            const int ln = __ffs(hit) - 1;
            if (lane <= ln && bin != skip_bin) big = max(big, v);
            edge = (unsigned)__shfl_sync(0xffffffffu, bin, ln);
            total = __shfl_sync(0xffffffffu, run + inc, ln) - first_rank;
            break;
        }
        if (bin != skip_bin) big = max(big, v);
        run += __shfl_sync(0xffffffffu, inc, 31);
        if (c == 0) total = run - first_rank;
    }
    biggest = __reduce_max_sync(0xffffffffu, big);
}

static __device__ Refined refine_candidates(const TopkArgs &A, int b, int n, unsigned long long *s) {
    __shared__ unsigned s_rh[kRefineBins], s_start[kRefineBins], s_fill[kRefineBins];
    __shared__ unsigned s_rh2[kRefineBins], s_start2[kRefineBins], s_fill2[kRefineBins];
    __shared__ unsigned s_edge, s_cnt, s_maxbin, s_top, s_wmax[kSelThreads / 32];
    const int tid = threadIdx.x;
    const unsigned base = __ldcg(A.thr + b);   // every candidate key is >= base
    const unsigned long long *cand = A.cand + (size_t)b * kCap;
    for (int i = tid; i < kRefineBins; i += blockDim.x) { s_rh[i] = 0; s_fill[i] = 0; s_rh2[i] = 0; s_fill2[i] = 0; }
    if (tid == 0) { s_edge = 0u; s_cnt = 0u; s_maxbin = 0u; }
    __syncthreads();
    unsigned long long mine[kCap / kSelThreads];
    auto bin_of = [&](unsigned long long key) {
        return min(((unsigned)(key >> 32) - base) >> kRefineShift, (unsigned)kRefineBins - 1u);
    };
    constexpr unsigned kTopBin = kRefineBins - 1;   // also catches everything above the sub-bin range
    unsigned vmax = 0u;
#pragma unroll
    for (int k = 0; k < kCap / kSelThreads; ++k) {
        const int i = tid + k * kSelThreads;
        mine[k] = i < n ? __ldcg(cand + i) : 0ull;
        if (i < n) {
            atomicAdd(&s_rh[bin_of(mine[k])], 1u);
            vmax = max(vmax, (unsigned)(mine[k] >> 32));
        }
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if ((tid & 31) == 0) s_wmax[tid >> 5] = vmax;
    __syncthreads();
    if (tid < 32) {
        unsigned edge, total, biggest;
        suffix_scan(s_rh, s_start, (unsigned)A.K, 0u, edge, total, biggest, (int)kTopBin);
        unsigned top = s_wmax[tid];
        top = __reduce_max_sync(0xffffffffu, top);
        if (tid == 0) { s_edge = edge; s_cnt = total; s_maxbin = biggest; s_top = top; }
    }
    __syncthreads();
    const unsigned edge = s_edge;   // keep sub-bins >= edge (edge 0: keep everything)
    const int m = (int)s_cnt;       // survivors (>= K: the select kernel only runs with n >= K)
    if (m > kSortSlots) return {m, false};
    // The top sub-bin also holds every key above the sub-bin range (the strongest scores, often across the
    // sign change where float bit patterns are sparse): it gets a second histogram that is linear in the VALUE
    // between the sub-bin's lower edge and the largest candidate.
    const float x_lo = val_of(base + (kTopBin << kRefineShift));
    const float x_hi = val_of(s_top);
    const float scale2 = x_hi > x_lo ? (float)(kRefineBins - 1) / (x_hi - x_lo) : 0.0f;
    auto bin2_of = [&](unsigned long long key) {
        const float t = (val_of((unsigned)(key >> 32)) - x_lo) * scale2;   // monotone in the key; NaN/inf -> 0
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
        if (tid < 32) {
            unsigned e2, t2, big2;
            suffix_scan(s_rh2, s_start2, 0xFFFFFFFFu, 0u, e2, t2, big2, -1);   // the top sub-bin starts at rank 0
            if (tid == 0) s_maxbin = max(s_maxbin, big2);
        }
        __syncthreads();
    } else if (tid == 0) {
        s_maxbin = max(s_maxbin, s_rh[kTopBin]);
    }
    __syncthreads();
    const bool ranked = s_maxbin <= (unsigned)kBucketMax;
    if (!ranked) {
        // compact in any order, the caller sorts
        __syncthreads();
        if (tid == 0) s_cnt = 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kCap / kSelThreads; ++k) {
            const int i = tid + k * kSelThreads;
            if (i < n && bin_of(mine[k]) >= edge) s[atomicAdd(&s_cnt, 1u)] = mine[k];
        }
        __syncthreads();
        return {m, false};
    }
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
            }
        }
    }
    __syncthreads();
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
    }
    __syncthreads();
    return {m, true};
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
