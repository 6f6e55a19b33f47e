// Device-side building blocks of detection generation (shared by odk_detect.cu and odk_post.cu).
// See odk_detect.cu for the design notes.
#pragma once
#include <string.h>

#include "odk_common.cuh"

namespace odk {

constexpr int kDetThreads = 1024;
constexpr int kDetMaxN = 8192;
constexpr int kDetWarps = kDetThreads / 32;
constexpr int kDetFirstWindow = 256;
constexpr int kSoftGroup = 8;          // warps that run the Soft-NMS rounds on the lazy window of sorted candidates   // first activated chunk of the lazy window (doubles up to kDetThreads)

struct DetSmem {
    float4 *box;          // [cap] class-offset xyxy boxes in processing order
    float *score;         // [cap]
    int *src;             // [cap] position in the caller's candidate list
    unsigned *alive;      // [cap/32]
};

__device__ __forceinline__ DetSmem carve(unsigned char *raw, int cap) {
    DetSmem s;
    s.box = reinterpret_cast<float4 *>(raw);
    s.score = reinterpret_cast<float *>(raw + (size_t)cap * 16);
    s.src = reinterpret_cast<int *>(raw + (size_t)cap * 20);
    s.alive = reinterpret_cast<unsigned *>(raw + (size_t)cap * 24);
    return s;
}
static inline size_t det_base_bytes(int cap) { return (size_t)cap * 24 + (size_t)(cap / 32) * 4 + 16; }
static inline size_t det_smem_bytes(int cap) { return det_base_bytes(cap) + 2048; }   // + hard-NMS window masks

// decode_box_outputs(output_xyxy=True) + optional clip, reference anchors.py:51-92 (fp32 op order)
__device__ __forceinline__ float4 decode_xyxy(float4 a, float4 r, bool clip, float lim_x, float lim_y) {
    const float yca = __fdiv_rn(__fadd_rn(a.x, a.z), 2.0f), xca = __fdiv_rn(__fadd_rn(a.y, a.w), 2.0f);
    const float ha = __fsub_rn(a.z, a.x), wa = __fsub_rn(a.w, a.y);
    const float w = __fmul_rn(expf(r.w), wa), h = __fmul_rn(expf(r.z), ha);
    const float yc = __fadd_rn(__fmul_rn(r.x, ha), yca), xc = __fadd_rn(__fmul_rn(r.y, wa), xca);
    const float hh = __fdiv_rn(h, 2.0f), hw = __fdiv_rn(w, 2.0f);
    float4 o = make_float4(__fsub_rn(xc, hw), __fsub_rn(yc, hh), __fadd_rn(xc, hw), __fadd_rn(yc, hh));
    if (clip) {
        o.x = fminf(fmaxf(o.x, 0.0f), lim_x); o.y = fminf(fmaxf(o.y, 0.0f), lim_y);
        o.z = fminf(fmaxf(o.z, 0.0f), lim_x); o.w = fminf(fmaxf(o.w, 0.0f), lim_y);
    }
    return o;
}
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// torchvision::nms overlap (CPU kernel order): inter / (area_i + area_j - inter)
__device__ __forceinline__ float iou_nms(float4 p, float ap, float4 q) {
    const float aq = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
    const float w = fmaxf(0.0f, __fsub_rn(fminf(p.z, q.z), fmaxf(p.x, q.x)));
    const float h = fmaxf(0.0f, __fsub_rn(fminf(p.w, q.w), fmaxf(p.y, q.y)));
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ap, aq), inter));
}
// soft_nms.py:23-38 pairwise_iou: inter > 0 ? inter / (a1 + a2 - inter) : 0
__device__ __forceinline__ float iou_soft(float4 p, float ap, float4 q) {
    const float aq = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
    const float w = fmaxf(__fsub_rn(fminf(p.z, q.z), fmaxf(p.x, q.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(p.w, q.w), fmaxf(p.y, q.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    return inter > 0.0f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(ap, aq), inter)) : 0.0f;
}

static __device__ void bitonic_sort_desc_u64(unsigned long long *s, int P) {
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const unsigned long long x = s[lo], y = s[hi];
                if ((x < y) == ((lo & k) == 0)) { s[lo] = y; s[hi] = x; }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ void init_alive(unsigned *alive, int n, int cap) {
    for (int w = threadIdx.x; w < cap / 32; w += blockDim.x) {
        const int lo = w * 32;
        alive[w] = lo + 32 <= n ? 0xFFFFFFFFu : (lo < n ? ((1u << (n - lo)) - 1u) : 0u);
    }
}

// Greedy NMS over candidates already in descending score order.  Returns the number kept
// (<= max_keep); kept[q] = candidate rank.  thr_f is the largest float <= the double threshold,
// so `iou > thr_f` equals torchvision's `(double)iou > thr`.
__device__ __forceinline__ bool nms_hit(float4 p, float ap, float4 q, float thr_f) {
    // boxes of other classes sit in other offset bands: when the x or y extents do not overlap the
    // intersection is 0 and the IoU cannot exceed a threshold >= 0, so the division is skipped
    const bool touch = (fminf(p.z, q.z) > fmaxf(p.x, q.x)) && (fminf(p.w, q.w) > fmaxf(p.y, q.y));
    return (touch || thr_f < 0.0f) && iou_nms(p, ap, q) > thr_f;
}

// Window bitmask: the candidates are taken kNmsWin at a time in score order.  For a window, (1) every candidate
// is tested against everything kept in EARLIER windows (8 threads share a candidate), (2) the suppression mask
// of the still-alive candidates among themselves is built by the whole block -- a warp per row, its lanes test 32
// later candidates at once and a ballot is the mask word (row i, bit c: kept i would suppress the later
// candidate c) -- and (3) one thread replays the greedy rule over the alive bits in order: keep the first alive
// candidate, clear everything its row suppresses, repeat -- which is exactly the one-at-a-time loop of
// torchvision::nms restricted to the window.  Only the first max_keep survivors are ever used (anchors.py:153),
// so the walk stops there: typically after one or two windows (~7 us) instead of ~max_keep/16 rounds of two
// block barriers with a dependent one-warp chain in between (31 us).
constexpr int kNmsWin = 128;
constexpr int kNmsWinWords = kNmsWin / 32;
constexpr int kNmsParts = kDetWarps / kNmsWinWords;   // threads that share a candidate in step 1
constexpr size_t kNmsMaskBytes = (size_t)kNmsWin * kNmsWinWords * 4;   // 2 KB of (dynamic) shared memory from the caller

static __device__ int hard_nms_rounds(const DetSmem &S, int n, float thr_f, int max_keep, int *kept, unsigned *mask_mem) {
    unsigned (*s_mask)[kNmsWinWords] = reinterpret_cast<unsigned (*)[kNmsWinWords]>(mask_mem);
    __shared__ unsigned s_dead[kNmsParts][kNmsWinWords];
    __shared__ unsigned s_alive[kNmsWinWords];
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int count = 0;
    for (int base = 0; base < n && count < max_keep; base += kNmsWin) {
        const int W = min(kNmsWin, n - base);
        const int nwords = (W + 31) / 32;
        // 1. against everything kept in earlier windows: warp -> (word of the window, slice of the kept list)
        {
            const int word = warp % kNmsWinWords, part = warp / kNmsWinWords;
            const int c = word * 32 + lane;
            bool dead = false;
            if (c < W && count > 0) {
                const float4 q = S.box[base + c];
                for (int j = part; j < count && !dead; j += kNmsParts) {
                    const float4 p = S.box[kept[j]];
                    dead = nms_hit(p, __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y)), q, thr_f);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, dead);
            if (lane == 0) s_dead[part][word] = bal;
        }
        __syncthreads();
        if (tid < kNmsWinWords) {
            const int lo = tid * 32;
            const unsigned valid = lo + 32 <= W ? 0xFFFFFFFFu : (lo < W ? ((1u << (W - lo)) - 1u) : 0u);
            unsigned d = 0u;
#pragma unroll
            for (int p = 0; p < kNmsParts; ++p) d |= s_dead[p][tid];
            s_alive[tid] = valid & ~d;
        }
        __syncthreads();
        // 2. suppression masks inside the window: alive rows only, words from the row's own word on (the scan
        //    below reads nothing else)
        for (int row = warp; row < W; row += kDetWarps) {
            if (!((s_alive[row >> 5] >> (row & 31)) & 1u)) continue;   // warp-uniform
            const float4 p = S.box[base + row];
            const float ap = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
            for (int word = row >> 5; word < nwords; ++word) {
                const int c = word * 32 + lane;
                bool hit = c > row && ((s_alive[word] >> lane) & 1u);
                if (hit) hit = nms_hit(p, ap, S.box[base + c], thr_f);
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_mask[row][word] = m;
            }
        }
        __syncthreads();
        // 3. the greedy rule over the alive bits in score order
        if (tid == 0) {
            unsigned a[kNmsWinWords];
#pragma unroll
            for (int w = 0; w < kNmsWinWords; ++w) a[w] = s_alive[w];
            int c = count;
#pragma unroll
            for (int w = 0; w < kNmsWinWords; ++w) {
                while (a[w] && c < max_keep) {
                    const int i = w * 32 + __ffs(a[w]) - 1;
                    a[w] &= a[w] - 1u;
                    kept[c++] = base + i;
#pragma unroll
                    for (int w2 = 0; w2 < kNmsWinWords; ++w2)
                        if (w2 >= w) a[w2] &= ~s_mask[i][w2];
                }
            }
            s_count = c;
        }
        __syncthreads();
        count = s_count;
    }
    __syncthreads();
    return count;
}

// Soft-NMS rounds (soft_nms.py:88-110).  Returns rounds run; `emit(q, rank, score)` is called by
// thread 0 for every pick, `picked[q]` receives the rank.  Lazy window (only valid when the input
// scores are non-increasing, i.e. `window < n` must not be used otherwise): the arg-max over the
// activated prefix is the global arg-max as long as it is >= the ORIGINAL score of the first
// un-activated candidate (scores only ever decay).  Otherwise the next chunk is activated by
// replaying, in order, the decays of all picks so far on each of its candidates -- the same fp32
// operations in the same order as if it had been active from the start.

// 64-bit arg-max over a warp with two hardware reductions (REDUX) instead of five 64-bit shuffle steps
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)(v & 0xFFFFFFFFull);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | (unsigned long long)ml;
}
__device__ __forceinline__ void group_sync(int group_warps) {   // named barrier 1 over the first group_warps warps
    asm volatile("bar.sync 1, %0;" ::"r"(group_warps * 32) : "memory");
}

// `group` = number of warps (from warp 0) that run the rounds; the others wait at the closing block barrier.
// A round is a dependent chain (pick -> IoU of everything alive with the pick -> decay -> next pick), so with
// the few hundred candidates of the lazy window a SMALL group is faster: its barrier is cheaper and the
// per-warp work is still one or two words.  Stand-alone calls on thousands of unsorted boxes use all 32.
template <class Emit>
static __device__ int soft_nms_rounds(const DetSmem &S, int n, bool gaussian, float sigma, float iou_thr, float score_thr,
                                      int max_rounds, int *picked, int window, int window_max, int group, Emit emit) {
    // One group barrier per round: a warp that rescales its candidates also notes its best survivor
    // (score key, ~rank); after the barrier every warp reduces the notes to the next pick itself.
    __shared__ unsigned long long s_best[2][kDetWarps];
    __shared__ int s_rounds;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // x / sigma == x * (1 / sigma) exactly when sigma is a power of two (the reference's 0.5): no second division
    int sig_e;
    const bool sig_pow2 = frexpf(sigma, &sig_e) == 0.5f;
    const float sig_inv = 1.0f / sigma;
    auto decay_of = [&](float4 p, float ap, float4 q) {
        // disjoint extents -> iou 0 -> decay exactly 1
        if (!((fminf(p.z, q.z) > fmaxf(p.x, q.x)) && (fminf(p.w, q.w) > fmaxf(p.y, q.y)))) return 1.0f;
        const float iou = iou_soft(p, ap, q);
        if (gaussian) {                                                       // soft_nms.py:96
            const float sq = -__fmul_rn(iou, iou);
            return expf(sig_pow2 ? __fmul_rn(sq, sig_inv) : __fdiv_rn(sq, sigma));
        }
        return iou > iou_thr ? __fsub_rn(1.0f, iou) : 1.0f;                   // :98-100
    };
    if (warp < group) {
    int limit = min(n, window);
    int count = 0, parity = 0;
    auto score_key = [&](int i) {   // order-preserving (scores may be <= 0 in the first round); never 0
        const unsigned u = __float_as_uint(S.score[i]);
        const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
        return ((unsigned long long)vk << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    };
    auto local_best = [&](int from_word, int nwords) {
        unsigned long long best = 0ull;
        for (int w = from_word + ((warp - from_word % group + group) % group); w < nwords; w += group) {
            const unsigned m = S.alive[w];
            if ((m >> lane) & 1u) {
                const unsigned long long key = score_key(w * 32 + lane);
                best = key > best ? key : best;
            }
        }
        return warp_max_u64(best);
    };
    {
        const unsigned long long b0 = local_best(0, (limit + 31) / 32);
        if (lane == 0) s_best[0][warp] = b0;
    }
    while (count < max_rounds) {
        group_sync(group);
        const int nwords = (limit + 31) / 32;
        const unsigned long long pick = warp_max_u64(lane < group ? s_best[parity][lane] : 0ull);
        if (limit < n) {
            // un-activated candidates still carry their original scores; the first one bounds them all
            const unsigned ub = __float_as_uint(S.score[limit]);
            const unsigned bk = ub ^ ((unsigned)((int)ub >> 31) | 0x80000000u);
            if (pick == 0ull || (unsigned)(pick >> 32) < bk) {
                window = min(2 * window, window_max);
                const int new_limit = min(n, limit + window);
                unsigned long long best = 0ull;
                const int w0 = limit / 32;
                for (int w = w0 + ((warp - w0 % group + group) % group); w < (new_limit + 31) / 32; w += group) {
                    const int i = w * 32 + lane;
                    bool ok = i < new_limit;
                    if (ok) {
                        const float4 q = S.box[i];
                        float sc = S.score[i];
                        for (int c = 0; c < count && ok; ++c) {
                            const float4 p = S.box[picked[c]];
                            const float d = decay_of(p, __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y)), q);
                            if (d != 1.0f) sc = __fmul_rn(sc, d);
                            ok = sc > score_thr;
                        }
                        S.score[i] = sc;
                        if (ok) { const unsigned long long key = score_key(i); best = key > best ? key : best; }
                    }
                    const unsigned k = __ballot_sync(0xffffffffu, ok);
                    if (lane == 0) S.alive[w] = k;
                }
                // merge with what this warp already had in the old prefix
                const unsigned long long old = local_best(0, nwords);
                best = warp_max_u64(best);
                if (lane == 0) s_best[parity ^ 1][warp] = best > old ? best : old;
                parity ^= 1;
                limit = new_limit;
                continue;
            }
        }
        if (pick == 0ull) break;
        const int top = (int)(0xFFFFFFFFu - (unsigned)(pick & 0xFFFFFFFFull));
        {
            // the pick's score is in its key (nobody may read score[top] now: its owner is about to decay it)
            const unsigned vk = (unsigned)(pick >> 32);
            const float top_score = __uint_as_float((vk & 0x80000000u) ? (vk ^ 0x80000000u) : ~vk);
            if (threadIdx.x == 0) { emit(count, top, top_score); picked[count] = top; }
        }
        ++count;
        const float4 p = S.box[top];
        const float ap = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
        unsigned long long best = 0ull;
        for (int w = warp; w < nwords; w += group) {
            unsigned m = S.alive[w];
            if (!m) continue;
            const int i = w * 32 + lane;
            bool kill = false;
            float sc = 0.f;
            if ((m >> lane) & 1u) {
                sc = S.score[i];
                const float d = decay_of(p, ap, S.box[i]);
                if (d != 1.0f) { sc = __fmul_rn(sc, d); S.score[i] = sc; }
                kill = !(sc > score_thr) || i == top;                                     // :103-104
            }
            const unsigned k = __ballot_sync(0xffffffffu, kill);
            m &= ~k;
            if (lane == 0 && k) S.alive[w] = m;
            if ((m >> lane) & 1u) {
                const unsigned u = __float_as_uint(sc);
                const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
                const unsigned long long key = ((unsigned long long)vk << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                best = key > best ? key : best;
            }
        }
        best = warp_max_u64(best);
        if (lane == 0) s_best[parity ^ 1][warp] = best;
        parity ^= 1;
    }
    if (threadIdx.x == 0) s_rounds = count;
    }
    __syncthreads();
    return s_rounds;
}

// OOD scores of one anchor by one warp: energy = -T * logsumexp(row / T), max_logit = max(row) over the C raw
// class logits of the anchor, read in place from the NCHW level (stride hw between classes).
__device__ __forceinline__ void ood_row(const Geo &g, const float *const *levels, const unsigned char *nhwc, int b, int C, long long anc,
                                        float T, int lane, float &e, float &m) {
    e = 0.f; m = 0.f;
    if (anc < 0 || anc >= g.A) return;
    const int l = geo_level(g, (int)anc);
    const int loc = (int)anc - g.off[l];
    const int sp = loc / g.na, a = loc - sp * g.na;
    // NCHW: the row's classes are hw apart; channels_last: they are contiguous
    const size_t cstride = nhwc[l] ? 1 : (size_t)g.hw[l];
    const float *row = nhwc[l] ? levels[l] + ((size_t)b * g.hw[l] + sp) * (size_t)(g.na * C) + (size_t)a * C
                               : levels[l] + ((size_t)(b * g.na + a) * C) * g.hw[l] + sp;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldg(row + (size_t)c * cstride));
    mx = warp_max(mx);
    float s = 0.f;
    const float invT = 1.0f / T;
    for (int c = lane; c < C; c += 32) s += expf((__ldg(row + (size_t)c * cstride) - mx) * invT);
    s = warp_sum(s);
    e = -T * (mx * invT + logf(s));
    m = mx;
}

static inline float float_at_or_below(double d) {
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
}

static inline int det_cap(int n) {
    int cap = (n + 1023) / 1024 * 1024;
    return cap < 1024 ? 1024 : cap;
}


}  // namespace odk
