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
static inline size_t det_smem_bytes(int cap) { return det_base_bytes(cap) + 8192; }   // + hard-NMS window masks (2 KB) / Soft-NMS batch scratch (8 KB)

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

// Batched Soft-NMS rounds for candidates in NON-INCREASING score order (the top-k output) -- same results as
// soft_nms_rounds, bit for bit, with far fewer block-wide steps.  A round of the reference is a dependent chain
// (arg-max -> IoU of everything alive with the pick -> decay -> next arg-max; ~0.5 us as a block-wide step), and
// a picture needs max_rounds of them.  Here the kSoftBatch best alive candidates (the LEADERS) are taken out
// together: their pairwise decay factors are computed at once, and one warp then replays the reference's rounds
// among them in registers.  A pick is the global arg-max as long as its key beats the BOUND = the best key
// outside the leader set when the batch was formed (those scores only decay) and the original score of the
// first un-activated candidate of the lazy window; the batch stops at the first pick that does not.  Leaders
// that no earlier leader touches keep their order, so that prefix is picked in ONE step; the rest goes one pick
// at a time (arg-max by REDUX, decay = one multiply by a table entry).  The picks of the batch are then applied,
// in pick order, to every other alive candidate in parallel -- the same fp32 multiplies in the same order as
// the one-pick-at-a-time loop.  Dead candidates carry score -inf (no alive words here).
//
// Forming a batch must not cost O(prefix^2): a value t that at least kSoftBatch+1 candidates reach is taken from
// the maxima of ~64-128 small groups of the prefix (their (kSoftBatch+1)-th largest), the candidates >= t are
// compacted (a few dozen) and only those are ranked against each other.  If more than kSoftCompact reach t
// (massive ties) the whole prefix is ranked instead.
//
// A batch is a handful of short phases (profiles/micro/barrier_micro.cu: ~450-600 cycles each even when trivial --
// dependent LDS / SHFL / ALU latencies, not the barrier, which is 78 cycles for 1024 threads).  The phases can
// be confined to the first G warps (named barrier; the others sleep at the block barrier that closes the batch
// and pick the outcome up from shared memory).  Measured on D3 B=32: G = 8 for short prefixes is slightly SLOWER
// than all 32 warps (dense 29-31 us vs 25-28 us per image, the pair table and the apply step have real work),
// so kSoftSmallPrefix is 0 and the knob stays only for re-measurement.
constexpr int kSoftBatch = 32;
constexpr int kSoftCompact = 256;
constexpr size_t kSoftScratchBytes = 8192;   // 512 + 3 * 1024 + 128 + 32 * 33 * 4 = 7936
#ifndef ODK_SOFT_SMALL_PREFIX
#define ODK_SOFT_SMALL_PREFIX 0
#endif
constexpr int kSoftSmallPrefix = ODK_SOFT_SMALL_PREFIX;   // prefixes up to this long are worked by kSoftSmallGroup warps
constexpr int kSoftSmallGroup = 8;

template <class Emit>
static __device__ __noinline__ int soft_nms_batched(const DetSmem &S, int n, bool gaussian, float sigma, float iou_thr, float score_thr,
                                       int max_rounds, int *picked, int window, int window_max, Emit emit, unsigned char *scratch) {
    // kSoftScratchBytes of (dynamic) shared memory from the caller, 16-byte aligned: static arrays here would push
    // the post-process tail kernel over the 196 KB shared-memory carve-out step and leave its gather phases 28 KB of L1
    // for their misses instead of 60 KB (measured: decode 28 us instead of 14 us at D5)
    __shared__ int s_lead[kSoftBatch];                    // candidate index by rank among the alive
    __shared__ unsigned s_touch[kSoftBatch];              // row a: bit b = leader a decays the LATER leader b
    float4 *s_pbox = reinterpret_cast<float4 *>(scratch);                                   // the batch's picks in order
    float *s_val = reinterpret_cast<float *>(scratch + 512);                                // group maxima, then the compacted scores
    int *s_cidx = reinterpret_cast<int *>(scratch + 512 + 1024);
    unsigned *s_leadmask = reinterpret_cast<unsigned *>(scratch + 512 + 2048);
    float *s_parea = reinterpret_cast<float *>(scratch + 512 + 3072);
    float (*s_dec)[kSoftBatch + 1] = reinterpret_cast<float (*)[kSoftBatch + 1]>(scratch + 512 + 3072 + 128);   // [a][b]: factor on leader b when leader a is picked
    __shared__ unsigned long long s_bound;
    __shared__ float s_t;
    __shared__ int s_m;
    __shared__ int s_state[2], s_np[2];                   // per batch parity: 0 picks made, 1 window widened, 2 nothing alive
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int sig_e;
    const bool sig_pow2 = frexpf(sigma, &sig_e) == 0.5f;
    const float sig_inv = 1.0f / sigma;
    auto overlap = [](float4 p, float4 q) { return (fminf(p.z, q.z) > fmaxf(p.x, q.x)) && (fminf(p.w, q.w) > fmaxf(p.y, q.y)); };
    auto decay_of = [&](float4 p, float ap, float4 q) {
        if (!overlap(p, q)) return 1.0f;                                      // disjoint extents -> iou 0 -> decay exactly 1
        const float iou = iou_soft(p, ap, q);
        if (gaussian) {                                                       // soft_nms.py:96
            const float sq = -__fmul_rn(iou, iou);
            return expf(sig_pow2 ? __fmul_rn(sq, sig_inv) : __fdiv_rn(sq, sigma));
        }
        return iou > iou_thr ? __fsub_rn(1.0f, iou) : 1.0f;                   // :98-100
    };
    auto area_of = [](float4 p) { return __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y)); };
    auto key_of = [](float sc, int i) {   // order-preserving, never 0 for a finite score
        const unsigned u = __float_as_uint(sc);
        const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
        return ((unsigned long long)vk << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    };
    auto before = [](float sj, int j, float si, int i) { return sj > si || (sj == si && j < i); };
    int limit = min(n, window);
    int count = 0, parity = 0;
    // the whole-prefix ranking reads the scores four at a time: pad the last group
    if (tid < 4 && n + tid < ((n + 3) & ~3)) S.score[n + tid] = -INFINITY;
    __syncthreads();
    while (count < max_rounds) {
        const int G = limit <= kSoftSmallPrefix ? kSoftSmallGroup : kDetWarps;
        if (warp < G) {
            const int nthr = G * 32;
            const int nwords = (limit + 31) / 32;
            // A. the kSoftBatch + 1 best alive candidates of the prefix
            if (tid < kSoftBatch) s_lead[tid] = -1;
            for (int w = tid; w < nwords; w += nthr) s_leadmask[w] = 0u;
            if (tid == 0) { s_bound = 0ull; s_m = 0; s_t = -INFINITY; }
            // A1. maxima of groups of gs consecutive candidates (at most kSoftCompact groups)
            int gs = 4;
            while (gs < 32 && limit > gs * (kSoftCompact / 2)) gs <<= 1;
            const int ngr = (limit + gs - 1) / gs;
            for (int i0 = 0; i0 < limit; i0 += nthr) {
                const int i = i0 + tid;
                float v = i < limit ? S.score[i] : -INFINITY;
                for (int o = 1; o < gs; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (i < limit && (lane & (gs - 1)) == 0) s_val[i / gs] = v;
            }
            group_sync(G);
            // A2. t = the (kSoftBatch+1)-th largest group maximum (4 threads share a group)
            if (ngr > kSoftBatch) {
                for (int g0 = 0; g0 < ngr; g0 += nthr / 4) {
                    const int g = g0 + (tid >> 2), part = tid & 3;
                    const bool has = g < ngr;
                    const float vg = has ? s_val[g] : 0.f;
                    int r = 0;
                    if (has)
                        for (int h = part; h < ngr; h += 4) r += before(s_val[h], h, vg, g);
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    if (has && part == 0 && r == kSoftBatch) s_t = vg;
                }
            }
            group_sync(G);
            // A3. compact the alive candidates that reach t
            const float t = s_t;
            for (int i0 = 0; i0 < limit; i0 += nthr) {
                const int i = i0 + tid;
                const float v = i < limit ? S.score[i] : -INFINITY;
                const bool in = v > -INFINITY && v >= t;
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                int base = 0;
                if (lane == 0 && bal) base = atomicAdd(&s_m, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                if (in && pos < kSoftCompact) { s_val[pos] = v; s_cidx[pos] = i; }
            }
            group_sync(G);
            const int m = s_m;
            if (m <= kSoftCompact) {
                // A4. rank the compacted candidates among themselves (4 threads share one)
                for (int a0 = 0; a0 < m; a0 += nthr / 4) {
                    const int a = a0 + (tid >> 2), part = tid & 3;
                    const bool has = a < m;
                    const float sa = has ? s_val[a] : 0.f;
                    const int ia = has ? s_cidx[a] : 0;
                    int r = 0;
                    if (has)
                        for (int h = part; h < m; h += 4) r += before(s_val[h], s_cidx[h], sa, ia);
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    if (has && part == 0) {
                        if (r < kSoftBatch) { s_lead[r] = ia; atomicOr(&s_leadmask[ia >> 5], 1u << (ia & 31)); }
                        else if (r == kSoftBatch) s_bound = key_of(sa, ia);
                    }
                }
            } else {
                // A4'. massive ties: rank the whole prefix (4 threads share a candidate, scores read four at a time)
                const int ngroups = (limit + 3) >> 2;
                for (int c0 = 0; c0 < limit; c0 += nthr / 4) {
                    const int i = c0 + (tid >> 2), part = tid & 3;
                    const float si = i < limit ? S.score[i] : -INFINITY;
                    const bool live = si > -INFINITY;
                    int r = 0;
                    if (__any_sync(0xffffffffu, live)) {
                        const float4 *sc4 = reinterpret_cast<const float4 *>(S.score);
                        for (int g = part; g < ngroups; g += 4) {
                            const float4 v = sc4[g];
                            const int j = g * 4;
                            r += before(v.x, j, si, i) + before(v.y, j + 1, si, i) + before(v.z, j + 2, si, i) + before(v.w, j + 3, si, i);
                        }
                    }
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    if (live && part == 0) {
                        if (r < kSoftBatch) { s_lead[r] = i; atomicOr(&s_leadmask[i >> 5], 1u << (i & 31)); }
                        else if (r == kSoftBatch) s_bound = key_of(si, i);
                    }
                }
            }
            group_sync(G);
            // B. the bound; widen the lazy window when the best leader does not beat the un-activated candidates
            unsigned long long bound = s_bound;
            const int lead0 = s_lead[0];
            bool widen = false;
            if (limit < n) {
                const unsigned long long bk = key_of(S.score[limit], limit);
                bound = bk > bound ? bk : bound;
                widen = lead0 < 0 || !(key_of(S.score[lead0], lead0) > bk);
            }
            if (widen) {
                const int new_limit = min(n, limit + min(2 * window, window_max));
                for (int i = limit + tid; i < new_limit; i += nthr) {
                    const float4 q = S.box[i];
                    float sc = S.score[i];
                    bool ok = true;
                    for (int c = 0; c < count && ok; ++c) {
                        const float4 p = S.box[picked[c]];
                        const float d = decay_of(p, area_of(p), q);
                        if (d != 1.0f) sc = __fmul_rn(sc, d);
                        ok = sc > score_thr;
                    }
                    S.score[i] = ok ? sc : -INFINITY;
                }
                if (tid == 0) s_state[parity] = 1;
            } else if (lead0 < 0) {
                if (tid == 0) s_state[parity] = 2;
            } else {
                // C. pairwise decay factors of the leaders; row masks of who touches a later leader
                for (int ra = warp; ra < kSoftBatch; ra += G) {
                    const int a = s_lead[ra], b = s_lead[lane];
                    float d = 1.0f;
                    if (a >= 0 && b >= 0 && a != b) {
                        const float4 p = S.box[a];
                        d = decay_of(p, area_of(p), S.box[b]);
                        s_dec[ra][lane] = d;
                    }
                    const unsigned touch = __ballot_sync(0xffffffffu, d != 1.0f && lane > ra);
                    if (lane == 0) s_touch[ra] = touch;
                }
                group_sync(G);
                // D. one warp replays the rounds among the leaders
                if (warp == 0) {
                    const int idx = s_lead[lane];
                    float cur = idx >= 0 ? S.score[idx] : 0.f;
                    const float4 mybox = idx >= 0 ? S.box[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
                    bool live = idx >= 0;
                    const unsigned my_touch = s_touch[lane];
                    // leaders before the first touched one keep their (descending) order: picked in one step
                    const unsigned touched = __reduce_or_sync(0xffffffffu, my_touch);
                    const unsigned stop = __ballot_sync(0xffffffffu, !live || !(key_of(cur, idx) > bound)) | touched;
                    int np = min(stop ? __ffs(stop) - 1 : 32, max_rounds - count);
                    if (lane < np) {
                        emit(count + lane, idx, cur);
                        picked[count + lane] = idx;
                        s_pbox[lane] = mybox;
                        s_parea[lane] = area_of(mybox);
                        live = false;
                    }
                    for (int a = 0; a < np; ++a) {
                        const unsigned row = __shfl_sync(0xffffffffu, my_touch, a);
                        if (live && ((row >> lane) & 1u)) cur = __fmul_rn(cur, s_dec[a][lane]);
                    }
                    if (live && np > 0) live = cur > score_thr;               // :103-104
                    // the others, one pick at a time
                    while (count + np < max_rounds) {
                        const unsigned long long key = live ? key_of(cur, idx) : 0ull;
                        const unsigned long long pk = warp_max_u64(key);
                        if (pk == 0ull || !(pk > bound)) break;
                        const int p = __ffs(__ballot_sync(0xffffffffu, key == pk)) - 1;
                        if (lane == p) {
                            emit(count + np, idx, cur);
                            picked[count + np] = idx;
                            s_pbox[np] = mybox;
                            s_parea[np] = area_of(mybox);
                            live = false;
                        } else if (live) {
                            const float d = s_dec[p][lane];
                            if (d != 1.0f) cur = __fmul_rn(cur, d);
                            live = cur > score_thr;
                        }
                        ++np;
                    }
                    if (idx >= 0) S.score[idx] = live ? cur : -INFINITY;
                    if (lane == 0) { s_np[parity] = np; s_state[parity] = 0; }
                }
                group_sync(G);
                // E. the batch's picks, in order, on every other alive candidate of the prefix
                const int np = s_np[parity];
                if (count + np < max_rounds) {
                    for (int i = tid; i < limit; i += nthr) {
                        if ((s_leadmask[i >> 5] >> (i & 31)) & 1u) continue;
                        float sc = S.score[i];
                        if (!(sc > -INFINITY)) continue;
                        const float4 q = S.box[i];
                        bool ok = true, changed = false;
                        for (int c = 0; c < np && ok; ++c) {
                            const float4 p = s_pbox[c];
                            if (!overlap(p, q)) continue;
                            const float d = decay_of(p, s_parea[c], q);
                            if (d != 1.0f) { sc = __fmul_rn(sc, d); ok = sc > score_thr; changed = true; }
                        }
                        ok = ok && sc > score_thr;
                        if (changed || !ok) S.score[i] = ok ? sc : -INFINITY;
                    }
                }
            }
        }
        __syncthreads();
        const int state = s_state[parity];
        parity ^= 1;
        if (state == 2) break;
        if (state == 1) {
            window = min(2 * window, window_max);
            limit = min(n, limit + window);
            continue;
        }
        count += s_np[parity ^ 1];
    }
    __syncthreads();
    return count;
}

// OOD scores of one anchor by one warp: energy = -T * logsumexp(row / T), max_logit = max(row) over the C raw
// class logits of the anchor, read in place from the NCHW level (stride hw between classes).
struct OodRow { const float *row; size_t cstride; };
__device__ __forceinline__ OodRow ood_row_of(const Geo &g, const float *const *levels, const unsigned char *nhwc, int b, int C, long long anc) {
    OodRow r;
    r.row = nullptr; r.cstride = 0;
    if (anc < 0 || anc >= g.A) return r;
    const int l = geo_level(g, (int)anc);
    const int loc = (int)anc - g.off[l];
    const int sp = loc / g.na, a = loc - sp * g.na;
    // NCHW: the row's classes are hw apart; channels_last: they are contiguous
    r.cstride = nhwc[l] ? 1 : (size_t)g.hw[l];
    r.row = nhwc[l] ? levels[l] + ((size_t)b * g.hw[l] + sp) * (size_t)(g.na * C) + (size_t)a * C
                    : levels[l] + ((size_t)(b * g.na + a) * C) * g.hw[l] + sp;
    return r;
}
// energy / max-logit from the values a lane holds (class c = lane + 32 * k); same operation order as the two-pass loop
template <int NV>
__device__ __forceinline__ void ood_reduce(const float (&v)[NV], int C, float T, int lane, float &e, float &m) {
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k)
        if (lane + 32 * k < C) mx = fmaxf(mx, v[k]);
    mx = warp_max(mx);
    float s = 0.f;
    const float invT = 1.0f / T;
#pragma unroll
    for (int k = 0; k < NV; ++k)
        if (lane + 32 * k < C) s += expf((v[k] - mx) * invT);
    s = warp_sum(s);
    e = -T * (mx * invT + logf(s));
    m = mx;
}
__device__ __forceinline__ void ood_row(const Geo &g, const float *const *levels, const unsigned char *nhwc, int b, int C, long long anc,
                                        float T, int lane, float &e, float &m) {
    e = 0.f; m = 0.f;
    const OodRow r = ood_row_of(g, levels, nhwc, b, C, anc);
    if (!r.row) return;
    if (C <= 128) {   // one round of loads, everything in registers
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = lane + 32 * k < C ? __ldg(r.row + (size_t)(lane + 32 * k) * r.cstride) : 0.f;
        ood_reduce(v, C, T, lane, e, m);
        return;
    }
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldg(r.row + (size_t)c * r.cstride));
    mx = warp_max(mx);
    float s = 0.f;
    const float invT = 1.0f / T;
    for (int c = lane; c < C; c += 32) s += expf((__ldg(r.row + (size_t)c * r.cstride) - mx) * invT);
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
