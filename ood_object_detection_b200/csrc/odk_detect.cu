// Detection generation (K4/K5/K6): box decode + score filter + class-aware NMS or Soft-NMS, and
// the per-detection OOD scores.  Replaces generate_detections / _batch_detection
// (reference effdet/anchors.py:95-172, effdet/bench.py:59-76), torchvision batched_nms
// (coordinate trick) and effdet/soft_nms.py.  See include/odk.h.
//
// One CTA (1024 threads) per image; all candidates of the image (<= 8192) live in shared
// memory.  Both suppressors are "frontier" algorithms that need at most max_det rounds instead
// of the O(n^2) mask of the classic bitmask NMS, because only the first max_det survivors are
// ever used (anchors.py:153):
//   hard NMS : round = take the first 16 still-alive candidates (score order), decide among them which are
//              kept (the greedy rule), kill every later candidate whose IoU with a newly kept one exceeds
//              the threshold; a warp owns 32 consecutive candidates = one word of the alive bitmask and
//              updates it with a ballot, so there are no atomics;
//   Soft-NMS : round = block-wide arg-max of the current scores (first index on ties), record it,
//              decay every alive score by exp(-iou^2/sigma), drop those at or below the score
//              threshold (soft_nms.py:88-110).
// All IoU arithmetic is in the reference's fp32 operation order on the class-offset boxes.
#include <string.h>

#include "odk_common.cuh"

namespace odk {

constexpr int kDetThreads = 1024;
constexpr int kDetMaxN = 8192;
constexpr int kDetWarps = kDetThreads / 32;
constexpr int kDetFirstWindow = 256;   // first activated chunk of the lazy window (doubles up to kDetThreads)

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
static size_t det_smem_bytes(int cap) { return (size_t)cap * 24 + (size_t)(cap / 32) * 4 + 16; }

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

__device__ void bitonic_sort_desc_u64(unsigned long long *s, int P) {
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
// Lazy window: only the first `window` candidates are "activated" at first; a later chunk is
// activated (each of its candidates tested against everything kept so far) only when the window
// runs out of alive candidates.  When the first max_keep survivors come from the first chunk --
// the common case -- the other candidates are never touched.  Same result as testing everything.
__device__ __forceinline__ bool nms_hit(float4 p, float ap, float4 q, float thr_f) {
    // boxes of other classes sit in other offset bands: when the x or y extents do not overlap the
    // intersection is 0 and the IoU cannot exceed a threshold >= 0, so the division is skipped
    const bool touch = (fminf(p.z, q.z) > fmaxf(p.x, q.x)) && (fminf(p.w, q.w) > fmaxf(p.y, q.y));
    return (touch || thr_f < 0.0f) && iou_nms(p, ap, q) > thr_f;
}

// Rounds are BATCHED: the first kLead alive candidates ("leaders", all earlier candidates are dead) are
// resolved among themselves in score order -- leader b is kept iff no kept leader a < b suppresses it, which
// is exactly what the one-at-a-time greedy loop decides for every candidate up to the last leader -- and
// then every later candidate is tested against all newly kept leaders in one pass.  ~max_keep / kLead
// rounds of two block barriers instead of max_keep rounds.
constexpr int kLead = 16;   // measured at D3 B=32: 8 / 16 / 32 leaders = 57 / 50 / 84 us per detect launch
constexpr int kPairs = kLead * (kLead - 1) / 2;   // 120 leader pairs (a < b), ordered by b then a
constexpr int kPairWords = (kPairs + 31) / 32;

__device__ int hard_nms_rounds(const DetSmem &S, int n, float thr_f, int max_keep, int *kept, int window, int window_max) {
    __shared__ int s_lead[kLead];
    __shared__ float4 s_lbox[kLead];
    __shared__ float s_larea[kLead];
    __shared__ int s_g;
    __shared__ unsigned s_keep;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int limit = min(n, window);   // activated prefix
    int count = 0, from = 0;      // words below `from` are dead
    while (count < max_keep) {
        __syncthreads();          // the alive words are final, last round's leader slots are free
        const int nwords = (limit + 31) / 32;
        if (warp == 0) {
            // ---- the first kLead alive candidates, in order ----
            int g = 0;
            for (int base = from; base < nwords && g < kLead; base += 32) {
                const int w = base + lane;
                const unsigned m = w < nwords ? S.alive[w] : 0u;
                const int c = __popc(m);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                unsigned mm = m;
                for (int r = g + incl - c; mm && r < kLead; ++r) {   // my set bits that rank below kLead
                    s_lead[r] = w * 32 + __ffs(mm) - 1;
                    mm &= mm - 1u;
                }
                g = min(kLead, g + __shfl_sync(0xffffffffu, incl, 31));
            }
            __syncwarp();
            if (g > 0) {
                if (lane < g) {
                    const float4 p = S.box[s_lead[lane]];
                    s_lbox[lane] = p;
                    s_larea[lane] = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
                }
                __syncwarp();
                // ---- resolve the leaders among themselves: pair p = b(b-1)/2 + a (a < b), lanes take p, p+32, ... ----
                unsigned pairs[kPairWords];
#pragma unroll
                for (int k = 0; k < kPairWords; ++k) {
                    const int p = k * 32 + lane;
                    int pb = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
                    if (pb * (pb - 1) / 2 > p) --pb;
                    if ((pb + 1) * pb / 2 <= p) ++pb;
                    const int pa = p - pb * (pb - 1) / 2;
                    bool hit = false;
                    if (p < kPairs && pb < g) hit = nms_hit(s_lbox[pa], s_larea[pa], s_lbox[pb], thr_f);
                    pairs[k] = __ballot_sync(0xffffffffu, hit);
                }
                unsigned keep_mask = 0u;
                int taken = 0;
                for (int bb = 0; bb < g; ++bb) {
                    const int bit = bb * (bb - 1) / 2;     // first pair of column bb
                    const int k = bit >> 5, sh = bit & 31;
                    unsigned lo = 0u, hi = 0u;
#pragma unroll
                    for (int j = 0; j < kPairWords; ++j) {
                        if (j == k) lo = pairs[j];
                        if (j == k + 1) hi = pairs[j];
                    }
                    const unsigned hb = __funnelshift_r(lo, hi, sh) & ((1u << bb) - 1u);   // kept a < bb that suppress bb
                    if (!(hb & keep_mask) && count + taken < max_keep) { keep_mask |= 1u << bb; ++taken; }
                }
                if (lane == 0) {
                    s_keep = keep_mask;
                    int c = count;
                    for (int bb = 0; bb < g; ++bb)
                        if ((keep_mask >> bb) & 1u) kept[c++] = s_lead[bb];
                }
            }
            if (lane == 0) s_g = g;
        }
        __syncthreads();
        const int g = s_g;
        if (g == 0) {
            if (limit >= n) break;
            // activate the next chunk: a candidate stays alive iff nothing kept so far suppresses it
            window = min(2 * window, window_max);   // chunks grow: 1 word per warp at most
            const int new_limit = min(n, limit + window);
            for (int w = limit / 32 + ((warp - limit / 32) & (kDetWarps - 1)); w < (new_limit + 31) / 32; w += kDetWarps) {
                const int i = w * 32 + lane;
                bool dead = i >= new_limit;
                if (!dead) {
                    const float4 q = S.box[i];
                    for (int c = 0; c < count && !dead; ++c) {
                        const float4 p = S.box[kept[c]];
                        dead = nms_hit(p, __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y)), q, thr_f);
                    }
                }
                const unsigned k = __ballot_sync(0xffffffffu, !dead);
                if (lane == 0) S.alive[w] = k;
            }
            from = limit / 32;
            limit = new_limit;
            continue;
        }
        const unsigned keep_mask = s_keep;
        const int last = s_lead[g - 1];
        count += __popc(keep_mask);
        // ---- suppression by the newly kept leaders; every leader leaves the alive set ----
        for (int w = from + ((warp - from) & (kDetWarps - 1)); w < nwords; w += kDetWarps) {
            unsigned m = S.alive[w];   // warp-uniform
            if (!m) continue;
            const int i = w * 32 + lane;
            bool kill = false;
            if ((m >> lane) & 1u) {
                if (i <= last) {
                    kill = true;       // a leader (everything else up to the last leader was dead already)
                } else {
                    const float4 q = S.box[i];
                    for (int bb = 0; bb < g && !kill; ++bb)
                        if ((keep_mask >> bb) & 1u) kill = nms_hit(s_lbox[bb], s_larea[bb], q, thr_f);
                }
            }
            const unsigned k = __ballot_sync(0xffffffffu, kill);
            if (lane == 0 && k) S.alive[w] = m & ~k;
        }
        from = last >> 5;
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
__device__ __forceinline__ float soft_decay(float4 p, float ap, float4 q, bool gaussian, float sigma, float iou_thr) {
    // disjoint extents -> iou 0 -> decay exactly 1
    if (!((fminf(p.z, q.z) > fmaxf(p.x, q.x)) && (fminf(p.w, q.w) > fmaxf(p.y, q.y)))) return 1.0f;
    const float iou = iou_soft(p, ap, q);
    if (gaussian) return expf(__fdiv_rn(-__fmul_rn(iou, iou), sigma));   // soft_nms.py:96
    return iou > iou_thr ? __fsub_rn(1.0f, iou) : 1.0f;                   // :98-100
}

template <class Emit>
__device__ int soft_nms_rounds(const DetSmem &S, int n, bool gaussian, float sigma, float iou_thr, float score_thr,
                               int max_rounds, int *picked, int window, int window_max, Emit emit) {
    // One block barrier per round: a warp that rescales its candidates also notes its best survivor
    // (score key, ~rank); after the barrier every warp reduces the 32 notes to the next pick itself.
    __shared__ unsigned long long s_best[2][kDetWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int limit = min(n, window);
    int count = 0, parity = 0;
    auto score_key = [&](int i) {   // order-preserving (scores may be <= 0 in the first round); never 0
        const unsigned u = __float_as_uint(S.score[i]);
        const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
        return ((unsigned long long)vk << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    };
    auto warp_max64 = [&](unsigned long long v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
            v = other > v ? other : v;
        }
        return v;
    };
    auto local_best = [&](int from_word, int nwords) {
        unsigned long long best = 0ull;
        for (int w = from_word + ((warp - from_word) & (kDetWarps - 1)); w < nwords; w += kDetWarps) {
            const unsigned m = S.alive[w];
            if ((m >> lane) & 1u) {
                const unsigned long long key = score_key(w * 32 + lane);
                best = key > best ? key : best;
            }
        }
        return warp_max64(best);
    };
    {
        const unsigned long long b0 = local_best(0, (limit + 31) / 32);
        if (lane == 0) s_best[0][warp] = b0;
    }
    while (count < max_rounds) {
        __syncthreads();
        const int nwords = (limit + 31) / 32;
        const unsigned long long pick = warp_max64(s_best[parity][lane]);
        if (limit < n) {
            // un-activated candidates still carry their original scores; the first one bounds them all
            const unsigned ub = __float_as_uint(S.score[limit]);
            const unsigned bk = ub ^ ((unsigned)((int)ub >> 31) | 0x80000000u);
            if (pick == 0ull || (unsigned)(pick >> 32) < bk) {
                window = min(2 * window, window_max);
                const int new_limit = min(n, limit + window);
                unsigned long long best = 0ull;
                for (int w = limit / 32 + ((warp - limit / 32) & (kDetWarps - 1)); w < (new_limit + 31) / 32; w += kDetWarps) {
                    const int i = w * 32 + lane;
                    bool ok = i < new_limit;
                    if (ok) {
                        const float4 q = S.box[i];
                        float sc = S.score[i];
                        for (int c = 0; c < count && ok; ++c) {
                            const float4 p = S.box[picked[c]];
                            const float d = soft_decay(p, __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y)), q, gaussian, sigma, iou_thr);
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
                best = warp_max64(best);
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
        for (int w = warp; w < nwords; w += kDetWarps) {
            unsigned m = S.alive[w];
            if (!m) continue;
            const int i = w * 32 + lane;
            bool kill = false;
            float sc = 0.f;
            if ((m >> lane) & 1u) {
                sc = S.score[i];
                const float d = soft_decay(p, ap, S.box[i], gaussian, sigma, iou_thr);
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
        best = warp_max64(best);
        if (lane == 0) s_best[parity ^ 1][warp] = best;
        parity ^= 1;
    }
    __syncthreads();
    return count;
}

struct DetArgs {
    const float *cls;       // [B,N]
    const float4 *box;      // [B,N]
    const long long *idx;   // [B,N] anchor index
    const long long *klass; // [B,N]
    const float4 *anchors;  // [A]
    const float *scale;     // [B] or null
    const float *size;      // [B,2] or null
    int N, cap;
    long long A;
    odk_detect_params p;
    float nms_thr_f;
    float *dets;            // [B,D,6]
    int *count;             // [B]
    int *src;               // [B,D]
};

__global__ void __launch_bounds__(kDetThreads) detect_kernel(const __grid_constant__ DetArgs A) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_wcnt[kDetWarps];
    __shared__ float s_wmax[kDetWarps];
    __shared__ int s_kept[1024];
    __shared__ float s_keptscore[1024];
    __shared__ int s_n, s_unsorted;
    const DetSmem S = carve(s_raw, A.cap);
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_raw);   // aliases S.box until step 4
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = A.N, D = A.p.max_det;
    const float *cls = A.cls + (size_t)b * N;
    const float4 *box = A.box + (size_t)b * N;
    const long long *idx = A.idx + (size_t)b * N;
    const long long *klass = A.klass + (size_t)b * N;
    const bool has_scale = A.scale != nullptr;
    const bool clip = has_scale && A.size != nullptr;                       // anchors.py:137
    const float scale = has_scale ? __ldg(A.scale + b) : 1.0f;
    const float lim_x = clip ? __fdiv_rn(__ldg(A.size + 2 * b), scale) : 0.f;
    const float lim_y = clip ? __fdiv_rn(__ldg(A.size + 2 * b + 1), scale) : 0.f;
    if (tid == 0) s_unsorted = 0;

    // 1. scores, score filter, order-preserving compaction (anchors.py:140-144)
    int total = 0;
    for (int base = 0; base < N; base += kDetThreads) {
        const int p = base + tid;
        float sc = 0.f;
        bool ok = false;
        if (p < N) { sc = sigmoid_ref(__ldg(cls + p)); ok = sc > A.p.score_min; }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = total, chunk = 0;
        for (int w = 0; w < kDetWarps; ++w) {
            const int c = s_wcnt[w];
            if (w < warp) off += c;
            chunk += c;
        }
        if (ok) {
            const int i = off + __popc(bal & ((1u << lane) - 1u));
            s_key[i] = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)p);
        }
        total += chunk;
        __syncthreads();
    }
    const int n = total;
    int kept_n = 0;
    if (n > 0) {
        // 2. torchvision::nms orders by descending score (stable); top-k output already is
        {
            int bad = 0;
            for (int i = tid; i + 1 < n; i += kDetThreads) bad |= (s_key[i] >> 32) < (s_key[i + 1] >> 32);
            if (bad) s_unsorted = 1;
            __syncthreads();
            if (s_unsorted && !A.p.soft_nms) {
                int P = 2;
                while (P < n) P <<= 1;
                for (int i = n + tid; i < P; i += kDetThreads) s_key[i] = 0ull;
                __syncthreads();
                bitonic_sort_desc_u64(s_key, P);
            }
        }
        // 3. keys -> registers, then decode boxes into the (aliased) box array
        unsigned long long mykeys[kDetMaxN / kDetThreads];
#pragma unroll
        for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
            const int i = tid + k * kDetThreads;
            mykeys[k] = i < n ? s_key[i] : 0ull;
        }
        __syncthreads();
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
            const int i = tid + k * kDetThreads;
            if (i < n) {
                const int p = (int)(0xFFFFFFFFu - (unsigned)(mykeys[k] & 0xFFFFFFFFull));
                const float4 o = decode_xyxy(__ldg(A.anchors + __ldg(idx + p)), __ldg(box + p), clip, lim_x, lim_y);
                S.box[i] = o;
                S.score[i] = __uint_as_float((unsigned)(mykeys[k] >> 32));
                S.src[i] = p;
                mx = fmaxf(mx, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
            }
        }
        // 4. coordinate trick: boxes + class * (max_coordinate + 1)  (torchvision boxes.py:105-108)
        mx = warp_max(mx);
        if (lane == 0) s_wmax[warp] = mx;
        init_alive(S.alive, n, A.cap);
        __syncthreads();
        mx = s_wmax[0];
        for (int w = 1; w < kDetWarps; ++w) mx = fmaxf(mx, s_wmax[w]);
        const float mul = __fadd_rn(mx, 1.0f);
        for (int i = tid; i < n; i += kDetThreads) {
            const float off = __fmul_rn((float)__ldg(klass + S.src[i]), mul);
            float4 o = S.box[i];
            o.x = __fadd_rn(o.x, off); o.y = __fadd_rn(o.y, off); o.z = __fadd_rn(o.z, off); o.w = __fadd_rn(o.w, off);
            S.box[i] = o;
        }
        __syncthreads();
        // 5. suppression, first D survivors
        if (A.p.soft_nms)
            // the window needs non-increasing scores: true for top-k output (checked below for API inputs)
            kept_n = soft_nms_rounds(S, n, true, A.p.soft_sigma, A.p.soft_iou, A.p.soft_score_thr, D, s_kept,
                                     s_unsorted ? n : kDetFirstWindow, s_unsorted ? n : kDetThreads,
                                     [&](int q, int i, float sc) { s_keptscore[q] = sc; });
        else
            kept_n = hard_nms_rounds(S, n, A.nms_thr_f, D, s_kept, kDetFirstWindow, kDetThreads);
        __syncthreads();
    }
    // 6. rows: boxes (re-decoded, unoffset) * img_scale, score, class + 1 (anchors.py:153-166)
    float *dets = A.dets + (size_t)b * D * 6;
    int *src = A.src + (size_t)b * D;
    for (int q = tid; q < D; q += kDetThreads) {
        float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int sp = -1;
        if (q < kept_n) {
            const int i = s_kept[q];
            sp = S.src[i];
            float4 o = decode_xyxy(__ldg(A.anchors + __ldg(idx + sp)), __ldg(box + sp), clip, lim_x, lim_y);
            if (has_scale) { o.x = __fmul_rn(o.x, scale); o.y = __fmul_rn(o.y, scale); o.z = __fmul_rn(o.z, scale); o.w = __fmul_rn(o.w, scale); }
            r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = o.w;
            r[4] = A.p.soft_nms ? s_keptscore[q] : S.score[i];
            r[5] = (float)(__ldg(klass + sp) + 1);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) dets[q * 6 + k] = r[k];
        src[q] = sp;
    }
    if (tid == 0) A.count[b] = kept_n;
}

// ---- stand-alone soft_nms / nms on one box set ---------------------------------------------------
__global__ void __launch_bounds__(kDetThreads)
soft_nms_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores, int n, int cap, int gaussian,
                float sigma, float iou_thr, float score_thr, int max_rounds, long long *__restrict__ idx_out,
                float *__restrict__ score_out, int *__restrict__ count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const DetSmem S = carve(s_raw, cap);
    for (int i = threadIdx.x; i < n; i += kDetThreads) { S.box[i] = __ldg(boxes + i); S.score[i] = __ldg(scores + i); }
    init_alive(S.alive, n, cap);
    __syncthreads();
    __shared__ int s_picked[kDetMaxN];   // 32 KB: ranks in pick order (needed to replay decays)
    const int c = soft_nms_rounds(S, n, gaussian != 0, sigma, iou_thr, score_thr, max_rounds, s_picked, n, n,
                                  [&](int q, int i, float sc) { idx_out[q] = i; score_out[q] = sc; });
    if (threadIdx.x == 0) *count = c;
}

__global__ void __launch_bounds__(kDetThreads)
nms_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores, int n, int cap, float thr_f,
           long long *__restrict__ keep, int *__restrict__ count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_kept[kDetMaxN];   // 32 KB: every box may survive
    const DetSmem S = carve(s_raw, cap);
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_raw);
    int P = 2;
    while (P < n) P <<= 1;
    // stable descending order on arbitrary (possibly negative) scores: order-preserving key
    for (int i = threadIdx.x; i < P; i += kDetThreads) {
        unsigned long long k = 0ull;
        if (i < n) {
            const unsigned u = __float_as_uint(__ldg(scores + i));
            const unsigned vk = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
            k = ((unsigned long long)vk << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        }
        s_key[i] = k;
    }
    __syncthreads();
    bitonic_sort_desc_u64(s_key, P);
    int mine[kDetMaxN / kDetThreads];
#pragma unroll
    for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
        const int i = threadIdx.x + k * kDetThreads;
        mine[k] = i < n ? (int)(0xFFFFFFFFu - (unsigned)(s_key[i] & 0xFFFFFFFFull)) : -1;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kDetMaxN / kDetThreads; ++k) {
        const int i = threadIdx.x + k * kDetThreads;
        if (i < n) { S.box[i] = __ldg(boxes + mine[k]); S.src[i] = mine[k]; }
    }
    init_alive(S.alive, n, cap);
    __syncthreads();
    const int c = hard_nms_rounds(S, n, thr_f, n, s_kept, n, n);
    __syncthreads();
    for (int q = threadIdx.x; q < c; q += kDetThreads) keep[q] = S.src[s_kept[q]];
    if (threadIdx.x == 0) *count = c;
}

// ---- OOD scores --------------------------------------------------------------------------------
// one warp per detection: energy = -T * logsumexp(row / T), max_logit = max(row)
struct LevelPtrs { const float *p[ODK_MAX_LEVELS]; };

__global__ void __launch_bounds__(256)
ood_kernel(const __grid_constant__ Geo g, const __grid_constant__ LevelPtrs lv, int B, int C, const long long *__restrict__ anchor_idx, int D, float T,
           float *__restrict__ energy, float *__restrict__ max_logit) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= B * D) return;
    const int b = wid / D;
    const long long anc = __ldg(anchor_idx + wid);
    float e = 0.f, m = 0.f;
    if (anc >= 0 && anc < g.A) {
        const int l = geo_level(g, (int)anc);
        const int loc = (int)anc - g.off[l];
        const int sp = loc / g.na, a = loc - sp * g.na;
        const float *row = lv.p[l] + ((size_t)(b * g.na + a) * C) * g.hw[l] + sp;   // stride hw between classes
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldg(row + (size_t)c * g.hw[l]));
        mx = warp_max(mx);
        float s = 0.f;
        const float invT = 1.0f / T;
        for (int c = lane; c < C; c += 32) s += expf((__ldg(row + (size_t)c * g.hw[l]) - mx) * invT);
        s = warp_sum(s);
        e = -T * (mx * invT + logf(s));
        m = mx;
    }
    if (lane == 0) { energy[wid] = e; max_logit[wid] = m; }
}

static float float_at_or_below(double d) {
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, -INFINITY);
    return f;
}

static int det_cap(int n) {
    int cap = (n + 1023) / 1024 * 1024;
    return cap < 1024 ? 1024 : cap;
}

}  // namespace odk

extern "C" {

int odk_detect(const float *cls_topk, const float *box_topk, const int64_t *indices, const int64_t *classes, int B,
               int N, const float *anchors, int64_t A, const float *img_scale, const float *img_size,
               const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, void *stream) {
    using namespace odk;
    if (!params) return set_error(ODK_EINVAL, "odk_detect: null params");
    if (B < 0 || N < 0) return set_error(ODK_EINVAL, "odk_detect: negative size");
    if (B == 0) return ODK_OK;
    if (!dets || !count || !src || !anchors || (N > 0 && (!cls_topk || !box_topk || !indices || !classes)))
        return set_error(ODK_EINVAL, "odk_detect: null pointer");
    if (N > kDetMaxN) return set_error(ODK_EUNSUPPORTED, "odk_detect: more than %d candidates per image", kDetMaxN);
    if (params->max_det < 1 || params->max_det > 1024) return set_error(ODK_EUNSUPPORTED, "odk_detect: max_det must be in [1,1024]");
    if (((uintptr_t)box_topk | (uintptr_t)anchors) & 15) return set_error(ODK_EINVAL, "odk_detect: box_topk / anchors must be 16-byte aligned");
    DetArgs a;
    memset(&a, 0, sizeof(a));
    a.cls = cls_topk; a.box = (const float4 *)box_topk; a.idx = (const long long *)indices; a.klass = (const long long *)classes;
    a.anchors = (const float4 *)anchors; a.scale = img_scale; a.size = img_size; a.N = N; a.cap = det_cap(N); a.A = A;
    a.p = *params; a.nms_thr_f = float_at_or_below(params->nms_iou);
    a.dets = dets; a.count = count; a.src = src;
    const size_t smem = det_smem_bytes(a.cap);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_smem_bytes(kDetMaxN));
        cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_smem_bytes(kDetMaxN));
        cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_smem_bytes(kDetMaxN));
        attr = true;
    }
    detect_kernel<<<B, kDetThreads, smem, (cudaStream_t)stream>>>(a);
    return check_launch("odk_detect");
}

int odk_soft_nms(const float *boxes, const float *scores, int n, int method_gaussian, float sigma, float iou_thr,
                 float score_thr, int max_rounds, int64_t *idx_out, float *score_out, int32_t *count, void *stream) {
    using namespace odk;
    if (n < 0 || !count) return set_error(ODK_EINVAL, "odk_soft_nms: bad arguments");
    if (n > kDetMaxN) return set_error(ODK_EUNSUPPORTED, "odk_soft_nms: more than %d boxes", kDetMaxN);
    if (n > 0 && (!boxes || !scores || !idx_out || !score_out)) return set_error(ODK_EINVAL, "odk_soft_nms: null pointer");
    if ((uintptr_t)boxes & 15) return set_error(ODK_EINVAL, "odk_soft_nms: boxes must be 16-byte aligned");
    if (max_rounds < 0 || max_rounds > n) max_rounds = n;
    const int cap = det_cap(n);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_smem_bytes(kDetMaxN));
        attr = true;
    }
    soft_nms_kernel<<<1, kDetThreads, det_smem_bytes(cap), (cudaStream_t)stream>>>(
        (const float4 *)boxes, scores, n, cap, method_gaussian, sigma, iou_thr, score_thr, max_rounds,
        (long long *)idx_out, score_out, count);
    return check_launch("odk_soft_nms");
}

size_t odk_nms_workspace_bytes(int n) { (void)n; return 16; }

int odk_nms(const float *boxes, const float *scores, int n, double iou_thr, int64_t *keep, int32_t *count,
            void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    (void)workspace; (void)workspace_bytes;
    if (n < 0 || !count) return set_error(ODK_EINVAL, "odk_nms: bad arguments");
    if (n > kDetMaxN) return set_error(ODK_EUNSUPPORTED, "odk_nms: more than %d boxes", kDetMaxN);
    if (n > 0 && (!boxes || !scores || !keep)) return set_error(ODK_EINVAL, "odk_nms: null pointer");
    if ((uintptr_t)boxes & 15) return set_error(ODK_EINVAL, "odk_nms: boxes must be 16-byte aligned");
    const int cap = det_cap(n);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)det_smem_bytes(kDetMaxN));
        attr = true;
    }
    nms_kernel<<<1, kDetThreads, det_smem_bytes(cap), (cudaStream_t)stream>>>((const float4 *)boxes, scores, n, cap,
                                                                             float_at_or_below(iou_thr), (long long *)keep, count);
    return check_launch("odk_nms");
}

int odk_ood(const void *const *cls_levels, int B, int C, const int32_t *level_hw, int num_levels, int na,
            const int64_t *anchor_idx, int D, float temperature, float *energy, float *max_logit, void *stream) {
    using namespace odk;
    Geo g;
    int rc = make_geo(&g, level_hw, num_levels, na);
    if (rc) return rc;
    if (B < 0 || D < 0 || C < 1) return set_error(ODK_EINVAL, "odk_ood: bad sizes");
    if (B == 0 || D == 0) return ODK_OK;
    if (!cls_levels || !anchor_idx || !energy || !max_logit) return set_error(ODK_EINVAL, "odk_ood: null pointer");
    if (!(temperature > 0.0f)) return set_error(ODK_EINVAL, "odk_ood: temperature must be positive");
    LevelPtrs lv;
    memset(&lv, 0, sizeof(lv));
    for (int l = 0; l < num_levels; ++l) {
        lv.p[l] = (const float *)cls_levels[l];
        if (!lv.p[l]) return set_error(ODK_EINVAL, "odk_ood: null level pointer (level %d)", l);
    }
    const long long warps = (long long)B * D;
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    ood_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, lv, B, C, (const long long *)anchor_idx, D, temperature, energy,
                                                          max_logit);
    return check_launch("odk_ood");
}

}  // extern "C"
