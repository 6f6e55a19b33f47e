// Fused detection loss (K2): focal/BCE classification loss + Huber box loss over all pyramid
// levels, forward and (optionally) the gradient of the total in the same pass.  HBM-bound: the
// [B, na*C, H, W] logits are streamed exactly once in their native NCHW layout; the one-hot
// target tensor the reference materialises (loss.py:182-186) never exists.
// See include/odk.h (odk_loss).
//
// Work decomposition: one item = (level, image b, anchor shape a, chunk of <=48 classes, 4
// consecutive (y,x) positions) = up to 48 rows of 16 bytes per thread.  Consecutive threads take
// consecutive position groups of one channel plane, so every warp access is 512 contiguous
// bytes.  The inner loop treats every element as a negative (target 0); the single positive class
// of a position (if any) is patched after the loop.  A persistent grid (a multiple of the SM
// count) walks the items; per-CTA partial sums are combined by the last CTA in a fixed order, so
// results are deterministic.
//
// Two kernels share all the arithmetic:
//   loss_kernel_ring : levels whose planes are 16-byte aligned rows (H*W % 4 == 0).  Each thread
//                      keeps a private ring of cp.async.ca (LDGSTS) 16-byte copies in shared memory,
//                      kRingDepth row-groups (4 rows = 64 B) ahead of the math, without spending
//                      registers on it.  Measured on B200 (profiles/r1_summary.md): the .cg form
//                      made L2 look every sector up twice (0.223 ms); with .ca the lines in flight
//                      live in L1, so a SMALLER ring (more L1 left of the 256 KB) is faster:
//                      4 / 3 / 2 slots = 0.196 / 0.192 / 0.190 ms (forward).
//   loss_kernel      : everything else (odd-sized levels such as D3's 7x7), plain register loads.
#include <stdlib.h>
#include <string.h>

#include "odk_common.cuh"
#include "odk_exchange.cuh"

namespace odk {

constexpr int kLossThreads = 256;
constexpr int kMaxPartials = 148 * 16;
constexpr int kMaxChunk = 48;                  // classes per work item
// row-groups of smem per thread: both passes are fastest with the smallest ring (most L1 left); measured
// fwd 0.190 ms and fwd+grad step 0.471 ms with 2 slots, 0.192 / 0.485 ms with 3
__host__ __device__ constexpr int ring_slots(bool grad) { return grad ? 2 : 2; }
constexpr int kMaxRingDepth = 1;               // row-groups in flight ahead of the math, at most
__host__ __device__ constexpr int ring_bytes(bool grad) { return ring_slots(grad) * 4 * kLossThreads * 16; }

enum LossMode { kNew = 0, kNewSmooth = 1, kLegacy = 2 };

struct LossArgs {
    Geo g;
    const float *cls[ODK_MAX_LEVELS];
    const float *box[ODK_MAX_LEVELS];
    float *gcls[ODK_MAX_LEVELS];
    float *gbox[ODK_MAX_LEVELS];
    unsigned item_off[ODK_MAX_LEVELS + 1];   // items of the levels THIS launch covers
    FastDiv div_nq[ODK_MAX_LEVELS];
    FastDiv div_nchunk, div_na;
    unsigned char cls_nhwc[ODK_MAX_LEVELS], box_nhwc[ODK_MAX_LEVELS];   // channels_last levels ([B, H, W, channels] in memory)
    int vec[ODK_MAX_LEVELS];   // 4 if the level's planes are 16-byte aligned rows of 4, else 1
    int nq[ODK_MAX_LEVELS];    // position groups per plane
    int B, C, cchunk, nchunk, Mmax;
    int keys_early;            // ring kernel: an item's assignment keys are requested when the load cursor enters it
    const int32_t *match;
    const float4 *anchors;
    const float4 *gt_boxes;
    const int32_t *gt_labels;
    const int64_t *cls_t;
    const float *box_t;
    const float *normalizer;
    odk_loss_params p;
    double *partials;     // [part_base + gridDim.x][2]
    int part_base;        // first partial slot of this launch
    int part_total;       // slots the finishing launch sums (0: this launch does not finish)
    int part_int_from;    // > 0: slots from here on hold fixed-point integer pairs (patch kernel)
    int finish_ctas;      // > 0: CTAs (of several concurrent launches) that report to the counter before the reduction is finished
    unsigned *counter;
    float *out;
    // optional fused exchange of the partial sums with the other data-parallel ranks (xworld == 0: off)
    Mailboxes xmb;
    int xworld, xrank;
    const float *xnpos;
    float *xout;
    int *xstatus;
    int xnormalized;
    unsigned xtimeout_ms;
    // optional: leave odk_assign_grid's workspace zero again after the last read of the keys (clear_cap > 0)
    unsigned long long *clr_keys;
    int32_t *clr_pos;
    const unsigned *clr_touched;
    unsigned tma_tile_off[ODK_MAX_LEVELS + 1];   // loss_flat_tma_kernel: 32 KB tiles per level, prefix sums
    int clr_in_patch;     // the patch kernel walks the labeler's list of matched anchors and zeroes the keys itself
    int patch_slices;     // CTAs that share an image's list
    unsigned *clr_done;
    int clr_cap;
};

// ---- element math ----------------------------------------------------------------------------
// softplus(x) = max(x,0) + log1p(e), e = exp(-|x|) in (0,1].  One MUFU.EX2 (flush-to-zero form: no
// denormal fix-up code); log1p(e) = e*(1 + e*Q(e)) with Q a degree-7 polynomial fitted at Chebyshev
// nodes of [0,1] (profiles/probes/log1p_poly.py: max relative error 3.1e-7 in fp32 Horner form, mean
// signed error 6e-10).  A MUFU.LG2 here would make the XU pipe (4 lanes/clk/SMSP) the kernel's limit
// and is biased near 1 (profiles/probes/lg2_probe.cu).
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#define ODK_Q0 -0.4999998211860657f
#define ODK_Q1 0.3333110213279724f
#define ODK_Q2 -0.2494998425245285f
#define ODK_Q3 0.19561563432216644f
#define ODK_Q4 -0.14697664976119995f
#define ODK_Q5 0.0910765677690506f
#define ODK_Q6 -0.03774333372712135f
#define ODK_Q7 0.007363723125308752f
__device__ __forceinline__ float log1p_unit(float e) {
    float t = fmaf(e, ODK_Q7, ODK_Q6);
    t = fmaf(e, t, ODK_Q5); t = fmaf(e, t, ODK_Q4); t = fmaf(e, t, ODK_Q3);
    t = fmaf(e, t, ODK_Q2); t = fmaf(e, t, ODK_Q1); t = fmaf(e, t, ODK_Q0);
    t = fmaf(e, t, 1.0f);
    return e * t;
}
__device__ __forceinline__ float softplus_fast(float x, float &e) {
    e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
    return fmaxf(x, 0.0f) + log1p_unit(e);
}
// Blackwell packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2: two fp32 results per issue slot).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// softplus of two values at once; same arithmetic as softplus_fast on packed pairs.
// Returns the packed softplus, e0/e1 = exp(-|x|).
__device__ __forceinline__ f32x2 softplus_fast2(float x0, float x1, float &e0, float &e1) {
    e0 = ex2_ftz(fabsf(x0) * -1.4426950408889634f);
    e1 = ex2_ftz(fabsf(x1) * -1.4426950408889634f);
    const f32x2 E = pk2(e0, e1);
    f32x2 T = fma2(E, pk2(ODK_Q7, ODK_Q7), pk2(ODK_Q6, ODK_Q6));
    T = fma2(E, T, pk2(ODK_Q5, ODK_Q5));
    T = fma2(E, T, pk2(ODK_Q4, ODK_Q4));
    T = fma2(E, T, pk2(ODK_Q3, ODK_Q3));
    T = fma2(E, T, pk2(ODK_Q2, ODK_Q2));
    T = fma2(E, T, pk2(ODK_Q1, ODK_Q1));
    T = fma2(E, T, pk2(ODK_Q0, ODK_Q0));
    T = fma2(E, T, pk2(1.0f, 1.0f));
    return fma2(E, T, pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}

__device__ __forceinline__ float sigmoid_from_e(float x, float e) {
    return __fdividef(x >= 0.0f ? 1.0f : e, 1.0f + e);
}

template <int VEC> struct Vec;
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load_stream(const float *p) { float4 t = ld_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void load(const float *p) { float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void store(float *p) const { st_stream4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load_stream(const float *p) { v[0] = ld_stream1(p); }
    __device__ __forceinline__ void load(const float *p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void store(float *p) const { st_stream1(p, v[0]); }
};

// ---- one work item ------------------------------------------------------------------------------
struct Item {
    int l, b, a, chunk, s0, c0, c1, hw;
    size_t plane0;   // element offset of (b, a, class c0, position s0) in the level's logits
};

template <int VEC>
__device__ __forceinline__ Item decode_item(const LossArgs &A, int l, unsigned local) {
    const Geo &g = A.g;
    Item it;
    it.l = l;
    it.hw = g.hw[l];
    const unsigned nq = (unsigned)A.nq[l];
    const unsigned t = fd_div(local, A.div_nq[l]);
    const unsigned q = local - t * nq;
    const unsigned t2 = fd_div(t, A.div_nchunk);
    it.chunk = (int)(t - t2 * (unsigned)A.nchunk);
    const unsigned t3 = fd_div(t2, A.div_na);
    it.a = (int)(t2 - t3 * (unsigned)g.na);
    it.b = (int)t3;
    it.s0 = (int)q * VEC;
    it.c0 = it.chunk * A.cchunk;
    it.c1 = min(it.c0 + A.cchunk, A.C);
    it.plane0 = ((size_t)(it.b * g.na + it.a) * A.C + it.c0) * it.hw + it.s0;
    return it;
}

__device__ __forceinline__ int level_of_item(const LossArgs &A, unsigned it) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < ODK_MAX_LEVELS; ++i)
        if (i < A.g.nlev && it >= A.item_off[i]) l = i;
    return l;
}

// class target of each of my positions: >=0 class, -1 background, -2 ignore (+ matched gt row)
template <int VEC>
__device__ __forceinline__ void targets_from_match(const LossArgs &A, const Item &it, int (&tc)[VEC], const int (&mt)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) tc[j] = mt[j] >= 0 ? __ldg(A.gt_labels + (size_t)it.b * A.Mmax + mt[j]) - 1 : -1;
}
__device__ __forceinline__ int match_of_key(unsigned long long k) {
    return k ? (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)) : -1;
}
template <int VEC, bool FUSED>
__device__ __forceinline__ void item_targets(const LossArgs &A, const Item &it, int (&tc)[VEC], int (&mt)[VEC]) {
    const Geo &g = A.g;
    if (FUSED) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const size_t mi = (size_t)it.b * g.Apad + g.off[it.l] + it.a * it.hw + it.s0 + j;
            if (A.p.match_is_key64) mt[j] = match_of_key(__ldg(reinterpret_cast<const unsigned long long *>(A.match) + mi));
            else mt[j] = __ldg(A.match + mi);
        }
        targets_from_match<VEC>(A, it, tc, mt);
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            mt[j] = -1;
            tc[j] = (int)__ldg(A.cls_t + (size_t)A.B * g.off[it.l] + ((size_t)it.b * it.hw + it.s0 + j) * g.na + it.a);
        }
    }
}

// one row (one class) of VEC positions, all treated as negatives
template <int VEC, int MODE, bool GRAD>
__device__ __forceinline__ void row_values(const LossArgs &A, const Vec<VEC> &x, Vec<VEC> &gr, float gneg, float (&acc)[VEC],
                                           float (&accx)[VEC]) {
    const float gamma = A.p.gamma, sm = A.p.label_smoothing;
    if (VEC == 4 && MODE != kLegacy) {
        float e[4];
        const f32x2 s01 = softplus_fast2(x.v[0], x.v[1], e[0], e[1]);
        const f32x2 s23 = softplus_fast2(x.v[2], x.v[3], e[2], e[3]);
        upk2(add2(pk2(acc[0], acc[1]), s01), acc[0], acc[1]);
        upk2(add2(pk2(acc[2], acc[3]), s23), acc[2], acc[3]);
        if (MODE == kNewSmooth) {
            upk2(add2(pk2(accx[0], accx[1]), pk2(x.v[0], x.v[1])), accx[0], accx[1]);
            upk2(add2(pk2(accx[2], accx[3]), pk2(x.v[2], x.v[3])), accx[2], accx[3]);
        }
        if (GRAD) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float sg = sigmoid_from_e(x.v[j], e[j]);
                gr.v[j] = gneg * (MODE == kNewSmooth ? sg - 0.5f * sm : sg);
            }
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        if (MODE == kLegacy) {
            // t = 0: modulator = exp(-gamma*softplus(-x)), bce = softplus(x)   (loss.py:40-47)
            const float xv = x.v[j];
            const float sp = fmaxf(xv, 0.f) + log1pf(expf(-fabsf(xv)));
            const float mod = expf(-gamma * (sp - xv));
            acc[j] += mod * sp;
            if (GRAD) {
                const float sg = 1.0f / (1.0f + expf(-xv));
                gr.v[j] = gneg * mod * (sg + sp * gamma * (1.0f - sg));
            }
        } else {
            float e;
            const float sp = softplus_fast(x.v[j], e);
            acc[j] += sp;
            if (MODE == kNewSmooth) accx[j] += x.v[j];
            if (GRAD) {
                const float sg = sigmoid_from_e(x.v[j], e);
                gr.v[j] = gneg * (MODE == kNewSmooth ? sg - 0.5f * sm : sg);
            }
        }
    }
}
template <int VEC, int MODE, bool GRAD>
__device__ __forceinline__ void row_compute(const LossArgs &A, const Vec<VEC> &x, float *gp, float gneg, float (&acc)[VEC],
                                            float (&accx)[VEC]) {
    Vec<VEC> gr;
    row_values<VEC, MODE, GRAD>(A, x, gr, gneg, acc, accx);
    if (GRAD) gr.store(gp);
}

// combine the row sums, patch the positive class, apply the ignore mask (loss.py:285), and do the
// Huber box loss once per (b, a, positions) in the first class chunk
template <int VEC, int MODE, bool GRAD, bool FUSED>
__device__ __forceinline__ void item_finish(const LossArgs &A, const Item &it, const int (&tc)[VEC], const int (&mt)[VEC],
                                            const float (&acc)[VEC], const float (&accx)[VEC], float inv_n, float &csum,
                                            float &bsum) {
    const Geo &g = A.g;
    const int l = it.l, hw = it.hw, c0 = it.c0, c1 = it.c1;
    const float alpha = A.p.alpha, sm = A.p.label_smoothing, gamma = A.p.gamma;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float val = (1.0f - alpha) * (MODE == kNewSmooth ? acc[j] - 0.5f * sm * accx[j] : acc[j]);
        if (tc[j] >= c0 && tc[j] < c1) {
            const size_t o = ((size_t)(it.b * g.na + it.a) * A.C + tc[j]) * hw + it.s0 + j;
            const float xp = __ldg(A.cls[l] + o);
            float gpos;
            if (MODE == kLegacy) {
                const float sp = fmaxf(xp, 0.f) + log1pf(expf(-fabsf(xp)));
                const float sg = 1.0f / (1.0f + expf(-xp));
                const float mod_n = expf(-gamma * (sp - xp));          // negative-target modulator
                const float mod_p = expf(-gamma * sp);                 // positive-target modulator
                const float bce_p = sp - xp;
                val += alpha * mod_p * bce_p - (1.0f - alpha) * mod_n * sp;
                gpos = alpha * inv_n * mod_p * ((sg - 1.0f) - bce_p * gamma * sg);
            } else {
                float e;
                const float sp = softplus_fast(xp, e);
                const float tp = MODE == kNewSmooth ? 1.0f - 0.5f * sm : 1.0f;  // smoothed positive target
                const float tn = MODE == kNewSmooth ? 0.5f * sm : 0.0f;
                val += alpha * (sp - tp * xp) - (1.0f - alpha) * (sp - tn * xp);
                gpos = alpha * inv_n * (sigmoid_from_e(xp, e) - tp);
            }
            if (GRAD) A.gcls[l][o] = gpos;
        }
        if (tc[j] == -2) {
            val = 0.f;
            if (GRAD) {
                float *pz = A.gcls[l] + it.plane0 + j;
                for (int c = c0; c < c1; ++c, pz += hw) *pz = 0.f;
            }
        }
        csum += val;
    }

    if (it.chunk == 0) {
        float tg[VEC][4];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (FUSED) {
                if (mt[j] >= 0)
                    t4 = encode_ref(__ldg(A.gt_boxes + (size_t)it.b * A.Mmax + mt[j]),
                                    __ldg(A.anchors + g.off[l] + (it.s0 + j) * g.na + it.a));
            } else {
                t4 = __ldg(reinterpret_cast<const float4 *>(A.box_t) + (size_t)A.B * g.off[l] +
                           ((size_t)it.b * hw + it.s0 + j) * g.na + it.a);
            }
            tg[j][0] = t4.x; tg[j][1] = t4.y; tg[j][2] = t4.z; tg[j][3] = t4.w;
        }
        const float delta = A.p.delta;
        const float gb = A.p.box_loss_weight * inv_n * 0.25f;
        const size_t bplane = ((size_t)(it.b * g.na + it.a) * 4) * hw + it.s0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            Vec<VEC> o, gr;
            o.load(A.box[l] + bplane + (size_t)k * hw);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float tv = tg[j][k];
                const float e = o.v[j] - tv;                       // loss.py:108-112
                const float ae = fabsf(e);
                const float qd = fminf(ae, delta);
                const float lv = 0.5f * qd * qd + delta * (ae - qd);
                const bool on = tv != 0.0f;                        // loss.py:177
                bsum += on ? lv : 0.f;
                if (GRAD) gr.v[j] = on ? gb * (ae <= delta ? e : copysignf(delta, e)) : 0.f;
            }
            if (GRAD) gr.store(A.gbox[l] + bplane + (size_t)k * hw);
        }
    }
}

// register-load path: any level
template <int VEC, int MODE, bool GRAD, bool FUSED>
__device__ __forceinline__ void loss_item(const LossArgs &A, int l, unsigned local, float inv_n, float &csum, float &bsum) {
    const Item it = decode_item<VEC>(A, l, local);
    int tc[VEC], mt[VEC];
    item_targets<VEC, FUSED>(A, it, tc, mt);
    const int hw = it.hw;
    const float *px = A.cls[l] + it.plane0;
    float *pg = GRAD ? A.gcls[l] + it.plane0 : nullptr;
    float acc[VEC], accx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { acc[j] = 0.f; accx[j] = 0.f; }
    const float gneg = (1.0f - A.p.alpha) * inv_n;
    int c = it.c0;
    for (; c + 4 <= it.c1; c += 4) {   // 4 independent loads in flight per thread
        Vec<VEC> x0, x1, x2, x3;
        x0.load_stream(px);
        x1.load_stream(px + hw);
        x2.load_stream(px + 2 * (size_t)hw);
        x3.load_stream(px + 3 * (size_t)hw);
        row_compute<VEC, MODE, GRAD>(A, x0, pg, gneg, acc, accx);
        row_compute<VEC, MODE, GRAD>(A, x1, pg + (GRAD ? hw : 0), gneg, acc, accx);
        row_compute<VEC, MODE, GRAD>(A, x2, pg + (GRAD ? 2 * (size_t)hw : 0), gneg, acc, accx);
        row_compute<VEC, MODE, GRAD>(A, x3, pg + (GRAD ? 3 * (size_t)hw : 0), gneg, acc, accx);
        px += 4 * (size_t)hw;
        if (GRAD) pg += 4 * (size_t)hw;
    }
    for (; c < it.c1; ++c) {
        Vec<VEC> x0;
        x0.load_stream(px);
        row_compute<VEC, MODE, GRAD>(A, x0, pg, gneg, acc, accx);
        px += hw;
        if (GRAD) pg += hw;
    }
    item_finish<VEC, MODE, GRAD, FUSED>(A, it, tc, mt, acc, accx, inv_n, csum, bsum);
}

// ---- deterministic two-stage reduction shared by both kernels --------------------------------
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// `fixed`: this CTA's sums are the 2^-32 fixed-point integers ci / bi (patch kernel); its partial slot then holds
// the two integers, and the finishing CTA adds all integer slots (>= part_int_from) exactly before converting once.
__device__ __forceinline__ void finish_block(const LossArgs &A, float csum, float bsum, float nrm, double cextra = 0.0, bool fixed = false,
                                             long long ci = 0, long long bi = 0) {
    __shared__ double s_c[kLossThreads / 32], s_b[kLossThreads / 32];
    __shared__ bool s_last;
    __shared__ float4 s_mine, s_recv[ODK_MAILBOX_MAX_WORLD];
    __shared__ long long s_ci[kLossThreads / 32], s_bi[kLossThreads / 32];
    double dc = warp_sum((double)csum + cextra), db = warp_sum((double)bsum);
    if (fixed) { ci = warp_sum_ll(ci); bi = warp_sum_ll(bi); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_c[warp] = dc; s_b[warp] = db; s_ci[warp] = ci; s_bi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0, bx = 0;
        for (int w = 0; w < kLossThreads / 32; ++w) { c += s_c[w]; bx += s_b[w]; }
        if (fixed) {
            long long ic = 0, ib = 0;
            for (int w = 0; w < kLossThreads / 32; ++w) { ic += s_ci[w]; ib += s_bi[w]; }
            c = __longlong_as_double(ic); bx = __longlong_as_double(ib);
        }
        const unsigned slot = blockIdx.y * gridDim.x + blockIdx.x;
        A.partials[2 * (A.part_base + slot)] = c;
        A.partials[2 * (A.part_base + slot) + 1] = bx;
        s_last = false;
        if (A.part_total > 0) {   // (a launch that does not finish the reduction leaves the counter alone)
            __threadfence();
            const unsigned done = atomicAdd(A.counter, 1u);
            s_last = done == (A.finish_ctas > 0 ? (unsigned)A.finish_ctas : gridDim.x * gridDim.y) - 1;
        }
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (A.clr_in_patch) {   // every CTA has read (and zeroed) its keys: the labeler's counters for the next step
            for (int i = threadIdx.x; i < A.B; i += kLossThreads) A.clr_pos[(size_t)i * kCtrStride] = 0;
            if (threadIdx.x == 0) *A.clr_done = 0u;
        }
        double c = 0, bx = 0;
        long long ic = 0, ib = 0;
        const int ifrom = A.part_int_from > 0 ? A.part_int_from : A.part_total;
        for (int i = threadIdx.x; i < A.part_total; i += kLossThreads) {
            const double pc = __ldcg(A.partials + 2 * i), pb = __ldcg(A.partials + 2 * i + 1);
            if (i >= ifrom) { ic += __double_as_longlong(pc); ib += __double_as_longlong(pb); }
            else { c += pc; bx += pb; }
        }
        c = warp_sum(c); bx = warp_sum(bx);
        ic = warp_sum_ll(ic); ib = warp_sum_ll(ib);
        if (lane == 0) { s_c[warp] = c; s_b[warp] = bx; s_ci[warp] = ic; s_bi[warp] = ib; }
        __syncthreads();
        if (threadIdx.x == 0) {
            c = 0; bx = 0; ic = 0; ib = 0;
            for (int w = 0; w < kLossThreads / 32; ++w) { c += s_c[w]; bx += s_b[w]; ic += s_ci[w]; ib += s_bi[w]; }
            c += (double)ic * (1.0 / 4294967296.0);
            bx += (double)ib * (1.0 / 4294967296.0);
            const double n = (double)nrm;
            const float cls_loss = (float)(c / n);
            const float box_loss = (float)(bx / (n * 4.0));                 // loss.py:176-179
            A.out[1] = cls_loss;
            A.out[2] = box_loss;
            A.out[0] = cls_loss + A.p.box_loss_weight * box_loss;         // loss.py:297
            *A.counter = 0;                                                 // ready for the next call
            s_mine = make_float4(cls_loss + A.p.box_loss_weight * box_loss, cls_loss, box_loss, 0.f);
        }
        if (A.xworld > 0) {
            // fused exchange: first collect the previous step's record set if one is outstanding (it landed
            // a whole step ago), then store this step's sums into every rank's mailbox over NVLink
            __syncthreads();
            if (threadIdx.x < 32) {
                unsigned *ctr = mailbox_counters(A.xmb.p[A.xrank], A.xworld);
                if (ctr[0] > ctr[1]) collect_records(A.xmb.p[A.xrank], A.xworld, s_recv, A.xout, A.xstatus, A.xtimeout_ms);
                float4 v = s_mine;
                v.w = __ldg(A.xnpos);
                if (A.xnormalized) { v.x *= v.w; v.y *= v.w; v.z *= v.w; }   // back to the un-normalised sums the records carry
                publish_record(A.xmb, A.xworld, A.xrank, v);
            }
        }
    }
}

template <int MODE, bool GRAD, bool FUSED>
__global__ void __launch_bounds__(kLossThreads, 3)
loss_kernel(const __grid_constant__ LossArgs A) {
    const float nrm = __ldg(A.normalizer);
    const float inv_n = 1.0f / nrm;
    float csum = 0.f, bsum = 0.f;
    const unsigned total = A.item_off[A.g.nlev];
    const unsigned stride = gridDim.x * kLossThreads;
    for (unsigned it = blockIdx.x * kLossThreads + threadIdx.x; it < total; it += stride) {
        const int l = level_of_item(A, it);
        const unsigned local = it - A.item_off[l];
        if (A.vec[l] == 4) loss_item<4, MODE, GRAD, FUSED>(A, l, local, inv_n, csum, bsum);
        else loss_item<1, MODE, GRAD, FUSED>(A, l, local, inv_n, csum, bsum);
    }
    finish_block(A, csum, bsum, nrm);
}

// ---- cp.async ring kernel (vec4 levels only) -----------------------------------------------------
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const float *gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MODE, bool GRAD, bool FUSED>
__global__ void __launch_bounds__(kLossThreads, 3)
loss_kernel_ring(const __grid_constant__ LossArgs A) {
    extern __shared__ __align__(16) unsigned char s_ring[];
    constexpr int kRingSlots = ring_slots(GRAD), kRingDepth = kRingSlots - 1;
    const float nrm = __ldg(A.normalizer);
    const float inv_n = 1.0f / nrm;
    float csum = 0.f, bsum = 0.f;
    const unsigned total = A.item_off[A.g.nlev];
    const unsigned stride = gridDim.x * kLossThreads;
    const unsigned first = blockIdx.x * kLossThreads + threadIdx.x;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(s_ring) + threadIdx.x * 16u;
    const float4 *ring_rd = reinterpret_cast<const float4 *>(s_ring) + threadIdx.x;
    // slot s, row r of this thread: ring + ((s * 4 + r) * kLossThreads) * 16 bytes (conflict-free)

    // load cursor: runs kRingDepth row-groups ahead of the compute cursor over the same sequence
    unsigned it_ld = first;
    const float *px_ld = nullptr;
    int rows_ld = 0, hw_ld = 0;
    // The 4 assignment keys (low words: ~gt index, 0 = unmatched) or match indices of an item are
    // requested when the load cursor enters it, >= kRingDepth row-groups before the math needs them.
    // (Registers, not the ring: more shared memory would shrink the L1 the cp.async.ca loads pass through.)
    const bool keys_early = FUSED && A.keys_early;
    unsigned nk[4] = {0u, 0u, 0u, 0u};
    auto enter_item = [&]() {
        const int l = level_of_item(A, it_ld);
        const Item t = decode_item<4>(A, l, it_ld - A.item_off[l]);
        px_ld = A.cls[l] + t.plane0; rows_ld = t.c1 - t.c0; hw_ld = t.hw;
        if (keys_early) {
            const size_t mi = (size_t)t.b * A.g.Apad + A.g.off[l] + t.a * t.hw + t.s0;
            const unsigned *kp = reinterpret_cast<const unsigned *>(A.match) + (A.p.match_is_key64 ? 2 * mi : mi);
            const int step = A.p.match_is_key64 ? 2 : 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) nk[j] = __ldg(kp + j * step);
        }
    };
    if (it_ld < total) enter_item();
    auto issue_group = [&](int slot) {
        if (it_ld < total) {
            const unsigned dst = ring0 + (unsigned)(slot * 4) * (kLossThreads * 16u);
            if (rows_ld >= 4) {   // common case: no per-row predicates
                cp_async16(dst, px_ld);
                cp_async16(dst + kLossThreads * 16u, px_ld + hw_ld);
                cp_async16(dst + 2 * kLossThreads * 16u, px_ld + 2 * (size_t)hw_ld);
                cp_async16(dst + 3 * kLossThreads * 16u, px_ld + 3 * (size_t)hw_ld);
            } else {
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (r < rows_ld) cp_async16(dst + (unsigned)r * (kLossThreads * 16u), px_ld + (size_t)r * hw_ld);
            }
            rows_ld -= 4;
            px_ld += 4 * (size_t)hw_ld;
            if (rows_ld <= 0) {
                it_ld += stride;
                if (it_ld < total) enter_item();
            }
        }
        cp_async_commit();   // always commit so group counting stays uniform
    };
#pragma unroll
    for (int s = 0; s < kRingDepth; ++s) issue_group(s);

    int slot = 0;
    const float gneg = (1.0f - A.p.alpha) * inv_n;
    for (unsigned itx = first; itx < total; itx += stride) {
        const int l = level_of_item(A, itx);
        const Item it = decode_item<4>(A, l, itx - A.item_off[l]);
        int tc[4], mt[4];
        if (keys_early) {
#pragma unroll
            for (int j = 0; j < 4; ++j) mt[j] = A.p.match_is_key64 ? (nk[j] ? (int)(0xFFFFFFFFu - nk[j]) : -1) : (int)nk[j];
            targets_from_match<4>(A, it, tc, mt);
        } else {
            item_targets<4, FUSED>(A, it, tc, mt);
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, accx[4] = {0.f, 0.f, 0.f, 0.f};
        float *pg = GRAD ? A.gcls[l] + it.plane0 : nullptr;
        for (int c = it.c0; c < it.c1; c += 4) {
            issue_group((slot + kRingDepth) % kRingSlots);
            cp_async_wait<kRingDepth>();   // the oldest group (this slot) has landed
            const int rows = it.c1 - c;
            const float4 *rd = ring_rd + (size_t)(slot * 4) * kLossThreads;
            if (rows >= 4) {
                Vec<4> x[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 q = rd[(size_t)r * kLossThreads];
                    x[r].v[0] = q.x; x[r].v[1] = q.y; x[r].v[2] = q.z; x[r].v[3] = q.w;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    row_compute<4, MODE, GRAD>(A, x[r], pg + (GRAD ? (size_t)r * it.hw : 0), gneg, acc, accx);
            } else {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (r < rows) {
                        const float4 q = rd[(size_t)r * kLossThreads];
                        Vec<4> x;
                        x.v[0] = q.x; x.v[1] = q.y; x.v[2] = q.z; x.v[3] = q.w;
                        row_compute<4, MODE, GRAD>(A, x, pg + (GRAD ? (size_t)r * it.hw : 0), gneg, acc, accx);
                    }
                }
            }
            if (GRAD) pg += 4 * (size_t)it.hw;
            slot = (slot + 1) % kRingSlots;
        }
        item_finish<4, MODE, GRAD, FUSED>(A, it, tc, mt, acc, accx, inv_n, csum, bsum);
    }
    cp_async_wait<0>();
    finish_block(A, csum, bsum, nrm);
}

// After the loss: zero the assignment keys the labeler set (it listed them), its positive counters and its CTA
// counter, so the next odk_assign_grid needs no 8-bytes-per-anchor memset (ODK_ASSIGN_WS_CLEAN).  One CTA per image.
__global__ void __launch_bounds__(256)
clear_keys_kernel(unsigned long long *keys, int32_t *pos, const unsigned *__restrict__ touched, unsigned *done, int Apad, int cap) {
    const int b = blockIdx.x;
    const int n = pos[(size_t)b * kCtrStride];
    unsigned long long *row = keys + (size_t)b * Apad;
    if (n <= cap) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) row[__ldg(touched + (size_t)b * cap + i)] = 0ull;
    } else {   // more positives than the list holds: the whole row
        for (int i = threadIdx.x; i < Apad; i += blockDim.x) row[i] = 0ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pos[(size_t)b * kCtrStride] = 0;
        if (b == 0) *done = 0u;
    }
}


// ---- channels_last inputs: a layout-agnostic stream + a per-anchor patch --------------------------------------
// With [B, H, W, na*C] logits a thread's 16 bytes are four CLASSES of one anchor, not four positions of one class,
// so the plane-walking items above do not apply.  The work splits instead into
//   loss_flat_kernel  : every logit of every level as a NEGATIVE (target 0): sum of (1 - alpha) * softplus(x) and,
//                       with gradients, (1 - alpha) / N * sigmoid(x) written in the same pass.  It needs no targets
//                       and no layout: a level is one flat array.  Box gradients are zero-filled here.
//   loss_patch_kernel : one thread per anchor; for the matched ones (a few hundred per image) the positive class'
//                       term and gradient are corrected and the Huber box loss / gradient of its four codes is added.
//                       Knows both layouts per level, so mixed batches of levels work too.  Finishes the reduction.
template <int MODE, bool GRAD>
__global__ void __launch_bounds__(kLossThreads, 4)
loss_flat_kernel(const __grid_constant__ LossArgs A) {
    const float nrm = __ldg(A.normalizer);
    const float inv_n = 1.0f / nrm;
    const float gneg = (1.0f - A.p.alpha) * inv_n;
    const size_t stride = (size_t)gridDim.x * kLossThreads;
    const size_t first = (size_t)blockIdx.x * kLossThreads + threadIdx.x;
    double dsum = 0.0;
    for (int l = 0; l < A.g.nlev; ++l) {
        const size_t len = (size_t)A.B * A.g.na * A.C * A.g.hw[l];
        const float *px = A.cls[l];
        float *pg = GRAD ? A.gcls[l] : nullptr;
        const bool vec4 = (((uintptr_t)px | (uintptr_t)pg) & 15) == 0;
        const size_t n4 = vec4 ? len / 4 : 0;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, accx[4] = {0.f, 0.f, 0.f, 0.f};
        int since = 0;
        auto flush = [&]() {   // fp32 partial sums stay short: a few thousand terms each
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { v += (MODE == kNewSmooth ? acc[j] - 0.5f * A.p.label_smoothing * accx[j] : acc[j]); acc[j] = 0.f; accx[j] = 0.f; }
            dsum += (double)v;
            since = 0;
        };
        // A CTA takes CONTIGUOUS 16 KB tiles (four 128-bit loads in flight per thread, 4 KB apart), whole rounds of the
        // grid; what is left of the level goes grid-stride so that every CTA gets the same share.  Measured with the
        // same arithmetic on a 1.13 GB read + write stream (profiles/micro/stream_micro.cu): tiles 0.391 ms, loads
        // a grid-stride apart 0.402 ms (cudaMemcpy 0.342, a plain LDG/STG copy kernel 0.378).
        constexpr int kU = 4;
        const size_t tile = (size_t)kLossThreads * kU;
        const size_t ntiles = (n4 / tile) / gridDim.x * gridDim.x;
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const size_t base = t * tile + threadIdx.x;
            Vec<4> x[kU];
#pragma unroll
            for (int j = 0; j < kU; ++j) x[j].load_stream(px + (base + (size_t)j * kLossThreads) * 4);
#pragma unroll
            for (int j = 0; j < kU; ++j)
                row_compute<4, MODE, GRAD>(A, x[j], GRAD ? pg + (base + (size_t)j * kLossThreads) * 4 : nullptr, gneg, acc, accx);
            if (++since == 256) flush();
        }
        for (size_t u = ntiles * tile + first; u < n4; u += stride) {
            Vec<4> x0;
            x0.load_stream(px + u * 4);
            row_compute<4, MODE, GRAD>(A, x0, GRAD ? pg + u * 4 : nullptr, gneg, acc, accx);
        }
        flush();
        float t1[1] = {0.f}, t1x[1] = {0.f};
        for (size_t e = n4 * 4 + first; e < len; e += stride) {   // unaligned level, or the 1-3 elements past the last float4
            Vec<1> x0;
            x0.load_stream(px + e);
            row_compute<1, MODE, GRAD>(A, x0, GRAD ? pg + e : nullptr, gneg, t1, t1x);
        }
        dsum += (double)(MODE == kNewSmooth ? t1[0] - 0.5f * A.p.label_smoothing * t1x[0] : t1[0]);
        if (GRAD) {   // box gradients: zero everywhere, the patch kernel writes the matched anchors'
            float *gb = A.gbox[l];
            const size_t blen = (size_t)A.B * A.g.na * 4 * A.g.hw[l];
            if (((uintptr_t)gb & 15) == 0) {
                for (size_t q = first; q < blen / 4; q += stride) st_stream4(gb + q * 4, make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
                for (size_t q = first; q < blen; q += stride) gb[q] = 0.f;
            }
        }
    }
    finish_block(A, (float)0.f, 0.f, nrm, (1.0 - (double)A.p.alpha) * dsum);
}

// ---- the gradient stream through TMA bulk copies ----------------------------------------------------------------
// Same work as loss_flat_kernel<MODE, true>, but a tile (32 KB of logits) travels global -> shared memory as ONE
// cp.async.bulk with an mbarrier, the arithmetic runs in place in shared memory, and the tile of gradients goes
// shared -> global as one bulk store: the SM issues two instructions per 32 KB instead of 4096 LDG/STG, and three tiles
// per CTA are in flight.  profiles/micro/stream_micro.cu measured this shape at 0.381 ms against 0.392 ms for the
// LDG/STG tiles on the 1.13 GB read + write stream (cudaMemcpy: 0.342 ms).  Levels whose arrays are not 16-byte
// aligned, the part of a level that does not fill a tile, and the box-gradient zero fill go the LDG/STG way below.
constexpr int kTmaTileF4 = 2048;                  // float4 per tile: 32 KB
constexpr int kTmaStages = 3;
constexpr size_t kTmaSmemBytes = (size_t)kTmaStages * kTmaTileF4 * 16;
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(kLossThreads, 2)
loss_flat_tma_kernel(const __grid_constant__ LossArgs A) {
    extern __shared__ __align__(128) unsigned char s_tiles[];
    __shared__ __align__(8) unsigned long long s_full[kTmaStages];
    const float nrm = __ldg(A.normalizer);
    const float inv_n = 1.0f / nrm;
    const float gneg = (1.0f - A.p.alpha) * inv_n;
    float4 *buf = reinterpret_cast<float4 *>(s_tiles);
    constexpr unsigned kTileBytes = kTmaTileF4 * 16;
    const unsigned total = A.tma_tile_off[A.g.nlev];
    const unsigned mine = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto locate = [&](unsigned i, int &l, size_t &elem) {   // tile i of this CTA -> level, first element
        const unsigned g = blockIdx.x + i * gridDim.x;
        l = 0;
#pragma unroll
        for (int k = 1; k < ODK_MAX_LEVELS; ++k)
            if (k < A.g.nlev && g >= A.tma_tile_off[k]) l = k;
        elem = (size_t)(g - A.tma_tile_off[l]) * kTmaTileF4 * 4;
    };
    auto load = [&](unsigned i) {   // thread 0
        int l;
        size_t elem;
        locate(i, l, elem);
        const int s = (int)(i % kTmaStages);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_full[s])), "r"(kTileBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(buf + (size_t)s * kTmaTileF4)),
                     "l"(A.cls[l] + elem), "r"(kTileBytes), "r"(smem_u32(&s_full[s]))
                     : "memory");
    };
    if (threadIdx.x == 0)
        for (unsigned i = 0; i + 1 < (unsigned)kTmaStages && i < mine; ++i) load(i);
    double dsum = 0.0;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, accx[4] = {0.f, 0.f, 0.f, 0.f};
    auto flush = [&]() {   // fp32 partial sums stay short
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { v += (MODE == kNewSmooth ? acc[j] - 0.5f * A.p.label_smoothing * accx[j] : acc[j]); acc[j] = 0.f; accx[j] = 0.f; }
        dsum += (double)v;
    };
    for (unsigned i = 0; i < mine; ++i) {
        const int s = (int)(i % kTmaStages);
        const unsigned parity = (i / kTmaStages) & 1u;
        if (threadIdx.x == 0 && i + kTmaStages - 1 < mine) {
            // the stage that held tile i-1 takes tile i+STAGES-1: its store must have finished READING shared memory
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            load(i + kTmaStages - 1);
        }
        unsigned ready = 0;
        while (!ready)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ready) : "r"(smem_u32(&s_full[s])), "r"(parity) : "memory");
        float4 *t = buf + (size_t)s * kTmaTileF4;
#pragma unroll
        for (int j = 0; j < kTmaTileF4 / kLossThreads; ++j) {
            const float4 q = t[threadIdx.x + j * kLossThreads];
            Vec<4> x, gr;
            x.v[0] = q.x; x.v[1] = q.y; x.v[2] = q.z; x.v[3] = q.w;
            row_values<4, MODE, true>(A, x, gr, gneg, acc, accx);
            t[threadIdx.x + j * kLossThreads] = make_float4(gr.v[0], gr.v[1], gr.v[2], gr.v[3]);
        }
        if ((i & 15u) == 15u) flush();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy writes above, before the bulk store reads them
        __syncthreads();
        if (threadIdx.x == 0) {
            int l;
            size_t elem;
            locate(i, l, elem);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A.gcls[l] + elem), "r"(smem_u32(t)), "r"(kTileBytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    flush();
    // what the tiles do not cover
    const size_t stride = (size_t)gridDim.x * kLossThreads;
    const size_t first = (size_t)blockIdx.x * kLossThreads + threadIdx.x;
    for (int l = 0; l < A.g.nlev; ++l) {
        const size_t len = (size_t)A.B * A.g.na * A.C * A.g.hw[l];
        const float *px = A.cls[l];
        float *pg = A.gcls[l];
        const size_t covered = (size_t)(A.tma_tile_off[l + 1] - A.tma_tile_off[l]) * kTmaTileF4 * 4;   // elements (0 if unaligned)
        const bool vec4 = (((uintptr_t)px | (uintptr_t)pg) & 15) == 0;
        const size_t n4 = vec4 ? len / 4 : 0;
        for (size_t u = covered / 4 + first; u < n4; u += stride) {
            Vec<4> x0;
            x0.load_stream(px + u * 4);
            row_compute<4, MODE, true>(A, x0, pg + u * 4, gneg, acc, accx);
        }
        flush();
        float t1[1] = {0.f}, t1x[1] = {0.f};
        for (size_t e = (vec4 ? n4 * 4 : 0) + first; e < len; e += stride) {
            Vec<1> x0;
            x0.load_stream(px + e);
            row_compute<1, MODE, true>(A, x0, pg + e, gneg, t1, t1x);
        }
        dsum += (double)(MODE == kNewSmooth ? t1[0] - 0.5f * A.p.label_smoothing * t1x[0] : t1[0]);
        float *gb = A.gbox[l];   // box gradients: zero everywhere, the patch kernel writes the matched anchors'
        const size_t blen = (size_t)A.B * A.g.na * 4 * A.g.hw[l];
        if (((uintptr_t)gb & 15) == 0) {
            for (size_t q = first; q < blen / 4; q += stride) st_stream4(gb + q * 4, make_float4(0.f, 0.f, 0.f, 0.f));
            for (size_t q = blen / 4 * 4 + first; q < blen; q += stride) gb[q] = 0.f;
        } else {
            for (size_t q = first; q < blen; q += stride) gb[q] = 0.f;
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the stores are complete before the CTA reports in
    finish_block(A, (float)0.f, 0.f, nrm, (1.0 - (double)A.p.alpha) * dsum);
}

// One matched anchor: the positive class' term / gradient corrected, the Huber loss / gradient of its four codes.
// The two sums are accumulated in 2^-32 fixed point: integer addition is associative, so the result does not
// depend on which thread takes which anchor -- the labeler's list comes in atomic-append order -- and the loss
// stays bit-reproducible from run to run (quantisation 2.3e-10 per anchor against sums of 1e2..1e5).
__device__ __forceinline__ long long fix32(float v) { return __float2ll_rn(v * 4294967296.0f); }

template <int MODE, bool GRAD>
__device__ __forceinline__ void patch_anchor(const LossArgs &A, int b, int p, int mt, float inv_n, long long &ci, long long &bi) {
    const Geo &g = A.g;
    const int l = geo_level(g, p);
    const int hw = g.hw[l], loc = p - g.off[l];
    const int a = loc / hw, pos = loc - a * hw;
    const int tc = __ldg(A.gt_labels + (size_t)b * A.Mmax + mt) - 1;
    const float alpha = A.p.alpha, sm = A.p.label_smoothing, gamma = A.p.gamma;
    if (tc >= 0 && tc < A.C) {
        const size_t o = A.cls_nhwc[l] ? ((size_t)b * hw + pos) * (size_t)(g.na * A.C) + (size_t)a * A.C + tc
                                       : ((size_t)(b * g.na + a) * A.C + tc) * hw + pos;
        const float xp = __ldg(A.cls[l] + o);
        float gpos, dc;
        if (MODE == kLegacy) {
            const float sp = fmaxf(xp, 0.f) + log1pf(expf(-fabsf(xp)));
            const float sg = 1.0f / (1.0f + expf(-xp));
            const float mod_n = expf(-gamma * (sp - xp)), mod_p = expf(-gamma * sp);
            const float bce_p = sp - xp;
            dc = alpha * mod_p * bce_p - (1.0f - alpha) * mod_n * sp;
            gpos = alpha * inv_n * mod_p * ((sg - 1.0f) - bce_p * gamma * sg);
        } else {
            float e;
            const float sp = softplus_fast(xp, e);
            const float tp = MODE == kNewSmooth ? 1.0f - 0.5f * sm : 1.0f;
            const float tn = MODE == kNewSmooth ? 0.5f * sm : 0.0f;
            dc = alpha * (sp - tp * xp) - (1.0f - alpha) * (sp - tn * xp);
            gpos = alpha * inv_n * (sigmoid_from_e(xp, e) - tp);
        }
        ci += fix32(dc);
        if (GRAD) A.gcls[l][o] = gpos;
    }
    const float4 t4 = encode_ref(__ldg(A.gt_boxes + (size_t)b * A.Mmax + mt), __ldg(A.anchors + g.off[l] + pos * g.na + a));
    const float tg[4] = {t4.x, t4.y, t4.z, t4.w};
    const float delta = A.p.delta, gb = A.p.box_loss_weight * inv_n * 0.25f;
    float db = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t o = A.box_nhwc[l] ? ((size_t)b * hw + pos) * (size_t)(g.na * 4) + (size_t)a * 4 + k
                                       : ((size_t)(b * g.na + a) * 4 + k) * hw + pos;
        const float tv = tg[k];
        const float e = __ldg(A.box[l] + o) - tv;          // loss.py:108-112
        const float ae = fabsf(e), qd = fminf(ae, delta);
        const bool on = tv != 0.0f;                        // loss.py:177
        db += on ? 0.5f * qd * qd + delta * (ae - qd) : 0.f;
        if (GRAD) A.gbox[l][o] = on ? gb * (ae <= delta ? e : copysignf(delta, e)) : 0.f;
    }
    bi += fix32(db);
}

// Work units: (image, slice).  With the labeler's list (clr_in_patch) a unit walks every patch_slices-th run of 256
// list entries and zeroes the keys it has used (ODK_ASSIGN_WS_CLEAN for the next step, no extra launch); without
// it -- int32 `match`, foreign keys, or an image with more positives than the list holds -- a unit scans a slice
// of the image's key / match row.
template <int MODE, bool GRAD>
__global__ void __launch_bounds__(kLossThreads)
loss_patch_kernel(const __grid_constant__ LossArgs A) {
    const Geo &g = A.g;
    const float nrm = __ldg(A.normalizer);
    const float inv_n = 1.0f / nrm;
    long long ci = 0, bi = 0;
    const int S = A.patch_slices;
    const long long units = (long long)A.B * S;
    const bool key64 = A.p.match_is_key64 != 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int b = (int)(u / S), slice = (int)(u - (long long)b * S);
        const int n = A.clr_in_patch ? A.clr_pos[(size_t)b * kCtrStride] : 0;
        if (A.clr_in_patch && n <= A.clr_cap) {
            unsigned long long *row = A.clr_keys + (size_t)b * g.Apad;
            for (int i = slice * kLossThreads + threadIdx.x; i < n; i += S * kLossThreads) {
                const int p = (int)__ldg(A.clr_touched + (size_t)b * A.clr_cap + i);
                const int mt = match_of_key(row[p]);
                row[p] = 0ull;
                if (mt >= 0) patch_anchor<MODE, GRAD>(A, b, p, mt, inv_n, ci, bi);
            }
        } else {
            for (int p = slice * kLossThreads + threadIdx.x; p < g.A; p += S * kLossThreads) {
                const size_t mi = (size_t)b * g.Apad + p;
                int mt;
                if (key64) {
                    const unsigned long long key = __ldcg(reinterpret_cast<const unsigned long long *>(A.match) + mi);
                    mt = match_of_key(key);
                    if (A.clr_in_patch && key) A.clr_keys[mi] = 0ull;
                } else {
                    mt = __ldg(A.match + mi);
                }
                if (mt >= 0) patch_anchor<MODE, GRAD>(A, b, p, mt, inv_n, ci, bi);
            }
        }
    }
    finish_block(A, 0.f, 0.f, nrm, 0.0, true, ci, bi);
}

static int sm_count() { return device_sm_count(); }

static int grid_for(long long items, int blocks_per_sm) {
    long long need = (items + kLossThreads - 1) / kLossThreads;
    long long grid = (long long)sm_count() * blocks_per_sm;
    if (grid > kMaxPartials / 2) grid = kMaxPartials / 2;
    if (grid > need) grid = need;
    return (int)(grid < 1 ? 1 : grid);
}

template <int MODE, bool GRAD, bool FUSED>
static int launch_loss(LossArgs &ring, LossArgs &plain, cudaStream_t st) {
    static int occ_ring_dev[kMaxDevices] = {0}, occ_plain_dev[kMaxDevices] = {0};   // per device: the opt-in is per device too
    const int dev = current_device();
    if (!occ_ring_dev[dev]) {
        int r = 0, pl = 0;
        cudaFuncSetAttribute(loss_kernel_ring<MODE, GRAD, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes(GRAD));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, loss_kernel_ring<MODE, GRAD, FUSED>, kLossThreads, ring_bytes(GRAD));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pl, loss_kernel<MODE, GRAD, FUSED>, kLossThreads, 0);
        occ_plain_dev[dev] = pl < 1 ? 1 : pl;
        occ_ring_dev[dev] = r < 1 ? 1 : r;
    }
    const int occ_ring = occ_ring_dev[dev], occ_plain = occ_plain_dev[dev];
    const long long n_ring = ring.item_off[ring.g.nlev], n_plain = plain.item_off[plain.g.nlev];
    // The finishing launch sums every partial slot; the other one (if any) runs first on the stream.
    int g_plain = 0, g_ring = 0;
    if (n_plain > 0) g_plain = grid_for(n_plain, occ_plain);
    if (n_ring > 0) g_ring = grid_for(n_ring, occ_ring);
    if (n_plain > 0) {
        plain.part_base = 0;
        plain.part_total = n_ring > 0 ? 0 : g_plain;
        loss_kernel<MODE, GRAD, FUSED><<<g_plain, kLossThreads, 0, st>>>(plain);
        int rc = check_launch("odk_loss/loss_kernel");
        if (rc) return rc;
        if (n_ring > 0) {   // the ring launch starts from a zero counter again
            cudaError_t e = cudaMemsetAsync(plain.counter, 0, sizeof(unsigned), st);
            if (e != cudaSuccess) return set_error((int)e, "odk_loss memset: %s", cudaGetErrorString(e));
        }
    }
    if (n_ring > 0) {
        ring.part_base = g_plain;
        ring.part_total = g_plain + g_ring;
        loss_kernel_ring<MODE, GRAD, FUSED><<<g_ring, kLossThreads, ring_bytes(GRAD), st>>>(ring);
        return check_launch("odk_loss/loss_kernel_ring");
    }
    return ODK_OK;
}

}  // namespace odk

extern "C" {

size_t odk_loss_workspace_bytes(void) { return (size_t)odk::kMaxPartials * 2 * sizeof(double) + 256; }

int odk_loss(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
             int num_levels, int na, const int32_t *match, const float *anchors, const float *gt_boxes,
             const int32_t *gt_labels, int Mmax, const int64_t *cls_targets, const float *box_targets,
             const float *normalizer, const odk_loss_params *params, float *out, void *const *grad_cls_levels,
             void *const *grad_box_levels, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace odk;
    LossArgs a;
    memset(&a, 0, sizeof(a));
    int rc = make_geo(&a.g, level_hw, num_levels, na);
    if (rc) return rc;
    if (!cls_levels || !box_levels || !normalizer || !params || !out)
        return set_error(ODK_EINVAL, "odk_loss: null pointer");
    if (B < 1 || C < 1) return set_error(ODK_EINVAL, "odk_loss: B and C must be positive");
    if ((size_t)B * na * (size_t)C > 0x7fffffffull) return set_error(ODK_EUNSUPPORTED, "odk_loss: B*na*C overflows int");
    const bool fused = match != nullptr;
    if (fused) {
        if (!anchors || (Mmax > 0 && (!gt_boxes || !gt_labels))) return set_error(ODK_EINVAL, "odk_loss: fused targets need anchors/gt");
        if (((uintptr_t)anchors | (uintptr_t)gt_boxes) & 15) return set_error(ODK_EINVAL, "odk_loss: anchors/gt_boxes must be 16-byte aligned");
    } else {
        if (!cls_targets || !box_targets) return set_error(ODK_EINVAL, "odk_loss: need match or cls_targets+box_targets");
        if ((uintptr_t)box_targets & 15) return set_error(ODK_EINVAL, "odk_loss: box_targets must be 16-byte aligned");
    }
    const bool grad = grad_cls_levels != nullptr || grad_box_levels != nullptr;
    if (grad && (!grad_cls_levels || !grad_box_levels)) return set_error(ODK_EINVAL, "odk_loss: need both gradient arrays");
    if (!workspace || workspace_bytes < odk_loss_workspace_bytes())
        return set_error(ODK_EWORKSPACE, "odk_loss: workspace too small (%zu < %zu)", workspace_bytes, odk_loss_workspace_bytes());
    if ((uintptr_t)workspace & 15) return set_error(ODK_EINVAL, "odk_loss: workspace must be 16-byte aligned");

    a.B = B; a.C = C; a.Mmax = Mmax;
    // class chunks: as few as possible with <= kMaxChunk classes each, so the per-item work (index
    // decode, target lookup, positive patch) is amortised over up to 192 elements per thread
    a.nchunk = (C + kMaxChunk - 1) / kMaxChunk;
    a.cchunk = (C + a.nchunk - 1) / a.nchunk;
    a.nchunk = (C + a.cchunk - 1) / a.cchunk;
    a.div_nchunk = make_fastdiv((unsigned)a.nchunk);
    a.div_na = make_fastdiv((unsigned)na);
    for (int l = 0; l < num_levels; ++l) {
        a.cls[l] = (const float *)cls_levels[l];
        a.box[l] = (const float *)box_levels[l];
        a.gcls[l] = grad ? (float *)grad_cls_levels[l] : nullptr;
        a.gbox[l] = grad ? (float *)grad_box_levels[l] : nullptr;
        if (!a.cls[l] || !a.box[l] || (grad && (!a.gcls[l] || !a.gbox[l])))
            return set_error(ODK_EINVAL, "odk_loss: null level pointer (level %d)", l);
        uintptr_t al = (uintptr_t)a.cls[l] | (uintptr_t)a.box[l] | (uintptr_t)a.gcls[l] | (uintptr_t)a.gbox[l];
        a.vec[l] = (a.g.hw[l] % 4 == 0 && (al & 15) == 0) ? 4 : 1;
        a.nq[l] = (a.g.hw[l] + a.vec[l] - 1) / a.vec[l];
        a.div_nq[l] = make_fastdiv((unsigned)a.nq[l]);
    }
    a.match = match; a.anchors = (const float4 *)anchors; a.gt_boxes = (const float4 *)gt_boxes; a.gt_labels = gt_labels;
    a.cls_t = cls_targets; a.box_t = box_targets; a.normalizer = normalizer; a.p = *params; a.out = out;
    a.xworld = 0;
    if (params->exchange) {
        const odk_exchange *x = params->exchange;
        if (x->world < 1 || x->world > ODK_MAILBOX_MAX_WORLD || x->rank < 0 || x->rank >= x->world)
            return set_error(ODK_EINVAL, "odk_loss: exchange needs 1 <= world <= %d and 0 <= rank < world", ODK_MAILBOX_MAX_WORLD);
        if (!x->num_pos_plus_1 || !x->global_out3 || !x->status) return set_error(ODK_EINVAL, "odk_loss: exchange has a null pointer");
        for (int r = 0; r < x->world; ++r) {
            if (!x->mailboxes[r] || ((uintptr_t)x->mailboxes[r] & 15)) return set_error(ODK_EINVAL, "odk_loss: bad mailbox pointer %d", r);
            a.xmb.p[r] = (unsigned char *)x->mailboxes[r];
        }
        a.xworld = x->world; a.xrank = x->rank; a.xnpos = x->num_pos_plus_1; a.xout = x->global_out3; a.xstatus = x->status;
        a.xnormalized = x->normalized; a.xtimeout_ms = x->timeout_ms;
    }
    a.clr_cap = 0;
    if (params->clear_keys) {
        if (!fused || !params->match_is_key64) return set_error(ODK_EINVAL, "odk_loss: clear_keys needs odk_assign_grid's 64-bit keys");
        const AssignGridWs w = assign_grid_layout(B, a.g.Apad);
        char *aws = (char *)match;   // the keys are the head of the labeler's workspace
        a.clr_keys = (unsigned long long *)aws; a.clr_pos = (int32_t *)(aws + w.pos);
        a.clr_touched = (const unsigned *)(aws + w.touched); a.clr_done = (unsigned *)(aws + w.done);
        a.clr_cap = w.touched_cap;
    }
    a.partials = (double *)workspace;
    a.counter = (unsigned *)((char *)workspace + (size_t)kMaxPartials * 2 * sizeof(double));

    // Early key loads need every item to span more row-groups than the ring is deep: the load cursor
    // must not enter item n+2 before the math has started item n+1 (one set of key registers).
    a.keys_early = (fused && C - (a.nchunk - 1) * a.cchunk > 4 * kMaxRingDepth) ? 1 : 0;

    // split the levels between the two kernels: item_off counts only the levels a launch covers
    LossArgs ring = a, plain = a;
    long long off_r = 0, off_p = 0;
    for (int l = 0; l < num_levels; ++l) {
        const long long items = (long long)B * na * a.nchunk * a.nq[l];
        ring.item_off[l] = (unsigned)off_r;
        plain.item_off[l] = (unsigned)off_p;
        if (a.vec[l] == 4) off_r += items; else off_p += items;
        if (off_r > 0x7fffffffll || off_p > 0x7fffffffll) return set_error(ODK_EUNSUPPORTED, "odk_loss: more than 2^31 work items");
    }
    for (int l = num_levels; l <= ODK_MAX_LEVELS; ++l) { ring.item_off[l] = (unsigned)off_r; plain.item_off[l] = (unsigned)off_p; }

    cudaStream_t st = (cudaStream_t)stream;
    int any_nhwc = 0;
    for (int l = 0; l < num_levels; ++l) {
        a.cls_nhwc[l] = (params->layout >> l) & 1;
        a.box_nhwc[l] = (params->layout >> (8 + l)) & 1;
        any_nhwc |= a.cls_nhwc[l] | a.box_nhwc[l];
    }
    // Targets from the labeler's match: a layout-agnostic stream over the logits + a patch of the matched anchors
    // (any layout).  ODK_LOSS_KERNEL=ring selects the plane-walking kernels below instead (NCHW only; kept for
    // targets given as tensors and for A/B measurements: fwd 190 us / fwd+grad 407 us against 170 + 5 / 378 + 5).
    const char *force = getenv("ODK_LOSS_KERNEL");
    const bool stream_path = fused && (any_nhwc || !(force && strcmp(force, "ring") == 0));
    if (any_nhwc && !fused)
        return set_error(ODK_EUNSUPPORTED, "odk_loss: channels_last inputs need the labeler's match (targets given as tensors: pass NCHW)");
    if (stream_path) {
        cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned), st);
        if (e != cudaSuccess) return set_error((int)e, "odk_loss memset: %s", cudaGetErrorString(e));
        const int mode = params->legacy_focal ? kLegacy : (params->label_smoothing > 0.0f ? kNewSmooth : kNew);
        LossArgs flat = a, patch = a;
        const int g_flat = sm_count() * 4;
        flat.part_base = 0; flat.part_total = 0;
        patch.clr_in_patch = a.clr_cap > 0 ? 1 : 0;
        // slices per image: ~32 matched anchors per gt box when the list is walked, 1024 anchors per slice when the row is scanned
        int S = patch.clr_in_patch ? (Mmax * 32 + kLossThreads - 1) / kLossThreads : (int)((a.g.A + 1023) / 1024);
        if (patch.clr_in_patch && S > 16) S = 16;
        if (S < 1) S = 1;
        patch.patch_slices = S;
        const long long units = (long long)B * S;
        const int g_patch = (int)(units < 1024 ? units : 1024);
        patch.part_base = g_flat; patch.part_total = g_flat + g_patch; patch.part_int_from = g_flat;
        // gradient pass: tiles through TMA bulk copies (ODK_LOSS_TMA=0 keeps the LDG/STG stream); same partial-slot count
        bool tma = grad;
        {
            const char *et = getenv("ODK_LOSS_TMA");
            if (et && et[0] == '0') tma = false;
        }
        if (tma) {
            unsigned long long tiles = 0;
            for (int l = 0; l < num_levels; ++l) {
                flat.tma_tile_off[l] = (unsigned)tiles;
                const bool aligned = (((uintptr_t)a.cls[l] | (uintptr_t)a.gcls[l]) & 15) == 0;
                if (aligned) tiles += (unsigned long long)B * na * C * a.g.hw[l] / ((unsigned long long)kTmaTileF4 * 4);
            }
            for (int l = num_levels; l <= ODK_MAX_LEVELS; ++l) flat.tma_tile_off[l] = (unsigned)tiles;
            if (tiles == 0 || tiles > 0xffffffffull) tma = false;
        }
        const int g_tma = sm_count() * 2;
        if (tma) { patch.part_base = g_tma; patch.part_total = g_tma + g_patch; patch.part_int_from = g_tma; }
        // Forward only: the patch writes nothing the stream touches, so it runs on a forked stream BESIDE it (~10 us of
        // dependent loads off the critical path); whichever CTA of the two launches reports last finishes the reduction.
        // (With gradients the patch overwrites elements the stream has written and must follow it.)
        const bool beside = !grad;
        SideLane lane;
        if (beside) {
            flat.part_total = patch.part_total; flat.part_int_from = patch.part_int_from;
            flat.finish_ctas = patch.finish_ctas = g_flat + g_patch;
            flat.clr_in_patch = patch.clr_in_patch;   // (only the finishing code looks at it in the stream kernel)
            rc = fork_side(st, &lane);
            if (rc) return rc;
        }
        cudaStream_t st_patch = beside ? lane.side : st;
        rc = -1;
#define ODK_STREAM_CASE(M)                                                                 \
        if (mode == M) {                                                                   \
            if (tma) {                                                                     \
                e = cudaFuncSetAttribute(loss_flat_tma_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes); \
                if (e != cudaSuccess) return set_error((int)e, "odk_loss: %zu bytes of shared memory: %s", kTmaSmemBytes, cudaGetErrorString(e)); \
                loss_flat_tma_kernel<M><<<g_tma, kLossThreads, kTmaSmemBytes, st>>>(flat); \
            }                                                                              \
            else if (grad) loss_flat_kernel<M, true><<<g_flat, kLossThreads, 0, st>>>(flat);    \
            else loss_flat_kernel<M, false><<<g_flat, kLossThreads, 0, st>>>(flat);        \
            rc = check_launch("odk_loss/loss_flat_kernel");                                \
            if (rc) return rc;                                                             \
            if (grad) loss_patch_kernel<M, true><<<g_patch, kLossThreads, 0, st_patch>>>(patch); \
            else loss_patch_kernel<M, false><<<g_patch, kLossThreads, 0, st_patch>>>(patch);     \
            rc = check_launch("odk_loss/loss_patch_kernel");                               \
        }
        ODK_STREAM_CASE(kNew)
        ODK_STREAM_CASE(kNewSmooth)
        ODK_STREAM_CASE(kLegacy)
#undef ODK_STREAM_CASE
        if (rc < 0) return set_error(ODK_EINVAL, "odk_loss: bad mode");
        if (beside) {
            const int rj = join_side(st, lane);
            if (!rc) rc = rj;
        }
        return rc;
    }
    // the counter must be zero on entry; the kernel re-zeroes it, but a fresh workspace is arbitrary
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return set_error((int)e, "odk_loss memset: %s", cudaGetErrorString(e));
    const int mode = params->legacy_focal ? kLegacy : (params->label_smoothing > 0.0f ? kNewSmooth : kNew);
    rc = -1;
#define ODK_LOSS_CASE(M)                                                                                       \
    if (mode == M) {                                                                                           \
        if (grad) rc = fused ? launch_loss<M, true, true>(ring, plain, st) : launch_loss<M, true, false>(ring, plain, st);   \
        else rc = fused ? launch_loss<M, false, true>(ring, plain, st) : launch_loss<M, false, false>(ring, plain, st);      \
    }
    ODK_LOSS_CASE(kNew)
    ODK_LOSS_CASE(kNewSmooth)
    ODK_LOSS_CASE(kLegacy)
#undef ODK_LOSS_CASE
    if (rc < 0) return set_error(ODK_EINVAL, "odk_loss: bad mode");
    if (rc == ODK_OK && a.clr_cap > 0) {
        clear_keys_kernel<<<B, 256, 0, st>>>(a.clr_keys, a.clr_pos, a.clr_touched, a.clr_done, a.g.Apad, a.clr_cap);
        rc = check_launch("odk_loss/clear_keys_kernel");
    }
    return rc;
}

}  // extern "C"
