// Fused detection loss (K2): focal/BCE classification loss + Huber box loss over all pyramid
// levels, forward and (optionally) the gradient of the total in the same pass.  HBM-bound: the
// [B, na*C, H, W] logits are streamed exactly once with 128-bit loads in their native NCHW
// layout; the one-hot target tensor the reference materialises (loss.py:182-186) never exists.
// See include/odk.h (odk_loss).
//
// Work decomposition: one item = (level, image b, anchor shape a, class chunk, 4 consecutive
// (y,x) positions).  Consecutive threads take consecutive position groups of one channel
// plane, so every warp load is 512 contiguous bytes.  The inner loop treats every element as
// a negative (target 0); the single positive class of a position (if any) is patched after
// the loop.  A persistent grid (a multiple of the SM count) walks the items; per-CTA partial
// sums are combined by the last CTA in a fixed order, so results are deterministic.
#include <string.h>
#include "odk_common.cuh"

namespace odk {

constexpr int kLossThreads = 256;
constexpr int kMaxPartials = 148 * 16;

enum LossMode { kNew = 0, kNewSmooth = 1, kLegacy = 2 };

struct LossArgs {
    Geo g;
    const float *cls[ODK_MAX_LEVELS];
    const float *box[ODK_MAX_LEVELS];
    float *gcls[ODK_MAX_LEVELS];
    float *gbox[ODK_MAX_LEVELS];
    unsigned item_off[ODK_MAX_LEVELS + 1];
    FastDiv div_nq[ODK_MAX_LEVELS];
    FastDiv div_nchunk, div_na;
    int vec[ODK_MAX_LEVELS];   // 4 if the level's planes are 16-byte aligned rows of 4, else 1
    int nq[ODK_MAX_LEVELS];    // position groups per plane
    int B, C, cchunk, nchunk, Mmax;
    const int32_t *match;
    const float4 *anchors;
    const float4 *gt_boxes;
    const int32_t *gt_labels;
    const int64_t *cls_t;
    const float *box_t;
    const float *normalizer;
    odk_loss_params p;
    double *partials;     // [gridDim.x][2]
    unsigned *counter;
    float *out;
};

// softplus(x) = max(x,0) + log1p(exp(-|x|)); also returns e = exp(-|x|).
// MUFU.EX2 + MUFU.LG2 with a 4-term series where 1+e would lose e's low bits.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float softplus_fast(float x, float &e) {
    e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
    const float series = e * (1.0f - e * (0.5f - e * (0.33333334f - 0.25f * e)));
    const float lg = lg2_ftz(1.0f + e) * 0.6931471805599453f;
    return fmaxf(x, 0.0f) + (e < 0.03125f ? series : lg);
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

template <int VEC, int MODE, bool GRAD, bool FUSED>
__device__ __forceinline__ void loss_item(const LossArgs &A, int l, unsigned local, float inv_n, float &csum,
                                          float &bsum) {
    const Geo &g = A.g;
    const int hw = g.hw[l], nq = A.nq[l];
    unsigned t = fd_div(local, A.div_nq[l]);
    const int q = (int)(local - t * (unsigned)nq);
    unsigned t2 = fd_div(t, A.div_nchunk);
    const int chunk = (int)(t - t2 * (unsigned)A.nchunk);
    const unsigned t3 = fd_div(t2, A.div_na);
    const int a = (int)(t2 - t3 * (unsigned)g.na);
    const int b = (int)t3;
    const int s0 = q * VEC;
    const int c0 = chunk * A.cchunk, c1 = min(c0 + A.cchunk, A.C);
    const float alpha = A.p.alpha, sm = A.p.label_smoothing, gamma = A.p.gamma;

    // ---- class target of each of my positions: >=0 class, -1 background, -2 ignore ----
    int tc[VEC], mt[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        if (FUSED) {
            mt[j] = __ldg(A.match + (size_t)b * g.Apad + g.off[l] + a * hw + s0 + j);
            tc[j] = mt[j] >= 0 ? __ldg(A.gt_labels + (size_t)b * A.Mmax + mt[j]) - 1 : -1;
        } else {
            mt[j] = -1;
            tc[j] = (int)__ldg(A.cls_t + (size_t)A.B * g.off[l] + ((size_t)b * hw + s0 + j) * g.na + a);
        }
    }

    // ---- stream my class chunk as negatives ----
    const size_t plane0 = ((size_t)(b * g.na + a) * A.C + c0) * hw + s0;
    const float *px = A.cls[l] + plane0;
    float *pg = GRAD ? A.gcls[l] + plane0 : nullptr;
    float acc[VEC], accx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { acc[j] = 0.f; accx[j] = 0.f; }
    const float gneg = (1.0f - alpha) * inv_n;
    auto one_plane = [&](const Vec<VEC> &x, float *gp) {
        Vec<VEC> gr;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float e;
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
                const float sp = softplus_fast(x.v[j], e);
                acc[j] += sp;
                if (MODE == kNewSmooth) accx[j] += x.v[j];
                if (GRAD) {
                    const float sg = sigmoid_from_e(x.v[j], e);
                    gr.v[j] = gneg * (MODE == kNewSmooth ? sg - 0.5f * sm : sg);
                }
            }
        }
        if (GRAD) gr.store(gp);
    };
    int c = c0;
    for (; c + 4 <= c1; c += 4) {   // 4 independent 128-bit loads in flight per thread
        Vec<VEC> x0, x1, x2, x3;
        x0.load_stream(px);
        x1.load_stream(px + hw);
        x2.load_stream(px + 2 * (size_t)hw);
        x3.load_stream(px + 3 * (size_t)hw);
        one_plane(x0, pg);
        one_plane(x1, pg + (GRAD ? hw : 0));
        one_plane(x2, pg + (GRAD ? 2 * (size_t)hw : 0));
        one_plane(x3, pg + (GRAD ? 3 * (size_t)hw : 0));
        px += 4 * (size_t)hw;
        if (GRAD) pg += 4 * (size_t)hw;
    }
    for (; c < c1; ++c) {
        Vec<VEC> x0;
        x0.load_stream(px);
        one_plane(x0, pg);
        px += hw;
        if (GRAD) pg += hw;
    }

    // ---- combine, patch the positive class, apply the ignore mask (loss.py:285) ----
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float val = (1.0f - alpha) * (MODE == kNewSmooth ? acc[j] - 0.5f * sm * accx[j] : acc[j]);
        if (tc[j] >= c0 && tc[j] < c1) {
            const size_t o = ((size_t)(b * g.na + a) * A.C + tc[j]) * hw + s0 + j;
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
                float *pz = A.gcls[l] + plane0 + j;
                for (int c = c0; c < c1; ++c, pz += hw) *pz = 0.f;
            }
        }
        csum += val;
    }

    // ---- Huber box loss: once per (b, a, positions), by the first class chunk ----
    if (chunk == 0) {
        float tg[VEC][4];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (FUSED) {
                if (mt[j] >= 0)
                    t4 = encode_ref(__ldg(A.gt_boxes + (size_t)b * A.Mmax + mt[j]),
                                    __ldg(A.anchors + g.off[l] + (s0 + j) * g.na + a));
            } else {
                t4 = __ldg(reinterpret_cast<const float4 *>(A.box_t) + (size_t)A.B * g.off[l] +
                           ((size_t)b * hw + s0 + j) * g.na + a);
            }
            tg[j][0] = t4.x; tg[j][1] = t4.y; tg[j][2] = t4.z; tg[j][3] = t4.w;
        }
        const float delta = A.p.delta;
        const float gb = A.p.box_loss_weight * inv_n * 0.25f;
        const size_t bplane = ((size_t)(b * g.na + a) * 4) * hw + s0;
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

template <int MODE, bool GRAD, bool FUSED>
__global__ void __launch_bounds__(kLossThreads, 3)
loss_kernel(const __grid_constant__ LossArgs A) {
    const float nrm = __ldg(A.normalizer);
    const float inv_n = 1.0f / nrm;
    float csum = 0.f, bsum = 0.f;
    const unsigned total = A.item_off[A.g.nlev];
    const unsigned stride = gridDim.x * kLossThreads;
    for (unsigned it = blockIdx.x * kLossThreads + threadIdx.x; it < total; it += stride) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < ODK_MAX_LEVELS; ++i)
            if (i < A.g.nlev && it >= A.item_off[i]) l = i;
        const unsigned local = it - A.item_off[l];
        if (A.vec[l] == 4) loss_item<4, MODE, GRAD, FUSED>(A, l, local, inv_n, csum, bsum);
        else loss_item<1, MODE, GRAD, FUSED>(A, l, local, inv_n, csum, bsum);
    }

    // ---- deterministic two-stage reduction ----
    __shared__ double s_c[kLossThreads / 32], s_b[kLossThreads / 32];
    __shared__ bool s_last;
    double dc = warp_sum((double)csum), db = warp_sum((double)bsum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_c[warp] = dc; s_b[warp] = db; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0, bx = 0;
        for (int w = 0; w < kLossThreads / 32; ++w) { c += s_c[w]; bx += s_b[w]; }
        A.partials[2 * blockIdx.x] = c;
        A.partials[2 * blockIdx.x + 1] = bx;
        __threadfence();
        const unsigned done = atomicAdd(A.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double c = 0, bx = 0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kLossThreads) {
            c += __ldcg(A.partials + 2 * i);
            bx += __ldcg(A.partials + 2 * i + 1);
        }
        c = warp_sum(c); bx = warp_sum(bx);
        if (lane == 0) { s_c[warp] = c; s_b[warp] = bx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            c = 0; bx = 0;
            for (int w = 0; w < kLossThreads / 32; ++w) { c += s_c[w]; bx += s_b[w]; }
            const double n = (double)nrm;
            const float cls_loss = (float)(c / n);
            const float box_loss = (float)(bx / (n * 4.0));                 // loss.py:176-179
            A.out[1] = cls_loss;
            A.out[2] = box_loss;
            A.out[0] = cls_loss + A.p.box_loss_weight * box_loss;         // loss.py:297
            *A.counter = 0;                                                 // ready for the next call
        }
    }
}

template <int MODE, bool GRAD, bool FUSED>
static int launch_loss(const LossArgs &args, int *grid_out, cudaStream_t st) {
    static int blocks_per_sm = 0, sms = 0;
    if (!blocks_per_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, loss_kernel<MODE, GRAD, FUSED>, kLossThreads, 0);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        if (sms < 1) sms = 148;
    }
    const long long total = args.item_off[args.g.nlev];
    long long need = (total + kLossThreads - 1) / kLossThreads;
    long long grid = (long long)sms * blocks_per_sm;
    if (grid > kMaxPartials) grid = kMaxPartials;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    *grid_out = (int)grid;
    loss_kernel<MODE, GRAD, FUSED><<<(unsigned)grid, kLossThreads, 0, st>>>(args);
    return check_launch("odk_loss/loss_kernel");
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
    a.cchunk = C < 16 ? C : 16;
    a.nchunk = (C + a.cchunk - 1) / a.cchunk;
    long long off = 0;
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
        a.item_off[l] = (unsigned)off;
        off += (long long)B * na * a.nchunk * a.nq[l];
        if (off > 0x7fffffffll) return set_error(ODK_EUNSUPPORTED, "odk_loss: more than 2^31 work items");
    }
    for (int l = num_levels; l <= ODK_MAX_LEVELS; ++l) a.item_off[l] = (unsigned)off;
    a.match = match; a.anchors = (const float4 *)anchors; a.gt_boxes = (const float4 *)gt_boxes; a.gt_labels = gt_labels;
    a.cls_t = cls_targets; a.box_t = box_targets; a.normalizer = normalizer; a.p = *params; a.out = out;
    a.partials = (double *)workspace;
    a.counter = (unsigned *)((char *)workspace + (size_t)kMaxPartials * 2 * sizeof(double));

    cudaStream_t st = (cudaStream_t)stream;
    // the counter must be zero on entry; the kernel re-zeroes it, but a fresh workspace is arbitrary
    cudaError_t e = cudaMemsetAsync(a.counter, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return set_error((int)e, "odk_loss memset: %s", cudaGetErrorString(e));
    const int mode = params->legacy_focal ? kLegacy : (params->label_smoothing > 0.0f ? kNewSmooth : kNew);
    int grid = 0;
#define ODK_LOSS_CASE(M)                                                                          \
    if (mode == M) {                                                                              \
        if (grad) return fused ? launch_loss<M, true, true>(a, &grid, st) : launch_loss<M, true, false>(a, &grid, st); \
        return fused ? launch_loss<M, false, true>(a, &grid, st) : launch_loss<M, false, false>(a, &grid, st);         \
    }
    ODK_LOSS_CASE(kNew)
    ODK_LOSS_CASE(kNewSmooth)
    ODK_LOSS_CASE(kLegacy)
#undef ODK_LOSS_CASE
    return set_error(ODK_EINVAL, "odk_loss: bad mode");
}

}  // extern "C"
