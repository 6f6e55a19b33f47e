// Shared host/device helpers for libodk (sm_100a).  See include/odk.h for the ABI.
#pragma once
#include <mutex>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/odk.h"

namespace odk {

// Pyramid geometry, passed by value to kernels (fits the 4 KB param space easily).
struct Geo {
    int nlev;
    int na;                         // anchors per location
    int hw[ODK_MAX_LEVELS];         // H_l * W_l
    int off[ODK_MAX_LEVELS + 1];    // na * sum_{l'<l} hw[l']  (same base in reference and planar order)
    int A;                          // anchors per image
    int Apad;                       // planar per-image stride (A rounded up to 4)
};

int set_error(int code, const char *fmt, ...);
int check_launch(const char *what);
int make_geo(Geo *g, const int32_t *level_hw, int num_levels, int na);

// Workspace of odk_assign_grid (also read by odk_loss when it clears the assignment keys):
//   keys [B][Apad] u64 | pos_count [B][kCtrStride] i32 | touched [B][touched_cap] u32 | done u32 (16 bytes)
// `touched` lists the planar positions whose key is non-zero, in the order the positive counter handed out
// slots; an image with more than touched_cap positives simply is not listed completely (the cleaner then clears
// its whole row).
constexpr int kCtrStride = 32;                   // ints between per-image counters: one 128 B line each
constexpr int kTouchedCapMax = 16384;
struct AssignGridWs { size_t keys, pos, touched, done, total; int touched_cap; };
inline AssignGridWs assign_grid_layout(int B, long long Apad) {
    AssignGridWs w;
    w.touched_cap = (int)(Apad < kTouchedCapMax ? Apad : kTouchedCapMax);
    w.keys = 0;
    w.pos = (size_t)B * (size_t)Apad * sizeof(unsigned long long);
    w.touched = w.pos + (size_t)B * kCtrStride * sizeof(int32_t);
    w.done = w.touched + (((size_t)B * w.touched_cap * sizeof(unsigned)) + 15) / 16 * 16;
    w.total = w.done + 16;
    return w;
}

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

// planar index p (off_l + a*HW + s) -> level, and reference index r (off_l + s*na + a)
__device__ __forceinline__ int geo_level(const Geo &g, int idx) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < ODK_MAX_LEVELS; ++i)
        if (i < g.nlev && idx >= g.off[i]) l = i;
    return l;
}
__device__ __forceinline__ int planar_to_ref(const Geo &g, int p, int &l) {
    l = geo_level(g, p);
    int loc = p - g.off[l];
    int a = loc / g.hw[l];
    int s = loc - a * g.hw[l];
    return g.off[l] + s * g.na + a;
}
__device__ __forceinline__ int ref_to_planar(const Geo &g, int r, int &l) {
    l = geo_level(g, r);
    int loc = r - g.off[l];
    int s = loc / g.na;
    int a = loc - s * g.na;
    return g.off[l] + a * g.hw[l] + s;
}

// IoU exactly as the reference forms it (region_similarity_calculator.py:48-73): separately
// rounded fp32 ops, no FMA contraction, IEEE division.  g/a are yxyx.
__device__ __forceinline__ float iou_ref(float gy0, float gx0, float gy1, float gx1, float garea, float ay0,
                                         float ax0, float ay1, float ax1, float aarea) {
    float h = fmaxf(__fsub_rn(fminf(gy1, ay1), fmaxf(gy0, ay0)), 0.0f);
    float w = fmaxf(__fsub_rn(fminf(gx1, ax1), fmaxf(gx0, ax0)), 0.0f);
    float inter = __fmul_rn(h, w);
    if (inter == 0.0f) return 0.0f;
    float uni = __fsub_rn(__fadd_rn(garea, aarea), inter);
    return __fdiv_rn(inter, uni);
}
__device__ __forceinline__ float area_ref(float y0, float x0, float y1, float x1) {
    return __fmul_rn(__fsub_rn(y1, y0), __fsub_rn(x1, x0));
}

// Faster-RCNN encode of gt box g against anchor a (box_list.py:152-164, box_coder.py:92-110).
__device__ __forceinline__ float4 encode_ref(float4 g, float4 a) {
    const float eps = 1e-8f;
    float wa = __fsub_rn(a.w, a.y), ha = __fsub_rn(a.z, a.x);
    float yca = __fadd_rn(a.x, __fdiv_rn(ha, 2.0f)), xca = __fadd_rn(a.y, __fdiv_rn(wa, 2.0f));
    float w = __fsub_rn(g.w, g.y), h = __fsub_rn(g.z, g.x);
    float yc = __fadd_rn(g.x, __fdiv_rn(h, 2.0f)), xc = __fadd_rn(g.y, __fdiv_rn(w, 2.0f));
    ha = __fadd_rn(ha, eps); wa = __fadd_rn(wa, eps); h = __fadd_rn(h, eps); w = __fadd_rn(w, eps);
    float4 t;
    t.x = __fdiv_rn(__fsub_rn(yc, yca), ha);
    t.y = __fdiv_rn(__fsub_rn(xc, xca), wa);
    t.z = logf(__fdiv_rn(h, ha));
    t.w = logf(__fdiv_rn(w, wa));
    return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Exact unsigned division by a runtime constant with one mul.hi + shift (divisor fixed per launch).
struct FastDiv {
    unsigned d, mul, shr;
};
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    unsigned l = 0;
    while ((1ull << l) < d) ++l;                       // ceil(log2 d)
    f.mul = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.shr = l - 1;
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv &f) {
    if (f.d <= 1) return n;
    const unsigned t = __umulhi(n, f.mul);
    return (t + ((n - t) >> 1)) >> f.shr;
}
#endif

// Streaming 128-bit load: read once, do not pollute L1.  `volatile` on purpose: it pins the loads
// where the source puts them, so a batch of loads issued before its consumers STAYS a batch
// (without it the compiler sinks each load next to its first use and only one is in flight).
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream1(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// asynchronous pull of one 32-byte sector into L2 (no register, no wait)
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream1(float *p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

#endif  // __CUDACC__
// SM count of the CURRENT device, cached per device ordinal (every cached launch parameter in this library is
// keyed by the device: one process may drive several GPUs).
constexpr int kMaxDevices = 64;
static inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
static inline int device_sm_count() {
    static int sms[kMaxDevices] = {0};
    const int dev = current_device();
    int v = sms[dev];
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v < 1) v = 148;
        sms[dev] = v;   // benign race: every thread computes the same value
    }
    return v;
}

// ---- a forked lane beside the caller's stream ------------------------------------------------------------------
// Per device: one non-blocking side stream and a ring of event pairs (an event may be re-recorded while an older wait on
// it is still pending: a wait captures the record that precedes it).  Works eagerly and under stream capture, where the
// record / wait pairs become the fork and join edges of the graph.
struct SideLane { cudaStream_t side; cudaEvent_t fork, join; };
static inline int fork_side(cudaStream_t st, SideLane *lane) {
    constexpr int kRing = 32, kSides = 8;   // concurrent calls on one device (threads, captures) get different side streams
    struct PerDevice { cudaStream_t side[kSides]; cudaEvent_t ev[2 * kRing]; unsigned next; bool ready; };
    static PerDevice devs[kMaxDevices];
    static std::mutex mu;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    PerDevice &d = devs[dev];
    if (!d.ready) {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < kSides && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&d.side[i], cudaStreamNonBlocking);
        for (int i = 0; i < 2 * kRing && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&d.ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return set_error((int)e, "side stream: %s", cudaGetErrorString(e));
        d.next = 0; d.ready = true;
    }
    const unsigned slot = d.next++ % kRing;
    lane->side = d.side[slot % kSides]; lane->fork = d.ev[2 * slot]; lane->join = d.ev[2 * slot + 1];
    cudaError_t e = cudaEventRecord(lane->fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(lane->side, lane->fork, 0);
    if (e != cudaSuccess) return set_error((int)e, "stream fork: %s", cudaGetErrorString(e));
    return ODK_OK;
}
static inline int join_side(cudaStream_t st, const SideLane &lane) {
    cudaError_t e = cudaEventRecord(lane.join, lane.side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, lane.join, 0);
    if (e != cudaSuccess) return set_error((int)e, "stream join: %s", cudaGetErrorString(e));
    return ODK_OK;
}


}  // namespace odk
