// What does a read-1-write-1 fp32 stream reach on this B200, with and without the focal-loss arithmetic, and how
// much do the access pattern / loads in flight / CTAs per SM matter?  (The question behind the fwd+grad loss kernel:
// is 5.5-5.8 TB/s the mixed read+write ceiling or the kernel's own limit?)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_micro stream_micro.cu && ./stream_micro
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float math1(float x, float g, float &acc) {
    const float e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
    float t = fmaf(e, 0.007363723125308752f, -0.03774333372712135f);
    t = fmaf(e, t, 0.0910765677690506f); t = fmaf(e, t, -0.14697664976119995f); t = fmaf(e, t, 0.19561563432216644f);
    t = fmaf(e, t, -0.2494998425245285f); t = fmaf(e, t, 0.3333110213279724f); t = fmaf(e, t, -0.4999998211860657f);
    t = fmaf(e, t, 1.0f);
    acc += fmaf(e, t, fmaxf(x, 0.f));
    return g * __fdividef(x >= 0.f ? 1.f : e, 1.f + e);
}
template <bool MATH>
__device__ __forceinline__ float4 work(float4 v, float g, float &acc) {
    if (!MATH) return make_float4(v.x * g, v.y * g, v.z * g, v.w * g);
    return make_float4(math1(v.x, g, acc), math1(v.y, g, acc), math1(v.z, g, acc), math1(v.w, g, acc));
}

// grid-stride: U loads in flight, `stride` float4 apart
template <bool MATH, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_stride(const float *__restrict__ x, float *__restrict__ y, size_t n4, float g, float *out) {
    const size_t stride = (size_t)gridDim.x * 256, first = (size_t)blockIdx.x * 256 + threadIdx.x;
    float acc = 0.f;
    size_t u = first;
    for (; u + (U - 1) * stride < n4; u += U * stride) {
        float4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld_stream4(x + (u + j * stride) * 4);
#pragma unroll
        for (int j = 0; j < U; ++j) st_stream4(y + (u + j * stride) * 4, work<MATH>(v[j], g, acc));
    }
    for (; u < n4; u += stride) st_stream4(y + u * 4, work<MATH>(ld_stream4(x + u * 4), g, acc));
    if (acc == 12345.f) *out = acc;
}
// tiles: a CTA takes a contiguous run of U*4 KB at a time (256 threads x 16 B x U)
template <bool MATH, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_tile(const float *__restrict__ x, float *__restrict__ y, size_t n4, float g, float *out) {
    float acc = 0.f;
    const size_t tile = 256 * U, ntiles = n4 / tile;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        float4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld_stream4(x + (base + j * 256) * 4);
#pragma unroll
        for (int j = 0; j < U; ++j) st_stream4(y + (base + j * 256) * 4, work<MATH>(v[j], g, acc));
    }
    if (acc == 12345.f) *out = acc;
}


// TMA bulk copies both ways: global -> shared (mbarrier), arithmetic in place in shared memory, shared -> global
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <bool MATH, int TILE_F4, int STAGES, int MINB, bool HINT = false>
__global__ void __launch_bounds__(256, MINB) k_tma(const float *__restrict__ x, float *__restrict__ y, size_t n4, float g, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    constexpr unsigned kTileBytes = TILE_F4 * 16;
    float4 *buf = reinterpret_cast<float4 *>(smem);
    unsigned long long pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t ntiles = n4 / TILE_F4;
    const size_t my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto load = [&](size_t i) {   // thread 0: tile i of this CTA into stage i % STAGES
        const int s = (int)(i % STAGES);
        const float *src = x + (blockIdx.x + i * gridDim.x) * (size_t)TILE_F4 * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(kTileBytes) : "memory");
        if (HINT)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                             smem_u32(buf + (size_t)s * TILE_F4)),
                         "l"(src), "r"(kTileBytes), "r"(smem_u32(&full[s])), "l"(pol)
                         : "memory");
        else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(buf + (size_t)s * TILE_F4)),
                     "l"(src), "r"(kTileBytes), "r"(smem_u32(&full[s]))
                     : "memory");
    };
    if (threadIdx.x == 0)
        for (size_t i = 0; i < (size_t)STAGES - 1 && i < my; ++i) load(i);
    float acc = 0.f;
    for (size_t i = 0; i < my; ++i) {
        const int s = (int)(i % STAGES);
        const unsigned parity = (unsigned)((i / STAGES) & 1);
        if (threadIdx.x == 0) {
            // the stage tile i+STAGES-1 goes into held tile i-1: its store must have finished reading shared memory
            if (i + STAGES - 1 < my) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                load(i + STAGES - 1);
            }
        }
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&full[s])), "r"(parity) : "memory");
        float4 *t = buf + (size_t)s * TILE_F4;
#pragma unroll
        for (int j = 0; j < TILE_F4 / 256; ++j) t[threadIdx.x + j * 256] = work<MATH>(t[threadIdx.x + j * 256], g, acc);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            float *dst = y + (blockIdx.x + i * gridDim.x) * (size_t)TILE_F4 * 4;
            if (HINT)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(t)), "r"(kTileBytes), "l"(pol) : "memory");
            else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(t)), "r"(kTileBytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 12345.f) *out = acc;
}

template <class F>
static float timed(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 10;
}

int main() {
    const size_t n = (size_t)64 * 49104 * 90;   // D0 B=64 class logits
    const size_t n4 = n / 4;
    float *x, *y, *out;
    cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&out, 4);
    cudaMemset(x, 0x3c, n * 4);
    const double gb = 2.0 * n * 4 / 1e9;
    auto rep = [&](const char *name, float ms) { printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, gb / ms * 1e3); };
    rep("cudaMemcpy D2D", timed([&] { cudaMemcpyAsync(y, x, n * 4, cudaMemcpyDeviceToDevice); }));
#define RUN(K, MATH, U, MINB)                                                                        \
    rep(#K " math=" #MATH " U=" #U " ctas/SM=" #MINB, timed([&] { K<MATH, U, MINB><<<148 * MINB, 256>>>(x, y, n4, 0.5f, out); }))
    RUN(k_stride, false, 4, 4); RUN(k_stride, false, 4, 8); RUN(k_stride, false, 8, 4);
    RUN(k_tile, false, 4, 4); RUN(k_tile, false, 4, 8); RUN(k_tile, false, 8, 4);
    RUN(k_stride, true, 4, 4); RUN(k_stride, true, 4, 6); RUN(k_stride, true, 4, 8); RUN(k_stride, true, 8, 4); RUN(k_stride, true, 2, 8);
    RUN(k_tile, true, 4, 4); RUN(k_tile, true, 4, 6); RUN(k_tile, true, 4, 8); RUN(k_tile, true, 8, 4); RUN(k_tile, true, 2, 8);
#define RUNTH(MATH, TILE, ST, MINB)                                                                             \
    {                                                                                                           \
        auto kf = k_tma<MATH, TILE, ST, MINB, true>;                                                            \
        cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16 * ST);                  \
        rep("k_tma+evict_first math=" #MATH " tile_f4=" #TILE " stages=" #ST " ctas/SM=" #MINB,                 \
            timed([&] { kf<<<148 * MINB, 256, TILE * 16 * ST>>>(x, y, n4, 0.5f, out); }));                      \
    }
#define RUNT(MATH, TILE, ST, MINB)                                                                              \
    {                                                                                                           \
        auto kf = k_tma<MATH, TILE, ST, MINB>;                                                                  \
        cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16 * ST);                  \
        rep("k_tma math=" #MATH " tile_f4=" #TILE " stages=" #ST " ctas/SM=" #MINB,                             \
            timed([&] { kf<<<148 * MINB, 256, TILE * 16 * ST>>>(x, y, n4, 0.5f, out); }));                      \
    }
    RUNT(false, 1024, 4, 2); RUNT(false, 1024, 3, 3); RUNT(false, 512, 4, 4); RUNT(false, 2048, 3, 2);
    RUNTH(false, 2048, 3, 2); RUNTH(true, 2048, 3, 2); RUNTH(true, 1024, 3, 3); RUNT(true, 1536, 4, 2); RUNT(true, 3072, 2, 2); RUNT(true, 2048, 2, 3);
    RUNT(true, 1024, 4, 2); RUNT(true, 1024, 3, 3); RUNT(true, 512, 4, 4); RUNT(true, 2048, 3, 2); RUNT(true, 1024, 6, 2);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
