// Micro-benchmark behind the Soft-NMS batch design: what does one short block-wide phase cost on an sm_100a SM?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o barrier_micro barrier_micro.cu && ./barrier_micro
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k(long long *out, float *buf, int gs) {
    __shared__ float s[2048];
    __shared__ float v[256];
    __shared__ int cnt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += 1024) s[i] = buf[i];
    if (tid == 0) cnt = 0;
    __syncthreads();
    long long t0, t1;
    // 1. bare block barrier
    t0 = clock64();
    for (int r = 0; r < 64; ++r) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0) / 64;
    // 2. named barrier over 8 warps
    t0 = clock64();
    if (warp < 8) for (int r = 0; r < 64; ++r) asm volatile("bar.sync 1, 256;" ::: "memory");
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / 64;
    __syncthreads();
    // 3. a "group maxima" phase: LDS, log2(gs) shuffles, STS, block barrier; all 32 warps
    t0 = clock64();
    for (int r = 0; r < 64; ++r) {
        float x = s[(tid + r) & 2047];
        for (int o = 1; o < gs; o <<= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
        if ((lane & (gs - 1)) == 0) v[tid / gs & 255] = x;
        __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0) / 64;
    // 4. the same with 8 warps and the named barrier
    t0 = clock64();
    if (warp < 8) for (int r = 0; r < 64; ++r) {
        float x = s[(tid + r) & 2047];
        for (int o = 1; o < gs; o <<= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
        if ((lane & (gs - 1)) == 0) v[tid / gs & 255] = x;
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / 64;
    __syncthreads();
    // 5. a ranking phase: 4 threads per value, 16 dependent-free LDS + compares, 2 shuffles, barrier (8 warps)
    t0 = clock64();
    if (warp < 8) for (int r = 0; r < 64; ++r) {
        const int g = tid >> 2, part = tid & 3;
        const float vg = v[g & 63];
        int rk = 0;
        for (int h = part; h < 64; h += 4) rk += v[h] > vg || (v[h] == vg && h < g);
        rk += __shfl_xor_sync(0xffffffffu, rk, 1);
        rk += __shfl_xor_sync(0xffffffffu, rk, 2);
        if (part == 0 && rk == 32) v[200] = vg;
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0) / 64;
    __syncthreads();
    // 6. shared atomicAdd by lane 0 of each of 8 warps + shfl + barrier
    t0 = clock64();
    if (warp < 8) for (int r = 0; r < 64; ++r) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&cnt, 3);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base == -1) v[0] = 1.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0) / 64;
    // 7. one warp: 64-bit arg-max by two REDUX + ballot + ffs + LDS + multiply (the replay step)
    t0 = clock64();
    if (warp == 0) {
        float cur = s[lane];
        bool live = true;
        for (int r = 0; r < 32; ++r) {
            const unsigned hi = live ? __float_as_uint(cur) : 0u, lo = 0xFFFFFFFFu - lane;
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            const int p = __ffs(__ballot_sync(0xffffffffu, hi == mh && lo == ml)) - 1;
            if (lane == p) live = false;
            else { const float d = s[64 + p * 33 + lane]; if (d != 1.0f) cur *= d; }
        }
        if (cur == -3.f) v[1] = cur;
    }
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0) / 32;
}

int main() {
    long long *out; float *buf;
    cudaMalloc(&out, 64); cudaMalloc(&buf, 2048 * 4);
    float h[2048];
    for (int i = 0; i < 2048; ++i) h[i] = 1.0f + (float)((i * 2654435761u) >> 8) * 1e-9f;
    for (int i = 64; i < 2048; ++i) h[i] = 1.0f;
    cudaMemcpy(buf, h, sizeof h, cudaMemcpyHostToDevice);
    for (int gs = 4; gs <= 32; gs *= 8) {
        k<<<1, 1024>>>(out, buf, gs);
        k<<<1, 1024>>>(out, buf, gs);
        long long r[8];
        cudaMemcpy(r, out, 56, cudaMemcpyDeviceToHost);
        printf("gs=%d cycles: block barrier %lld | 8-warp named barrier %lld | maxima phase 32 warps %lld | 8 warps %lld | rank phase %lld | "
               "atomic phase %lld | replay step %lld\n", gs, r[0], r[1], r[2], r[3], r[4], r[5], r[6]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
