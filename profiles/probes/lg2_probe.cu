// Probe: accuracy of MUFU.LG2 / MUFU.EX2 (approx.ftz) for softplus(x) = max(x,0) + log1p(exp(-|x|)).
// Compares candidate fp32 formulations against a double reference over x in [-20, 20].
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void probe(int n, double *out) {
    // out[0..]: per-variant max rel err, sum of signed rel err, weighted bias (sum err / sum val)
    double mx[3] = {0, 0, 0}, se[3] = {0, 0, 0}, sv = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = -20.0f + 40.0f * (float)i / (float)n;
        const double ref = fmax((double)x, 0.0) + log1p(exp(-fabs((double)x)));
        const float e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
        const float u = 1.0f + e;
        // A: plain lg2(1+e)
        const float a = fmaxf(x, 0.f) + lg2_ftz(u) * 0.6931471805599453f;
        // B: lg2(1+e) with the rounding error of 1+e added back
        const float d = e - (u - 1.0f);
        const float b = fmaxf(x, 0.f) + (lg2_ftz(u) * 0.6931471805599453f + d);
        // C: series below 1/32, lg2 above (current kernel)
        const float ser = e * (1.0f - e * (0.5f - e * (0.33333334f - 0.25f * e)));
        const float c = fmaxf(x, 0.f) + (e < 0.03125f ? ser : lg2_ftz(u) * 0.6931471805599453f);
        const float v[3] = {a, b, c};
        for (int k = 0; k < 3; ++k) {
            const double err = ((double)v[k] - ref);
            const double rel = fabs(err) / ref;
            if (rel > mx[k]) mx[k] = rel;
            se[k] += err;
        }
        sv += ref;
    }
    for (int k = 0; k < 3; ++k) { atomicAdd(out + 3 + k, se[k]); }
    atomicAdd(out + 6, sv);
    // max via atomicMax on bit pattern (non-negative doubles order as integers)
    for (int k = 0; k < 3; ++k) atomicMax((unsigned long long *)(out + k), (unsigned long long)__double_as_longlong(mx[k]));
}

// bias on a realistic logit distribution N(-4.6, 1.5): relative error of the SUM
__global__ void probe_sum(int n, double *out, float mean, float stdv) {
    double sa = 0, sb = 0, sc = 0, sr = 0;
    unsigned s = 1234567u + 7919u * (blockIdx.x * blockDim.x + threadIdx.x);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // Box-Muller from an LCG
        s = s * 1664525u + 1013904223u; const float u1 = ((s >> 8) + 1) * (1.0f / 16777217.0f);
        s = s * 1664525u + 1013904223u; const float u2 = (s >> 8) * (1.0f / 16777216.0f);
        const float x = mean + stdv * sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
        const double ref = fmax((double)x, 0.0) + log1p(exp(-fabs((double)x)));
        const float e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
        const float u = 1.0f + e;
        const float l = lg2_ftz(u) * 0.6931471805599453f;
        const float d = e - (u - 1.0f);
        const float ser = e * (1.0f - e * (0.5f - e * (0.33333334f - 0.25f * e)));
        sa += fmaxf(x, 0.f) + l; sb += fmaxf(x, 0.f) + (l + d); sc += fmaxf(x, 0.f) + (e < 0.03125f ? ser : l); sr += ref;
    }
    atomicAdd(out + 0, sa); atomicAdd(out + 1, sb); atomicAdd(out + 2, sc); atomicAdd(out + 3, sr);
}

int main() {
    double *d, h[8];
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    probe<<<296, 256>>>(1 << 26, d);
    cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
    printf("variant          max_rel_err     sum_err/sum_val\n");
    const char *nm[3] = {"A lg2(1+e)      ", "B lg2(1+e)+d    ", "C series|lg2    "};
    for (int k = 0; k < 3; ++k) printf("%s %.3e   %.3e\n", nm[k], h[k], h[3 + k] / h[6]);
    const float ms[6][2] = {{-4.6f, 1.5f}, {-7.f, 1.f}, {-10.f, 2.f}, {0.f, 3.f}, {-14.f, 1.f}, {3.f, 1.f}};
    for (int q = 0; q < 6; ++q) {
        cudaMemset(d, 0, 64);
        probe_sum<<<296, 256>>>(1 << 28, d, ms[q][0], ms[q][1]);
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("N(%5.1f,%3.1f) sum rel err: A %.3e  B %.3e  C %.3e\n", ms[q][0], ms[q][1], (h[0] - h[3]) / h[3], (h[1] - h[3]) / h[3], (h[2] - h[3]) / h[3]);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
