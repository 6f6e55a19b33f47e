// Device side of the peer-mailbox exchange (include/odk.h): used by the stand-alone publish / collect
// kernels (odk_exchange.cu) and by the finishing CTA of the loss kernel (odk_loss.cu).
#pragma once
#include "odk_common.cuh"

namespace odk {

constexpr int kRecBytes = 32;   // float4 data, u32 sequence number, padding

struct Mailboxes { unsigned char *p[ODK_MAILBOX_MAX_WORLD]; };

__host__ __device__ inline size_t counters_offset(int world) { return (size_t)2 * world * kRecBytes; }

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// counters in the local mailbox: [0] records published, [1] record sets collected (only this rank's
// own kernels touch them, in stream order)
__device__ __forceinline__ unsigned *mailbox_counters(unsigned char *local, int world) {
    return reinterpret_cast<unsigned *>(local + counters_offset(world));
}

// Called by one full warp: lane r stores `v` into slot (seq & 1, rank) of rank r's mailbox.
__device__ __forceinline__ void publish_record(const Mailboxes &mb, int world, int rank, float4 v) {
    unsigned *ctr = mailbox_counters(mb.p[rank], world);
    const int lane = threadIdx.x & 31;
    const unsigned seq = ctr[0] + 1u;
    __syncwarp();
    if (lane == 0) ctr[0] = seq;
    if (lane < world) {
        unsigned char *rec = mb.p[lane] + ((size_t)(seq & 1u) * world + rank) * kRecBytes;
        *reinterpret_cast<float4 *>(rec) = v;
        st_release_sys(reinterpret_cast<unsigned *>(rec + 16), seq);   // orders the data store before the flag
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned long long exchange_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Called by one full warp: waits for the next record set in the local mailbox (up to timeout_ms of wall clock;
// 0 = 30 s -- like a collective, a step only completes when every rank has contributed, so the bound exists to
// turn a dead peer into an error instead of a hung GPU, not to ride out a slow one), sums it in rank order and
// writes the normalised global losses.  On timeout the losses are NaN, *status = 1 + seq (sticky: only ever
// written on failure) and the collected counter is NOT advanced: the next collect waits for the same set, so a
// stale record is never summed.  s_v: shared scratch of ODK_MAILBOX_MAX_WORLD float4.
__device__ __forceinline__ void collect_records(unsigned char *local, int world, float4 *s_v, float *out3, int *status,
                                                unsigned timeout_ms) {
    unsigned *ctr = mailbox_counters(local, world);
    const int lane = threadIdx.x & 31;
    const unsigned seq = ctr[1] + 1u;
    const unsigned long long budget = (unsigned long long)(timeout_ms ? timeout_ms : 30000u) * 1000000ull;
    bool ok = true;
    if (lane < world) {
        const unsigned char *rec = local + ((size_t)(seq & 1u) * world + lane) * kRecBytes;
        const unsigned *flag = reinterpret_cast<const unsigned *>(rec + 16);
        const unsigned long long t0 = exchange_now_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(flag) != seq) {
            if ((++spins & 63u) == 0u && exchange_now_ns() - t0 > budget) { ok = false; break; }
            __nanosleep(200);
        }
        s_v[lane] = __ldcg(reinterpret_cast<const float4 *>(rec));   // L2: never a stale L1 line of an older step
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (lane == 0) {
        if (all_ok) {
            float t = 0.f, c = 0.f, b = 0.f, n = 0.f;
            for (int r = 0; r < world; ++r) { t += s_v[r].x; c += s_v[r].y; b += s_v[r].z; n += s_v[r].w; }
            n -= (float)(world - 1);   // every rank contributed sum(num_positives) + 1
            out3[0] = t / n; out3[1] = c / n; out3[2] = b / n;
            ctr[1] = seq;
        } else {
            const float nan = __int_as_float(0x7fc00000);
            out3[0] = nan; out3[1] = nan; out3[2] = nan;
            if (*status == 0) *status = (int)(1u + seq);
        }
    }
    __syncwarp();
}

}  // namespace odk
