// Image-major streaming model of the class logits (shared by the sample kernel and the fused
// post-process kernel, odk_post.cu).
//
// Per (image b, level l) the NCHW head output is ONE contiguous block of lb = na*C*H*W floats.  A block is
// cut into 16-byte GROUPS on the address grid (the block of image b starts `mis` elements past a 16-byte
// boundary, mis = 0..3, so at most its first and last group are partial) and a TASK is a run of up to
// kGroupsPerTask groups (32 KB): every task of a level has the same size whatever the plane size is, which
// the per-plane segments of odk_topk.cu's collect kernel do not give (D3: half of its tasks are < 1 KB).
// Tasks are numbered image-major: g = b * ntask_img + t, so a grid that takes them in order finishes the
// images in order.
#pragma once
#include "odk_topk.cuh"

namespace odk {

constexpr int kGroupsPerTask = 2048;   // 32 KB of logits
constexpr int kUnitsPerTask = kGroupsPerTask / 32;   // 512-byte warp units; the sample takes at most one per task

struct StreamGeo {
    const float *cls[ODK_MAX_LEVELS];
    unsigned lb[ODK_MAX_LEVELS];        // elements per image block: na * C * hw
    int hw[ODK_MAX_LEVELS];
    int off[ODK_MAX_LEVELS];            // first reference anchor index of the level
    int task_off[ODK_MAX_LEVELS + 1];   // tasks per image, prefix over levels
    FastDiv div_hw[ODK_MAX_LEVELS];
    FastDiv div_C, div_ntask;
    int nlev, na, C, ntask_img;
    unsigned char nhwc[ODK_MAX_LEVELS];   // the level is stored channels_last: memory order is the reference's flat order
};

struct STask {
    const float *blk;    // first element of the (image, level) block
    int l, mis;          // level; elements between the previous 16-byte boundary and blk
    unsigned lb;
    int g0, g1;          // group range of the task (g1 <= groups of the block; may be empty)
    int f0, f1;          // the full (4 valid elements, 16-byte aligned) groups among them
};

__device__ __forceinline__ STask stream_task(const StreamGeo &G, int b, int t) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < ODK_MAX_LEVELS; ++i)
        if (i < G.nlev && t >= G.task_off[i]) l = i;
    STask k;
    k.l = l;
    k.lb = G.lb[l];
    k.blk = G.cls[l] + (size_t)b * k.lb;
    k.mis = (int)(((uintptr_t)k.blk >> 2) & 3u);
    const int groups = (int)((k.mis + k.lb + 3u) >> 2);
    k.g0 = (t - G.task_off[l]) * kGroupsPerTask;
    k.g1 = min(k.g0 + kGroupsPerTask, groups);
    k.f0 = k.g0 + ((k.g0 == 0 && k.mis != 0) ? 1 : 0);
    k.f1 = k.g1 - ((k.g1 == groups && ((k.mis + k.lb) & 3u) != 0) ? 1 : 0);
    return k;
}

// reference flat top-k index (anchor * C + class, bench.py:37,44) of element e of a level block
__device__ __forceinline__ unsigned stream_flat(const StreamGeo &G, int l, unsigned e) {
    if (G.nhwc[l]) return (unsigned)G.off[l] * (unsigned)G.C + e;
    const unsigned ch = fd_div(e, G.div_hw[l]);
    const unsigned pos = e - ch * (unsigned)G.hw[l];
    const unsigned a = fd_div(ch, G.div_C);
    const unsigned c = ch - a * (unsigned)G.C;
    return ((unsigned)G.off[l] + pos * (unsigned)G.na + a) * (unsigned)G.C + c;
}

__device__ __forceinline__ unsigned stream_hash(unsigned b, unsigned t) {
    unsigned h = (t + b * 0x9E3779B1u) * 2654435761u;
    h ^= h >> 15;
    return h * 2246822519u;
}

// host side (odk_post.cu)
int make_stream_geo(StreamGeo *G, const Geo &g, const void *const *cls_levels, int C, int layout);
int fill_topk_levels(TopkArgs *a, const void *const *cls_levels, const void *const *box_levels, int na, int layout, int *ntasks,
                     const char *who);
struct SampleLaunch {
    StreamGeo G;
    int B, K;
    long long N;
    unsigned *slots;      // unused (the slots live in the leader CTA's shared memory)
    int slot_stride;
    int tps;              // tasks per slot group (power of two)
    int shift;            // 2^-shift of the 512-byte units are sampled (set by launch_sample from N)
    unsigned *thr;        // [B] out: collect threshold (value key)
    unsigned *thr_hi;     // [B] out: estimate of the key ~K/16 elements exceed (select's bin edge hint)
    unsigned *zero0;      // up to four [B] arrays and a few consecutive scalars zeroed for the kernels that follow
    unsigned *zero1;
    unsigned *zero2;
    unsigned *zero3;
    unsigned *zero_scalar;
    int zero_scalars;
    unsigned long long *zero_u64;   // one 64-bit word, or null
};
size_t sample_slot_stride(const StreamGeo &G, int *tps_out);
int launch_sample(const SampleLaunch &s, cudaStream_t st);

}  // namespace odk
