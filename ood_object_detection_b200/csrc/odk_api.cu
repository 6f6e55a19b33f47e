// Error reporting, geometry helpers and small utility entry points of libodk.
#include "odk_common.cuh"

namespace odk {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
    return ODK_OK;
}

int make_geo(Geo *g, const int32_t *level_hw, int num_levels, int na) {
    if (!level_hw || num_levels < 1 || num_levels > ODK_MAX_LEVELS) return set_error(ODK_EINVAL, "num_levels must be in [1,%d]", ODK_MAX_LEVELS);
    if (na < 1 || na > 64) return set_error(ODK_EINVAL, "anchors per location must be in [1,64]");
    int64_t off = 0;
    for (int l = 0; l < ODK_MAX_LEVELS; ++l) {
        g->hw[l] = l < num_levels ? level_hw[l] : 0;
        if (l < num_levels && level_hw[l] < 1) return set_error(ODK_EINVAL, "level_hw[%d] must be positive", l);
        g->off[l] = (int)off;
        off += (int64_t)g->hw[l] * na;
        if (off > (int64_t)1 << 30) return set_error(ODK_EINVAL, "too many anchors");
    }
    g->off[ODK_MAX_LEVELS] = (int)off;
    g->nlev = num_levels;
    g->na = na;
    g->A = (int)off;
    g->Apad = (int)odk_planar_stride(off);
    return ODK_OK;
}

__global__ void scale_inplace_kernel(float *__restrict__ buf, int64_t n, const float *__restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.0f) return;  // the common case (loss.backward()): no memory traffic at all
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    const bool al = ((uintptr_t)buf & 15) == 0;
    for (; i < n; i += stride) {
        if (al && i + 3 < n) {
            float4 v = *reinterpret_cast<float4 *>(buf + i);
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            *reinterpret_cast<float4 *>(buf + i) = v;
        } else {
            for (int k = 0; k < 4 && i + k < n; ++k) buf[i + k] *= s;
        }
    }
}

struct ScaleMulti {
    float *ptr[ODK_SCALE_MAX];
    long long n[ODK_SCALE_MAX];
    int count;
};

__global__ void scale_multi_kernel(const ScaleMulti m, const float *__restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.0f) return;  // loss.backward(): nothing to do, no memory traffic
    float *buf = m.ptr[blockIdx.y];
    const long long n = m.n[blockIdx.y];
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const bool al = ((uintptr_t)buf & 15) == 0;
    for (; i < n; i += stride) {
        if (al && i + 3 < n) {
            float4 v = *reinterpret_cast<float4 *>(buf + i);
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            *reinterpret_cast<float4 *>(buf + i) = v;
        } else {
            for (int k = 0; k < 4 && i + k < n; ++k) buf[i + k] *= s;
        }
    }
}

}  // namespace odk

extern "C" {

int odk_version(void) { return ODK_VERSION; }
const char *odk_last_error(void) { return odk::g_err; }
int64_t odk_planar_stride(int64_t A) { return (A + 3) & ~(int64_t)3; }

int odk_scale_inplace(float *buf, int64_t n, const float *scale, void *stream) {
    if (!buf || !scale || n < 0) return odk::set_error(ODK_EINVAL, "odk_scale_inplace: bad arguments");
    if (n == 0) return ODK_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    odk::scale_inplace_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(buf, n, scale);
    return odk::check_launch("odk_scale_inplace");
}

int odk_scale_inplace_multi(void *const *bufs, const int64_t *sizes, int count, const float *scale, void *stream) {
    using namespace odk;
    if (count < 0 || count > ODK_SCALE_MAX || !scale || (count > 0 && (!bufs || !sizes)))
        return set_error(ODK_EINVAL, "odk_scale_inplace_multi: bad arguments (at most %d buffers)", ODK_SCALE_MAX);
    if (count == 0) return ODK_OK;
    ScaleMulti m;
    for (int i = 0; i < count; ++i) {
        if (!bufs[i] || sizes[i] < 0) return set_error(ODK_EINVAL, "odk_scale_inplace_multi: bad buffer %d", i);
        m.ptr[i] = (float *)bufs[i];
        m.n[i] = sizes[i];
    }
    m.count = count;
    dim3 grid(148 * 4, count);
    scale_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m, scale);
    return check_launch("odk_scale_inplace_multi");
}

}  // extern "C"
