// Loss partial sums between data-parallel ranks over peer memory (NVLink): see include/odk.h.
// Replaces the all-reduce of the loss scalars (reference effdet/distributed.py reduce_tensor) with
// remote stores into peer mailboxes: nothing waits on the wire, no collective kernel occupies an SM
// while the persistent loss grid runs, and the ranks are coupled with one step of slack instead of
// in lockstep.
#include <string.h>

#include "odk_exchange.cuh"

namespace odk {

__global__ void __launch_bounds__(32) publish_kernel(const float4 *__restrict__ src, const __grid_constant__ Mailboxes mb,
                                                     int world, int rank) {
    publish_record(mb, world, rank, *src);
}

__global__ void __launch_bounds__(32) collect_kernel(unsigned char *local, int world, float *__restrict__ out3,
                                                     int *status, unsigned timeout_ms) {
    __shared__ float4 s_v[ODK_MAILBOX_MAX_WORLD];
    collect_records(local, world, s_v, out3, status, timeout_ms);
}

}  // namespace odk

extern "C" {

size_t odk_mailbox_bytes(int world) {
    if (world < 1 || world > ODK_MAILBOX_MAX_WORLD) return 0;
    return odk::counters_offset(world) + 64;
}

int odk_partials_publish(const float *partials4, void *const *mailboxes, int world, int rank, void *stream) {
    using namespace odk;
    if (world < 1 || world > ODK_MAILBOX_MAX_WORLD || rank < 0 || rank >= world)
        return set_error(ODK_EINVAL, "odk_partials_publish: need 1 <= world <= %d and 0 <= rank < world", ODK_MAILBOX_MAX_WORLD);
    if (!partials4 || !mailboxes || ((uintptr_t)partials4 & 15))
        return set_error(ODK_EINVAL, "odk_partials_publish: partials4 must be non-null and 16-byte aligned");
    Mailboxes mb;
    memset(&mb, 0, sizeof(mb));
    for (int r = 0; r < world; ++r) {
        mb.p[r] = (unsigned char *)mailboxes[r];
        if (!mb.p[r] || ((uintptr_t)mb.p[r] & 15)) return set_error(ODK_EINVAL, "odk_partials_publish: bad mailbox pointer %d", r);
    }
    publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float4 *)partials4, mb, world, rank);
    return check_launch("odk_partials_publish");
}

int odk_partials_collect(void *mailbox_local, int world, float *out3, int32_t *status, uint32_t timeout_ms, void *stream) {
    using namespace odk;
    if (world < 1 || world > ODK_MAILBOX_MAX_WORLD) return set_error(ODK_EINVAL, "odk_partials_collect: bad world size");
    if (!mailbox_local || !out3 || !status || ((uintptr_t)mailbox_local & 15))
        return set_error(ODK_EINVAL, "odk_partials_collect: null or misaligned pointer");
    collect_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned char *)mailbox_local, world, out3, status, timeout_ms);
    return check_launch("odk_partials_collect");
}

}  // extern "C"
