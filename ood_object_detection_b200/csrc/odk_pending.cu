// Entry points whose kernels are not written yet: fail loudly (never fall back).
#include "odk_common.cuh"
extern "C" {
size_t odk_topk_workspace_bytes(int, int) { return 0; }
int odk_topk(const void *const *, const void *const *, int, int, const int32_t *, int, int, int, float *, float *, int64_t *,
             int64_t *, void *, size_t, void *) { return odk::set_error(ODK_EUNSUPPORTED, "odk_topk: not implemented"); }
int odk_detect(const float *, const float *, const int64_t *, const int64_t *, int, int, const float *, int64_t, const float *,
               const float *, const odk_detect_params *, float *, int32_t *, int32_t *, void *) { return odk::set_error(ODK_EUNSUPPORTED, "odk_detect: not implemented"); }
int odk_soft_nms(const float *, const float *, int, int, float, float, float, int, int64_t *, float *, int32_t *, void *) { return odk::set_error(ODK_EUNSUPPORTED, "odk_soft_nms: not implemented"); }
size_t odk_nms_workspace_bytes(int) { return 0; }
int odk_nms(const float *, const float *, int, double, int64_t *, int32_t *, void *, size_t, void *) { return odk::set_error(ODK_EUNSUPPORTED, "odk_nms: not implemented"); }
int odk_ood(const void *const *, int, int, const int32_t *, int, int, const int64_t *, int, float, float *, float *, void *) { return odk::set_error(ODK_EUNSUPPORTED, "odk_ood: not implemented"); }
}
