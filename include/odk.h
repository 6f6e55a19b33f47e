/*
 * odk.h -- C ABI of libodk.so: the B200 (sm_100a) kernels for the dense per-anchor hot path of
 * DavidPetrus/ood_object_detection (an effdet fork).
 *
 * The reference has no FFI layer: its boundary for this path is the python API of
 * effdet/anchors.py, effdet/loss.py, effdet/bench.py and effdet/soft_nms.py.  Every entry point
 * below replaces the torch/torchvision op sequence behind one of those functions; the python
 * shims in ood_object_detection_b200/ keep the reference signatures and call these through ctypes
 * (INTEGRATION.md shows the binding).  Citations are relative to the reference root.
 *
 * Conventions
 *   - plain pointers and sizes only; every DEVICE buffer (inputs, outputs, workspace) is owned
 *     by the caller; the library never allocates, frees or retains device memory;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - pyramid geometry: `level_hw[l]` = H_l*W_l for l in [0, num_levels) (HOST array),
 *     `na` = anchors per location (9).  A = na * sum(level_hw).  Reference anchor order
 *     (anchors.py:264-299): r = off_l + (y*W_l + x)*na + a.
 *   - "planar" order used by odk_assign/odk_loss for per-anchor integers: p = off_l + a*HW_l +
 *     (y*W_l + x), the order of the NCHW head outputs; the per-image stride is
 *     odk_planar_stride(A) (A rounded up to 4).
 *   - `cls_levels` / `box_levels`: HOST arrays of num_levels DEVICE pointers to contiguous
 *     NCHW fp32 tensors [B, na*C, H_l, W_l] / [B, na*4, H_l, W_l] (efficientdet.py:410-414).
 *   - `layout` (post-process entry points): bit l set = class level l is stored channels_last, i.e.
 *     [B, H_l, W_l, na*C] in memory (what a channels_last / AMP head writes); bit 8 + l the same for
 *     box level l ([B, H_l, W_l, na*4], 16-byte aligned).  0 = everything NCHW.  Both are read in place.
 *   - return 0 on success, <0 for argument errors, >0 = cudaError_t of a failed launch;
 *     odk_last_error() returns a thread-local message.  No exceptions cross the ABI.
 */
#ifndef ODK_H_
#define ODK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODK_MAX_LEVELS 8
#define ODK_VERSION 3

/* error codes */
#define ODK_OK 0
#define ODK_EINVAL (-1)
#define ODK_EWORKSPACE (-2)
#define ODK_EUNSUPPORTED (-3)

int odk_version(void);
const char *odk_last_error(void);
/* per-image stride (elements) of planar per-anchor arrays */
int64_t odk_planar_stride(int64_t A);

/* ---- target assignment -------------------------------------------------------------------
 * Replaces AnchorLabeler.batch_label_anchors' per-image loop (anchors.py:393-434):
 * IouSimilarity.compare (region_similarity_calculator.py:59-73), ArgMaxMatcher.match with
 * force_match_for_each_row (argmax_matcher.py:105-146), thresholds matched == unmatched ==
 * match_thr (anchors.py:321-325).
 *   anchors   [A,4] yxyx fp32, reference order
 *   gt_boxes  [B,Mmax,4] yxyx fp32; gt_labels [B,Mmax] int32 (1-based classes)
 *   gt_count  [B] int32 or NULL (= Mmax rows each)
 *   filter_valid != 0: rows with label < 0 are skipped (anchors.py:405-408)
 * Outputs: match [B, odk_planar_stride(A)] int32, planar order: gt row (index into the
 * UNFILTERED Mmax rows) or -1; num_pos [B] fp32 (anchors.py:434).
 * Workspace: odk_assign_workspace_bytes(B, Mmax).
 */
size_t odk_assign_workspace_bytes(int B, int Mmax);
int odk_assign(const float *anchors, const float *gt_boxes, const int32_t *gt_labels, const int32_t *gt_count, int B,
               int Mmax, const int32_t *level_hw, int num_levels, int na, float match_thr, int filter_valid,
               int32_t *match, float *num_pos, void *workspace, size_t workspace_bytes, void *stream);

/* Same result as odk_assign for anchors that are REGULAR GRIDS per (level, shape) -- the pyramid
 * of anchors.py:264-299 -- in time proportional to the anchors each gt box can touch instead of
 * B*M*A (a CTA per gt enumerates only the cells of the planes that can matter).
 *   plane_desc  DEVICE [num_levels*na][12] fp32, plane k = level*na + shape:
 *               cy0, cx0 (centre of cell (0,0)), sy, sx (cell pitch), hy, hx (half sizes), area,
 *               W, H, off_l, shape, level
 *   plane_gen   DEVICE [num_levels*na][6] fp64 or NULL: the generator of the same grids -- cy0, cx0, sy, sx,
 *               half_y, half_x exactly as anchors.py:264-299 computes them in float64.  When given, a cell's anchor is
 *               recomputed from it (bit-identical to the table, see odk_anchor_table) instead of gathered.
 * match_thr must be > 0.  Workspace: odk_assign_grid_workspace_bytes(B, A). */
size_t odk_assign_grid_workspace_bytes(int B, int64_t A);
#define ODK_ASSIGN_WS_CLEAN 1   /* flags: the workspace is all zero (fresh, or left so by odk_loss' clear_keys): no memset */
int odk_assign_grid(const float *anchors, const float *plane_desc, const double *plane_gen, int num_planes, const float *gt_boxes,
                    const int32_t *gt_labels, const int32_t *gt_count, int B, int Mmax, const int32_t *level_hw,
                    int num_levels, int na, float match_thr, int filter_valid, int32_t *match, float *num_pos,
                    float *normalizer, int flags, void *workspace, size_t workspace_bytes, void *stream);
/* `match` may be NULL: the assignment then stays in the workspace as 64-bit keys
 * ([B, odk_planar_stride(A)] uint64 at offset 0; 0 = unmatched, else low 32 bits = ~gt_row), which
 * odk_loss consumes directly (odk_loss_params.match_is_key64) and odk_keys_to_match converts on
 * demand.  `normalizer` (nullable) receives sum(num_pos) + 1 (loss.py:261) as one fp32. */
int odk_keys_to_match(const void *keys, int B, int64_t A, int32_t *match, void *stream);
/* The anchor table [A,4] (yxyx fp32, reference order) from plane_desc + plane_gen: what Anchors._generate_boxes
 * (anchors.py:264-299) produces on the host, bit for bit. */
int odk_anchor_table(const float *plane_desc, const double *plane_gen, int num_planes, const int32_t *level_hw, int num_levels,
                     int na, float *anchors_out, void *stream);

/* Pairwise IoU matrix out[n,m] of yxyx boxes: IouSimilarity.compare
 * (region_similarity_calculator.py:59-101), same fp32 operation order. */
int odk_iou_matrix(const float *boxes1, int n, const float *boxes2, int m, float *out, void *stream);

/* Materialises the reference's target tensors from `match` (matcher.py:170-179,
 * box_coder.py:81-110, target_assigner.py:168-220, anchors.py:416-432):
 *   cls_targets: one int64 block per level, [B, H_l, W_l, na], blocks concatenated
 *                (block l starts at element B*off_l);  value = label-1, background -1
 *   box_targets: same blocking, fp32 [B, H_l, W_l, na*4]
 */
int odk_targets(const float *anchors, const float *gt_boxes, const int32_t *gt_labels, int B, int Mmax,
                const int32_t *level_hw, int num_levels, int na, const int32_t *match, int64_t *cls_targets,
                float *box_targets, void *stream);

/* ---- detection loss ----------------------------------------------------------------------
 * Replaces loss_fn (loss.py:224-298): one_hot (:182-186), new_focal_loss (:49-95) or
 * focal_loss_legacy (:15-47), huber/_box_loss (:104-118,:171-179), the per-level loop and sums.
 * Targets come either from odk_assign (`match` != NULL: class and encoded box are recomputed on
 * the fly, nothing per-anchor but `match` is read) or from reference-layout tensors
 * (`cls_targets`/`box_targets` blocks as written by odk_targets; `match` == NULL).
 *   normalizer  device fp32 scalar = sum(num_positives)+1 (loss.py:261)
 *   out         device fp32 [3] = total, cls_loss, box_loss (loss.py:295-298)
 *   grad_cls_levels / grad_box_levels: HOST arrays of DEVICE pointers (same shapes as the
 *   inputs) receiving d total / d input, or NULL for forward only.
 * Workspace: odk_loss_workspace_bytes().
 * With `match` the work is a layout-agnostic stream over the class logits (every element as a negative)
 * plus a patch of the matched anchors; the loss sums are bit-reproducible from call to call.  The
 * environment variable ODK_LOSS_KERNEL=ring selects the plane-walking kernels instead (NCHW only; they
 * also serve targets given as tensors).
 */
#define ODK_MAILBOX_MAX_WORLD 32
/* Optional fused exchange of the loss partial sums between data-parallel ranks (see the mailbox
 * section below): the CTA of odk_loss that finishes the reduction first collects the record set
 * of the PREVIOUS step if one is outstanding (-> global_out3, status), then stores this launch's
 * {out[0], out[1], out[2], *num_pos_plus_1} into every rank's mailbox.  Use with a unit normaliser. */
typedef struct odk_exchange {
    void *mailboxes[ODK_MAILBOX_MAX_WORLD]; /* device pointers to every rank's mailbox; entry `rank` is local */
    int32_t world, rank;
    const float *num_pos_plus_1;  /* device: this rank's sum(num_positives) + 1 (odk_assign_grid's `normalizer`) */
    float *global_out3;           /* device: {total, cls_loss, box_loss} of the previous step's GLOBAL batch */
    int32_t *status;              /* device, STICKY: 0, or 1 + the first sequence number whose record set did not arrive in
                                     time (never cleared by the library; global_out3 is NaN for that step) */
    int32_t normalized;           /* out[0..2] of this launch are already divided by *num_pos_plus_1 (local normaliser,
                                     what a gradient step uses): the published sums are multiplied back */
    uint32_t timeout_ms;          /* how long a collect waits for the peers' records; 0 = 30 000 ms */
} odk_exchange;

typedef struct odk_loss_params {
    float alpha;
    float gamma;
    float delta;
    float box_loss_weight;
    float label_smoothing;
    int32_t legacy_focal;
    int32_t match_is_key64;   /* `match` points at odk_assign_grid's 64-bit keys instead of int32 rows */
    int32_t clear_keys;       /* with match_is_key64 (`match` = the labeler's workspace): after the last read, zero the keys the
                                 labeler set and its counters, so the next odk_assign_grid can run with ODK_ASSIGN_WS_CLEAN */
    int32_t layout;           /* bit l: class level l is channels_last ([B,H,W,na*C] in memory), bit 8 + l: box level l; 0 = NCHW.
                                 channels_last levels need `match` (fused targets) */
    const odk_exchange *exchange;   /* NULL = no exchange (HOST pointer, read during the call) */
} odk_loss_params;

size_t odk_loss_workspace_bytes(void);
int odk_loss(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
             int num_levels, int na, const int32_t *match, const float *anchors, const float *gt_boxes,
             const int32_t *gt_labels, int Mmax, const int64_t *cls_targets, const float *box_targets,
             const float *normalizer, const odk_loss_params *params, float *out, void *const *grad_cls_levels,
             void *const *grad_box_levels, void *workspace, size_t workspace_bytes, void *stream);

/* In-place `buf[i] *= *scale` over n floats; returns immediately on device when *scale == 1.
 * Used by the autograd shim when the upstream gradient of the loss is not 1. */
int odk_scale_inplace(float *buf, int64_t n, const float *scale, void *stream);
/* Same for up to ODK_SCALE_MAX separate buffers in ONE launch (bufs / sizes: HOST arrays). */
#define ODK_SCALE_MAX 16
int odk_scale_inplace_multi(void *const *bufs, const int64_t *sizes, int count, const float *scale, void *stream);

/* ---- data-parallel loss partial sums over peer memory --------------------------------------
 * The only exchange of the sharded path (the reference's counterpart is the
 * all-reduce of its loss scalars, effdet/distributed.py reduce_tensor): every rank owns a MAILBOX
 * of odk_mailbox_bytes(world) device bytes that all peers can write (CUDA VMM / IPC mapping, e.g.
 * torch symmetric memory), zero-filled once before first use.
 *   odk_partials_publish: stores this rank's 4 floats {cls + w*box, cls, box, sum(num_pos) + 1}
 *     (what odk_assign_grid + odk_loss leave in one buffer against a unit normaliser) into slot
 *     (seq & 1, rank) of EVERY rank's mailbox with release semantics; seq is a device-side counter
 *     in the local mailbox, so the call can be replayed from a CUDA graph.  mailboxes: HOST array
 *     of `world` device pointers (entry `rank` is the local mailbox).  No waiting.
 *   odk_partials_collect: waits (bounded spin) until all `world` records of the next sequence
 *     number are in the LOCAL mailbox, sums them in rank order (deterministic) and writes
 *     out3 = {total, cls_loss, box_loss} of the global batch (sums / (sum(num_pos) + 1), loss.py:261,297).
 *     If a peer's record does not arrive within timeout_ms (wall clock, %globaltimer; 0 = 30 s) the step's
 *     out3 is NaN, *status becomes 1 + its sequence number and STAYS set (sticky: the caller decides when
 *     to look), and the collected counter is NOT advanced, so the same record set is waited for again by
 *     the next collect instead of summing whatever stale record is in the slot.
 * Every rank must alternate collect(j-1) ... publish(j) in stream order (that ordering is the flow
 * control that makes two slots enough) and publish once before its first collect.
 * odk_loss with odk_loss_params.exchange does both inside the loss kernel (collect(j-1) when a record
 * is outstanding, then publish(j)); odk_partials_collect then only drains the last step. */
size_t odk_mailbox_bytes(int world);
int odk_partials_publish(const float *partials4, void *const *mailboxes, int world, int rank, void *stream);
int odk_partials_collect(void *mailbox_local, int world, float *out3, int32_t *status, uint32_t timeout_ms, void *stream);

/* ---- post-process: top-k -------------------------------------------------------------------
 * Replaces _post_process (bench.py:12-56): concat/permute of the levels, torch.topk over
 * [B, A*C] (k = K, sorted descending; ties broken by ascending flat index), index split and
 * the three gathers.  Outputs: cls_topk [B,K] fp32 (the selected logit), box_topk [B,K,4],
 * indices [B,K] int64 (anchor = flat // C), classes [B,K] int64 (flat % C).
 * Workspace: odk_topk_workspace_bytes(B, C, level_hw, num_levels, na, K).
 */
size_t odk_topk_workspace_bytes(int B, int C, const int32_t *level_hw, int num_levels, int na, int K);
int odk_topk(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
             int num_levels, int na, int K, int layout, float *cls_topk, float *box_topk, int64_t *indices, int64_t *classes,
             void *workspace, size_t workspace_bytes, void *stream);

/* ---- post-process: detections ---------------------------------------------------------------
 * Replaces _batch_detection / generate_detections (bench.py:59-76, anchors.py:95-172):
 * anchor gather, decode_box_outputs (anchors.py:51-85), optional clip (anchors.py:88-92),
 * sigmoid, score > score_min filter, then either torchvision batched_nms (coordinate trick,
 * iou > nms_thr) or batched_soft_nms (soft_nms.py:115-169, gaussian), first max_det kept,
 * class+1, optional rescale.
 *   cls_topk [B,N] logits, box_topk [B,N,4], indices/classes [B,N] int64 (odk_topk layout)
 *   img_scale [B] or NULL; img_size [B,2] or NULL (clip only when both given)
 * Outputs: dets [B,max_det,6] (x0,y0,x1,y1,score,class; zero padded), count [B] int32,
 * src [B,max_det] int32 = position in the N list each detection came from (-1 padded).
 */
typedef struct odk_detect_params {
    int32_t max_det;
    int32_t soft_nms;       /* 0: hard NMS, 1: gaussian soft-NMS */
    float score_min;        /* 0.01 (anchors.py:141) */
    double nms_iou;         /* 0.3  (anchors.py:150) */
    float soft_sigma;       /* 0.5  */
    float soft_iou;         /* 0.3  (unused by the gaussian method) */
    float soft_score_thr;   /* 0.001 */
    int32_t pipeline;       /* odk_postprocess only: ODK_PIPELINE_STAGED (default) or ODK_PIPELINE_PERSISTENT */
} odk_detect_params;
#define ODK_PIPELINE_STAGED 0       /* sample -> one-wave collect -> one tail CTA per image */
#define ODK_PIPELINE_PERSISTENT 1   /* sample -> one persistent kernel: image-major stream, tails overlap the later images */

int odk_detect(const float *cls_topk, const float *box_topk, const int64_t *indices, const int64_t *classes, int B,
               int N, const float *anchors, int64_t A, const float *img_scale, const float *img_size,
               const odk_detect_params *params, float *dets, int32_t *count, int32_t *src, void *stream);

/* ---- post-process: the whole chain in one pipeline ------------------------------------------------
 * Replaces DetBenchPredict.forward's post-process (bench.py:93-100: _post_process then
 * _batch_detection): same results as odk_topk followed by odk_detect (and odk_ood), but the logits are
 * streamed image-major by one persistent kernel and everything else of an image (select, decode,
 * suppression, OOD scores) runs while the later images are still streaming.  K <= 6144.
 *   dets [B,max_det,6], count [B], src [B,max_det] as odk_detect;
 *   det_anchor [B,max_det] int64 anchor index of every detection (-1 padded), or NULL;
 *   energy / max_logit [B,max_det] as odk_ood (both or neither; need det_anchor), temperature > 0;
 *   cls_topk / box_topk / indices / classes: the top-k tensors as odk_topk, all four or all NULL.
 * Workspace: odk_postprocess_workspace_bytes(B, C, level_hw, num_levels, na, K). */
size_t odk_postprocess_workspace_bytes(int B, int C, const int32_t *level_hw, int num_levels, int na, int K);
/* Diagnostics: byte offset inside the workspace of a uint32 [B] array that holds, after the call, 1 for every
 * image that left the sampled-threshold path (exact radix select + stand-alone detect), else 0. */
size_t odk_postprocess_flags_offset(int B, int C, const int32_t *level_hw, int num_levels, int na, int K);
/* Diagnostics: byte offset of a uint64 [B*8 + 2] array of %globaltimer nanoseconds: per image {its last logit was
 * streamed, a CTA took its tail, the tail started, the tail ended, select done, filter + decode done, class offsets
 * done, suppression done}; then the start and the end of the kernel. */
size_t odk_postprocess_timeline_offset(int B, int C, const int32_t *level_hw, int num_levels, int na, int K);
int odk_postprocess(const void *const *cls_levels, const void *const *box_levels, int B, int C, const int32_t *level_hw,
                    int num_levels, int na, int K, int layout, const float *anchors, const float *img_scale, const float *img_size,
                    const odk_detect_params *params, float temperature, float *dets, int32_t *count, int32_t *src,
                    int64_t *det_anchor, float *energy, float *max_logit, float *cls_topk, float *box_topk,
                    int64_t *indices, int64_t *classes, void *workspace, size_t workspace_bytes, void *stream);

/* Stand-alone soft_nms (soft_nms.py:42-112) on one box set [n,4] xyxy: runs until no box is
 * left or max_rounds is reached.  idx_out [n] int64, score_out [n] fp32, count [1] int32. */
int odk_soft_nms(const float *boxes, const float *scores, int n, int method_gaussian, float sigma, float iou_thr,
                 float score_thr, int max_rounds, int64_t *idx_out, float *score_out, int32_t *count, void *stream);

/* Stand-alone greedy NMS with torchvision::nms semantics on [n,4] xyxy (scores need not be
 * sorted).  keep [n] int64 in descending score order, count [1] int32. */
size_t odk_nms_workspace_bytes(int n);
int odk_nms(const float *boxes, const float *scores, int n, double iou_thr, int64_t *keep, int32_t *count,
            void *workspace, size_t workspace_bytes, void *stream);

/* ---- OOD score --------------------------------------------------------------------------------
 * Not in the reference.  For anchor_idx [B,D] int64 (entries < 0 skipped ->
 * outputs 0): energy = -T*logsumexp(logits[anchor,:]/T), max_logit = max_c logits[anchor,c],
 * read in place from the NCHW levels (the row the reference gathers at bench.py:51-52). */
int odk_ood(const void *const *cls_levels, int B, int C, const int32_t *level_hw, int num_levels, int na, int layout,
            const int64_t *anchor_idx, int D, float temperature, float *energy, float *max_logit, void *stream);

/* ---- evaluation: per-image true/false positives and CorLoc ------------------------------------
 * Replaces PerImageEvaluation.compute_object_detection_metrics (effdet/evaluation/per_image_evaluation.py:29-92;
 * _compute_tp_fp :177-240, _compute_tp_fp_for_single_class :305-470, _compute_cor_loc :93-175,
 * _remove_invalid_boxes :512-536; np_box_list.py iou / ioa / non_max_suppression), the numpy loop the reference's
 * evaluators run per image on the host (detection_evaluator.py:268-305), for a whole batch on the device.
 *   dets [B,D,6] rows x0,y0,x1,y1,score,class as odk_detect writes them; count [B] rows in use or NULL (= D);
 *   gt_boxes [B,M,4] yxyx fp32; gt_labels [B,M] int32 in the same numbering as dets' class column, < 0 = padding;
 *   gt_difficult / gt_group_of [B,M] uint8 or NULL; class index = label - label_offset must be in [0, num_classes);
 *   match_iou 0.5, nms_iou 1.0 (= off) and nms_max 10000 are the ObjectDetectionEvaluation defaults.
 * Outputs: label [B,D] int8 per detection slot: 1 true positive, 0 false positive, -1 ignored (matched a difficult
 * or group-of box; group_of_weight 0), -2 not evaluated (padding, invalid box, removed by the NMS);
 * corloc [B,num_classes] uint8.  Decisions are bit-exact (numpy's mixed fp32 / float64 arithmetic is reproduced);
 * equal scores are taken later-detection-first (numpy's argsort()[::-1] leaves that order unspecified). */
int odk_match_detections(const float *dets, const int32_t *count, int B, int D, const float *gt_boxes, const int32_t *gt_labels,
                         const uint8_t *gt_difficult, const uint8_t *gt_group_of, int M, int num_classes, int label_offset,
                         double match_iou, double nms_iou, int nms_max, int8_t *label, uint8_t *corloc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ODK_H_ */
