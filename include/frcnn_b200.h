/*
 * frcnn_b200.h -- C ABI of libfrcnn_b200.so: the B200 (sm_100a) proposal-and-RoI hot path of a
 * two-stage detector, as hand-written CUDA kernels.
 *
 * The reference (3SAILab/two_stage_object_detection) has no FFI; its boundary for this path is a
 * set of Python callables.  Each entry point below names the reference
 * callable (file:line, relative to the reference tree) it replaces; the Python mirror in
 * two_stage_object_detection_b200/{utils,nets}/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - fp32 boxes are (x1, y1, x2, y2); tensors are dense row-major in the stated shape;
 *   - the caller allocates every output and the workspace (query *_workspace_bytes first);
 *     workspaces need 256-byte alignment and no initialisation;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises with the host;
 *   - return value: 0 on success, a negative frcnn_status otherwise; frcnn_last_error() returns a
 *     thread-local description of the last failure;
 *   - per-image `status` outputs carry the reference's data-dependent IndexError conditions
 *     (FRCNN_IMG_*), because raising them needs no host sync on the hot path;
 *   - arithmetic is fp32 with every multiply/add rounded separately (no FMA contraction), in the
 *     reference's operation order, so index/label outputs are bit-identical to the reference.
 */
#ifndef FRCNN_B200_H_
#define FRCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRCNN_ABI_VERSION 2

typedef void* frcnn_stream_t; /* cudaStream_t */

enum frcnn_status {
    FRCNN_OK = 0,
    FRCNN_ERR_INVALID_ARG = -1,
    FRCNN_ERR_WORKSPACE = -2,
    FRCNN_ERR_CUDA = -3,
    FRCNN_ERR_UNSUPPORTED = -4
};

/* per-image status bits written by the batched ops */
#define FRCNN_IMG_OK 0
#define FRCNN_IMG_PAD_INDEX_ERROR 1     /* nets/rpn.py:65-69 would raise IndexError            */
#define FRCNN_IMG_SCATTER_INDEX_ERROR 2 /* nets/frcnn_training.py:175 would raise IndexError   */

#define FRCNN_MAX_BASE_ANCHORS 64

int frcnn_abi_version(void);
const char* frcnn_last_error(void);
/* sm count / compute capability of the current device (host-side query) */
int frcnn_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Measurement hooks (bench.py): kernels launched by this library in this process so far, and the template
 * instance the last RoI forward call of the calling thread picked (e.g. "roi_pool_tab_kernel<14,392,4,2,...>"). */
uint64_t frcnn_launch_count(void);
const char* frcnn_last_roi_kernel(void);

/* How anchors reach a kernel: either an explicit [N,4] tensor, or generated in registers from the
 * base anchors (utils/basic_anchors.py:27-57: anchor[(y*W+x)*A+a] = base[a] + (x*s, y*s, x*s, y*s)). */
typedef struct frcnn_anchor_spec {
    const float* anchors; /* [N,4] device, or NULL to generate                                 */
    const float* base;    /* [A,4] device (used when anchors == NULL)                          */
    int32_t num_base;     /* A  (<= FRCNN_MAX_BASE_ANCHORS)                                    */
    int32_t feat_stride;
    int32_t height;       /* feature-map H                                                     */
    int32_t width;        /* feature-map W                                                     */
} frcnn_anchor_spec;

/* ---- anchors -------------------------------------------------------------------------------
 * generate_basic_anchor  utils/basic_anchors.py:11-23.  sizes_host[j] = fp32(base_size*scale_j),
 * ratios_host[i] = fp32(r_i), inv_ratios_host[i] = fp32(1/r_i) (host does the Python-double part);
 * out [R*S,4], row i*S+j = (-w/2,-h/2,w/2,h/2), h = size*sqrt(r), w = size*sqrt(1/r).            */
int frcnn_base_anchors(const float* ratios_host, const float* inv_ratios_host, int32_t num_ratios,
                       const float* sizes_host, int32_t num_sizes, float* out, frcnn_stream_t stream);
/* enumerate_shifted_anchor  utils/basic_anchors.py:27-57.  out [H*W*A,4].                        */
int frcnn_shifted_anchors(const float* base, int32_t num_base, int32_t feat_stride, int32_t height,
                          int32_t width, float* out, frcnn_stream_t stream);

/* ---- box math ------------------------------------------------------------------------------
 * loc2bbox  utils/loc_bbox_iou.py:29-61.  src [R,4], loc [R,4*groups] -> out [R,4*groups].       */
int frcnn_loc2bbox(const float* src, const float* loc, int64_t rows, int32_t groups, float* out,
                   frcnn_stream_t stream);
/* bbox2loc  utils/loc_bbox_iou.py:63-89.   src [R,4], dst [R,4] -> out [R,4].                    */
int frcnn_bbox2loc(const float* src, const float* dst, int64_t rows, float* out, frcnn_stream_t stream);
/* bbox_iou  utils/loc_bbox_iou.py:4-27.    a [Na,4], b [Nb,4] -> out [Na,Nb].                    */
int frcnn_bbox_iou(const float* a, const float* b, int64_t na, int64_t nb, float* out,
                   frcnn_stream_t stream);

/* ---- RPN proposal layer (batched over images) ----------------------------------------------
 * ProposalCreator.__call__  nets/rpn.py:36-70  +  the per-image loop of
 * RegionProposalNetwork.forward  nets/rpn.py:129-139  +  torchvision.ops.nms (nets/rpn.py:63).   */
typedef struct frcnn_proposal_params {
    int32_t batch;      /* B images                                                            */
    int32_t num_anchors;/* N per image                                                         */
    int32_t n_pre_nms;  /* <= 0: keep all valid                                                */
    int32_t n_post_nms; /* rows of the output per image                                        */
    float clip_x_max;   /* img_size[1] as the reference indexes it                             */
    float clip_y_max;   /* img_size[2]                                                         */
    float min_size;     /* fp32(min_size * scale)                                              */
    double nms_thresh;  /* compared as torchvision does: (double)iou > thresh                  */
    int32_t score_mode; /* 0: score is fg probability [B,N]; 1: score is logits [B,N,2] and the
                           kernel computes softmax(...)[1] (nets/rpn.py:115-118); 2: loc and score
                           are the RPN's conv outputs as cuDNN wrote them, loc [B,4A,H,W] and logits
                           [B,2A,H,W] (NCHW), read in place: the permute(0,2,3,1).contiguous() passes
                           of nets/rpn.py:107-113 never run (needs num_base / height / width in the
                           anchor spec; boxes and keys are bit-identical to mode 1 on the permuted
                           tensors)                                                            */
    int32_t boxes_are_decoded; /* 1: `loc` already holds decoded boxes (skip loc2bbox)          */
    int32_t nms_superblock;    /* 0 = library default                                           */
} frcnn_proposal_params;

size_t frcnn_proposals_workspace_bytes(const frcnn_proposal_params* p);
/* loc [B,N,4]; score [B,N] or [B,N,2]; outputs: rois [B,n_post,4]; roi_src [B,n_post] original
 * anchor index of every output row (nullable); n_keep [B] NMS survivors before pad/truncate,
 * capped at n_post (nullable); status [B] FRCNN_IMG_* (required).                                */
int frcnn_proposals(const frcnn_proposal_params* p, const frcnn_anchor_spec* anchors,
                    const float* loc, const float* score, float* rois, int32_t* roi_src,
                    int32_t* n_keep, int32_t* status, void* workspace, size_t workspace_bytes,
                    frcnn_stream_t stream);

/* The same pipeline, stage by stage (used by the stage-isolated parity tests and by users who want
 * the intermediates).  boxes [B,N,4] clipped; keys [B,N]: 0 = filtered out, else order-preserving
 * image of the score (bigger = better).                                                        */
int frcnn_decode_clip_score(const frcnn_proposal_params* p, const frcnn_anchor_spec* anchors,
                            const float* loc, const float* score, float* boxes, uint32_t* keys,
                            float* fg_out /* [B,N] nullable */, frcnn_stream_t stream);
size_t frcnn_topk_workspace_bytes(int32_t batch, int32_t n);
/* order [B,k_cap]: anchor indices by (key desc, index asc), -1 past n_sel; n_sel [B] =
 * min(#keys != 0, k_cap); sorted_boxes [B,k_cap,4] = boxes[order] (nullable together with boxes). */
int frcnn_topk_sorted(const uint32_t* keys, const float* boxes, int32_t batch, int32_t n,
                      int32_t k_cap, int32_t* order, int32_t* n_sel, float* sorted_boxes,
                      void* workspace, size_t workspace_bytes, frcnn_stream_t stream);
size_t frcnn_nms_sorted_workspace_bytes(int32_t batch, int32_t n_rows, int32_t keep_cap,
                                        int32_t superblock);
/* greedy NMS over boxes already in score order.  sorted_boxes [B,row_stride,4], n_sel [B] rows
 * valid per image; keep [B,keep_cap] positions (ascending), n_keep [B] (stops at keep_cap).      */
int frcnn_nms_sorted(const float* sorted_boxes, const int32_t* n_sel, int32_t batch,
                     int32_t row_stride, double thresh, int32_t keep_cap, int32_t superblock,
                     int32_t* keep, int32_t* n_keep, void* workspace, size_t workspace_bytes,
                     frcnn_stream_t stream);
/* torchvision.ops.nms(boxes, scores, thr) for one set (nets/rpn.py:63, nets/frcnn_training.py:454):
 * keep [n] int64 original indices, score-descending (stable); n_keep [1].                       */
size_t frcnn_nms_workspace_bytes(int32_t n);
int frcnn_nms(const float* boxes, const float* scores, int32_t n, double thresh, int64_t* keep,
              int32_t* n_keep, void* workspace, size_t workspace_bytes, frcnn_stream_t stream);

/* ---- training targets (batched over images) -------------------------------------------------
 * AnchorTargetCreator.__call__  nets/frcnn_training.py:29-101.  bbox [B,Gmax,4], n_gt [B] device;
 * outputs loc [B,N,4], label [B,N] int64 in {-1,0,1}, argmax [B,N] int32 (nullable).             */
typedef struct frcnn_anchor_target_params {
    int32_t batch, num_anchors, max_gt;
    int32_t n_sample;      /* 256 */
    float pos_iou_thresh;  /* 0.7 */
    float neg_iou_thresh;  /* 0.3 */
    int32_t n_pos;         /* int(pos_ratio * n_sample), host-computed */
} frcnn_anchor_target_params;
size_t frcnn_anchor_targets_workspace_bytes(const frcnn_anchor_target_params* p);
int frcnn_anchor_targets(const frcnn_anchor_target_params* p, const frcnn_anchor_spec* anchors,
                         const float* bbox, const int32_t* n_gt, float* loc, int64_t* label,
                         int32_t* argmax, void* workspace, size_t workspace_bytes,
                         frcnn_stream_t stream);

/* ProposalTargetCreator.__call__  nets/frcnn_training.py:122-177.  roi [B,R,4] (row stride R),
 * bbox [B,Gmax,4], gt_label [B,Gmax] int64, n_gt [B]; outputs sample_roi / gt_loc [B,n_sample,4],
 * out_label [B,n_sample] int64, n_out [B] valid rows, status [B].                               */
typedef struct frcnn_proposal_target_params {
    int32_t batch, num_roi, max_gt;
    int32_t n_sample;          /* 128 */
    int32_t pos_per_image;     /* int(n_sample * pos_ratio) */
    float pos_iou_thresh;      /* 0.5 */
    float neg_iou_thresh_high; /* 0.5 */
    float neg_iou_thresh_low;  /* 0.0 */
} frcnn_proposal_target_params;
int frcnn_proposal_targets(const frcnn_proposal_target_params* p, const float* roi,
                           const float* bbox, const int64_t* gt_label, const int32_t* n_gt,
                           float* sample_roi, float* gt_loc, int64_t* out_label, int32_t* n_out,
                           int32_t* status, frcnn_stream_t stream);

/* ---- after the head: detections --------------------------------------------------
 * Post-head decode  nets/frcnn_training.py:311-320.  roi [T,4], roi_cls_loc [T,4*C], roi_score [T,C]
 * (T = n*R rows), label [T] int64 or NULL: the loc row of class label[t] (the reference passes
 * gt_roi_label) or, with NULL, of the best class is applied to the RoI with loc2bbox;
 * cls_score / cls_index = torch.max(roi_score, dim=1) (first index on ties, NaN wins).
 * bad_label [1] int32 (device) is set when a label is outside [0,C) (the reference's IndexError).   */
int frcnn_detection_decode(const float* roi, const float* roi_cls_loc, const float* roi_score,
                           const int64_t* label, int64_t total, int32_t n_class, float* boxes,
                           float* cls_score, int64_t* cls_index, int32_t* bad_label,
                           frcnn_stream_t stream);
/* The evaluator's per-class NMS  nets/frcnn_training.py:441-454 (for each class c: torchvision nms on
 * the rows with classes == c), all classes and images in one launch: boxes [B,R,4], scores [B,R],
 * classes [B,R] int64 (NULL: one class = multi_inference.py:84), n_valid [B] (NULL: R rows each).
 * keep [B,R] int32: kept original row indices ordered by (score desc, index asc), -1 padded -- the rows
 * of class c, in that order, are the reference's keep list for c; n_keep [B].  R <= 1024 per image,
 * else FRCNN_ERR_UNSUPPORTED (frcnn_nms handles long single-class lists).                            */
int frcnn_nms_by_class(const float* boxes, const float* scores, const int64_t* classes,
                       const int32_t* n_valid, int32_t batch, int32_t rows, double iou_threshold,
                       int32_t* keep, int32_t* n_keep, frcnn_stream_t stream);

/* ---- RoI head gather -------------------------------------------------------------------------
 * HarNetRoIHead.forward coordinate map + index concat  nets/classify.py:29-38:
 * rois [n*R,4] image coords, roi_indices [n] (int32) -> rois5 [n*R,5] = (idx, x/d1*Wf, y/d0*Hf..). */
int frcnn_roi_head_coords(const float* rois, const int32_t* roi_indices, int32_t n_images,
                          int32_t rois_per_image, float img_size0, float img_size1, int32_t feat_h,
                          int32_t feat_w, float* rois5, frcnn_stream_t stream);

/* torchvision RoIPool forward (nets/classify.py:17,43).  feat [B,C,H,W]; rois5 [K,5];
 * out [K,C,PH,PW]; argmax [K,C,PH,PW] int32 (nullable; needed only for backward).
 * rois_per_image: 0 = RoIs in any order (bucketed by their batch index on the device); R > 0 = the
 * caller guarantees K == B*R and rows [b*R,(b+1)*R) belong to image b, as the head builds them
 * (nets/classify.py:38), which skips the bucketing pass.
 * The workspace (frcnn_roi_workspace_bytes) is needed either way: besides the bucketing arrays it holds
 * the per-RoI bin / sample geometry that the 7x7 / 14x14 kernels compute once per RoI instead of once
 * per channel slab.                                                                              */
size_t frcnn_roi_workspace_bytes(int32_t batch, int32_t num_rois);
int frcnn_roi_pool_forward(const float* feat, int32_t batch, int32_t channels, int32_t height,
                           int32_t width, const float* rois5, int32_t num_rois, int32_t rois_per_image,
                           int32_t pooled_h, int32_t pooled_w, float spatial_scale, float* out, int32_t* argmax,
                           void* workspace, size_t workspace_bytes, frcnn_stream_t stream);
/* RoIPool followed by the HarDNet head's classifier, which is only a global average over the bins
 * (models/hardnet.py:203-212 AdaptiveAvgPool2d(1)+Flatten, called at nets/classify.py:43-46):
 * out [K,C] = mean over the PH*PW bins of roi_pool(feat, rois5), without materialising [K,C,PH,PW].
 * Fixed summation order (run-to-run identical); agrees with pool().mean() to fp32 rounding.  7x7 and 14x14
 * bins on maps whose max tables fit in shared memory, else FRCNN_ERR_UNSUPPORTED (use the two steps).  */
int frcnn_roi_pool_mean_forward(const float* feat, int32_t batch, int32_t channels, int32_t height,
                                int32_t width, const float* rois5, int32_t num_rois, int32_t rois_per_image,
                                int32_t pooled_h, int32_t pooled_w, float spatial_scale, float* out,
                                void* workspace, size_t workspace_bytes, frcnn_stream_t stream);
/* The same fusion for RoIAlign (the RoIAlign 7x7 HarDNet configuration): out [K,C] = mean over the bins
 * of roi_align(feat, rois5, sampling_ratio > 0).  The mean is linear and separable in the features, so the
 * kernel reads every pixel of a RoI's window once with a per-axis weight instead of 4*sr*sr taps per bin.
 * Needs pooled*sampling_ratio <= 32 per side and a map of at most 64x64, else FRCNN_ERR_UNSUPPORTED.
 * Agrees with roi_align().mean() to fp32 rounding (1e-5 relative).                                       */
size_t frcnn_roi_align_mean_workspace_bytes(int32_t batch, int32_t num_rois);
int frcnn_roi_align_mean_forward(const float* feat, int32_t batch, int32_t channels, int32_t height,
                                 int32_t width, const float* rois5, int32_t num_rois, int32_t rois_per_image,
                                 int32_t pooled_h, int32_t pooled_w, float spatial_scale,
                                 int32_t sampling_ratio, int32_t aligned, float* out, void* workspace,
                                 size_t workspace_bytes, frcnn_stream_t stream);
/* Backward of torchvision RoIPool w.r.t. the features (the gradient the head's loss sends into the
 * extractor through nets/classify.py:43).  grad_in [B,C,H,W] must be zero-initialised by the caller; RoIs
 * whose batch index is outside [0,batch) contribute nothing.  Maps whose 4-channel slab fits in shared
 * memory: one CTA per (image, slab) accumulates there and adds to grad_in once (no global atomics).  */
int frcnn_roi_pool_backward(const float* grad_out, const int32_t* argmax, const float* rois5,
                            int32_t num_rois, int32_t batch, int32_t channels, int32_t height, int32_t width,
                            int32_t pooled_h, int32_t pooled_w, float* grad_in, frcnn_stream_t stream);
/* torchvision roi_align forward (BASELINE.json RoIAlign 7x7 configuration; the reference itself only calls
 * RoIPool, nets/classify.py:17,43).  exact = 1: the reference's operation order without FMA, bit-identical to
 * torchvision's CPU kernel.  exact = 0: the fast variant where one exists (sampling_ratio 2, 7x7 / 14x14) --
 * merged separable weights and FMA, within 1e-5 of the largest tap magnitude of the reference's value (the
 * tolerance north_star states for RoIAlign) -- and the exact kernels everywhere else.                   */
int frcnn_roi_align_forward(const float* feat, int32_t batch, int32_t channels, int32_t height,
                            int32_t width, const float* rois5, int32_t num_rois, int32_t rois_per_image,
                            int32_t pooled_h, int32_t pooled_w, float spatial_scale, int32_t sampling_ratio,
                            int32_t aligned, int32_t exact, float* out, void* workspace, size_t workspace_bytes,
                            frcnn_stream_t stream);
/* Backward of torchvision roi_align w.r.t. the features; grad_in [B,C,H,W] zero-initialised by the caller. */
int frcnn_roi_align_backward(const float* grad_out, const float* rois5, int32_t num_rois, int32_t batch,
                             int32_t channels, int32_t height, int32_t width, int32_t pooled_h,
                             int32_t pooled_w, float spatial_scale, int32_t sampling_ratio,
                             int32_t aligned, float* grad_in, frcnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FRCNN_B200_H_ */
