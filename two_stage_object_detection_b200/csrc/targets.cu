// targets.cu -- AnchorTargetCreator / ProposalTargetCreator (nets/frcnn_training.py:19-177) as
// batched sm_100a kernels: tiled IoU with the image's GT boxes staged in shared memory, fused
// row max/argmax, column argmax through packed 64-bit atomicMax, deterministic first-k sampling
// by prefix counts, and bbox2loc.  The reference's quirks are reproduced on purpose.
#include "common.cuh"

namespace frcnn {

constexpr int AT_TILE = 256;

struct AnchorTargetArgs {
    AnchorGen gen;
    const float4* bbox;  // [B,Gmax]
    const int* n_gt;     // [B]
    int batch, n, max_gt, tiles;
    int n_sample, n_pos;
    float pos_thr, neg_thr;
    // workspace
    float* max_iou;              // [B,N]
    int* argmax;                 // [B,N]  (bit 30 = forced by a GT's best anchor)
    unsigned long long* colbest; // [B,Gmax] packed (key<<32 | ~anchor)
    int* cnt_pos;                // [B,tiles]
    int* cnt_neg;                // [B,tiles]
    // outputs
    float4* loc;
    long long* label;
    int* argmax_out;
};

constexpr int FORCED_BIT = 0x40000000;

// K1: per anchor row max / argmax over the image's GT boxes; per GT column argmax (first index).
__global__ void __launch_bounds__(AT_TILE) anchor_iou_kernel(AnchorTargetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sgt = reinterpret_cast<float4*>(smem_raw);
    float* sarea = reinterpret_cast<float*>(sgt + a.max_gt);
    unsigned long long* sbest = reinterpret_cast<unsigned long long*>(sarea + ((a.max_gt + 1) & ~1));
    const int b = blockIdx.y, tile = blockIdx.x;
    const int G = min(a.n_gt[b], a.max_gt);
    const int lane = threadIdx.x & 31;
    for (int g = threadIdx.x; g < G; g += AT_TILE) {
        float4 v = __ldg(a.bbox + (size_t)b * a.max_gt + g);
        sgt[g] = v;
        sarea[g] = box_area(v);
        sbest[g] = 0ull;
    }
    __syncthreads();
    const int i = tile * AT_TILE + threadIdx.x;
    const bool valid = i < a.n;
    float4 an = valid ? load_anchor(a.gen, i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float aa = box_area(an);
    float best = 0.f;
    int besti = 0;
    for (int g = 0; g < G; ++g) {
        float v = iou_eps(an, aa, sgt[g], sarea[g]);
        if (g == 0 || beats(v, best)) {
            best = v;
            besti = g;
        }
        // column argmax: biggest key, then smallest anchor index (lanes are in index order)
        uint32_t key = valid ? score_key(v) : 0u;
        uint32_t m = __reduce_max_sync(0xFFFFFFFFu, key);
        uint32_t who = __ballot_sync(0xFFFFFFFFu, key == m);
        if (m != 0u && lane == __ffs(who) - 1) {
            unsigned long long packed = ((unsigned long long)m << 32) | (uint32_t)(~(uint32_t)i);
            atomicMax(&sbest[g], packed);
        }
    }
    if (valid) {
        a.max_iou[(size_t)b * a.n + i] = best;
        a.argmax[(size_t)b * a.n + i] = besti;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += AT_TILE)
        if (sbest[g]) atomicMax(a.colbest + (size_t)b * a.max_gt + g, sbest[g]);
}

__device__ __forceinline__ int label_before_cap(float max_iou, bool forced, float pos_thr, float neg_thr) {
    int l = -1;
    if (max_iou < neg_thr) l = 0;
    if (max_iou >= pos_thr) l = 1;
    if (forced) l = 1;
    return l;
}

__device__ __forceinline__ int block_sum_256(int v, int* scratch) {
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < AT_TILE / 32; ++w) t += scratch[w];
    __syncthreads();
    return t;
}

// K2: "each GT's best anchor takes that GT; later GT wins" (frcnn_training.py:60-62, :82) + counts
__global__ void __launch_bounds__(AT_TILE) anchor_force_count_kernel(AnchorTargetArgs a) {
    __shared__ int forced[AT_TILE];
    __shared__ int scratch[AT_TILE / 32];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int G = min(a.n_gt[b], a.max_gt);
    const int t0 = tile * AT_TILE;
    forced[threadIdx.x] = -1;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += AT_TILE) {
        unsigned long long p = a.colbest[(size_t)b * a.max_gt + g];
        int anchor = (int)(~(uint32_t)(p & 0xFFFFFFFFull));
        if (p != 0ull && anchor >= t0 && anchor < t0 + AT_TILE) atomicMax(&forced[anchor - t0], g);
    }
    __syncthreads();
    const int i = t0 + threadIdx.x;
    int pos = 0, neg = 0;
    if (i < a.n) {
        size_t o = (size_t)b * a.n + i;
        int f = forced[threadIdx.x];
        if (f >= 0) a.argmax[o] = f | FORCED_BIT;
        int l = label_before_cap(a.max_iou[o], f >= 0, a.pos_thr, a.neg_thr);
        pos = l == 1;
        neg = l == 0;
    }
    int tp = block_sum_256(pos, scratch);
    int tn = block_sum_256(neg, scratch);
    if (threadIdx.x == 0) {
        a.cnt_pos[b * a.tiles + tile] = tp;
        a.cnt_neg[b * a.tiles + tile] = tn;
    }
}

// K3: first-k cap on positives, the len()-of-a-tuple negative rule, bbox2loc for every anchor
__global__ void __launch_bounds__(AT_TILE) anchor_label_kernel(AnchorTargetArgs a) {
    __shared__ int scratch[AT_TILE / 32];
    __shared__ int wsum_pos[AT_TILE / 32], wsum_neg[AT_TILE / 32];
    __shared__ int s_pre_pos, s_pre_neg, s_tot_pos, s_tot_neg;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // totals and prefixes over tiles
    int pp = 0, pn = 0, tp = 0, tn = 0;
    for (int t = threadIdx.x; t < a.tiles; t += AT_TILE) {
        int cp = a.cnt_pos[b * a.tiles + t], cn = a.cnt_neg[b * a.tiles + t];
        tp += cp;
        tn += cn;
        if (t < tile) {
            pp += cp;
            pn += cn;
        }
    }
    pp = block_sum_256(pp, scratch);
    pn = block_sum_256(pn, scratch);
    tp = block_sum_256(tp, scratch);
    tn = block_sum_256(tn, scratch);
    if (threadIdx.x == 0) {
        s_pre_pos = pp;
        s_pre_neg = pn;
        s_tot_pos = tp;
        s_tot_neg = tn;
    }
    __syncthreads();
    const int G = min(a.n_gt[b], a.max_gt);
    const int i = tile * AT_TILE + threadIdx.x;
    const bool valid = i < a.n;
    size_t o = (size_t)b * a.n + (valid ? i : 0);
    int am = valid ? a.argmax[o] : 0;
    bool forced = (am & FORCED_BIT) != 0;
    am &= ~FORCED_BIT;
    int l = valid ? label_before_cap(a.max_iou[o], forced, a.pos_thr, a.neg_thr) : -1;
    // exclusive ranks inside the tile
    uint32_t bp = __ballot_sync(0xFFFFFFFFu, l == 1), bn = __ballot_sync(0xFFFFFFFFu, l == 0);
    if (lane == 0) {
        wsum_pos[warp] = __popc(bp);
        wsum_neg[warp] = __popc(bn);
    }
    __syncthreads();
    int rp = s_pre_pos + __popc(bp & lanemask_lt()), rn = s_pre_neg + __popc(bn & lanemask_lt());
    for (int w = 0; w < warp; ++w) {
        rp += wsum_pos[w];
        rn += wsum_neg[w];
    }
    const int tot_pos = s_tot_pos, tot_neg = s_tot_neg;
    if (l == 1 && tot_pos > a.n_pos && rp >= a.n_pos) l = -1;  // frcnn_training.py:85-91
    const int pos_len = tot_pos > a.n_pos ? a.n_pos : tot_pos;
    const int n_neg = a.n_sample - pos_len;
    if (l == 0 && 1 > n_neg) {  // frcnn_training.py:96-99: len(neg_index) is 1
        int start = n_neg == 0 ? 0 : max(tot_neg + n_neg, 0);
        if (rn >= start) l = -1;
    }
    if (!valid) return;
    a.label[o] = (long long)l;
    if (a.argmax_out) a.argmax_out[o] = am;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos_len > 0 && G > 0) {  // (label > 0).any()
        float4 an = load_anchor(a.gen, i);
        float4 gt = __ldg(a.bbox + (size_t)b * a.max_gt + am);
        out = encode_box(an, gt);
    }
    a.loc[o] = out;
}

// ---------------------------------------------------------------------------------------------
// ProposalTargetCreator: one CTA per image
// ---------------------------------------------------------------------------------------------
constexpr int PT_THREADS = 512;
constexpr int PT_MAX_SAMPLE = 1024;

struct ProposalTargetArgs {
    const float4* roi;   // [B,R0]
    const float4* bbox;  // [B,Gmax]
    const long long* gt_label;
    const int* n_gt;
    int batch, n_roi, max_gt, n_sample, pos_per_image;
    float pos_thr, neg_hi, neg_lo;
    float4* sample_roi;
    float4* gt_loc;
    long long* out_label;
    int* n_out;
    int* status;
};

__device__ __forceinline__ float4 pt_row(const ProposalTargetArgs& a, int b, int r) {
    return r < a.n_roi ? __ldg(a.roi + (size_t)b * a.n_roi + r)
                       : __ldg(a.bbox + (size_t)b * a.max_gt + (r - a.n_roi));
}

__device__ __forceinline__ void pt_best(const float4& bx, const float4* sgt, const float* sarea, int G,
                                        float& best, int& besti) {
    float aa = box_area(bx);
    best = 0.f;
    besti = 0;
    for (int g = 0; g < G; ++g) {
        float v = iou_eps(bx, aa, sgt[g], sarea[g]);
        if (g == 0 || beats(v, best)) {
            best = v;
            besti = g;
        }
    }
}

__global__ void __launch_bounds__(PT_THREADS) proposal_target_kernel(ProposalTargetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sgt = reinterpret_cast<float4*>(smem_raw);
    float* sarea = reinterpret_cast<float*>(sgt + a.max_gt);
    __shared__ int pos_list[PT_MAX_SAMPLE], neg_list[PT_MAX_SAMPLE];
    __shared__ int wpos[PT_THREADS / 32], wneg[PT_THREADS / 32];
    __shared__ int s_pos, s_neg, s_err;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = min(a.n_gt[b], a.max_gt);
    const int R = a.n_roi + G;  // torch.cat((roi, bbox))
    for (int g = tid; g < G; g += PT_THREADS) {
        float4 v = __ldg(a.bbox + (size_t)b * a.max_gt + g);
        sgt[g] = v;
        sarea[g] = box_area(v);
    }
    if (tid == 0) {
        s_pos = 0;
        s_neg = 0;
        s_err = 0;
    }
    __syncthreads();
    // ordered compaction of positives / negatives (first-k by index)
    for (int base = 0; base < R; base += PT_THREADS) {
        int r = base + tid;
        bool isp = false, isn = false;
        if (r < R) {
            float best;
            int besti;
            pt_best(pt_row(a, b, r), sgt, sarea, G, best, besti);
            isp = best >= a.pos_thr;
            isn = (best < a.neg_hi) && (best >= a.neg_lo);
        }
        uint32_t bp = __ballot_sync(0xFFFFFFFFu, isp), bn = __ballot_sync(0xFFFFFFFFu, isn);
        if (lane == 0) {
            wpos[warp] = __popc(bp);
            wneg[warp] = __popc(bn);
        }
        __syncthreads();
        int op = s_pos + __popc(bp & lanemask_lt()), on = s_neg + __popc(bn & lanemask_lt());
        int tp = 0, tn = 0;
        for (int w = 0; w < PT_THREADS / 32; ++w) {
            if (w < warp) {
                op += wpos[w];
                on += wneg[w];
            }
            tp += wpos[w];
            tn += wneg[w];
        }
        if (isp && op < a.pos_per_image) pos_list[op] = r;
        if (isn && on < a.n_sample) neg_list[on] = r;
        __syncthreads();
        if (tid == 0) {
            s_pos += tp;
            s_neg += tn;
        }
        __syncthreads();
    }
    const int n_pos = min(s_pos, a.pos_per_image);
    const int n_neg = min(s_neg, max(a.n_sample - n_pos, 0));
    const int n_keep = n_pos + n_neg;
    float4* o_roi = a.sample_roi + (size_t)b * a.n_sample;
    float4* o_loc = a.gt_loc + (size_t)b * a.n_sample;
    long long* o_lab = a.out_label + (size_t)b * a.n_sample;
    for (int j = tid; j < a.n_sample; j += PT_THREADS) {
        float4 sr = make_float4(0.f, 0.f, 0.f, 0.f), gl = sr;
        long long lab = -1;
        if (j < n_keep) {
            int r = j < n_pos ? pos_list[j] : neg_list[j - n_pos];
            sr = pt_row(a, b, r);
            lab = 0;
            if (G > 0) {
                float best;
                int besti;
                pt_best(sr, sgt, sarea, G, best, besti);
                gl = encode_box(sr, sgt[besti]);
                lab = a.gt_label[(size_t)b * a.max_gt + besti] + 1;
            }
        }
        o_roi[j] = sr;
        o_loc[j] = gl;
        o_lab[j] = lab;
    }
    __syncthreads();
    // frcnn_training.py:173-175: gt_roi_label[neg_index] = 0 with ORIGINAL indices on the sampled array
    if (G > 0) {
        for (int j = tid; j < n_neg; j += PT_THREADS) {
            int v = neg_list[j];
            if (v < n_keep) o_lab[v] = 0;
            else s_err = 1;
        }
    }
    __syncthreads();
    if (tid == 0) {
        a.n_out[b] = n_keep;
        a.status[b] = s_err ? FRCNN_IMG_SCATTER_INDEX_ERROR : FRCNN_IMG_OK;
    }
}

static size_t at_smem(int max_gt) {
    return (size_t)max_gt * sizeof(float4) + (size_t)((max_gt + 1) & ~1) * sizeof(float) +
           (size_t)max_gt * sizeof(unsigned long long);
}

struct AtWs {
    float* max_iou;
    int* argmax;
    unsigned long long* colbest;
    int* cnt_pos;
    int* cnt_neg;
};

static size_t at_layout(Workspace& ws, const frcnn_anchor_target_params* p, AtWs* out) {
    int tiles = cdiv(p->num_anchors, AT_TILE);
    AtWs w;
    w.max_iou = ws.take<float>((size_t)p->batch * p->num_anchors);
    w.argmax = ws.take<int>((size_t)p->batch * p->num_anchors);
    w.colbest = ws.take<unsigned long long>((size_t)p->batch * (p->max_gt > 0 ? p->max_gt : 1));
    w.cnt_pos = ws.take<int>((size_t)p->batch * tiles);
    w.cnt_neg = ws.take<int>((size_t)p->batch * tiles);
    if (out) *out = w;
    return ws.off;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

size_t frcnn_anchor_targets_workspace_bytes(const frcnn_anchor_target_params* p) {
    if (!p || p->batch <= 0 || p->num_anchors <= 0) return 0;
    Workspace ws(nullptr, 0);
    return at_layout(ws, p, nullptr);
}

int frcnn_anchor_targets(const frcnn_anchor_target_params* p, const frcnn_anchor_spec* anchors,
                         const float* bbox, const int32_t* n_gt, float* loc, int64_t* label, int32_t* argmax,
                         void* workspace, size_t workspace_bytes, frcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FRCNN_CHECK_ARG(p && p->batch > 0 && p->num_anchors > 0 && p->max_gt >= 0, "frcnn_anchor_targets: bad shape");
    FRCNN_CHECK_ARG(p->max_gt <= 4096, "frcnn_anchor_targets: max_gt %d > 4096", p->max_gt);
    FRCNN_CHECK_ARG(n_gt && loc && label && (bbox || p->max_gt == 0), "frcnn_anchor_targets: null pointer");
    if (!anchors) {
        set_error("frcnn_anchor_targets: anchor spec is null");
        return FRCNN_ERR_INVALID_ARG;
    }
    if (!anchors->anchors) {
        FRCNN_CHECK_ARG(anchors->base && anchors->num_base > 0 && anchors->width > 0 &&
                            (int64_t)anchors->num_base * anchors->height * anchors->width == p->num_anchors,
                        "frcnn_anchor_targets: generated anchors need H*W*A == N");
    }
    Workspace ws(workspace, workspace_bytes);
    AtWs w;
    at_layout(ws, p, &w);
    if (!ws.ok()) {
        set_error("frcnn_anchor_targets: workspace too small or misaligned (%zu needed, %zu given)", ws.off,
                  workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    AnchorTargetArgs a;
    memset(&a, 0, sizeof(a));
    a.gen.anchors = (const float4*)anchors->anchors;
    a.gen.base = (const float4*)anchors->base;
    a.gen.num_base = anchors->num_base;
    a.gen.stride = anchors->feat_stride;
    a.gen.height = anchors->height;
    a.gen.width = anchors->width;
    a.bbox = (const float4*)bbox;
    a.n_gt = n_gt;
    a.batch = p->batch;
    a.n = p->num_anchors;
    a.max_gt = p->max_gt;
    a.tiles = cdiv(p->num_anchors, AT_TILE);
    a.n_sample = p->n_sample;
    a.n_pos = p->n_pos;
    a.pos_thr = p->pos_iou_thresh;
    a.neg_thr = p->neg_iou_thresh;
    a.max_iou = w.max_iou;
    a.argmax = w.argmax;
    a.colbest = w.colbest;
    a.cnt_pos = w.cnt_pos;
    a.cnt_neg = w.cnt_neg;
    a.loc = (float4*)loc;
    a.label = (long long*)label;
    a.argmax_out = argmax;
    FRCNN_CUDA(cudaMemsetAsync(w.colbest, 0, sizeof(unsigned long long) * p->batch * (p->max_gt > 0 ? p->max_gt : 1),
                               stream));
    size_t smem = at_smem(p->max_gt);
    if (smem > 48 * 1024)
        FRCNN_CUDA(cudaFuncSetAttribute(anchor_iou_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(a.tiles, p->batch);
    anchor_iou_kernel<<<grid, AT_TILE, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    anchor_force_count_kernel<<<grid, AT_TILE, 0, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    anchor_label_kernel<<<grid, AT_TILE, 0, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_proposal_targets(const frcnn_proposal_target_params* p, const float* roi, const float* bbox,
                           const int64_t* gt_label, const int32_t* n_gt, float* sample_roi, float* gt_loc,
                           int64_t* out_label, int32_t* n_out, int32_t* status, frcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FRCNN_CHECK_ARG(p && p->batch > 0 && p->num_roi >= 0 && p->max_gt >= 0, "frcnn_proposal_targets: bad shape");
    FRCNN_CHECK_ARG(p->n_sample > 0 && p->n_sample <= PT_MAX_SAMPLE && p->pos_per_image >= 0 &&
                        p->pos_per_image <= PT_MAX_SAMPLE,
                    "frcnn_proposal_targets: n_sample must be in [1,%d]", PT_MAX_SAMPLE);
    FRCNN_CHECK_ARG(p->max_gt <= 4096, "frcnn_proposal_targets: max_gt %d > 4096", p->max_gt);
    FRCNN_CHECK_ARG(n_gt && sample_roi && gt_loc && out_label && n_out && status &&
                        (roi || p->num_roi == 0) && ((bbox && gt_label) || p->max_gt == 0),
                    "frcnn_proposal_targets: null pointer");
    ProposalTargetArgs a;
    a.roi = (const float4*)roi;
    a.bbox = (const float4*)bbox;
    a.gt_label = (const long long*)gt_label;
    a.n_gt = n_gt;
    a.batch = p->batch;
    a.n_roi = p->num_roi;
    a.max_gt = p->max_gt;
    a.n_sample = p->n_sample;
    a.pos_per_image = p->pos_per_image;
    a.pos_thr = p->pos_iou_thresh;
    a.neg_hi = p->neg_iou_thresh_high;
    a.neg_lo = p->neg_iou_thresh_low;
    a.sample_roi = (float4*)sample_roi;
    a.gt_loc = (float4*)gt_loc;
    a.out_label = (long long*)out_label;
    a.n_out = n_out;
    a.status = status;
    size_t smem = (size_t)p->max_gt * (sizeof(float4) + sizeof(float));
    if (smem > 32 * 1024)
        FRCNN_CUDA(cudaFuncSetAttribute(proposal_target_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    proposal_target_kernel<<<p->batch, PT_THREADS, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

}  // extern "C"
