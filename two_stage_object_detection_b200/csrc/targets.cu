// targets.cu -- AnchorTargetCreator / ProposalTargetCreator (nets/frcnn_training.py:19-177) as
// batched sm_100a kernels: tiled IoU with the image's GT boxes staged in shared memory, fused
// row max/argmax, column argmax through packed 64-bit atomicMax, deterministic first-k sampling
// by prefix counts, and bbox2loc.  The reference's quirks are reproduced on purpose.
#include "common.cuh"

namespace frcnn {

constexpr int AT_THREADS = 256;
constexpr int AT_PER = 4;                       // anchors per thread, strided by AT_THREADS inside the tile
constexpr int AT_TILE = AT_THREADS * AT_PER;    // 1024 anchors per CTA
constexpr int AT_WARPS = AT_THREADS / 32;

struct AnchorTargetArgs {
    AnchorGen gen;
    const float4* bbox;  // [B,Gmax]
    const int* n_gt;     // [B]
    int batch, n, max_gt, tiles;
    int n_sample, n_pos;
    float pos_thr, neg_thr;
    // workspace
    int* packed;                 // [B,N]  argmax | class << 28 | forced << 30
    unsigned long long* colbest; // [B,Gmax] packed (key<<32 | ~anchor), zeroed per call
    int* done;                   // [B] tiles finished (zeroed per call, same memset as colbest)
    int* cnt_pos;                // [B,tiles]
    int* cnt_neg;                // [B,tiles]
    // outputs
    float4* loc;
    long long* label;
    int* argmax_out;
};

// one 32-bit word per anchor between the two launches (4 B written + 4 B read instead of max_iou + argmax)
constexpr int AT_FORCED_BIT = 0x40000000;  // this anchor is some GT's best anchor (frcnn_training.py:56-62)
constexpr int AT_CLS_SHIFT = 28;           // 0: ignore (-1), 1: negative (0), 2: positive (1), from the thresholds only
constexpr int AT_ARG_MASK = 0xFFFF;        // argmax over the GT boxes (max_gt <= 4096)

__device__ __forceinline__ int label_class(float max_iou, float pos_thr, float neg_thr) {
    int c = 0;
    if (max_iou < neg_thr) c = 1;   // frcnn_training.py:79
    if (max_iou >= pos_thr) c = 2;  // :80
    return c;
}

// K1: per anchor, row max / argmax over the image's GT boxes and the label class the thresholds give;
// per GT, the column argmax (largest IoU, then smallest anchor index: torch.argmax's first index) through
// REDUX + packed 64-bit atomicMax.  The CTA that finishes an image last applies "each GT's best anchor
// takes that GT; later GT wins" to at most G words and corrects the per-tile counts, so no kernel has to
// re-read every anchor for it.
__global__ void __launch_bounds__(AT_THREADS) anchor_iou_kernel(AnchorTargetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sgt = reinterpret_cast<float4*>(smem_raw);
    float* sarea = reinterpret_cast<float*>(sgt + a.max_gt);
    unsigned long long* sbest = reinterpret_cast<unsigned long long*>(sarea + ((a.max_gt + 1) & ~1));
    __shared__ int s_cnt[2][AT_WARPS];
    __shared__ int s_last;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int G = min(a.n_gt[b], a.max_gt);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int g = tid; g < G; g += AT_THREADS) {
        float4 v = __ldg(a.bbox + (size_t)b * a.max_gt + g);
        sgt[g] = v;
        sarea[g] = box_area(v);
        sbest[g] = 0ull;
    }
    __syncthreads();
    const int i0 = tile * AT_TILE + tid;
    float4 an[AT_PER];
    float aa[AT_PER];
    uint32_t kbest[AT_PER];  // row max as an order-preserving key: NaN largest, first index wins ties (torch.max)
    int besti[AT_PER];
#pragma unroll
    for (int k = 0; k < AT_PER; ++k) {
        const int i = i0 + k * AT_THREADS;
        an[k] = i < a.n ? load_anchor(a.gen, i) : make_float4(0.f, 0.f, 0.f, 0.f);
        aa[k] = box_area(an[k]);
        kbest[k] = 0u;
        besti[k] = 0;
    }
    for (int g = 0; g < G; ++g) {
        const float4 gt = sgt[g];
        const float ga = sarea[g];
        uint32_t kmax = 0u;
        int imax = 0x7FFFFFFF;
#pragma unroll
        for (int k = 0; k < AT_PER; ++k) {
            const int i = i0 + k * AT_THREADS;
            uint32_t key = score_key(iou_eps(an[k], aa[k], gt, ga));
            if (key > kbest[k]) {  // keys are never 0: g = 0 always enters
                kbest[k] = key;
                besti[k] = g;
            }
            if (i >= a.n) key = 0u;
            if (key > kmax) {  // ascending i: the first of equal keys stays
                kmax = key;
                imax = i;
            }
        }
        // column argmax: biggest key, then smallest anchor index
        const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, kmax);
        if (m != 0u) {
            const int cand = kmax == m ? imax : 0x7FFFFFFF;
            const int first = __reduce_min_sync(0xFFFFFFFFu, cand);
            if (cand == first) atomicMax(&sbest[g], ((unsigned long long)m << 32) | (uint32_t)(~(uint32_t)first));
        }
    }
    int pos = 0, neg = 0;
#pragma unroll
    for (int k = 0; k < AT_PER; ++k) {
        const int i = i0 + k * AT_THREADS;
        if (i < a.n) {
            const int c = label_class(G > 0 ? key_score(kbest[k]) : 0.f, a.pos_thr, a.neg_thr);
            pos += c == 2;
            neg += c == 1;
            a.packed[(size_t)b * a.n + i] = besti[k] | (c << AT_CLS_SHIFT);
        }
    }
    pos = __reduce_add_sync(0xFFFFFFFFu, pos);
    neg = __reduce_add_sync(0xFFFFFFFFu, neg);
    if (lane == 0) {
        s_cnt[0][warp] = pos;
        s_cnt[1][warp] = neg;
    }
    __syncthreads();  // sbest complete, s_cnt visible
    for (int g = tid; g < G; g += AT_THREADS)
        if (sbest[g]) atomicMax(a.colbest + (size_t)b * a.max_gt + g, sbest[g]);
    if (tid < 2) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < AT_WARPS; ++w) t += s_cnt[tid][w];
        (tid == 0 ? a.cnt_pos : a.cnt_neg)[b * a.tiles + tile] = t;
    }
    // last CTA of the image: everything the others wrote is visible after their fence + ticket
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.done + b, 1) == a.tiles - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int g = tid; g < G; g += AT_THREADS) {
        const unsigned long long p = __ldcg(a.colbest + (size_t)b * a.max_gt + g);
        if (p == 0ull) continue;
        const int anchor = (int)(~(uint32_t)(p & 0xFFFFFFFFull));
        // forced words outrank every unforced one; among forced ones the larger g (the later GT) wins
        const int old = atomicMax(a.packed + (size_t)b * a.n + anchor, g | AT_FORCED_BIT);
        if (!(old & AT_FORCED_BIT)) {  // first GT to claim this anchor: its label becomes 1 (:82)
            const int c = (old >> AT_CLS_SHIFT) & 3, t = anchor / AT_TILE;
            if (c != 2) atomicAdd(a.cnt_pos + b * a.tiles + t, 1);
            if (c == 1) atomicSub(a.cnt_neg + b * a.tiles + t, 1);
        }
    }
}

// K2: first-k cap on positives, the len()-of-a-tuple negative rule, bbox2loc for every anchor
__global__ void __launch_bounds__(AT_THREADS) anchor_label_kernel(AnchorTargetArgs a) {
    __shared__ int s_red[4][AT_WARPS];
    __shared__ int s_wpos[AT_PER * AT_WARPS], s_wneg[AT_PER * AT_WARPS];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // totals and prefixes over tiles
    int pp = 0, pn = 0, tp = 0, tn = 0;
    for (int t = tid; t < a.tiles; t += AT_THREADS) {
        int cp = a.cnt_pos[b * a.tiles + t], cn = a.cnt_neg[b * a.tiles + t];
        tp += cp;
        tn += cn;
        if (t < tile) {
            pp += cp;
            pn += cn;
        }
    }
    pp = __reduce_add_sync(0xFFFFFFFFu, pp);
    pn = __reduce_add_sync(0xFFFFFFFFu, pn);
    tp = __reduce_add_sync(0xFFFFFFFFu, tp);
    tn = __reduce_add_sync(0xFFFFFFFFu, tn);
    if (lane == 0) {
        s_red[0][warp] = pp;
        s_red[1][warp] = pn;
        s_red[2][warp] = tp;
        s_red[3][warp] = tn;
    }
    const int G = min(a.n_gt[b], a.max_gt);
    const int i0 = tile * AT_TILE + tid;
    int am[AT_PER], l[AT_PER];
    uint32_t bp[AT_PER], bn[AT_PER];
#pragma unroll
    for (int k = 0; k < AT_PER; ++k) {
        const int i = i0 + k * AT_THREADS;
        const int w = i < a.n ? __ldcg(a.packed + (size_t)b * a.n + i) : 0;
        const int c = (w & AT_FORCED_BIT) ? 2 : ((w >> AT_CLS_SHIFT) & 3);
        am[k] = w & AT_ARG_MASK;
        l[k] = i < a.n ? c - 1 : -1;
        bp[k] = __ballot_sync(0xFFFFFFFFu, l[k] == 1);
        bn[k] = __ballot_sync(0xFFFFFFFFu, l[k] == 0);
        if (lane == 0) {
            s_wpos[k * AT_WARPS + warp] = __popc(bp[k]);
            s_wneg[k * AT_WARPS + warp] = __popc(bn[k]);
        }
    }
    __syncthreads();
    pp = pn = tp = tn = 0;
#pragma unroll
    for (int w = 0; w < AT_WARPS; ++w) {
        pp += s_red[0][w];
        pn += s_red[1][w];
        tp += s_red[2][w];
        tn += s_red[3][w];
    }
    // exclusive ranks inside the tile: chunk k of the tile precedes chunk k+1, warps in order inside a chunk
    static_assert(AT_PER * AT_WARPS == 32, "one warp scans the per-(chunk,warp) counts");
    if (warp == 0) {
        int vp = s_wpos[lane], vn = s_wneg[lane];
        int ip = vp, in = vn;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int up = __shfl_up_sync(0xFFFFFFFFu, ip, d), un = __shfl_up_sync(0xFFFFFFFFu, in, d);
            if (lane >= d) {
                ip += up;
                in += un;
            }
        }
        s_wpos[lane] = ip - vp;
        s_wneg[lane] = in - vn;
    }
    __syncthreads();
    const int pos_len = tp > a.n_pos ? a.n_pos : tp;
    const int n_neg = a.n_sample - pos_len;
    const int neg_start = n_neg == 0 ? 0 : max(tn + n_neg, 0);
    const bool any_pos = pos_len > 0 && G > 0;  // (label > 0).any()
#pragma unroll
    for (int k = 0; k < AT_PER; ++k) {
        const int i = i0 + k * AT_THREADS;
        if (i >= a.n) continue;
        const int rp = pp + s_wpos[k * AT_WARPS + warp] + __popc(bp[k] & lanemask_lt());
        const int rn = pn + s_wneg[k * AT_WARPS + warp] + __popc(bn[k] & lanemask_lt());
        int lab = l[k];
        if (lab == 1 && tp > a.n_pos && rp >= a.n_pos) lab = -1;  // frcnn_training.py:85-91
        if (lab == 0 && 1 > n_neg && rn >= neg_start) lab = -1;   // :96-99: len(neg_index) is 1
        const size_t o = (size_t)b * a.n + i;
        __stcs(a.label + o, (long long)lab);
        if (a.argmax_out) __stcs(a.argmax_out + o, am[k]);
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (any_pos) out = encode_box(load_anchor(a.gen, i), __ldg(a.bbox + (size_t)b * a.max_gt + am[k]));
        __stcs(a.loc + o, out);
    }
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
    int* packed;
    unsigned long long* colbest;
    int* done;
    size_t zero_bytes;  // colbest and done are one region, cleared by one memset
    int* cnt_pos;
    int* cnt_neg;
};

static size_t at_layout(Workspace& ws, const frcnn_anchor_target_params* p, AtWs* out) {
    int tiles = cdiv(p->num_anchors, AT_TILE);
    AtWs w;
    w.packed = ws.take<int>((size_t)p->batch * p->num_anchors);
    const size_t ncol = (size_t)p->batch * (p->max_gt > 0 ? p->max_gt : 1);
    w.zero_bytes = ncol * sizeof(unsigned long long) + (size_t)p->batch * sizeof(int);
    w.colbest = reinterpret_cast<unsigned long long*>(ws.take<unsigned char>(w.zero_bytes));
    w.done = reinterpret_cast<int*>(w.colbest + ncol);
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
    a.gen = make_anchor_gen(anchors->anchors, anchors->base, anchors->num_base, anchors->feat_stride,
                            anchors->height, anchors->width);
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
    a.packed = w.packed;
    a.colbest = w.colbest;
    a.done = w.done;
    a.cnt_pos = w.cnt_pos;
    a.cnt_neg = w.cnt_neg;
    a.loc = (float4*)loc;
    a.label = (long long*)label;
    a.argmax_out = argmax;
    FRCNN_CUDA(cudaMemsetAsync(w.colbest, 0, w.zero_bytes, stream));
    size_t smem = at_smem(p->max_gt);
    if (smem > 48 * 1024)
        FRCNN_SMEM(anchor_iou_kernel, smem);
    dim3 grid(a.tiles, p->batch);
    anchor_iou_kernel<<<grid, AT_THREADS, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    anchor_label_kernel<<<grid, AT_THREADS, 0, stream>>>(a);
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
        FRCNN_SMEM(proposal_target_kernel, smem);
    proposal_target_kernel<<<p->batch, PT_THREADS, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

}  // extern "C"
