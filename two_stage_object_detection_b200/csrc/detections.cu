// detections.cu -- the step after the RoI head: class-specific box decode + best class
// (nets/frcnn_training.py:311-320) and the evaluator's per-class NMS (:441-454, torchvision nms on the
// rows of each class) / multi_inference.py:84's class-agnostic NMS, batched over images.
#include "common.cuh"

namespace frcnn {

// ---------------------------------------------------------------------------------------------
// decode: one warp per RoI.  Lanes stride over the C class scores (coalesced), keep (key, first index)
// of their maximum, a 64-bit shuffle tree picks the row maximum with torch.max's rules (NaN largest,
// first index on ties); lane 0 gathers the 4 loc values of the chosen class and applies loc2bbox.
// ---------------------------------------------------------------------------------------------
struct DetDecodeArgs {
    const float4* roi;        // [T]
    const float* cls_loc;     // [T, C*4]
    const float* score;       // [T, C]
    const long long* label;   // [T] or nullptr (use the best class)
    int total, n_class;
    float4* boxes;            // [T]
    float* cls_score;         // [T]
    long long* cls_index;     // [T]
    int* bad_label;           // set to 1 when a label is outside [0, C)
};

__global__ void __launch_bounds__(256) detection_decode_kernel(DetDecodeArgs a) {
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= a.total) return;
    const int lane = threadIdx.x & 31;
    const float* s = a.score + (size_t)t * a.n_class;
    unsigned long long best = 0ull;  // key << 32 | ~class: larger key first, then smaller class
    for (int c = lane; c < a.n_class; c += 32) {
        const unsigned long long v = ((unsigned long long)score_key(__ldg(s + c)) << 32) | (uint32_t)(~(uint32_t)c);
        best = v > best ? v : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, d);
        best = o > best ? o : best;
    }
    if (lane != 0) return;
    const int ci = (int)(~(uint32_t)(best & 0xFFFFFFFFull));
    int pick = ci;
    if (a.label) {
        const long long l = a.label[t];
        if (l < 0 || l >= a.n_class) {
            *a.bad_label = 1;  // the reference would raise IndexError on the gather
            pick = 0;
        } else {
            pick = (int)l;
        }
    }
    const float4 loc = __ldg(reinterpret_cast<const float4*>(a.cls_loc + ((size_t)t * a.n_class + pick) * 4));
    a.boxes[t] = decode_box(__ldg(a.roi + t), loc);
    a.cls_score[t] = __ldg(s + ci);
    a.cls_index[t] = ci;
}

// ---------------------------------------------------------------------------------------------
// per-class NMS: one CTA per image, everything in shared memory (R <= 1024 detections per image).
//   1. rank by (score desc, index asc) with an O(R^2) count -- R is the head's RoI count, a few hundred
//   2. IoU > thr bitmask, upper triangle, only between boxes of the same class
//   3. one warp resolves the greedy chain: lane w owns the removed-word of candidates 32w..32w+31
// Output: kept original row indices in score order, -1 padded, and their count.
// ---------------------------------------------------------------------------------------------
constexpr int DN_THREADS = 1024;
constexpr int DN_MAX = 1024;

struct DetNmsArgs {
    const float4* boxes;      // [B,R]
    const float* scores;      // [B,R]
    const long long* classes; // [B,R] or nullptr
    const int* n_valid;       // [B] or nullptr (all R rows)
    int R;
    float thr;
    int* keep;                // [B,R]
    int* n_keep;              // [B]
};

__global__ void __launch_bounds__(DN_THREADS) nms_by_class_kernel(DetNmsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = a.R, W32 = (R + 31) / 32;
    float4* sbox = reinterpret_cast<float4*>(smem_raw);            // sorted boxes
    float* sarea = reinterpret_cast<float*>(sbox + R);
    uint32_t* skey = reinterpret_cast<uint32_t*>(sarea + R);       // keys in original order
    int* scls = reinterpret_cast<int*>(skey + R);                  // sorted classes
    int* sorder = scls + R;                                        // sorted position -> original row
    uint32_t* mask = reinterpret_cast<uint32_t*>(sorder + R);      // [R][W32]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = a.n_valid ? min(max(a.n_valid[b], 0), R) : R;
    const float4* gb = a.boxes + (size_t)b * R;
    for (int i = tid; i < n; i += DN_THREADS) skey[i] = score_key(__ldg(a.scores + (size_t)b * R + i));
    __syncthreads();
    for (int i = tid; i < n; i += DN_THREADS) {
        const uint32_t k = skey[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const uint32_t kj = skey[j];
            rank += (kj > k) || (kj == k && j < i);
        }
        const float4 v = __ldg(gb + i);
        sbox[rank] = v;
        sarea[rank] = box_area(v);
        scls[rank] = a.classes ? (int)a.classes[(size_t)b * R + i] : 0;
        sorder[rank] = i;
    }
    __syncthreads();
    // word (i, w): which of the candidates 32w..32w+31 (later than i, same class) does candidate i suppress
    for (int it = tid; it < n * W32; it += DN_THREADS) {
        const int i = it / W32, w = it - i * W32;
        uint32_t bits = 0u;
        if (32 * w + 31 > i) {
            const float4 bi = sbox[i];
            const float ai = sarea[i];
            const int ci = scls[i];
            const int j0 = max(32 * w, i + 1), j1 = min(32 * w + 32, n);
            for (int j = j0; j < j1; ++j)
                if (scls[j] == ci && nms_suppresses(bi, ai, sbox[j], sarea[j], a.thr)) bits |= 1u << (j & 31);
        }
        mask[it] = bits;
    }
    __syncthreads();
    if (tid < 32) {
        uint32_t removed = 0u;  // lane w: candidates 32w..32w+31
        int kept = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t word = __shfl_sync(0xFFFFFFFFu, removed, i >> 5);
            if (!((word >> (i & 31)) & 1u)) {  // uniform
                if (tid == 0) a.keep[(size_t)b * R + kept] = sorder[i];
                ++kept;
                if (tid < W32) removed |= mask[i * W32 + tid];
            }
        }
        for (int i = kept + tid; i < R; i += 32) a.keep[(size_t)b * R + i] = -1;
        if (tid == 0) a.n_keep[b] = kept;
    }
}

static size_t det_nms_smem(int R) {
    const int W32 = (R + 31) / 32;
    return (size_t)R * (sizeof(float4) + sizeof(float) + sizeof(uint32_t) + 2 * sizeof(int)) +
           (size_t)R * W32 * sizeof(uint32_t);
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_detection_decode(const float* roi, const float* roi_cls_loc, const float* roi_score, const int64_t* label,
                           int64_t total, int32_t n_class, float* boxes, float* cls_score, int64_t* cls_index,
                           int32_t* bad_label, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(total >= 0 && total < (1ll << 31) && n_class > 0, "frcnn_detection_decode: bad shape");
    if (total == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(roi && roi_cls_loc && roi_score && boxes && cls_score && cls_index && bad_label,
                    "frcnn_detection_decode: null pointer");
    DetDecodeArgs a;
    a.roi = (const float4*)roi;
    a.cls_loc = roi_cls_loc;
    a.score = roi_score;
    a.label = (const long long*)label;
    a.total = (int)total;
    a.n_class = n_class;
    a.boxes = (float4*)boxes;
    a.cls_score = cls_score;
    a.cls_index = (long long*)cls_index;
    a.bad_label = bad_label;
    FRCNN_CUDA(cudaMemsetAsync(bad_label, 0, sizeof(int32_t), (cudaStream_t)stream));
    detection_decode_kernel<<<cdiv(total, 8), 256, 0, (cudaStream_t)stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_nms_by_class(const float* boxes, const float* scores, const int64_t* classes, const int32_t* n_valid,
                       int32_t batch, int32_t rows, double iou_threshold, int32_t* keep, int32_t* n_keep,
                       frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(batch >= 0 && rows >= 0, "frcnn_nms_by_class: bad shape");
    if (batch == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(keep && n_keep && ((boxes && scores) || rows == 0), "frcnn_nms_by_class: null pointer");
    if (rows > DN_MAX) {
        set_error("frcnn_nms_by_class: %d rows per image > %d (use frcnn_nms per class)", rows, DN_MAX);
        return FRCNN_ERR_UNSUPPORTED;
    }
    DetNmsArgs a;
    a.boxes = (const float4*)boxes;
    a.scores = scores;
    a.classes = (const long long*)classes;
    a.n_valid = n_valid;
    a.R = rows;
    a.thr = float_threshold(iou_threshold);
    a.keep = keep;
    a.n_keep = n_keep;
    const size_t smem = det_nms_smem(rows > 0 ? rows : 1);
    FRCNN_SMEM(nms_by_class_kernel, smem);
    nms_by_class_kernel<<<batch, DN_THREADS, smem, (cudaStream_t)stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

}  // extern "C"
