// boxmath.cu -- anchors (utils/basic_anchors.py) and box arithmetic (utils/loc_bbox_iou.py) of the
// reference as sm_100a kernels, plus the library's runtime helpers.
#include <stdarg.h>
#include <algorithm>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace frcnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::mutex g_smem_mu;
static std::map<std::pair<int, const void*>, size_t> g_smem_set;
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_smem_mu);
    size_t& have = g_smem_set[std::make_pair(dev, func)];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

static thread_local char g_roi_kernel[160] = "";
void note_roi_kernel(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_roi_kernel, sizeof(g_roi_kernel), fmt, ap);
    va_end(ap);
    // names come from stringified template argument lists: fold the one negated literal they contain
    for (const char* pat : {"!(true)", "!(false)"}) {
        const char* to = pat[2] == 't' ? "false" : "true";
        for (char* q = strstr(g_roi_kernel, pat); q; q = strstr(q, pat)) {
            const size_t lp = strlen(pat), lt = strlen(to);
            memmove(q + lt, q + lp, strlen(q + lp) + 1);
            memcpy(q, to, lt);
        }
    }
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---------------------------------------------------------------------------------------------
// utils/basic_anchors.py:11-23
// ---------------------------------------------------------------------------------------------
struct BaseAnchorArgs {
    float ratio[FRCNN_MAX_BASE_ANCHORS];
    float inv_ratio[FRCNN_MAX_BASE_ANCHORS];
    float size[FRCNN_MAX_BASE_ANCHORS];
    int num_ratios, num_sizes;
};

__global__ void base_anchor_kernel(BaseAnchorArgs a, float4* __restrict__ out) {
    int t = threadIdx.x;
    if (t >= a.num_ratios * a.num_sizes) return;
    int i = t / a.num_sizes, j = t % a.num_sizes;
    float h = a.size[j] * __fsqrt_rn(a.ratio[i]);
    float w = a.size[j] * __fsqrt_rn(a.inv_ratio[i]);
    float hw = w / 2.f, hh = h / 2.f;
    out[t] = make_float4(-hw, -hh, hw, hh);
}

// utils/basic_anchors.py:27-57 -- one float4 store per anchor, fully coalesced
__global__ void shifted_anchor_kernel(AnchorGen g, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = load_anchor(g, i);
}

// ---------------------------------------------------------------------------------------------
// utils/loc_bbox_iou.py:29-61   loc [R, 4*groups]; thread per (row, group)
// ---------------------------------------------------------------------------------------------
__global__ void loc2bbox_kernel(const float4* __restrict__ src, const float4* __restrict__ loc,
                                int64_t total, int groups, float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float4 a = __ldg(src + i / groups);
    out[i] = decode_box(a, __ldg(loc + i));
}

__global__ void bbox2loc_kernel(const float4* __restrict__ src, const float4* __restrict__ dst,
                                int64_t rows, float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    out[i] = encode_box(__ldg(src + i), __ldg(dst + i));
}

// ---------------------------------------------------------------------------------------------
// utils/loc_bbox_iou.py:4-27   dense [Na,Nb].  The output is the only large stream (4*Na*Nb bytes against
// 16*(Na+Nb) read), so threads walk it FLAT: a thread owns four consecutive elements of the row-major
// matrix and writes them with one 16-byte store whatever Nb is (the callers' Nb is the GT count: 1-90, so a
// [rows x Nb-tile] mapping would leave most lanes idle and store 4*Nb-byte runs).  The b boxes and their
// areas are staged in shared memory once per CTA (grid-stride loop), the a box is re-read only when the
// flat index crosses a row.
// ---------------------------------------------------------------------------------------------
constexpr int IOU_THREADS = 256;
constexpr int IOU_STAGE = 1024;  // b boxes kept in shared memory (20 KB); larger Nb reads b through L1

template <bool STAGED, bool VEC>
__global__ void __launch_bounds__(IOU_THREADS)
bbox_iou_kernel(const float4* __restrict__ a, const float4* __restrict__ b, int64_t na, int nb, int64_t total,
                float* __restrict__ out) {
    __shared__ float4 sb[STAGED ? IOU_STAGE : 1];
    __shared__ float sarea[STAGED ? IOU_STAGE : 1];
    if (STAGED) {
        for (int j = threadIdx.x; j < nb; j += IOU_THREADS) {
            float4 v = __ldg(b + j);
            sb[j] = v;
            sarea[j] = box_area(v);
        }
        __syncthreads();
    }
    const int64_t quads = (total + 3) >> 2;
    for (int64_t q = (int64_t)blockIdx.x * IOU_THREADS + threadIdx.x; q < quads; q += (int64_t)gridDim.x * IOU_THREADS) {
        const int64_t idx = q << 2;
        int64_t i;
        int j;
        if (total < (1ll << 31)) {  // uniform: 32-bit division whenever it is enough
            uint32_t ii = (uint32_t)idx / (uint32_t)nb;
            i = ii;
            j = (int)((uint32_t)idx - ii * (uint32_t)nb);
        } else {
            i = idx / nb;
            j = (int)(idx - i * nb);
        }
        float4 av = __ldg(a + i);
        float aa = box_area(av);
        float r[4];
        const int nvalid = (int)min((int64_t)4, total - idx);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = 0.f;
            if (k < nvalid) {
                float4 bv;
                float ba;
                if (STAGED) {
                    bv = sb[j];
                    ba = sarea[j];
                } else {
                    bv = __ldg(b + j);
                    ba = box_area(bv);
                }
                r[k] = iou_eps(av, aa, bv, ba);
                if (++j == nb && k < 3) {
                    j = 0;
                    if (++i < na) {
                        av = __ldg(a + i);
                        aa = box_area(av);
                    }
                }
            }
        }
        if (VEC && nvalid == 4) {
            __stcs(reinterpret_cast<float4*>(out + idx), make_float4(r[0], r[1], r[2], r[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nvalid) out[idx + k] = r[k];
        }
    }
}

// Fast path for Nb % 4 == 0 (and a 16-byte aligned output): a thread's four elements never cross a row,
// so there is no per-element range or wrap test and the row index comes from one multiply-high.  With
// Nb % 8 == 0 (QUADS = 2) a thread takes two quads of one row (half a row apart): one index division and one a-box
// load per eight outputs.  b boxes and their areas are staged once per CTA.
template <bool STAGED, int QUADS>
__global__ void __launch_bounds__(IOU_THREADS)
bbox_iou_rows4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, int nb, FastDiv by_items_per_row,
                      int64_t items, float* __restrict__ out) {
    __shared__ float4 sb[STAGED ? IOU_STAGE : 1];
    __shared__ float sba[STAGED ? IOU_STAGE : 1];
    const int ipr = by_items_per_row.d;  // items (QUADS quads each) per row
    const int qpr = ipr * QUADS;
    if (STAGED) {  // box j lives at (j % 4) * qpr + j / 4: lanes reading their k-th box hit consecutive float4s
        for (int j = threadIdx.x; j < nb; j += IOU_THREADS) {
            const float4 v = __ldg(b + j);
            sb[(j & 3) * qpr + (j >> 2)] = v;
            sba[(j & 3) * qpr + (j >> 2)] = box_area(v);
        }
        __syncthreads();
    }
    for (int64_t t = (int64_t)blockIdx.x * IOU_THREADS + threadIdx.x; t < items; t += (int64_t)gridDim.x * IOU_THREADS) {
        int64_t i;
        int ji;
        if (items < (1ll << 31)) {
            const int ii = fast_div((int)t, by_items_per_row);
            i = ii;
            ji = (int)t - ii * ipr;
        } else {
            i = t / ipr;
            ji = (int)(t - i * ipr);
        }
        const float4 av = __ldg(a + i);
        const float aa = box_area(av);
#pragma unroll
        for (int u = 0; u < QUADS; ++u) {
            const int jq = ji + u * ipr;  // lanes stay on consecutive quads: conflict-free LDS, coalesced stores
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float4 bv;
                float ba;
                if (STAGED) {
                    bv = sb[k * qpr + jq];
                    ba = sba[k * qpr + jq];
                } else {
                    bv = __ldg(b + 4 * jq + k);
                    ba = box_area(bv);
                }
                r[k] = iou_eps(av, aa, bv, ba);
            }
            __stcs(reinterpret_cast<float4*>(out) + (i * qpr + jq), make_float4(r[0], r[1], r[2], r[3]));
        }
    }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_abi_version(void) { return FRCNN_ABI_VERSION; }
const char* frcnn_last_error(void) { return g_err; }
uint64_t frcnn_launch_count(void) { return (uint64_t)g_launches.load(std::memory_order_relaxed); }
const char* frcnn_last_roi_kernel(void) { return g_roi_kernel; }

int frcnn_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    FRCNN_CUDA(cudaGetDevice(&dev));
    if (sm) FRCNN_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    if (major) FRCNN_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
    if (minor) FRCNN_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
    return FRCNN_OK;
}

int frcnn_base_anchors(const float* ratios, const float* inv_ratios, int32_t nr, const float* sizes,
                       int32_t ns, float* out, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(ratios && inv_ratios && sizes && out, "frcnn_base_anchors: null pointer");
    FRCNN_CHECK_ARG(nr > 0 && ns > 0 && nr <= FRCNN_MAX_BASE_ANCHORS && ns <= FRCNN_MAX_BASE_ANCHORS &&
                        nr * ns <= FRCNN_MAX_BASE_ANCHORS,
                    "frcnn_base_anchors: need 0 < ratios*scales <= %d", FRCNN_MAX_BASE_ANCHORS);
    BaseAnchorArgs a;
    memset(&a, 0, sizeof(a));
    memcpy(a.ratio, ratios, nr * sizeof(float));
    memcpy(a.inv_ratio, inv_ratios, nr * sizeof(float));
    memcpy(a.size, sizes, ns * sizeof(float));
    a.num_ratios = nr;
    a.num_sizes = ns;
    base_anchor_kernel<<<1, FRCNN_MAX_BASE_ANCHORS, 0, (cudaStream_t)stream>>>(a, (float4*)out);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_shifted_anchors(const float* base, int32_t A, int32_t stride, int32_t H, int32_t W,
                          float* out, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(A > 0 && H >= 0 && W >= 0, "frcnn_shifted_anchors: bad shape");
    int64_t n64 = (int64_t)A * H * W;
    FRCNN_CHECK_ARG(n64 < (1ll << 31), "frcnn_shifted_anchors: too many anchors");
    int n = (int)n64;
    if (n == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(base && out, "frcnn_shifted_anchors: null pointer");
    AnchorGen g = make_anchor_gen(nullptr, base, A, stride, H, W);
    shifted_anchor_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(g, n, (float4*)out);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_loc2bbox(const float* src, const float* loc, int64_t rows, int32_t groups, float* out,
                   frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(rows >= 0 && groups > 0, "frcnn_loc2bbox: bad shape");
    if (rows == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(src && loc && out, "frcnn_loc2bbox: null pointer");
    int64_t total = rows * groups;
    loc2bbox_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)src, (const float4*)loc, total, groups, (float4*)out);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_bbox2loc(const float* src, const float* dst, int64_t rows, float* out, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(rows >= 0, "frcnn_bbox2loc: bad shape");
    if (rows == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(src && dst && out, "frcnn_bbox2loc: null pointer");
    bbox2loc_kernel<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)src, (const float4*)dst, rows, (float4*)out);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_bbox_iou(const float* a, const float* b, int64_t na, int64_t nb, float* out,
                   frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(na >= 0 && nb >= 0, "frcnn_bbox_iou: bad shape");
    if (na == 0 || nb == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(a && b && out, "frcnn_bbox_iou: null pointer");
    FRCNN_CHECK_ARG(nb < (1ll << 31) && na <= (1ll << 40) / nb, "frcnn_bbox_iou: matrix too large");
    const int64_t total = na * nb;
    const int64_t quads = (total + 3) / 4;
    const int grid = (int)std::min<int64_t>((quads + IOU_THREADS - 1) / IOU_THREADS, (int64_t)sm_count() * 16);
    const bool vec = ((uintptr_t)out & 15) == 0;
    const bool staged = nb <= IOU_STAGE;
    auto* a4 = (const float4*)a;
    auto* b4 = (const float4*)b;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec && nb % 4 == 0) {
        const int quads_per_item = nb % 8 == 0 ? 2 : 1;
        const int64_t items = quads / quads_per_item;
        const FastDiv fd = make_fastdiv((int)(nb / (4 * quads_per_item)));
        const int g2 = (int)std::min<int64_t>((items + IOU_THREADS - 1) / IOU_THREADS, (int64_t)sm_count() * 16);
        if (quads_per_item == 2) {
            if (staged) bbox_iou_rows4_kernel<true, 2><<<g2, IOU_THREADS, 0, st>>>(a4, b4, (int)nb, fd, items, out);
            else bbox_iou_rows4_kernel<false, 2><<<g2, IOU_THREADS, 0, st>>>(a4, b4, (int)nb, fd, items, out);
        } else {
            if (staged) bbox_iou_rows4_kernel<true, 1><<<g2, IOU_THREADS, 0, st>>>(a4, b4, (int)nb, fd, items, out);
            else bbox_iou_rows4_kernel<false, 1><<<g2, IOU_THREADS, 0, st>>>(a4, b4, (int)nb, fd, items, out);
        }
        FRCNN_LAUNCH_CHECK();
        return FRCNN_OK;
    }
    if (staged && vec) bbox_iou_kernel<true, true><<<grid, IOU_THREADS, 0, st>>>(a4, b4, na, (int)nb, total, out);
    else if (staged) bbox_iou_kernel<true, false><<<grid, IOU_THREADS, 0, st>>>(a4, b4, na, (int)nb, total, out);
    else if (vec) bbox_iou_kernel<false, true><<<grid, IOU_THREADS, 0, st>>>(a4, b4, na, (int)nb, total, out);
    else bbox_iou_kernel<false, false><<<grid, IOU_THREADS, 0, st>>>(a4, b4, na, (int)nb, total, out);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

}  // extern "C"
