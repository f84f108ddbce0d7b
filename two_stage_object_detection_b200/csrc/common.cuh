// common.cuh -- shared helpers for the sm_100a kernels behind include/frcnn_b200.h.
// Compiled with -fmad=false: every fp32 multiply/add is a separate IEEE rounding, as the
// reference's ATen/torchvision CPU kernels do them.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/frcnn_b200.h"

namespace frcnn {

void set_error(const char* fmt, ...);

#define FRCNN_CHECK_ARG(cond, ...)                    \
    do {                                              \
        if (!(cond)) {                                \
            ::frcnn::set_error(__VA_ARGS__);          \
            return FRCNN_ERR_INVALID_ARG;             \
        }                                             \
    } while (0)

#define FRCNN_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ::frcnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                               __FILE__, __LINE__);                                           \
            return FRCNN_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

// every kernel launch of the library goes through this: the counter is what bench.py reports as
// `gpu_launches` (frcnn_launch_count), counted, not computed
void count_launch();
#define FRCNN_LAUNCH_CHECK()            \
    do {                                \
        ::frcnn::count_launch();        \
        FRCNN_CUDA(cudaGetLastError()); \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (device, kernel) -- and again only when a
// launch needs more than any earlier one did -- instead of a driver call in front of every launch
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);
#define FRCNN_SMEM(kernel, bytes) FRCNN_CUDA(::frcnn::ensure_dynamic_smem((const void*)(kernel), (size_t)(bytes)))

// name of the gather kernel variant the last RoI forward call picked (frcnn_last_roi_kernel)
void note_roi_kernel(const char* fmt, ...);

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int sm_count();

// Bump allocator over the caller-provided workspace.
struct Workspace {
    char* base;
    size_t cap;
    size_t off;
    Workspace(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap && ((uintptr_t)base % 256) == 0; }
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
struct BaseAnchors {
    float v[FRCNN_MAX_BASE_ANCHORS][4];
};

// Division by a launch-invariant divisor: quotient = umulhi(n, mul) >> shr, exact for 0 <= n < 2^31
// (the runtime `/` and `%` of the anchor index cost ~20 instructions each; three of them per anchor made
// the streaming kernels issue-bound).
struct FastDiv {
    uint32_t mul, shr;
    int d;
};
static inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = d;
    f.mul = 0;
    f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1ll << lg) < d) ++lg;
        const int p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1) / (uint64_t)d);
        f.shr = (uint32_t)(p - 32);
    }
    return f;
}
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
    return f.d > 1 ? (int)(__umulhi((uint32_t)n, f.mul) >> f.shr) : n;
}

struct AnchorGen {
    const float4* anchors;  // explicit [N,4] or nullptr
    const float4* base;     // [A,4]
    int num_base, stride, height, width;
    FastDiv by_base, by_width;
};
static inline AnchorGen make_anchor_gen(const float* anchors, const float* base, int num_base, int stride,
                                        int height, int width) {
    AnchorGen g;
    g.anchors = (const float4*)anchors;
    g.base = (const float4*)base;
    g.num_base = num_base;
    g.stride = stride;
    g.height = height;
    g.width = width;
    g.by_base = make_fastdiv(num_base > 0 ? num_base : 1);
    g.by_width = make_fastdiv(width > 0 ? width : 1);
    return g;
}

__device__ __forceinline__ float4 load_anchor(const AnchorGen& g, int i) {
    if (g.anchors) return __ldg(g.anchors + i);
    const int k = fast_div(i, g.by_base);
    const int a = i - k * g.num_base;
    const int y = fast_div(k, g.by_width);
    const int x = k - y * g.width;
    float4 b = __ldg(g.base + a);
    float sx = (float)(x * g.stride), sy = (float)(y * g.stride);
    return make_float4(b.x + sx, b.y + sy, b.z + sx, b.w + sy);
}

// Order-preserving map fp32 -> uint32 (bigger score = bigger key), matching torch.sort's view of
// floats: NaN is the largest, -0 == +0.  Result is never 0 (0 is reserved for "filtered out").
__device__ __forceinline__ uint32_t score_key(float s) {
    if (s != s) return 0xFFFFFFFFu;
    if (s == 0.f) s = 0.f;
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// torch.max / argmax semantics: NaN beats everything, first index wins ties.
__device__ __forceinline__ bool beats(float v, float best) {
    return (v > best) || (v != v && best == best);
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// utils/loc_bbox_iou.py:18-26 -- inter / (((area_a + area_b) - inter) + 1e-8f)
// torch.maximum / torch.minimum propagate NaN (fmaxf/fminf do not): one FMNMX.NAN each
__device__ __forceinline__ float tmax(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float tmin(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// IEEE fp32 division for quotients that are mostly 0 / positive (boxes that do not overlap).  The inlined
// division takes its out-of-line slow path whenever FCHK dislikes an operand, and a zero numerator is one
// of those (measured: ~90 instead of ~35 instructions per IoU).  0 / (positive) is exactly the zero it
// started from, sign included, so those lanes divide 1 by 1 on the fast path and keep their numerator.
__device__ __forceinline__ float div_mostly_zero(float num, float den) {
    const bool zero = (num == 0.f) && (den > 0.f);
    float n = zero ? 1.f : num, d = zero ? 1.f : den;
    asm("" : "+f"(n), "+f"(d));  // opaque: otherwise the selects are folded back into num / den
    const float q = n / d;
    return zero ? num : q;
}

__device__ __forceinline__ float iou_eps(const float4& a, float area_a, const float4& b, float area_b) {
    float tlx = tmax(a.x, b.x), tly = tmax(a.y, b.y);
    float brx = tmin(a.z, b.z), bry = tmin(a.w, b.w);
    float w = tmax(brx - tlx, 0.f), h = tmax(bry - tly, 0.f);  // clamp_(min=0); NaN stays NaN
    float inter = w * h;
    float uni = area_a + area_b;
    uni = uni - inter;
    uni = uni + 1e-8f;
    return div_mostly_zero(inter, uni);
}

// inverse of score_key for keys it produced (0xFFFFFFFF -> a NaN)
__device__ __forceinline__ float key_score(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__device__ __forceinline__ float box_area(const float4& b) { return (b.z - b.x) * (b.w - b.y); }

// utils/loc_bbox_iou.py:63-89
__device__ __forceinline__ float4 encode_box(const float4& s, const float4& d) {
    float w = s.z - s.x, h = s.w - s.y;
    float cx = s.x + 0.5f * w, cy = s.y + 0.5f * h;
    float bw = d.z - d.x, bh = d.w - d.y;
    float bcx = d.x + 0.5f * bw, bcy = d.y + 0.5f * bh;
    const float eps = 1.1920928955078125e-07f;
    // torch.maximum propagates NaN
    w = (w != w) ? w : (w > eps ? w : eps);
    h = (h != h) ? h : (h > eps ? h : eps);
    float4 o;
    o.x = (bcx - cx) / w;
    o.y = (bcy - cy) / h;
    o.z = logf(bw / w);
    o.w = logf(bh / h);
    return o;
}

// utils/loc_bbox_iou.py:36-59
__device__ __forceinline__ float4 decode_box(const float4& a, const float4& l) {
    float w = a.z - a.x, h = a.w - a.y;
    float cx = a.x + 0.5f * w, cy = a.y + 0.5f * h;
    float ncx = l.x * w + cx, ncy = l.y * h + cy;
    float nw = expf(l.z) * w, nh = expf(l.w) * h;
    float hw = 0.5f * nw, hh = 0.5f * nh;
    return make_float4(ncx - hw, ncy - hh, ncx + hw, ncy + hh);
}

// torchvision nms_kernel: inter / (area_i + area_j - inter) > thr  (no eps; NaN never suppresses).
// `thr` is the largest float <= the double threshold, so (float)ovr > thr <=> (double)ovr > thr_d.
__device__ __forceinline__ bool nms_suppresses(const float4& r, float ra, const float4& c, float ca, float thr) {
    float xx1 = fmaxf(r.x, c.x), yy1 = fmaxf(r.y, c.y);
    float xx2 = fminf(r.z, c.z), yy2 = fminf(r.w, c.w);
    float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
    float inter = w * h;
    float uni = ra + ca;
    uni = uni - inter;
    return div_mostly_zero(inter, uni) > thr;
}


static inline float float_threshold(double thr) {
    // largest float T with T <= thr, so that for every float x: x > T  <=>  (double)x > thr
    float t = (float)thr;
    if ((double)t > thr) t = nextafterf(t, -INFINITY);
    return t;
}

}  // namespace frcnn
