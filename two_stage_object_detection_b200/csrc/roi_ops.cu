// roi_ops.cu -- RoI feature gathering for the classifier head (nets/classify.py:29-43):
// torchvision RoIPool / roi_align forward (+ backward) as sm_100a kernels.
//
// Data flow of the forward kernels: RoIs are bucketed per image on the device; one CTA owns
// (image, channel slab, RoI group), stages the slab's full H x W feature planes into shared memory
// with TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP), then every thread produces output
// elements from shared memory and stores them coalesced along the contiguous [C,P,P] axis of the
// RoI.  Features are read from L2/HBM once per (slab, group); the dominant HBM traffic is the
// K*C*P*P*4 B output write.
#include <float.h>

#include <algorithm>


#include "common.cuh"

namespace frcnn {

// ---------------------------------------------------------------------------------------------
// nets/classify.py:33-38
// ---------------------------------------------------------------------------------------------
__global__ void roi_head_coords_kernel(const float4* __restrict__ rois, const int* __restrict__ idx, int total,
                                       int per_image, float d0, float d1, float fh, float fw,
                                       float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float4 r = __ldg(rois + i);
    float* o = out + (size_t)i * 5;
    o[0] = (float)__ldg(idx + i / per_image);
    o[1] = (r.x / d1) * fw;
    o[2] = (r.y / d0) * fh;
    o[3] = (r.z / d1) * fw;
    o[4] = (r.w / d0) * fh;
}

// ---------------------------------------------------------------------------------------------
// bucket RoIs by image: perm[offs[b] .. offs[b+1]) = RoI ids of image b
// ---------------------------------------------------------------------------------------------
constexpr int BUCKET_THREADS = 1024;

// RoIs whose batch index is outside [0,B) belong to no image: they are listed after the valid ones,
// perm[offs[B] .. offs[B] + offs[B+1]), and roi_fill_dropped_kernel gives their output rows the value the
// direct kernels give them (0, argmax -1) instead of leaving them unwritten.
__global__ void __launch_bounds__(BUCKET_THREADS)
roi_bucket_kernel(const float* __restrict__ rois5, int K, int B, int* __restrict__ perm, int* __restrict__ offs) {
    extern __shared__ int sb[];  // cnt[B], cur[B]
    int* cnt = sb;
    int* cur = sb + B;
    __shared__ int chunk_tot[BUCKET_THREADS];
    __shared__ int s_total, s_bad;
    const int tid = threadIdx.x;
    for (int i = tid; i < B; i += BUCKET_THREADS) cnt[i] = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int k = tid; k < K; k += BUCKET_THREADS) {
        int b = (int)__ldg(rois5 + (size_t)k * 5);
        if (b >= 0 && b < B) atomicAdd(&cnt[b], 1);
    }
    __syncthreads();
    // exclusive scan: each thread owns a contiguous chunk of images
    int per = (B + BUCKET_THREADS - 1) / BUCKET_THREADS;
    int lo = tid * per, hi = min(lo + per, B);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    chunk_tot[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < BUCKET_THREADS; ++i) {
            int t = chunk_tot[i];
            chunk_tot[i] = run;
            run += t;
        }
        offs[B] = run;
        s_total = run;
    }
    __syncthreads();
    int run = chunk_tot[tid];
    for (int i = lo; i < hi; ++i) {
        cur[i] = run;
        offs[i] = run;
        run += cnt[i];
    }
    __syncthreads();
    for (int k = tid; k < K; k += BUCKET_THREADS) {
        int b = (int)__ldg(rois5 + (size_t)k * 5);
        if (b >= 0 && b < B) perm[atomicAdd(&cur[b], 1)] = k;
        else perm[s_total + atomicAdd(&s_bad, 1)] = k;
    }
    __syncthreads();
    if (tid == 0) offs[B + 1] = s_bad;
}

__global__ void __launch_bounds__(256)
roi_fill_dropped_kernel(const int* __restrict__ perm, const int* __restrict__ offs, int B, size_t per_roi,
                        float* __restrict__ out, int* __restrict__ argmax) {
    const int base = offs[B], n_bad = offs[B + 1];
    for (int i = blockIdx.x; i < n_bad; i += gridDim.x) {
        const size_t o = (size_t)perm[base + i] * per_roi;
        for (size_t j = threadIdx.x; j < per_roi; j += 256) {
            out[o + j] = 0.f;
            if (argmax) argmax[o + j] = -1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA bulk-copy helpers (1-D cp.async.bulk global -> shared, completion on an mbarrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Stage `count` floats (contiguous in global memory) into shared memory.  TMA path when the source
// is 16-byte aligned and the size a multiple of 16 bytes; plain coalesced loads otherwise.
__device__ __forceinline__ void stage_slab(float* sdst, const float* gsrc, int count, uint64_t* bar) {
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(gsrc) & 15) == 0) && ((count & 3) == 0);
    if (tma_ok) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t bytes = (uint32_t)count * 4u;
            mbar_expect_tx(bar, bytes);
            const uint32_t chunk = 32768u;
            for (uint32_t off = 0; off < bytes; off += chunk) {
                uint32_t nb = min(chunk, bytes - off);
                bulk_g2s(reinterpret_cast<char*>(sdst) + off, reinterpret_cast<const char*>(gsrc) + off, nb, bar);
            }
        }
        mbar_wait(bar, 0);
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) sdst[i] = __ldg(gsrc + i);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// RoIPool forward
// ---------------------------------------------------------------------------------------------
constexpr int ROI_THREADS = 512;
constexpr int ROI_NB = 8;        // RoIs per inner batch (one table build + one sync per batch)
constexpr int ROI_MAX_P = 32;    // largest pooled size served by the staged kernels

struct RoiArgs {
    const float* feat;
    const float* rois5;
    const int* perm;
    const int* offs;
    int B, C, H, W, K, PH, PW;
    float scale;
    int CS;          // channels per slab
    int groups;      // RoI groups per image
    float* out;
    int* argmax;
    int sampling_ratio, aligned;
    int per_image;   // > 0: RoIs are grouped, rows [b*per_image, (b+1)*per_image) belong to image b
    int pitch;       // table kernels: row pitch of the shared-memory tables (>= W)
    const int2* ent; // inference table kernels: per-RoI bin geometry [K][2P] from roi_pool_entries_kernel
    const int* perm2; // streaming RoIAlign: RoI rows of every image ordered by row-program length
};

__device__ __forceinline__ int round_half_away(float v) { return (int)roundf(v); }

// RoI rows of image b, and the RoI id stored at row r (identity when the caller grouped the RoIs)
__device__ __forceinline__ void roi_range(const RoiArgs& a, int b, int& r_begin, int& r_end) {
    if (a.per_image > 0) {
        r_begin = b * a.per_image;
        r_end = r_begin + a.per_image;
    } else {
        r_begin = a.offs[b];
        r_end = a.offs[b + 1];
    }
}
__device__ __forceinline__ int roi_at(const RoiArgs& a, int r) { return a.per_image > 0 ? r : a.perm[r]; }

template <int P, bool WITH_ARGMAX>
__global__ void __launch_bounds__(ROI_THREADS, 2) roi_pool_staged_kernel(RoiArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_roi[ROI_NB];
    __shared__ int th[ROI_NB][ROI_MAX_P];  // hstart | hend << 16
    __shared__ int tw[ROI_NB][ROI_MAX_P];  // wstart | wend << 16
    float* sfeat = reinterpret_cast<float*>(smem_raw);
    const int PH = P > 0 ? P : a.PH, PW = P > 0 ? P : a.PW;
    const int PP = PH * PW;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.CS;
    const int cs = min(a.CS, a.C - c0);
    const int HW = a.H * a.W;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int first = r_begin + blockIdx.x;
    if (first >= r_end) return;
    stage_slab(sfeat, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);

    const int per_roi = cs * PP;
    for (int rb = first; rb < r_end; rb += a.groups * ROI_NB) {
        // RoIs of this batch: rb, rb+groups, ...  (strided so groups stay balanced)
        int nb = 0;
        for (int j = 0; j < ROI_NB; ++j)
            if (rb + j * a.groups < r_end) nb = j + 1;
        __syncthreads();  // previous batch done with the tables
        if (threadIdx.x < nb * (PH + PW)) {
            int j = threadIdx.x / (PH + PW), e = threadIdx.x % (PH + PW);
            int k = roi_at(a, rb + j * a.groups);
            const float* r = a.rois5 + (size_t)k * 5;
            if (e == 0) s_roi[j] = k;
            if (e < PH) {
                int sh = round_half_away(__ldg(r + 2) * a.scale), eh = round_half_away(__ldg(r + 4) * a.scale);
                int rh = max(eh - sh + 1, 1);
                float bin = (float)rh / (float)PH;
                int hs = (int)floorf((float)e * bin), he = (int)ceilf((float)(e + 1) * bin);
                hs = min(max(hs + sh, 0), a.H);
                he = min(max(he + sh, 0), a.H);
                th[j][e] = hs | (he << 16);
            } else {
                int p = e - PH;
                int sw = round_half_away(__ldg(r + 1) * a.scale), ew = round_half_away(__ldg(r + 3) * a.scale);
                int rw = max(ew - sw + 1, 1);
                float bin = (float)rw / (float)PW;
                int ws = (int)floorf((float)p * bin), we = (int)ceilf((float)(p + 1) * bin);
                ws = min(max(ws + sw, 0), a.W);
                we = min(max(we + sw, 0), a.W);
                tw[j][p] = ws | (we << 16);
            }
        }
        __syncthreads();
        const int total = nb * per_roi;
        for (int it = threadIdx.x; it < total; it += ROI_THREADS) {
            int j = it / per_roi, rem = it - j * per_roi;
            int c = rem / PP, bin = rem - c * PP;
            int ph = bin / PW, pw = bin - ph * PW;
            int hh = th[j][ph], ww = tw[j][pw];
            int hs = hh & 0xFFFF, he = hh >> 16, ws = ww & 0xFFFF, we = ww >> 16;
            const float* plane = sfeat + c * HW;
            bool empty = (he <= hs) || (we <= ws);
            float best = empty ? 0.f : -FLT_MAX;
            int besti = -1;
            for (int h = hs; h < he; ++h) {
                const float* row = plane + h * a.W;
                for (int w = ws; w < we; ++w) {
                    float v = row[w];
                    if (WITH_ARGMAX) {
                        if (v > best) {
                            best = v;
                            besti = h * a.W + w;
                        }
                    } else {
                        best = v > best ? v : best;
                    }
                }
            }
            size_t o = ((size_t)s_roi[j] * a.C + c0) * PP + rem;
            a.out[o] = best;
            if (WITH_ARGMAX) a.argmax[o] = besti;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RoIPool forward, table kernel: roi_pool_tab_kernel<P, threads, CS, MINB, ARGMAX, LV, BPT>.
//
// LV = 2 (inference).  A bin's max over [hs,he) x [ws,we) is the max of four lookups into 2-D "sparse max
// tables":
//   T[a][b][y][x] = max of the a x b window anchored at (y,x),  a,b in {1,2}
// because any window up to 4 x 4 is covered by the (at most) four a x b windows placed in its corners
// (max is idempotent, overlaps do not matter).  The CTA builds the four tables once for its CS-channel
// slab -- channel-interleaved, one float4 (CS = 4) / float2 / float per pixel, so one LDS.128 serves four
// channels -- and then a bin costs 1-4 LDS.128 + FMNMX instead of a data-dependent double loop.  On the 7x7
// grid bins 5..8 long take four 2-windows per axis (tab_mid_bin); longer ones a scan of T[1][1]
// (tab_big_bin).  Values are first clamped with fmaxf(v, -FLT_MAX), which reproduces the reference's
// `v > best` scan exactly (NaN / -inf never win); empty bins give 0.
//
// LV = 1 (training, ARGMAX).  Only the interleaved pixels are kept; one scan per bin tracks the maximum and
// the first position attaining it (the reference's order).  With the 128 sampled RoIs per image of a
// training step the table build would cost more than it saves.
//
// smem: LV*LV tables x HWp elements (38x38, float4, LV = 2: 92 KB, two CTAs per SM).  The raw NCHW planes
// are TMA-staged into the region that later holds the last table built (LV = 1: a region of their own).
// The row pitch is padded by one pixel when a row's byte length is a multiple of 64 (vertically adjacent
// bins would share banks).
//
// Thread mapping: 392 threads = 2 * 14*14 = 8 * 7*7 (784 when only one CTA fits per SM).  A thread owns
// ONE output bin (ph,pw) -- or, BPT = 2 on the 14x14 grid, two horizontally adjacent ones -- for good and
// walks the CTA's RoIs; consecutive threads are consecutive bins of the same RoI, so every warp-level
// store is a contiguous 128-byte (BPT = 2: 256-byte, 8 bytes per lane) run of the [K,C,P,P] output
// (measured: the store pattern, not the lookups, bounded the first versions).  Per-RoI bin geometry is
// precomputed for a batch of RoIs into double-buffered shared-memory tables (two entries per thread, one
// barrier per batch), and the next batch's RoI boxes are prefetched.
// ---------------------------------------------------------------------------------------------
constexpr int TAB_NE_BIT = 0x80000000;   // entry .y bit 31: bin row / column is non-empty
constexpr int TAB_BIG_BIT = 0x40000000;  // entry .y bit 30: longer than the tables cover -> loop path
constexpr int TAB_MID_BIT = 0x20000000;  // entry .y bit 29 (tables 1,2 only): 5..8 long -> four 2-windows
constexpr int TAB_OFF_MASK = 0x1FFFFFFF;
constexpr int TAB_LVL_BIT = 0x10000000;  // entry .y bit 28 (DIAG tables only): this axis uses the 2-long window
constexpr int TAB_DIAG_MASK = 0x0FFFFFFF;

// CS channel-interleaved floats per pixel: float4 (LDS.128) for CS = 4, float2 (LDS.64) for CS = 2.
// Native vector types on purpose: a struct-of-array wrapper made ptxas split the predicated loads.
template <int CS>
struct VecT;
template <>
struct VecT<4> {
    typedef float4 type;
};
template <>
struct VecT<2> {
    typedef float2 type;
};
template <>
struct VecT<1> {
    typedef float type;
};
__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ void vsplat(float& v, float x) { v = x; }
__device__ __forceinline__ void vgather(float& v, const float* raw, int, int p, int) { v = fmaxf(raw[p], -FLT_MAX); }
__device__ __forceinline__ float4 vmax(const float4& a, const float4& b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float2 vmax(const float2& a, const float2& b) {
    return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y));
}
__device__ __forceinline__ void vsplat(float4& v, float x) { v = make_float4(x, x, x, x); }
__device__ __forceinline__ void vsplat(float2& v, float x) { v = make_float2(x, x); }
// planes -> interleaved pixel, clamped with fmaxf(., -FLT_MAX); channels >= cs read as -FLT_MAX
__device__ __forceinline__ void vgather(float4& v, const float* raw, int HW, int p, int cs) {
    v.x = fmaxf(raw[p], -FLT_MAX);
    v.y = cs > 1 ? fmaxf(raw[HW + p], -FLT_MAX) : -FLT_MAX;
    v.z = cs > 2 ? fmaxf(raw[2 * HW + p], -FLT_MAX) : -FLT_MAX;
    v.w = cs > 3 ? fmaxf(raw[3 * HW + p], -FLT_MAX) : -FLT_MAX;
}
__device__ __forceinline__ void vgather(float2& v, const float* raw, int HW, int p, int cs) {
    v.x = fmaxf(raw[p], -FLT_MAX);
    v.y = cs > 1 ? fmaxf(raw[HW + p], -FLT_MAX) : -FLT_MAX;
}
// masked store of the channels of one bin (stride = P*P floats between channels)
template <bool FULL>
__device__ __forceinline__ void vstore(float* o, int stride, const float4& v, unsigned m, int cs) {
    o[0] = __uint_as_float(__float_as_uint(v.x) & m);
    if (FULL || cs > 1) o[stride] = __uint_as_float(__float_as_uint(v.y) & m);
    if (FULL || cs > 2) o[2 * stride] = __uint_as_float(__float_as_uint(v.z) & m);
    if (FULL || cs > 3) o[3 * stride] = __uint_as_float(__float_as_uint(v.w) & m);
}
template <bool FULL>
__device__ __forceinline__ void vstore(float* o, int stride, const float2& v, unsigned m, int cs) {
    o[0] = __uint_as_float(__float_as_uint(v.x) & m);
    if (FULL || cs > 1) o[stride] = __uint_as_float(__float_as_uint(v.y) & m);
}

template <bool FULL>
__device__ __forceinline__ void vstore(float* o, int, float v, unsigned m, int) {
    o[0] = __uint_as_float(__float_as_uint(v) & m);
}

struct RoiBox {
    int k;
    float x1, y1, x2, y2;
};

__device__ __forceinline__ RoiBox load_roi(const RoiArgs& a, int r, int r_end) {
    RoiBox q;
    q.k = -1;
    q.x1 = q.y1 = q.x2 = q.y2 = 0.f;
    if (r < r_end) {
        q.k = roi_at(a, r);
        const float* rp = a.rois5 + (size_t)q.k * 5;
        q.x1 = __ldg(rp + 1);
        q.y1 = __ldg(rp + 2);
        q.x2 = __ldg(rp + 3);
        q.y2 = __ldg(rp + 4);
    }
    return q;
}

// What one geometry thread needs of a RoI: the two coordinates of ITS axis (rows: y1,y2; columns: x1,x2) and
// the RoI id -- 3 registers per prefetched RoI instead of the 5 of a whole box.
struct RoiAxis {
    int k;
    float c1, c2;
};
__device__ __forceinline__ RoiAxis load_roi_axis(const RoiArgs& a, int r, int r_end, bool rows) {
    RoiAxis q;
    q.k = -1;
    q.c1 = q.c2 = 0.f;
    if (r < r_end) {
        q.k = roi_at(a, r);
        const float* rp = a.rois5 + (size_t)q.k * 5 + (rows ? 2 : 1);
        q.c1 = __ldg(rp);
        q.c2 = __ldg(rp + 2);
    }
    return q;
}

// Inference table kernels: the geometry entry `ti` of RoI row r was computed once for all channel slabs by
// roi_pool_entries_kernel; it travels in the RoiAxis registers (bit patterns).  Only the thread that writes
// the RoI's output offset (ti == 0) needs the RoI id.
__device__ __forceinline__ RoiAxis load_roi_entry(const RoiArgs& a, int r, int r_end, int ti, int per_roi) {
    RoiAxis q;
    q.k = -1;
    q.c1 = q.c2 = 0.f;
    if (r < r_end) {
        q.k = ti == 0 ? roi_at(a, r) : 0;
        const int2 e = __ldg(a.ent + (size_t)r * per_roi + ti);
        q.c1 = __int_as_float(e.x);
        q.c2 = __int_as_float(e.y);
    }
    return q;
}

// One axis of the bin grid: [lo,hi) of bin `i`, as (byte offset of first corner, byte offset of second
// corner | flags); `unit` = table elements per step along this axis (row pitch for rows, 1 for columns),
// `tsel` = table stride per window level along this axis, `esz` = bytes per table element.
template <int LV, bool MID = false, bool DIAG = false>
__device__ __forceinline__ int2 tab_entry(int i, int P, float c1, float c2, float scale, int limit, int unit,
                                          int tsel, int esz, int* raw) {
    const int s = round_half_away(c1 * scale), e = round_half_away(c2 * scale);
    const float bin = (float)max(e - s + 1, 1) / (float)P;
    const int lo = min(max((int)floorf((float)i * bin) + s, 0), limit);
    const int hi = min(max((int)ceilf((float)(i + 1) * bin) + s, 0), limit);
    const int len = hi - lo;
    *raw = lo | (hi << 16);
    const bool empty = len <= 0;
    // window level: 2 when the run is at least 2 long; two such windows placed at both ends cover up to 4
    const int lvl = (LV >= 2 && len >= 2) ? 1 : 0;
    const int a = 1 << lvl;
    const int lo_ = empty ? 0 : lo, hi_ = empty ? a : hi;
    int2 r;
    // DIAG: the level is a flag, not a table offset (the table is chosen per bin from both axes' levels)
    const int tbase = DIAG ? 0 : lvl * tsel;
    r.x = (tbase + lo_ * unit) * esz;
    r.y = ((tbase + (hi_ - a) * unit) * esz) | (empty ? 0 : TAB_NE_BIT) |
          (len > (LV == 1 ? 1 : (MID ? 8 : 4)) ? TAB_BIG_BIT : 0) |
          ((MID && len > 4 && len <= 8) ? TAB_MID_BIT : 0) | ((DIAG && lvl) ? TAB_LVL_BIT : 0);
    return r;
}

// A run of 5..8 pixels is covered by the 2-long windows at lo, lo+2, hi-4, hi-2 (the first and the last are
// the regular two lookups).  `h0,h1` / `w0,w1` are the regular corner offsets, `hm` / `wm` say which axis is
// 5..8 long, `rstep` / `cstep` = byte distance of two pixels along the axis.  Twelve extra lookups at most.
template <typename V>
__device__ __noinline__ V tab_mid_bin(V v, const unsigned char* smem, int h0, int h1, bool hm, int w0, int w1,
                                         bool wm, int rstep, int cstep) {
    // only the positions an axis really adds are read: four lookups when one axis is long (the usual case),
    // twelve when both are
    const int h2 = h0 + rstep, h3 = h1 - rstep, w2 = w0 + cstep, w3 = w1 - cstep;
    auto at = [&](int o) { return *reinterpret_cast<const V*>(smem + o); };
    if (hm) v = vmax(vmax(v, vmax(at(h2 + w0), at(h2 + w1))), vmax(at(h3 + w0), at(h3 + w1)));
    if (wm) v = vmax(vmax(v, vmax(at(h0 + w2), at(h0 + w3))), vmax(at(h1 + w2), at(h1 + w3)));
    if (hm && wm) v = vmax(vmax(v, vmax(at(h2 + w2), at(h2 + w3))), vmax(at(h3 + w2), at(h3 + w3)));
    return v;
}

// Loop path for one bin longer than 4 in some direction.  Deliberately tiny and out of line: the call
// sits inside the unrolled fast loop, and a larger callee (e.g. one tiling the bin with 2 x 2 table
// windows) raises the register pressure at every call site enough to cost the fast path 6 % (measured).
template <typename V>
__device__ __noinline__ V tab_big_bin(const V* tab, int hr, int wr, int pitch) {
    V v;
    vsplat(v, -FLT_MAX);
    for (int y = hr & 0xFFFF; y < (hr >> 16); ++y)
        for (int x = wr & 0xFFFF; x < (wr >> 16); ++x) v = vmax(v, tab[y * pitch + x]);
    return v;
}

// value + argmax of one bin (training): the reference keeps the first element, in row-major order, that
// is `> best` starting from -FLT_MAX; pixels are pre-clamped with fmaxf(., -FLT_MAX), so NaN / -inf (and a
// genuine -FLT_MAX) are never selected and leave the index at -1.
__device__ __forceinline__ void scan_first_max(float4& v, int4& idx, const float4& t, int p) {
    if (t.x > v.x) v.x = t.x, idx.x = p;
    if (t.y > v.y) v.y = t.y, idx.y = p;
    if (t.z > v.z) v.z = t.z, idx.z = p;
    if (t.w > v.w) v.w = t.w, idx.w = p;
}
__device__ __forceinline__ void scan_first_max(float2& v, int2& idx, const float2& t, int p) {
    if (t.x > v.x) v.x = t.x, idx.x = p;
    if (t.y > v.y) v.y = t.y, idx.y = p;
}
__device__ __forceinline__ void scan_first_max(float& v, int& idx, float t, int p) {
    if (t > v) v = t, idx = p;
}
// The same update under a per-lane predicate (a pixel column the bin may not have): the predicate joins the compare
// (FSETP.GT.AND), so a column costs the same three instructions per channel whether it exists or not and the row
// body stays straight-line.
__device__ __forceinline__ void scan_first_max_if(float4& v, int4& idx, const float4& t, int p, bool on) {
    const bool cx = on & (t.x > v.x), cy = on & (t.y > v.y), cz = on & (t.z > v.z), cw = on & (t.w > v.w);
    v.x = cx ? t.x : v.x, idx.x = cx ? p : idx.x;
    v.y = cy ? t.y : v.y, idx.y = cy ? p : idx.y;
    v.z = cz ? t.z : v.z, idx.z = cz ? p : idx.z;
    v.w = cw ? t.w : v.w, idx.w = cw ? p : idx.w;
}
__device__ __forceinline__ void scan_first_max_if(float2& v, int2& idx, const float2& t, int p, bool on) {
    const bool cx = on & (t.x > v.x), cy = on & (t.y > v.y);
    v.x = cx ? t.x : v.x, idx.x = cx ? p : idx.x;
    v.y = cy ? t.y : v.y, idx.y = cy ? p : idx.y;
}
__device__ __forceinline__ void scan_first_max_if(float& v, int& idx, float t, int p, bool on) {
    const bool c = on & (t > v);
    v = c ? t : v, idx = c ? p : idx;
}
// Rows of a training bin, NW pixel columns per row as straight-line predicated code (row-major order, strict `>`:
// the reference's first maximum).  `rp` = the window's first pixel in the interleaved table, `p` = its flat index
// y * W + x in the plane; a row step is WP table elements / W plane elements.  Columns past the window are loaded
// (they lie inside the table or the staging region behind it) and ignored.  LONG: windows wider than NW finish
// each row in a loop.
template <int NW, bool LONG, typename V, typename I>
__device__ __forceinline__ void scan_rows(V& v, I& idx, const V* rp, int p, int hh, int ww, int WP, int W) {
    const bool w1 = ww > 0, w2 = ww > 1, w3 = ww > 2, w4 = ww > 3;
#pragma unroll 1
    for (int y = 0; y < hh; ++y, rp += WP, p += W) {
        // all loads first
        const V t0 = rp[0], t1 = rp[1], t2 = rp[NW > 2 ? 2 : 0], t3 = rp[NW > 3 ? 3 : 0];
        scan_first_max_if(v, idx, t0, p, w1);
        scan_first_max_if(v, idx, t1, p + 1, w2);
        if (NW > 2) scan_first_max_if(v, idx, t2, p + 2, w3);
        if (NW > 3) scan_first_max_if(v, idx, t3, p + 3, w4);
        if (LONG)
            for (int x = NW; x < ww; ++x) scan_first_max(v, idx, rp[x], p + x);
    }
}
__device__ __forceinline__ void vneg(int& i) { i = -1; }
__device__ __forceinline__ void vneg(int4& i) { i = make_int4(-1, -1, -1, -1); }
__device__ __forceinline__ void vneg(int2& i) { i = make_int2(-1, -1); }
__device__ __forceinline__ void istore(int* o, int, int i, int) { o[0] = i; }
__device__ __forceinline__ void istore(int* o, int stride, const int4& i, int cs) {
    o[0] = i.x;
    if (cs > 1) o[stride] = i.y;
    if (cs > 2) o[2 * stride] = i.z;
    if (cs > 3) o[3 * stride] = i.w;
}
__device__ __forceinline__ void istore(int* o, int stride, const int2& i, int cs) {
    o[0] = i.x;
    if (cs > 1) o[stride] = i.y;
}
template <int CS>
struct IdxT;
template <>
struct IdxT<4> {
    typedef int4 type;
};
template <>
struct IdxT<2> {
    typedef int2 type;
};
template <>
struct IdxT<1> {
    typedef int type;
};

// Builds the max tables of one channel slab from the staged planes `raw` [cs][H*W]:
// T[lr][lc][y][x] = max of the (1<<lr) x (1<<lc) window anchored at (y,x), at tab + (lr*LV + lc)*HWp, row
// pitch WP.  Pixels are first clamped with fmaxf(., -FLT_MAX).  Ends without a barrier.
template <typename V, int LV, int THREADS, bool DIAG = false>
__device__ __forceinline__ void build_max_tables(V* tab, const float* raw, int cs, int H, int W, int WP, int HWp,
                                                 int tid) {
    const int HW = H * W;
    // T[lr][lc][y][x] = max of the (1<<lr) x (1<<lc) window anchored at (y,x), at tab + (lr*LV + lc)*HWp.
    // Windows that would leave the map are clamped; the lookups never use those entries.
    // Pixels p = tid, tid + THREADS, ... with (y,x) carried along instead of divided out each time (the
    // build is ~5 % of a CTA's instructions on the 14x14 configuration).
    const int step_y = THREADS / W, step_x = THREADS - step_y * W;
    const int y_first = tid / W, x_first = tid - y_first * W;
    auto for_pixels = [&](auto&& f) {
        int y = y_first, x = x_first;
        for (int p = tid; p < HW; p += THREADS) {
            f(p, y, x, y * WP + x);
            x += step_x;
            y += step_y;
            if (x >= W) {
                x -= W;
                ++y;
            }
        }
    };
    for_pixels([&](int p, int, int, int q) {
        V v;
        vgather(v, raw, HW, p, cs);
        tab[q] = v;
    });
    __syncthreads();
    if (LV == 1) {
        // raw pixels only: every bin is scanned
    } else if (DIAG) {
        // two tables only: pixels and 2 x 2 windows (the mixed 1 x 2 / 2 x 1 cases read the pixels twice)
        for_pixels([&](int, int y, int x, int q) {
            const int dx = x + 1 < W ? 1 : 0, qd = y + 1 < H ? q + WP : q;
            tab[HWp + q] = vmax(vmax(tab[q], tab[q + dx]), vmax(tab[qd], tab[qd + dx]));
        });
    } else if (LV == 2) {
        for_pixels([&](int, int y, int x, int q) {
            const int qr = x + 1 < W ? q + 1 : q, qd = y + 1 < H ? q + WP : q;
            const V v = tab[q];
            tab[HWp + q] = vmax(v, tab[qr]);      // 1 x 2
            tab[2 * HWp + q] = vmax(v, tab[qd]);  // 2 x 1
        });
        __syncthreads();
        for_pixels([&](int, int y, int, int q) {
            const int qd = y + 1 < H ? q + WP : q;
            tab[3 * HWp + q] = vmax(tab[HWp + q], tab[HWp + qd]);  // 2 x 2
        });
    }
}

// Recovers [lo,hi) of a bin axis from its table entry (PIPE variant: the raw extents are not kept in shared
// memory).  o1 / o2 = the two corner offsets in bytes, tsel / unit as given to tab_entry, esz = element size.
__device__ __forceinline__ int tab_decode_range_diag(int o1, int o2, bool lvl, int unit, int esz) {
    return (o1 / (unit * esz)) | ((o2 / (unit * esz) + (lvl ? 2 : 1)) << 16);
}
__device__ __forceinline__ int tab_decode_range(int o1, int o2, int tsel, int unit, int esz) {
    o1 /= esz;
    o2 /= esz;
    const int lvl = o1 >= tsel ? 1 : 0;
    const int base = lvl ? tsel : 0;
    const int lo = (o1 - base) / unit, hi = (o2 - base) / unit + (lvl ? 2 : 1);
    return lo | (hi << 16);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Bin geometry of every RoI row, once for all channel slabs: ent[r][0..P) = row entries, [P..2P) = column entries,
// exactly what the table kernel's geometry threads used to compute per CTA (same tab_entry, same constants).
template <int LV, bool MID, bool DIAG>
__global__ void __launch_bounds__(256) roi_pool_entries_kernel(RoiArgs a, int2* __restrict__ ent, int P, int esz,
                                                               int HWp) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.K * 2 * P) return;
    const int r = idx / (2 * P), ti = idx - r * 2 * P;
    const float* rp = a.rois5 + (size_t)roi_at(a, r) * 5;
    int unused;
    ent[idx] = ti < P ? tab_entry<LV, MID, DIAG>(ti, P, __ldg(rp + 2), __ldg(rp + 4), a.scale, a.H, a.pitch, LV * HWp,
                                                 esz, &unused)
                      : tab_entry<LV, MID, DIAG>(ti - P, P, __ldg(rp + 1), __ldg(rp + 3), a.scale, a.W, 1, HWp, esz,
                                                 &unused);
}

template <int P, int TAB_THREADS, int CS, int MINB, bool ARGMAX, int LV, int BPT = 1, bool PIPE = false,
          bool DIAG = false>
__global__ void __launch_bounds__(TAB_THREADS, MINB) roi_pool_tab_kernel(RoiArgs a) {
    typedef typename VecT<CS>::type V;
    constexpr int BINS = P * P;
    constexpr int SLOTS = BINS / BPT;            // thread slots per RoI: BPT horizontally adjacent bins per thread
    constexpr int RPI = TAB_THREADS / SLOTS;     // RoIs per iteration
    constexpr int NB = TAB_THREADS / P;          // RoIs per batch: two table entries per thread
    constexpr int ITERS = NB / RPI;
    static_assert(BPT == 1 || (BPT == 2 && P % 2 == 0 && CS == 4 && LV == 2 && !ARGMAX), "bin pairs: 14x14 inference");
    // PIPE: three geometry buffers and one mbarrier per buffer instead of a CTA barrier per batch.  A thread
    // waits for batch b's geometry, computes batch b+1's, then pools batch b: warps may drift one batch apart
    // (measured upper bound of removing the per-batch barrier on the 14x14 configuration: ~10 %).
    constexpr int NBUF = PIPE ? 3 : 2;
    constexpr int NWARPS = (TAB_THREADS + 31) / 32;
    // tables: (row level, column level), levels 1 (, 2); DIAG keeps only (1,1) and (2,2) -- half the shared
    // memory, so a map too large for four 4-channel tables still gets four channels per lookup -- and a bin
    // with one axis 1 long and the other 2..4 reads the pixels along the long axis instead
    constexpr int NT = DIAG ? 2 : LV * LV;
    static_assert(LV == 1 || LV == 2, "table levels");
    static_assert(!DIAG || (LV == 2 && BPT == 1 && !ARGMAX && PIPE), "two-table form: single-bin inference kernels");
    // 5..8-long bins through four 2-windows: 7x7 bins only (a 14x14 grid needs a RoI > 56 pixels wide for
    // one, and the extra call site costs the 14x14 fast path registers)
    constexpr bool MID = LV == 2 && P == 7;
    // inference: the per-RoI geometry is the same for every channel slab, so it is computed once per RoI by
    // roi_pool_entries_kernel and only copied into the shared-memory buffers here (computing it per CTA was
    // 28 % of this kernel's instructions on the 7x7 grid)
    // (not the 14x14 bin-pair kernel: its geometry is 11 % of the instructions, it is bound by the LSU pipe and
    // HBM instead, and loading the entries measured 5 % slower than computing them there)
    constexpr bool PRE = PIPE && !ARGMAX && BPT == 1;
    static_assert(RPI * SLOTS == TAB_THREADS && NB * P == TAB_THREADS && ITERS * RPI == NB, "thread mapping");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) int2 s_th[NBUF][NB][P], s_tw[NBUF][NB][P];  // per RoI: row / column corner offsets + flags
    // raw [lo,hi) extents: kept for the training scan (every bin uses them); the inference kernels recover them
    // from the entries on their rare long-bin path, which leaves room for the third geometry buffer
    constexpr bool RAW = ARGMAX || !PIPE;
    __shared__ int s_hraw[RAW ? NBUF : 1][RAW ? NB : 1][RAW ? P : 1], s_wraw[RAW ? NBUF : 1][RAW ? NB : 1][RAW ? P : 1];
    __shared__ size_t s_ob[NBUF][NB];                // per RoI: byte offset of its [CS,P,P] output block
    __shared__ __align__(8) uint64_t s_full[PIPE ? 3 : 1];  // PIPE: geometry of buffer i complete
    V* tab = reinterpret_cast<V*>(smem_raw);
    const int H = a.H, W = a.W, HW = H * W;
    const int WP = a.pitch, HWp = (H * WP + 3) & ~3;  // row pitch (odd when rows would alias banks), table stride
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CS;
    const int cs = min(CS, a.C - c0);
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int stride = a.groups * NB;
    int r0 = r_begin + blockIdx.x * NB;  // first RoI of this CTA's current batch
    if (r0 >= r_end) return;
    const int tid = threadIdx.x;
    // table-entry role: axis entry ti (rows first, then columns) of RoIs tj and tj + NB/2 of the batch
    const int tj = tid / (2 * P), ti = tid % (2 * P);
    const bool trow = ti < P;  // this thread computes a row (else a column) entry
    auto load_geo = [&](int r) {
        return PRE ? load_roi_entry(a, r, r_end, ti, 2 * P) : load_roi_axis(a, r, r_end, trow);
    };
    RoiAxis nx0 = load_geo(r0 + tj), nx1 = load_geo(r0 + tj + NB / 2);

    // staged planes [cs][HW]: where the last-built table will be (LV = 1: a region of their own)
    float* raw = reinterpret_cast<float*>(tab + (LV == 1 ? 1 : NT - 1) * HWp);
    if (PIPE && tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], NWARPS);  // one arrival per warp; visible after the
    }                                                                // barriers inside stage_slab
    stage_slab(raw, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);
    build_max_tables<V, LV, TAB_THREADS, DIAG>(tab, raw, cs, H, W, WP, HWp, tid);

    // compute role: bin e = (ph,pw) (and its right neighbour when BPT = 2) of the (tid / SLOTS)-th RoI of
    // each iteration
    const int slot = tid % SLOTS, ej = tid / SLOTS;
    const int ph = slot / (P / BPT), pw = (slot % (P / BPT)) * BPT;
    const int e = ph * P + pw;
    auto fill_tables = [&](int buf, int j, const RoiAxis& q) {
        int unused;
        if (PRE) {
            const int2 e = make_int2(__float_as_int(q.c1), __float_as_int(q.c2));
            if (trow) s_th[buf][j][ti] = e;
            else s_tw[buf][j][ti - P] = e;
        } else if (trow)
            s_th[buf][j][ti] = tab_entry<LV, MID, DIAG>(ti, P, q.c1, q.c2, a.scale, H, WP, LV * HWp, sizeof(V),
                                                  RAW ? &s_hraw[RAW ? buf : 0][RAW ? j : 0][RAW ? ti : 0] : &unused);
        else
            s_tw[buf][j][ti - P] = tab_entry<LV, MID, DIAG>(ti - P, P, q.c1, q.c2, a.scale, W, 1, HWp, sizeof(V),
                                                      RAW ? &s_wraw[RAW ? buf : 0][RAW ? j : 0][RAW ? ti - P : 0]
                                                          : &unused);
        if (ti == 0)
            s_ob[buf][j] = (((size_t)max(q.k, 0) * a.C + c0) * BINS) * sizeof(float);
    };
    fill_tables(0, tj, nx0);
    fill_tables(0, tj + NB / 2, nx1);
    nx0 = load_geo(r0 + stride + tj);
    nx1 = load_geo(r0 + stride + tj + NB / 2);
    int cur = 0;
    uint32_t batch = 0;
    if (PIPE) {
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&s_full[0]);  // this warp's share of batch 0's geometry is written
        __syncthreads();                               // the last table pass is complete
    }
    for (; r0 < r_end; r0 += stride, ++batch) {
        int nxt;
        if (PIPE) {
            // every warp has written batch `batch` (so it has also finished pooling batch - 2, whose buffer
            // batch + 1 overwrites below); warps may be one batch apart, never two
            cur = (int)(batch % 3u);
            nxt = (int)((batch + 1u) % 3u);
            mbar_wait(&s_full[cur], (batch / 3u) & 1u);
        } else {
            cur = (int)(batch & 1u);
            nxt = cur ^ 1;
            __syncthreads();  // tables[cur] (and, first time, T22) complete; tables[cur^1] no longer read
        }
        fill_tables(nxt, tj, nx0);  // geometry of the next batch
        fill_tables(nxt, tj + NB / 2, nx1);
        if (PIPE) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&s_full[nxt]);
        }
        bool prefetched = false;
        auto prefetch_boxes = [&]() {  // boxes of the batch after the next one
            nx0 = load_geo(r0 + 2 * stride + tj);
            nx1 = load_geo(r0 + 2 * stride + tj + NB / 2);
            prefetched = true;
        };
        if (!(BPT == 2 && PIPE)) prefetch_boxes();
        const int nb = min(NB, r_end - r0);
        // one bin (this thread's ph,pw) of RoI j of the batch, all CS channels
        auto one_bin = [&](int j, bool full, bool valid) {
            const int2 h = s_th[cur][j][ph], w = s_tw[cur][j][pw];
            if (ARGMAX) {
                // training: one scan of the window gives the maximum and the first position that attains it
                // (the reference's `v > best` order); clamped NaN / -inf pixels are never selected
                V v;
                typename IdxT<CS>::type idx;
                vsplat(v, -FLT_MAX);
                vneg(idx);
                const unsigned m = (unsigned)((h.y & w.y) >> 31);  // all ones iff the bin is non-empty
                // [lo,hi) per axis, hi >= lo: an empty axis gives zero rows or only predicated-off columns
                const int hr = s_hraw[RAW ? cur : 0][RAW ? j : 0][RAW ? ph : 0];
                const int wr = s_wraw[RAW ? cur : 0][RAW ? j : 0][RAW ? pw : 0];
                const int y0 = hr & 0xFFFF, x0 = wr & 0xFFFF, hh = (hr >> 16) - y0, ww = (wr >> 16) - x0;
                // the nested loops the compiler made of "for y, for x" spent two thirds of their instructions on
                // control flow and re-derived addresses (450 per 32 bins x 4 channels, of which 120 compare /
                // select); the row body is chosen per warp by the widest window among its lanes
                const V* rp = tab + (y0 * WP + x0);
                const int p0 = y0 * W + x0;
                if (!__any_sync(0xFFFFFFFFu, ww > 2)) scan_rows<2, false>(v, idx, rp, p0, hh, ww, WP, W);
                else if (!__any_sync(0xFFFFFFFFu, ww > 3)) scan_rows<3, false>(v, idx, rp, p0, hh, ww, WP, W);
                else if (!__any_sync(0xFFFFFFFFu, ww > 4)) scan_rows<4, false>(v, idx, rp, p0, hh, ww, WP, W);
                else scan_rows<4, true>(v, idx, rp, p0, hh, ww, WP, W);
                float* o = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.out) + s_ob[cur][j]) + e;
                if (valid) {
                    vstore<false>(o, BINS, v, m, cs);
                    istore(reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(a.argmax) + s_ob[cur][j]) + e, BINS,
                           idx, cs);
                }
                return;
            }
            const int hy = h.y & (DIAG ? TAB_DIAG_MASK : TAB_OFF_MASK), wy = w.y & (DIAG ? TAB_DIAG_MASK : TAB_OFF_MASK);
            // DIAG: both axes 2-windows -> table (2,2); otherwise the pixels, and the axis that is 2..4 long is
            // covered by the pixel pairs at its two anchors (dr / dc = one pixel along that axis)
            const bool lr = DIAG && (h.y & TAB_LVL_BIT) != 0, lc = DIAG && (w.y & TAB_LVL_BIT) != 0;
            const unsigned char* tb = smem_raw + ((lr && lc) ? HWp * (int)sizeof(V) : 0);
            V v = *reinterpret_cast<const V*>(tb + (w.x + h.x));
            if (DIAG) {
                const int dr = (lr && !lc) ? WP * (int)sizeof(V) : 0, dc = (lc && !lr) ? (int)sizeof(V) : 0;
                const bool wide = w.x != wy || dr != 0, tall = h.x != hy || dc != 0;
                const bool any_wide = __any_sync(0xFFFFFFFFu, wide), any_tall = __any_sync(0xFFFFFFFFu, tall);
                if (any_wide) {
                    if (wide) v = vmax(v, *reinterpret_cast<const V*>(tb + (wy + h.x + dr)));
                }
                if (any_tall) {
                    if (tall) v = vmax(v, *reinterpret_cast<const V*>(tb + (w.x + hy + dc)));
                    if (any_wide) {
                        if (wide && tall) v = vmax(v, *reinterpret_cast<const V*>(tb + (wy + hy + dr + dc)));
                    }
                }
            } else if (LV > 1) {
                // lookups that coincide with the first one are skipped: warp-uniformly when no lane needs
                // them (saves the issue slots), per lane otherwise (idle lanes cost no LSU wavefronts)
                const bool wide = w.x != wy, tall = h.x != hy;
                const bool any_wide = __any_sync(0xFFFFFFFFu, wide), any_tall = __any_sync(0xFFFFFFFFu, tall);
                if (any_wide) {
                    if (wide) v = vmax(v, *reinterpret_cast<const V*>(smem_raw + (wy + h.x)));
                }
                if (any_tall) {
                    if (tall) v = vmax(v, *reinterpret_cast<const V*>(smem_raw + (w.x + hy)));
                    if (any_wide) {
                        if (wide && tall) v = vmax(v, *reinterpret_cast<const V*>(smem_raw + (wy + hy)));
                    }
                }
            }
            // DIAG: the four 2-windows per axis need table (2,2); a 5..8-long bin that is 1 thick is scanned
            const bool mid_ok = !DIAG || (lr && lc);
            if (MID) {
                const bool mid = ((h.y | w.y) & TAB_MID_BIT) != 0 && mid_ok;
                if (__any_sync(0xFFFFFFFFu, mid)) {
                    if (mid)
                        v = tab_mid_bin<V>(v, tb, h.x, hy, (h.y & TAB_MID_BIT) != 0, w.x, wy,
                                           (w.y & TAB_MID_BIT) != 0, 2 * WP * (int)sizeof(V), 2 * (int)sizeof(V));
                }
            }
            const bool big = ((h.y | w.y) & TAB_BIG_BIT) != 0 || (MID && !mid_ok && ((h.y | w.y) & TAB_MID_BIT) != 0);
            if (__any_sync(0xFFFFFFFFu, big)) {
                if (big) {
                    if (DIAG)
                        v = tab_big_bin<V>(tab, tab_decode_range_diag(h.x, hy, lr, WP, sizeof(V)),
                                           tab_decode_range_diag(w.x, wy, lc, 1, sizeof(V)), WP);
                    else
                        v = tab_big_bin<V>(tab,
                                           RAW ? s_hraw[RAW ? cur : 0][RAW ? j : 0][RAW ? ph : 0]
                                               : tab_decode_range(h.x, hy, LV * HWp, WP, sizeof(V)),
                                           RAW ? s_wraw[RAW ? cur : 0][RAW ? j : 0][RAW ? pw : 0]
                                               : tab_decode_range(w.x, wy, HWp, 1, sizeof(V)),
                                           WP);
                }
            }
            const unsigned m = (unsigned)((h.y & w.y) >> 31);  // all ones iff the bin is non-empty
            float* o = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.out) + s_ob[cur][j]) + e;
            if (full) vstore<true>(o, BINS, v, m, cs);
            else if (valid) vstore<false>(o, BINS, v, m, cs);
        };
        // two horizontally adjacent bins (pw, pw+1) of RoI j, all four channels: one row entry, one 16-byte
        // load for both column entries, one vote per decision, 8-byte stores (a 14x14 block is 784 bytes and
        // even bins sit on 8-byte boundaries): ~20 % fewer instructions and a quarter fewer store wavefronts
        // per output than two single bins
        auto one_pair = [&](int j, bool full, bool valid) {
          if constexpr (BPT == 2) {
            const int2 h = s_th[cur][j][ph];
            const int4 w = *reinterpret_cast<const int4*>(&s_tw[cur][j][pw]);  // (x0, y0 | flags, x1, y1 | flags)
            const int hy = h.y & TAB_OFF_MASK, wy0 = w.y & TAB_OFF_MASK, wy1 = w.w & TAB_OFF_MASK;
            const bool tall = h.x != hy, wide0 = w.x != wy0, wide1 = w.z != wy1;
            const bool any_wide = __any_sync(0xFFFFFFFFu, wide0 || wide1), any_tall = __any_sync(0xFFFFFFFFu, tall);
            float4 v0 = *reinterpret_cast<const float4*>(smem_raw + (w.x + h.x));
            float4 v1 = *reinterpret_cast<const float4*>(smem_raw + (w.z + h.x));
            if (any_wide) {
                if (wide0) v0 = vmax(v0, *reinterpret_cast<const float4*>(smem_raw + (wy0 + h.x)));
                if (wide1) v1 = vmax(v1, *reinterpret_cast<const float4*>(smem_raw + (wy1 + h.x)));
            }
            if (any_tall) {
                if (tall) {
                    v0 = vmax(v0, *reinterpret_cast<const float4*>(smem_raw + (w.x + hy)));
                    v1 = vmax(v1, *reinterpret_cast<const float4*>(smem_raw + (w.z + hy)));
                }
                if (any_wide) {
                    if (wide0 && tall) v0 = vmax(v0, *reinterpret_cast<const float4*>(smem_raw + (wy0 + hy)));
                    if (wide1 && tall) v1 = vmax(v1, *reinterpret_cast<const float4*>(smem_raw + (wy1 + hy)));
                }
            }
            const bool big0 = ((h.y | w.y) & TAB_BIG_BIT) != 0, big1 = ((h.y | w.w) & TAB_BIG_BIT) != 0;
            if (__any_sync(0xFFFFFFFFu, big0 || big1)) {
                const int hr = RAW ? s_hraw[RAW ? cur : 0][RAW ? j : 0][RAW ? ph : 0]
                                   : tab_decode_range(h.x, hy, LV * HWp, WP, sizeof(V));
                if (big0)
                    v0 = tab_big_bin<float4>(tab, hr, RAW ? s_wraw[RAW ? cur : 0][RAW ? j : 0][RAW ? pw : 0]
                                                          : tab_decode_range(w.x, wy0, HWp, 1, sizeof(V)), WP);
                if (big1)
                    v1 = tab_big_bin<float4>(tab, hr, RAW ? s_wraw[RAW ? cur : 0][RAW ? j : 0][RAW ? pw + 1 : 0]
                                                          : tab_decode_range(w.z, wy1, HWp, 1, sizeof(V)), WP);
            }
            const unsigned m0 = (unsigned)((h.y & w.y) >> 31), m1 = (unsigned)((h.y & w.w) >> 31);
            float2* o = reinterpret_cast<float2*>(
                reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.out) + s_ob[cur][j]) + e);
            auto put = [&](int c, float x0, float x1) {
                o[c * (BINS / 2)] = make_float2(__uint_as_float(__float_as_uint(x0) & m0),
                                                __uint_as_float(__float_as_uint(x1) & m1));
            };
            if (full || valid) {
                put(0, v0.x, v1.x);
                if (full || cs > 1) put(1, v0.y, v1.y);
                if (full || cs > 2) put(2, v0.z, v1.z);
                if (full || cs > 3) put(3, v0.w, v1.w);
            }
          }
        };
        if (nb == NB && cs == CS) {
            if constexpr (ARGMAX) {
                // rolled: unrolled over the 7 RoIs of a thread the training scan ran 0.170 ms instead of 0.100 ms
                // (instruction footprint); sharing one copy of the bin code with the ragged path below costs 6 %
#pragma unroll 1
                for (int it = 0; it < ITERS; ++it) one_bin(it * RPI + ej, true, true);
            } else if constexpr (BPT == 2) {
                // not unrolled (registers); prefetching the next RoI's entries by hand was measured slower too
#pragma unroll 1
                for (int it = 0; it < ITERS; ++it) {
                    // issued after the first RoI: at the top of the batch the loads shared a scoreboard with
                    // the loop's first constant loads, which then waited an L2 round trip every batch
                    if (PIPE && it == 1) prefetch_boxes();
                    one_pair(it * RPI + ej, true, true);
                }
            } else {
#pragma unroll
                for (int it = 0; it < ITERS; ++it) one_bin(it * RPI + ej, true, true);
            }
        } else {
            for (int it = 0; it * RPI < nb; ++it) {
                const int j = it * RPI + ej;
                if (BPT == 2) one_pair(j < nb ? j : 0, false, j < nb);
                else one_bin(j < nb ? j : 0, false, j < nb);
            }
        }
        if (!prefetched) prefetch_boxes();
    }
}

// Loads through 32-bit shared-window addresses kept in registers: with generic pointers into __shared__ arrays the
// compiler re-derived the window base (S2R SR_CgaCtaId, MOV, LEA) in every iteration of the bin loop.
__device__ __forceinline__ int lds_i32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
// ---------------------------------------------------------------------------------------------
// RoIPool forward (inference) from per-bin lookup lists.
//
// roi_pool_tab_kernel decides per bin and per slab CTA which table and which corners to read (window levels, the
// thin-bin cases of the two-table form, coinciding lookups, 5..8-long bins): ~130 warp instructions per 32 bins x 4
// channels, of which 24 are the lookups, maxima and stores themselves (ncu, 64 x 64 map).  None of that depends on
// the channel.  Here roi_pool_desc_kernel works it out ONCE per (RoI, bin) and writes the result as a list of up to
// 16 table positions (16-bit offsets into [pixels | 2 x 2 windows | a zero pixel | a -FLT_MAX pixel]); the gather of
// a slab only loads the list, takes the maximum of what it points at and stores.  A bin's first four positions
// are read without a predicate (unused slots point at the -FLT_MAX pixel, an empty bin at the zero pixel), the
// other twelve only by the lanes that have them, a bin needing more is scanned.  max is exact and commutative, so
// the value is the same bit pattern whatever covers the bin.
//   thin bin (1 x L or L x 1, L <= 16): its pixels;  otherwise 2-windows anchored at lo, lo + 2, ... and hi - 2 per
//   axis (ceil(L / 2) of them): the product of the two axes' anchors, at most 16 (L <= 8 on both axes always fits).
// ---------------------------------------------------------------------------------------------
constexpr unsigned PD_MASK = 0x7FFFu;

struct PoolDescTab {
    int HWp;   // table stride: pixels at 0, 2 x 2 windows at HWp, zero pixel at 2 * HWp, -FLT_MAX pixel at 2 * HWp + 1
    int WP;
};

// k-th window of side s covering [lo, hi), hi - lo >= s: lo, lo + s, ... and hi - s for the remainder
__device__ __forceinline__ int pd_anchor(int lo, int hi, int s, int full, int k) { return k < full ? lo + k * s : hi - s; }

// smax = side of the largest window table the gather builds (2 or 3)
__global__ void __launch_bounds__(256) roi_pool_desc_kernel(RoiArgs a, uint2* __restrict__ desc,
                                                            uint2* __restrict__ extra, int P, int HWp, int smax) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int BINS = P * P;
    if (idx >= a.K * BINS) return;
    const int r = idx / BINS, e = idx - r * BINS, ph = e / P, pw = e - ph * P;
    const float* rp = a.rois5 + (size_t)roi_at(a, r) * 5;
    int hr, wr;
    tab_entry<1>(ph, P, __ldg(rp + 2), __ldg(rp + 4), a.scale, a.H, 1, 0, 4, &hr);
    tab_entry<1>(pw, P, __ldg(rp + 1), __ldg(rp + 3), a.scale, a.W, 1, 0, 4, &wr);
    const int y0 = hr & 0xFFFF, y1 = hr >> 16, x0 = wr & 0xFFFF, x1 = wr >> 16, hh = y1 - y0, ww = x1 - x0;
    const int WP = a.pitch;
    const unsigned zero = (unsigned)(smax * HWp), none = zero + 1u;
    // square windows of side s = min(smax, hh, ww) from table s - 1: a 3 x 3 bin is one lookup, 2 x 3 two, ...
    const bool empty = hh <= 0 || ww <= 0;
    const int s = empty ? 1 : min(smax, min(hh, ww));
    // (s is 1, 2 or 3: no general integer division)
    auto div_s = [s](int v) { return s == 1 ? v : s == 2 ? v >> 1 : v / 3; };
    const int fy = div_s(hh), fx = div_s(ww), nry = div_s(hh + s - 1), ncx = div_s(ww + s - 1);
    const int n = empty ? 0 : nry * ncx;
    const bool big = n > 16;
    const int base = (s - 1) * HWp;
    unsigned o[16];  // static indices only: registers
    int i = 0, j = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        o[t] = (t < n && !big) ? (unsigned)(base + pd_anchor(y0, y1, s, fy, i) * WP + pd_anchor(x0, x1, s, fx, j)) : none;
        if (++j == ncx) j = 0, ++i;
        if (t == 3 && n <= 4) break;  // (the other twelve are only stored for longer lists)
    }
    if (empty) o[0] = zero;
    const bool more = n > 4 && !big;
    desc[idx] = make_uint2(o[0] | (o[1] << 16) | (more ? 0x80000000u : 0u), o[2] | (o[3] << 16) | (big ? 0x80000000u : 0u));
    if (more) {
        uint2* x = extra + (size_t)idx * 3;
        x[0] = make_uint2(o[4] | (o[5] << 16), o[6] | (o[7] << 16));
        x[1] = make_uint2(o[8] | (o[9] << 16), o[10] | (o[11] << 16));
        x[2] = make_uint2(o[12] | (o[13] << 16), o[14] | (o[15] << 16));
    }
}

// positions 5..16 of a bin's list (the lanes that have them): stops at the first unused slot
__device__ __noinline__ float4 pool_desc_more(float4 v, uint32_t tab_s, const uint2* __restrict__ x, unsigned none) {
    // all three words at once (the list kernel wrote them all): one L2 round trip, not one per word
    const uint2 w[3] = {__ldg(x), __ldg(x + 1), __ldg(x + 2)};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const unsigned o[4] = {w[q].x & 0xFFFFu, w[q].x >> 16, w[q].y & 0xFFFFu, w[q].y >> 16};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (o[i] == none) return v;
            v = vmax(v, lds_f4<0>(tab_s + o[i] * 16u));
        }
    }
    return v;
}
// a bin too large for a list: its ranges again from the RoI, every pixel read
__device__ __noinline__ float4 pool_desc_scan(const float* __restrict__ rp, float scale, int HW16, int pitch,
                                              uint32_t tab_s, int ph, int pw, int P) {
    int hr, wr;
    tab_entry<1>(ph, P, __ldg(rp + 2), __ldg(rp + 4), scale, HW16 >> 16, 1, 0, 4, &hr);
    tab_entry<1>(pw, P, __ldg(rp + 1), __ldg(rp + 3), scale, HW16 & 0xFFFF, 1, 0, 4, &wr);
    float4 v = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int y = hr & 0xFFFF; y < (hr >> 16); ++y)
        for (int x = wr & 0xFFFF; x < (wr >> 16); ++x) v = vmax(v, lds_f4<0>(tab_s + (uint32_t)(y * pitch + x) * 16u));
    return v;
}

// the slab's window tables: pixels (clamped), 2 x 2 (, 3 x 3) windows, the zero pixel and the -FLT_MAX pixel
template <int THREADS, int NT>
__device__ __forceinline__ void pool_list_tables(const RoiArgs& a, float4* tab, int b, int c0, uint64_t* bar) {
    typedef float4 V;
    const int H = a.H, W = a.W, HW = H * W, WP = a.pitch, HWp = (H * WP + 3) & ~3;
    const int tid = threadIdx.x;
    // planes staged where the last window table will be (dead once the pixels are interleaved)
    float* raw = reinterpret_cast<float*>(tab + (NT - 1) * HWp);
    stage_slab(raw, a.feat + ((size_t)b * a.C + c0) * HW, 4 * HW, bar);
    build_max_tables<V, 1, THREADS>(tab, raw, 4, H, W, WP, HWp, tid);  // pixels, clamped; ends with a CTA barrier
    for (int p = tid; p < HW; p += THREADS) {
        // windows anchored at (y, x), clamped at the map's edge (the lists never point at a clamped one)
        const int y = p / W, x = p - y * W, q = y * WP + x;
        const int d1 = x + 1 < W ? 1 : 0, d2 = x + 2 < W ? 2 : d1;
        const int q1 = y + 1 < H ? q + WP : q, q2 = y + 2 < H ? q + 2 * WP : q1;
        const V m2 = vmax(vmax(tab[q], tab[q + d1]), vmax(tab[q1], tab[q1 + d1]));
        tab[HWp + q] = m2;
        if (NT == 3) {
            const V c = vmax(vmax(tab[q + d2], tab[q1 + d2]), vmax(tab[q2], tab[q2 + d1]));
            tab[2 * HWp + q] = vmax(vmax(m2, c), tab[q2 + d2]);
        }
    }
    if (tid == 0) {
        tab[NT * HWp] = make_float4(0.f, 0.f, 0.f, 0.f);
        tab[NT * HWp + 1] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    }
    __syncthreads();
}

// maximum over one bin's list (called by whole warps: the votes skip what no lane needs)
__device__ __forceinline__ float4 pool_list_max(const uint2& d, uint32_t tab_s, unsigned none, const uint2* __restrict__ xtra,
                                                const float* __restrict__ roi, float scale, int HW16, int WP, int e, int P) {
    typedef float4 V;
    const V t0 = lds_f4<0>(tab_s + (d.x & PD_MASK) * 16u), t1 = lds_f4<0>(tab_s + ((d.x >> 16) & PD_MASK) * 16u);
    V v = vmax(t0, t1);
    if (__any_sync(0xFFFFFFFFu, (d.y & PD_MASK) != none)) {  // slots 3 and 4: only when some lane has a third position
        const V t2 = lds_f4<0>(tab_s + (d.y & PD_MASK) * 16u), t3 = lds_f4<0>(tab_s + ((d.y >> 16) & PD_MASK) * 16u);
        v = vmax(v, vmax(t2, t3));
    }
    const bool more = (int)d.x < 0, big = (int)d.y < 0;
    if (__any_sync(0xFFFFFFFFu, more)) {
        if (more) v = pool_desc_more(v, tab_s, xtra, none);
    }
    if (__any_sync(0xFFFFFFFFu, big)) {
        if (big) v = pool_desc_scan(roi, scale, HW16, WP, tab_s, e / P, e % P, P);
    }
    return v;
}

template <int P, int THREADS, int MINB, int NT>
__global__ void __launch_bounds__(THREADS, MINB)
roi_pool_gather_kernel(RoiArgs a, const uint2* __restrict__ desc, const uint2* __restrict__ extra) {
    typedef float4 V;
    constexpr int BINS = P * P;
    static_assert(NT == 2 || NT == 3, "window tables: 1, 2 (, 3)");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int H = a.H, W = a.W, WP = a.pitch, HWp = (H * WP + 3) & ~3;
    V* tab = reinterpret_cast<V*>(smem_raw);
    const int tid = threadIdx.x, b = blockIdx.z, c0 = blockIdx.y * 4;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int n_task = (r_end - r_begin) * BINS, step = a.groups * THREADS;
    if ((int)blockIdx.x * THREADS >= n_task) return;
    pool_list_tables<THREADS, NT>(a, tab, b, c0, &bar);
    uint32_t tab_s = smem_u32(tab);
    asm volatile("mov.u32 %0, %0;" : "+r"(tab_s));
    const unsigned none = (unsigned)(NT * HWp) + 1u;
    const uint2 idle = make_uint2(none | (none << 16), none | (none << 16));
    const uint2* const dbase = desc + (size_t)r_begin * BINS;
    unsigned char* const obase = reinterpret_cast<unsigned char*>(a.out + (size_t)c0 * BINS);
    const uint32_t kstride_b = (uint32_t)a.C * BINS * 4u;
    int i = blockIdx.x * THREADS + tid;
    uint2 d = i < n_task ? __ldg(dbase + i) : idle;
    // two tasks per trip: the second one's store addresses get registers of their own instead of waiting for the first
    // one's STGs to leave the memory-instruction queue (0.547 -> 0.541 ms on the 64 x 64 map)
#pragma unroll 2
    for (int i0 = blockIdx.x * THREADS; i0 < n_task; i0 += step, i += step) {
        const bool valid = i < n_task;
        const uint2 dn = i + step < n_task ? __ldg(dbase + i + step) : idle;  // the next task's list
        const int rl = i / BINS, e = i - rl * BINS;
        const int k = valid ? roi_at(a, r_begin + rl) : 0;
        const V v = pool_list_max(d, tab_s, none, extra + ((size_t)r_begin * BINS + i) * 3, a.rois5 + (size_t)k * 5, a.scale,
                                  (H << 16) | W, WP, e, P);
        if (valid) {
            float* po = reinterpret_cast<float*>(obase + ((unsigned long long)(unsigned)k * kstride_b + (unsigned)(e * 4)));
            po[0] = v.x, po[BINS] = v.y, po[2 * BINS] = v.z, po[3 * BINS] = v.w;
        }
        d = dn;
    }
}

// RoIPool + the head's global average from the same lists: [K,C] instead of [K,C,P,P].  A warp owns a RoI: lane l
// takes bins l and l + 32, the 49 bin maxima are added in a fixed order (own two bins, then an xor tree), so the
// result does not depend on how the RoIs are grouped and is run-to-run identical; it agrees with pool().mean() to
// fp32 rounding, not bit for bit.
template <int P, int THREADS, int MINB, int NT>
__global__ void __launch_bounds__(THREADS, MINB)
roi_pool_mean_list_kernel(RoiArgs a, const uint2* __restrict__ desc, const uint2* __restrict__ extra) {
    typedef float4 V;
    constexpr int BINS = P * P, WARPS = THREADS / 32;
    static_assert(BINS > 32 && BINS <= 64, "two bins per lane");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int H = a.H, W = a.W, WP = a.pitch, HWp = (H * WP + 3) & ~3;
    V* tab = reinterpret_cast<V*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, b = blockIdx.z, c0 = blockIdx.y * 4;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int n_roi = r_end - r_begin, step = a.groups * WARPS;
    if ((int)blockIdx.x * WARPS >= n_roi) return;
    pool_list_tables<THREADS, NT>(a, tab, b, c0, &bar);
    uint32_t tab_s = smem_u32(tab);
    asm volatile("mov.u32 %0, %0;" : "+r"(tab_s));
    const unsigned none = (unsigned)(NT * HWp) + 1u;
    const uint2 idle = make_uint2(none | (none << 16), none | (none << 16));
    const bool second = lane + 32 < BINS;
    int rl = blockIdx.x * WARPS + (tid >> 5);
    auto list = [&](int r, int e, bool on) {
        return on && r < n_roi ? __ldg(desc + (size_t)(r_begin + r) * BINS + e) : idle;
    };
    uint2 d0 = list(rl, lane, true), d1 = list(rl, lane + 32, second);
#pragma unroll 1
    for (; rl < n_roi; rl += step) {
        const uint2 n0 = list(rl + step, lane, true), n1 = list(rl + step, lane + 32, second);  // the next RoI's lists
        const int k = roi_at(a, r_begin + rl);
        const float* roi = a.rois5 + (size_t)k * 5;
        const uint2* x = extra + ((size_t)(r_begin + rl) * BINS + lane) * 3;
        V s = pool_list_max(d0, tab_s, none, x, roi, a.scale, (H << 16) | W, WP, lane, P);
        const V v1 = pool_list_max(d1, tab_s, none, x + 32 * 3, roi, a.scale, (H << 16) | W, WP, second ? lane + 32 : 0, P);
        if (second) s.x += v1.x, s.y += v1.y, s.z += v1.z, s.w += v1.w;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            s.x += __shfl_xor_sync(0xFFFFFFFFu, s.x, m);
            s.y += __shfl_xor_sync(0xFFFFFFFFu, s.y, m);
            s.z += __shfl_xor_sync(0xFFFFFFFFu, s.z, m);
            s.w += __shfl_xor_sync(0xFFFFFFFFu, s.w, m);
        }
        if (lane == 0) {
            float* o = a.out + (size_t)k * a.C + c0;
            o[0] = s.x / (float)BINS, o[1] = s.y / (float)BINS, o[2] = s.z / (float)BINS, o[3] = s.w / (float)BINS;
        }
        d0 = n0, d1 = n1;
    }
}

// ---------------------------------------------------------------------------------------------
// RoIPool forward for training (value + argmax, 7x7 / 14x14): persistent CTAs over (image, 4-channel slab) items.
//
// The table kernel above spends half of its time outside the scan when an image has only ~128 sampled RoIs
// (ncu, training configuration: 24 % of the stall samples in the per-CTA prologue -- waiting for the slab -- and
// 26 % in the per-batch bin geometry + CTA barrier, which every one of the 128 slab CTAs of an image repeats).
// Here (1) the [lo,hi) ranges of every bin row / column are computed once per RoI by roi_pool_ranges_kernel and
// a CTA copies the ranges of up to TR_CHUNK RoIs into shared memory in one go: no geometry arithmetic, no
// per-batch barrier, no double buffering; (2) a CTA keeps taking items from an atomic counter and asks for the
// NEXT item's planes (cp.async.bulk into the other of two staging buffers) before it touches the current one,
// so the slab latency is paid once per CTA instead of once per item; (3) the bin scan is the predicated
// straight-line row body (scan_rows).  Same mapping as before: thread = one bin of one RoI (392 = 8 x 49 =
// 2 x 196), consecutive lanes = consecutive bins, value and argmax stored in 128-byte runs.
// ---------------------------------------------------------------------------------------------
// scan_rows on a shared-window address (`row_bytes` = table pitch in bytes)
template <int NW, bool LONG>
__device__ __forceinline__ void scan_rows_s(float4& v, int4& idx, uint32_t addr, int p, int hh, int ww, int row_bytes,
                                            int W) {
    const bool w1 = ww > 0, w2 = ww > 1, w3 = ww > 2, w4 = ww > 3;
#pragma unroll 1
    for (int y = 0; y < hh; ++y, addr += row_bytes, p += W) {
        constexpr int O2 = NW > 2 ? 32 : 0, O3 = NW > 3 ? 48 : 0;
        const float4 t0 = lds_f4<0>(addr), t1 = lds_f4<16>(addr), t2 = lds_f4<O2>(addr), t3 = lds_f4<O3>(addr);
        scan_first_max_if(v, idx, t0, p, w1);
        scan_first_max_if(v, idx, t1, p + 1, w2);
        if (NW > 2) scan_first_max_if(v, idx, t2, p + 2, w3);
        if (NW > 3) scan_first_max_if(v, idx, t3, p + 3, w4);
        if (LONG)
            for (int x = NW; x < ww; ++x) scan_first_max(v, idx, lds_f4<0>(addr + 16 * x), p + x);
    }
}

constexpr int TR_CHUNK = 128;  // RoIs whose ranges are resident at a time (a training image has 128 samples)

// rng[r][0..P) = bin rows, [P..2P) = bin columns of the RoI at position r of the image-ordered list, lo | hi << 16
// (tab_entry: the same rounding / division / floor / ceil / clamp as every RoIPool kernel of this file).  Also
// resets the work counter of the gather that follows.
__global__ void __launch_bounds__(256) roi_pool_ranges_kernel(RoiArgs a, int* __restrict__ rng, int P,
                                                              int* __restrict__ counter, int first_item) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx == 0) *counter = first_item;
    if (idx >= a.K * 2 * P) return;
    const int r = idx / (2 * P), ti = idx - r * 2 * P;
    const float* rp = a.rois5 + (size_t)roi_at(a, r) * 5;
    int raw;
    if (ti < P) tab_entry<1>(ti, P, __ldg(rp + 2), __ldg(rp + 4), a.scale, a.H, 1, 0, 4, &raw);
    else tab_entry<1>(ti - P, P, __ldg(rp + 1), __ldg(rp + 3), a.scale, a.W, 1, 0, 4, &raw);
    rng[idx] = raw;
}

template <int P, int TR_THREADS>
__global__ void __launch_bounds__(TR_THREADS, 1024 / TR_THREADS)
roi_pool_train_kernel(RoiArgs a, const int* __restrict__ rng, int* __restrict__ counter) {
    typedef float4 V;
    constexpr int CS = 4, BINS = P * P;
    static_assert(TR_THREADS % 128 == 0, "whole warps, evenly over the four schedulers");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_rng[TR_CHUNK][2 * P];
    __shared__ int s_k[TR_CHUNK];
    __shared__ int s_next[2];
    const int H = a.H, W = a.W, HW = H * W, WP = a.pitch, HWp = (H * WP + 3) & ~3;
    V* tab = reinterpret_cast<V*>(smem_raw);
    const int raw_elems = (CS * HW + 3) & ~3;
    float* raw0 = reinterpret_cast<float*>(tab + HWp);
    const int tid = threadIdx.x;
    const int slabs = (a.C + CS - 1) / CS, items = a.B * slabs;
    int item = blockIdx.x;
    if (item >= items) return;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncthreads();
    auto slab_src = [&](int it, int& cs) {
        const int b = it / slabs, c0 = (it - b * slabs) * CS;
        cs = min(CS, a.C - c0);
        return a.feat + ((size_t)b * a.C + c0) * HW;
    };
    // bulk copies need a 16-byte aligned source and size
    auto tma_ok = [&](const float* src, int cs) {
        return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (((cs * HW) & 3) == 0);
    };
    auto request = [&](int it, int buf) {  // thread 0
        int cs;
        const float* src = slab_src(it, cs);
        if (!tma_ok(src, cs)) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = (uint32_t)(cs * HW) * 4u;
        mbar_expect_tx(&bar[buf], bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u)
            bulk_g2s(reinterpret_cast<char*>(raw0 + buf * raw_elems) + off, reinterpret_cast<const char*>(src) + off,
                     min(32768u, bytes - off), &bar[buf]);
    };
    int pending = 0;  // thread 0: the item after the next one
    if (tid == 0) {
        request(item, 0);
        pending = atomicAdd(counter, 1);
    }
    uint32_t uses = 0;  // bit b: parity of staging buffer b's next completion
    uint32_t tab_s = smem_u32(tab), rng_s = smem_u32(&s_rng[0][0]), k_s = smem_u32(&s_k[0]);
    // opaque copies: otherwise the window base is re-derived per bin instead of kept in a register
    asm volatile("mov.u32 %0, %0;" : "+r"(tab_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(rng_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(k_s));
    const uint32_t kstride_b = (uint32_t)a.C * BINS * 4u;  // bytes between two RoIs' output blocks (host: < 2^32)
    for (int n = 0; item < items; ++n) {
        const int buf = n & 1;
        const int b = item / slabs, c0 = (item - b * slabs) * CS;
        int cs;
        const float* src = slab_src(item, cs);
        float* raw = raw0 + buf * raw_elems;
        if (tid == 0) {
            // the other staging buffer was read by the previous item's interleave pass, which every thread left
            // before the barrier that closed that item.  The counter is read one item ahead of its use, so the
            // atomic's round trip is not on any item's critical path.
            const int nxt = pending;
            s_next[buf] = nxt;  // read after this item's closing barrier, rewritten two items later
            if (nxt < items) request(nxt, buf ^ 1);
            pending = atomicAdd(counter, 1);
        }
        unsigned char* const obase = reinterpret_cast<unsigned char*>(a.out + (size_t)c0 * BINS);
        unsigned char* const abase = reinterpret_cast<unsigned char*>(a.argmax + (size_t)c0 * BINS);
        int r_begin, r_end;
        roi_range(a, b, r_begin, r_end);
        const int n_roi = r_end - r_begin;
        auto load_ranges = [&](int first, int nc) {
            const int* g = rng + (size_t)(r_begin + first) * (2 * P);
            int* d = &s_rng[0][0];
            for (int i = tid; i < nc * 2 * P; i += TR_THREADS) d[i] = __ldg(g + i);
            for (int i = tid; i < nc; i += TR_THREADS) s_k[i] = roi_at(a, r_begin + first + i);
        };
        load_ranges(0, min(n_roi, TR_CHUNK));
        if (tma_ok(src, cs)) {
            mbar_wait(&bar[buf], (uses >> buf) & 1u);
            uses ^= 1u << buf;
        } else {
            for (int i = tid; i < cs * HW; i += TR_THREADS) raw[i] = __ldg(src + i);
            __syncthreads();
        }
        build_max_tables<V, 1, TR_THREADS>(tab, raw, cs, H, W, WP, HWp, tid);  // ends with a CTA barrier
        for (int first = 0; first < n_roi; first += TR_CHUNK) {
            const int nc = min(n_roi - first, TR_CHUNK);
            if (first > 0) {
                __syncthreads();
                load_ranges(first, nc);
                __syncthreads();
            }
            // task i = (RoI j, bin e) of the chunk, i = j * BINS + e: consecutive lanes are consecutive bins, a warp may
            // straddle two RoIs; whole warps whatever BINS is (392 = 8 x 49 threads left a 13th warp of 8 lanes and
            // 4 + 3 + 3 + 3 warps on the SM's four schedulers: the closing barrier waited for the first one's)
#pragma unroll 1
            for (int i0 = 0; i0 < nc * BINS; i0 += TR_THREADS) {
                const bool valid = i0 + tid < nc * BINS;
                const int i = valid ? i0 + tid : 0;
                const int j = i / BINS, e = i - j * BINS, ph = e / P, pw = e - ph * P;
                const uint32_t rj = rng_s + (uint32_t)j * (2 * P * 4);
                const int hr = lds_i32(rj + ph * 4), wr = lds_i32(rj + (P + pw) * 4);
                const int y0 = hr & 0xFFFF, x0 = wr & 0xFFFF, hh = (hr >> 16) - y0, ww = (wr >> 16) - x0;
                // an empty bin (hi >= lo on both axes: zero rows, or only predicated-off columns) scans nothing and
                // keeps its initial value: 0 and argmax -1; a non-empty one starts from -FLT_MAX like the reference
                V v;
                int4 idx;
                vsplat(v, (hh > 0 && ww > 0) ? -FLT_MAX : 0.f);
                vneg(idx);
                const uint32_t addr = tab_s + (uint32_t)(y0 * WP + x0) * 16u;
                const int p0 = y0 * W + x0;
                // row body by the widest window among the warp's lanes
                if (!__any_sync(0xFFFFFFFFu, ww > 3)) {
                    if (!__any_sync(0xFFFFFFFFu, ww > 2)) scan_rows_s<2, false>(v, idx, addr, p0, hh, ww, WP * 16, W);
                    else scan_rows_s<3, false>(v, idx, addr, p0, hh, ww, WP * 16, W);
                } else if (!__any_sync(0xFFFFFFFFu, ww > 4)) scan_rows_s<4, false>(v, idx, addr, p0, hh, ww, WP * 16, W);
                else scan_rows_s<4, true>(v, idx, addr, p0, hh, ww, WP * 16, W);
                if (valid) {
                    const unsigned long long o =
                        (unsigned long long)(unsigned)lds_i32(k_s + j * 4) * kstride_b + (unsigned)(e * 4);
                    float* po = reinterpret_cast<float*>(obase + o);
                    int* pa = reinterpret_cast<int*>(abase + o);
                    if (cs == CS) {
                        po[0] = v.x, po[BINS] = v.y, po[2 * BINS] = v.z, po[3 * BINS] = v.w;
                        pa[0] = idx.x, pa[BINS] = idx.y, pa[2 * BINS] = idx.z, pa[3 * BINS] = idx.w;
                    } else {
                        vstore<false>(po, BINS, v, 0xFFFFFFFFu, cs);
                        istore(pa, BINS, idx, cs);
                    }
                }
            }
        }
        __syncthreads();  // table, ranges and this item's staging buffer are free
        item = s_next[buf];
    }
}

// ---------------------------------------------------------------------------------------------
// RoIPool + global average pool in one kernel.  The HarDNet head's classifier is only
// AdaptiveAvgPool2d(1) + Flatten (models/hardnet.py:203-212), so HarNetRoIHead.forward reads nothing of the
// [K,C,P,P] tensor but its mean over the bins: this kernel writes [K,C] and the P*P*4 bytes per (RoI,
// channel) never exist (3.85 GB on the 14x14 configuration, plus the pass that would re-read them).
//
// Same tables as roi_pool_tab_kernel; the mapping is chosen for the reduction instead of for the stores:
// a group of LPR = 8 / 16 lanes owns one RoI, lane l owns column pw = l of its bin grid and walks the P rows,
// accumulating in registers; the row geometry lives in the lanes too (lane l computes row l's entry, the
// others fetch it by shuffle), so the main loop has no shared-memory geometry tables and no barrier.  The
// P column sums are combined by an xor tree.  Summation order: rows top to bottom inside a column, then the
// tree over columns -- fixed, so results are run-to-run identical; they agree with pool().mean() to fp32
// rounding (1e-6 relative to the largest bin value), not bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int PM_THREADS = 512;

__device__ __forceinline__ void vacc(float4& s, const float4& v, unsigned m) {
    s.x += __uint_as_float(__float_as_uint(v.x) & m);
    s.y += __uint_as_float(__float_as_uint(v.y) & m);
    s.z += __uint_as_float(__float_as_uint(v.z) & m);
    s.w += __uint_as_float(__float_as_uint(v.w) & m);
}
__device__ __forceinline__ void vacc(float2& s, const float2& v, unsigned m) {
    s.x += __uint_as_float(__float_as_uint(v.x) & m);
    s.y += __uint_as_float(__float_as_uint(v.y) & m);
}
__device__ __forceinline__ void vxor_add(float4& s, int d) {
    s.x += __shfl_xor_sync(0xFFFFFFFFu, s.x, d);
    s.y += __shfl_xor_sync(0xFFFFFFFFu, s.y, d);
    s.z += __shfl_xor_sync(0xFFFFFFFFu, s.z, d);
    s.w += __shfl_xor_sync(0xFFFFFFFFu, s.w, d);
}
__device__ __forceinline__ void vxor_add(float2& s, int d) {
    s.x += __shfl_xor_sync(0xFFFFFFFFu, s.x, d);
    s.y += __shfl_xor_sync(0xFFFFFFFFu, s.y, d);
}
__device__ __forceinline__ void vmean_store(float* o, const float4& s, float n, int cs) {
    o[0] = s.x / n;
    if (cs > 1) o[1] = s.y / n;
    if (cs > 2) o[2] = s.z / n;
    if (cs > 3) o[3] = s.w / n;
}
__device__ __forceinline__ void vmean_store(float* o, const float2& s, float n, int cs) {
    o[0] = s.x / n;
    if (cs > 1) o[1] = s.y / n;
}

template <int P, int CS, int LV, bool DIAG = false, int THREADS = PM_THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) roi_pool_mean_kernel(RoiArgs a) {
    typedef typename VecT<CS>::type V;
    constexpr int LPR = P <= 8 ? 8 : 16;  // lanes per RoI
    constexpr int RPW = 32 / LPR;         // RoIs per warp
    constexpr int NW = THREADS / 32;
    constexpr int NT = DIAG ? 2 : LV * LV;  // DIAG: pixels + 2 x 2 windows only (see roi_pool_tab_kernel)
    constexpr bool MID = LV == 2 && P == 7;
    constexpr int OFF_MASK = DIAG ? TAB_DIAG_MASK : TAB_OFF_MASK;
    static_assert(P <= 16, "one lane per bin column");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    V* tab = reinterpret_cast<V*>(smem_raw);
    const int H = a.H, W = a.W, HW = H * W;
    const int WP = a.pitch, HWp = (H * WP + 3) & ~3;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CS;
    const int cs = min(CS, a.C - c0);
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (r_begin + blockIdx.x * NW * RPW >= r_end) return;
    float* raw = reinterpret_cast<float*>(tab + (NT - 1) * HWp);
    stage_slab(raw, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);
    build_max_tables<V, LV, THREADS, DIAG>(tab, raw, cs, H, W, WP, HWp, tid);
    __syncthreads();

    const int sub = lane / LPR, l = lane % LPR, le = min(l, P - 1);
    const int stride = a.groups * NW * RPW;
    // bin geometry from roi_pool_entries_kernel (once per RoI, not once per channel slab): lane l carries the row
    // entry l and the column entry l of its RoI, loaded one RoI ahead
    struct Geo {
        int2 row, col;
        int k;
    };
    auto load_geo = [&](int r) {
        Geo g;
        g.row = g.col = make_int2(0, 0);
        g.k = -1;
        if (r < r_end) {
            const int2* e = a.ent + (size_t)r * (2 * P);
            g.row = __ldg(e + le);
            g.col = __ldg(e + P + le);
            g.k = roi_at(a, r);
        }
        return g;
    };
    const int rw0 = r_begin + (blockIdx.x * NW + warp) * RPW;
    Geo nxt = load_geo(rw0 + sub);
    for (int rw = rw0; rw < r_end; rw += stride) {  // warp-uniform
        const Geo q = nxt;
        nxt = load_geo(rw + stride + sub);
        const int2 row = q.row, w = q.col;
        const int wy = w.y & OFF_MASK;
        const bool lc = DIAG && (w.y & TAB_LVL_BIT) != 0;
        const bool wide_ = w.x != wy;
        const bool any_wide_ = __any_sync(0xFFFFFFFFu, wide_);
        V acc;
        vsplat(acc, 0.f);
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int src = sub * LPR + ph;
            const int hx = __shfl_sync(0xFFFFFFFFu, row.x, src), hyf = __shfl_sync(0xFFFFFFFFu, row.y, src);
            const int hy = hyf & OFF_MASK;
            // DIAG: table (2,2) when both axes are 2-windows, else the pixels, the 2..4-long axis covered by the
            // pixel pairs at its two anchors (dr / dc = one pixel along that axis)
            const bool lr = DIAG && (hyf & TAB_LVL_BIT) != 0;
            const unsigned char* tb = smem_raw + ((lr && lc) ? HWp * (int)sizeof(V) : 0);
            const int dr = (lr && !lc) ? WP * (int)sizeof(V) : 0, dc = (lc && !lr) ? (int)sizeof(V) : 0;
            const bool wide = DIAG ? (wide_ || dr != 0) : wide_, tall = hx != hy || dc != 0;
            const bool any_wide = DIAG ? __any_sync(0xFFFFFFFFu, wide) : any_wide_;
            const bool any_tall = __any_sync(0xFFFFFFFFu, tall);
            V v = *reinterpret_cast<const V*>(tb + (w.x + hx));
            if (any_wide) {
                if (wide) v = vmax(v, *reinterpret_cast<const V*>(tb + (wy + hx + dr)));
            }
            if (any_tall) {
                if (tall) v = vmax(v, *reinterpret_cast<const V*>(tb + (w.x + hy + dc)));
                if (any_wide) {
                    if (wide && tall) v = vmax(v, *reinterpret_cast<const V*>(tb + (wy + hy + dr + dc)));
                }
            }
            const bool mid_ok = !DIAG || (lr && lc);  // the four 2-windows per axis need table (2,2)
            if (MID) {
                const bool mid = ((hyf | w.y) & TAB_MID_BIT) != 0 && mid_ok;
                if (__any_sync(0xFFFFFFFFu, mid)) {
                    if (mid)
                        v = tab_mid_bin<V>(v, tb, hx, hy, (hyf & TAB_MID_BIT) != 0, w.x, wy,
                                           (w.y & TAB_MID_BIT) != 0, 2 * WP * (int)sizeof(V), 2 * (int)sizeof(V));
                }
            }
            const bool big = ((hyf | w.y) & TAB_BIG_BIT) != 0 || (MID && !mid_ok && ((hyf | w.y) & TAB_MID_BIT) != 0);
            if (__any_sync(0xFFFFFFFFu, big)) {
                if (big)
                    v = DIAG ? tab_big_bin<V>(tab, tab_decode_range_diag(hx, hy, lr, WP, sizeof(V)),
                                              tab_decode_range_diag(w.x, wy, lc, 1, sizeof(V)), WP)
                             : tab_big_bin<V>(tab, tab_decode_range(hx, hy, LV * HWp, WP, sizeof(V)),
                                              tab_decode_range(w.x, wy, HWp, 1, sizeof(V)), WP);
            }
            vacc(acc, v, (unsigned)((hyf & w.y) >> 31));  // empty bins pool to 0
        }
        if (l >= P) vsplat(acc, 0.f);  // spare lanes of the group shadowed column P-1
#pragma unroll
        for (int d = LPR / 2; d > 0; d >>= 1) vxor_add(acc, d);
        if (l == 0 && q.k >= 0) vmean_store(a.out + (size_t)q.k * a.C + c0, acc, (float)(P * P), cs);
    }
}

// Fallback for feature planes too large to stage: one thread per output, straight from global.
template <bool WITH_ARGMAX>
__global__ void roi_pool_direct_kernel(RoiArgs a) {
    size_t total = (size_t)a.K * a.C * a.PH * a.PW;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        int pw = o % a.PW, ph = (o / a.PW) % a.PH;
        int c = (o / ((size_t)a.PW * a.PH)) % a.C;
        int k = o / ((size_t)a.PW * a.PH * a.C);
        const float* r = a.rois5 + (size_t)k * 5;
        int b = (int)r[0];
        float best = 0.f;
        int besti = -1;
        if (b >= 0 && b < a.B) {
            int sw = round_half_away(r[1] * a.scale), sh = round_half_away(r[2] * a.scale);
            int ew = round_half_away(r[3] * a.scale), eh = round_half_away(r[4] * a.scale);
            int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
            float bh = (float)rh / (float)a.PH, bw = (float)rw / (float)a.PW;
            int hs = min(max((int)floorf((float)ph * bh) + sh, 0), a.H);
            int he = min(max((int)ceilf((float)(ph + 1) * bh) + sh, 0), a.H);
            int ws = min(max((int)floorf((float)pw * bw) + sw, 0), a.W);
            int we = min(max((int)ceilf((float)(pw + 1) * bw) + sw, 0), a.W);
            bool empty = (he <= hs) || (we <= ws);
            best = empty ? 0.f : -FLT_MAX;
            const float* plane = a.feat + ((size_t)b * a.C + c) * a.H * a.W;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    float v = __ldg(plane + h * a.W + w);
                    if (v > best) {
                        best = v;
                        besti = h * a.W + w;
                    }
                }
        }
        a.out[o] = best;
        if (WITH_ARGMAX) a.argmax[o] = besti;
    }
}

__global__ void roi_pool_backward_kernel(const float* __restrict__ go, const int* __restrict__ argmax,
                                         const float* __restrict__ rois5, size_t total, int B, int C, int HW,
                                         int PP, float* __restrict__ gi) {
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    int am = __ldg(argmax + o);
    if (am < 0) return;
    size_t kc = o / PP;
    int c = kc % C;
    size_t k = kc / C;
    int b = (int)__ldg(rois5 + k * 5);
    if (b < 0 || b >= B || am >= HW) return;  // a RoI of no image contributes nothing (its forward row is 0 / -1)
    atomicAdd(gi + ((size_t)b * C + c) * HW + am, __ldg(go + o));
}

// ---------------------------------------------------------------------------------------------
// RoIAlign forward (torchvision roi_align)
// ---------------------------------------------------------------------------------------------
struct AlignRoi {
    float sx, sy, bin_w, bin_h, count;
    int gw, gh, k;
};

__device__ __forceinline__ AlignRoi make_align_roi(const float* r, int k, float scale, int PH, int PW,
                                                   int sampling_ratio, int aligned) {
    AlignRoi q;
    float off = aligned ? 0.5f : 0.f;
    q.sx = r[1] * scale - off;
    q.sy = r[2] * scale - off;
    float ex = r[3] * scale - off, ey = r[4] * scale - off;
    float rw = ex - q.sx, rh = ey - q.sy;
    if (!aligned) {
        rw = fmaxf(rw, 1.f);
        rh = fmaxf(rh, 1.f);
    }
    q.bin_h = rh / (float)PH;
    q.bin_w = rw / (float)PW;
    q.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
    q.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
    q.count = (float)max(q.gh * q.gw, 1);
    q.k = k;
    return q;
}

// one output element; `plane` may point to shared or global memory
__device__ __forceinline__ float align_one(const float* plane, const AlignRoi& q, int ph, int pw, int H, int W) {
    float acc = 0.f;
    for (int iy = 0; iy < q.gh; ++iy) {
        float yy = q.sy + (float)ph * q.bin_h;
        yy = yy + ((float)iy + .5f) * q.bin_h / (float)q.gh;
        for (int ix = 0; ix < q.gw; ++ix) {
            float xx = q.sx + (float)pw * q.bin_w;
            xx = xx + ((float)ix + .5f) * q.bin_w / (float)q.gw;
            float y = yy, x = xx;
            if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
            if (y <= 0.f) y = 0.f;
            if (x <= 0.f) x = 0.f;
            int yl = (int)y, xl = (int)x, yh, xh;
            if (yl >= H - 1) {
                yh = yl = H - 1;
                y = (float)yl;
            } else {
                yh = yl + 1;
            }
            if (xl >= W - 1) {
                xh = xl = W - 1;
                x = (float)xl;
            } else {
                xh = xl + 1;
            }
            float ly = y - (float)yl, lx = x - (float)xl;
            float hy = 1.f - ly, hx = 1.f - lx;
            float t = (hy * hx) * plane[yl * W + xl];
            t = t + (hy * lx) * plane[yl * W + xh];
            t = t + (ly * hx) * plane[yh * W + xl];
            t = t + (ly * lx) * plane[yh * W + xh];
            acc = acc + t;
        }
    }
    return acc / q.count;
}

__global__ void __launch_bounds__(ROI_THREADS, 2) roi_align_staged_kernel(RoiArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ AlignRoi sq[ROI_NB];
    float* sfeat = reinterpret_cast<float*>(smem_raw);
    const int PH = a.PH, PW = a.PW, PP = PH * PW;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.CS;
    const int cs = min(a.CS, a.C - c0);
    const int HW = a.H * a.W;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int first = r_begin + blockIdx.x;
    if (first >= r_end) return;
    stage_slab(sfeat, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);
    const int per_roi = cs * PP;
    for (int rb = first; rb < r_end; rb += a.groups * ROI_NB) {
        int nb = 0;
        for (int j = 0; j < ROI_NB; ++j)
            if (rb + j * a.groups < r_end) nb = j + 1;
        __syncthreads();
        if (threadIdx.x < nb) {
            int k = roi_at(a, rb + threadIdx.x * a.groups);
            sq[threadIdx.x] = make_align_roi(a.rois5 + (size_t)k * 5, k, a.scale, PH, PW, a.sampling_ratio, a.aligned);
        }
        __syncthreads();
        const int total = nb * per_roi;
        for (int it = threadIdx.x; it < total; it += ROI_THREADS) {
            int j = it / per_roi, rem = it - j * per_roi;
            int c = rem / PP, bin = rem - c * PP;
            int ph = bin / PW, pw = bin - ph * PW;
            const AlignRoi q = sq[j];
            a.out[((size_t)q.k * a.C + c0) * PP + rem] = align_one(sfeat + c * HW, q, ph, pw, a.H, a.W);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RoIAlign forward, interleaved kernel (fixed sampling_ratio SR per bin side).
//
// Same skeleton as roi_pool_tab_kernel: the CTA owns (image, 4-channel slab, RoI group), TMA-stages the
// slab's NCHW planes, re-lays them channel-interleaved (one float4 per pixel: one LDS.128 = one
// bilinear tap for four channels), and a thread owns one output bin for good, so warp stores are
// contiguous 128-byte runs.  Per RoI and axis the P*SR sample positions are reduced once to
// {low offset, high offset, l, h} entries (16 B, double-buffered shared-memory tables); a bin then needs
// 2*SR entry loads + 4*SR*SR taps.  Samples outside [-1, limit] get zero weights, which adds exactly 0
// like the reference's `continue`.  Arithmetic order per channel is the reference's:
// ((w1*v1 + w2*v2) + w3*v3) + w4*v4, samples accumulated iy-outer / ix-inner, divided by the count.
// ---------------------------------------------------------------------------------------------
struct AlignEntry {
    int lohi;  // low pixel index | high pixel index << 16 along this axis (in pixels); bit 31: out of range
    float l;   // weight of the high pixel (the low one gets 1 - l); 0 when out of range
};

__device__ __forceinline__ AlignEntry align_entry(int p, int i, int P, int SR, float c1, float c2, float scale,
                                                  int aligned, int limit, int unit) {
    const float off = aligned ? 0.5f : 0.f;
    const float start = c1 * scale - off;
    float size = (c2 * scale - off) - start;
    if (!aligned) size = fmaxf(size, 1.f);
    const float bin = size / (float)P;
    float c = start + (float)p * bin;
    c = c + ((float)i + .5f) * bin / (float)SR;
    AlignEntry e;
    if (c < -1.0f || c > (float)limit) {
        e.lohi = (int)0x80000000;
        e.l = 0.f;
        return e;
    }
    if (c <= 0.f) c = 0.f;
    int lo = (int)c, hi;
    if (lo >= limit - 1) {
        hi = lo = limit - 1;
        c = (float)lo;
    } else {
        hi = lo + 1;
    }
    e.lohi = (lo * unit) | ((hi * unit) << 16);
    e.l = c - (float)lo;
    return e;
}

// Sample geometry of every RoI row, once for all channel slabs (like roi_pool_entries_kernel): ent[r][0..P*SR) =
// row samples, [P*SR..2*P*SR) = column samples.  Computing it per CTA was 18 % of the gather's instructions.
__global__ void __launch_bounds__(256) roi_align_entries_kernel(RoiArgs a, AlignEntry* __restrict__ ent, int P, int SR) {
    const int EPR = 2 * P * SR;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.K * EPR) return;
    const int r = idx / EPR, e = idx - r * EPR;
    const float* rp = a.rois5 + (size_t)roi_at(a, r) * 5;
    const bool is_row = e < P * SR;
    const int pe = is_row ? e : e - P * SR;
    ent[idx] = is_row ? align_entry(pe / SR, pe % SR, P, SR, __ldg(rp + 2), __ldg(rp + 4), a.scale, a.aligned, a.H, a.W)
                      : align_entry(pe / SR, pe % SR, P, SR, __ldg(rp + 1), __ldg(rp + 3), a.scale, a.aligned, a.W, 1);
}

template <int P, int SR, int AL_THREADS, int MINB>
__global__ void __launch_bounds__(AL_THREADS, MINB) roi_align_tab_kernel(RoiArgs a) {
    constexpr int BINS = P * P;
    constexpr int RPI = AL_THREADS / BINS;                  // RoIs per iteration
    constexpr int EPR = 2 * P * SR;                         // table entries per RoI (rows then columns)
    constexpr int EPT = 3;                                  // table entries per thread and batch
    constexpr int NB = (EPT * AL_THREADS / EPR) / RPI * RPI;  // RoIs per batch
    constexpr int ITERS = NB / RPI;
    static_assert(RPI * BINS == AL_THREADS && NB > 0 && NB <= AL_THREADS, "thread mapping");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    // geometry in three buffers handed over through one mbarrier each (see roi_pool_tab_kernel): a warp waits
    // for batch b's entries, computes batch b+1's, then gathers batch b; warps may drift one batch apart
    constexpr int NWARPS = (AL_THREADS + 31) / 32;
    __shared__ __align__(16) AlignEntry s_ent[3][NB][EPR];
    __shared__ size_t s_ob[3][NB];
    __shared__ __align__(8) uint64_t s_full[3];
    float4* tab = reinterpret_cast<float4*>(smem_raw);
    const int H = a.H, W = a.W, HW = H * W, HWp = (HW + 3) & ~3;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 4;
    const int cs = min(4, a.C - c0);
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int stride = a.groups * NB;
    int r0 = r_begin + blockIdx.x * NB;
    if (r0 >= r_end) return;
    const int tid = threadIdx.x;
    // The entries of a batch are EPR * nb consecutive 8-byte words of a.ent (roi_align_entries_kernel): a thread
    // carries its (at most EPT) words of the NEXT batch in registers, loaded one batch ahead, and thread j < NB
    // the id of RoI j for the output offset.
    const int2* gent = reinterpret_cast<const int2*>(a.ent);
    int2 pre[EPT];
    int prek;
    auto load_entries = [&](int rr) {  // batch starting at RoI row rr
        const int lim = rr < r_end ? min(NB, r_end - rr) * EPR : 0;
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int idx = tid + q * AL_THREADS;
            pre[q] = idx < lim ? __ldg(gent + (size_t)rr * EPR + idx) : make_int2((int)0x80000000, 0);
        }
        prek = (tid < NB && rr + tid < r_end) ? roi_at(a, rr + tid) : 0;
    };
    load_entries(r0);

    float* raw = reinterpret_cast<float*>(tab + HWp);  // [cs][HW] planes, staged next to the table
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], NWARPS);  // visible after the barriers inside stage_slab
    }
    stage_slab(raw, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);
    for (int p = tid; p < HW; p += AL_THREADS) {
        float4 v;
        v.x = raw[p];
        v.y = cs > 1 ? raw[HW + p] : 0.f;
        v.z = cs > 2 ? raw[2 * HW + p] : 0.f;
        v.w = cs > 3 ? raw[3 * HW + p] : 0.f;
        tab[p] = v;
    }
    __syncthreads();

    const int e = tid % BINS, ej = tid / BINS;
    const int ph = e / P, pw = e % P;
    auto fill_tables = [&](int buf) {  // geometry buffer `buf` from the prefetched registers
        int2* dst = reinterpret_cast<int2*>(&s_ent[buf][0][0]);
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int idx = tid + q * AL_THREADS;
            if (idx < NB * EPR) dst[idx] = pre[q];
        }
        if (tid < NB) s_ob[buf][tid] = (((size_t)prek * a.C + c0) * BINS) * sizeof(float);
    };
    fill_tables(0);
    load_entries(r0 + stride);
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&s_full[0]);
    uint32_t batch = 0;
    for (; r0 < r_end; r0 += stride, ++batch) {
        const int cur = (int)(batch % 3u), nbuf = (int)((batch + 1u) % 3u);
        // every warp has written batch `batch` (hence finished gathering batch - 2): buffer batch + 1 may be
        // overwritten
        mbar_wait(&s_full[cur], (batch / 3u) & 1u);
        fill_tables(nbuf);  // next batch: its entries were loaded one batch ago
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&s_full[nbuf]);
        load_entries(r0 + 2 * stride);
        const int nb = min(NB, r_end - r0);
        for (int it = 0; it < ITERS; ++it) {
            const int j = it * RPI + ej;
            if (it * RPI >= nb) break;
            const bool valid = j < nb;
            const int jc = valid ? j : 0;
            const int2* rows = reinterpret_cast<const int2*>(&s_ent[cur][jc][ph * SR]);
            const int2* cols = reinterpret_cast<const int2*>(&s_ent[cur][jc][P * SR + pw * SR]);
            int2 cx[SR];
#pragma unroll
            for (int ix = 0; ix < SR; ++ix) cx[ix] = cols[ix];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 v1, v2, v3, v4;  // the pixel quad of the current sample: (ylo,xlo) (ylo,xhi) (yhi,xlo) (yhi,xhi)
            auto interp = [&](const int2& ry, const int2& rx) -> float4 {
                const float ly = __int_as_float(ry.y), lx = __int_as_float(rx.y);
                const float hy = ry.x < 0 ? 0.f : 1.f - ly, hx = rx.x < 0 ? 0.f : 1.f - lx;
                const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                float4 t;
                t.x = w1 * v1.x; t.x = t.x + w2 * v2.x; t.x = t.x + w3 * v3.x; t.x = t.x + w4 * v4.x;
                t.y = w1 * v1.y; t.y = t.y + w2 * v2.y; t.y = t.y + w3 * v3.y; t.y = t.y + w4 * v4.y;
                t.z = w1 * v1.z; t.z = t.z + w2 * v2.z; t.z = t.z + w3 * v3.z; t.z = t.z + w4 * v4.z;
                t.w = w1 * v1.w; t.w = t.w + w2 * v2.w; t.w = t.w + w3 * v3.w; t.w = t.w + w4 * v4.w;
                return t;
            };
            auto add = [&](const float4& t) {
                acc.x = acc.x + t.x;
                acc.y = acc.y + t.y;
                acc.z = acc.z + t.z;
                acc.w = acc.w + t.w;
            };
            if constexpr (SR == 2) {
                // The four samples of a bin are visited (0,0) (0,1) (1,1) (1,0): every move changes one axis, and
                // the next sample's pixel cell along that axis is usually the same or the neighbouring one, so
                // half of the quad (or all of it) is already in registers.  ~8 loads per bin instead of ~12 with
                // whole-quad reuse only (16 without any).  The values and the order of the additions are the
                // reference's -- (0,0) + (0,1) + (1,0) + (1,1) -- so the result is unchanged bit for bit.
                const int2 ry0 = rows[0], ry1 = rows[1];
                const int y0l = ry0.x & 0xFFFF, y0h = (ry0.x >> 16) & 0x7FFF, y1l = ry1.x & 0xFFFF, y1h = (ry1.x >> 16) & 0x7FFF;
                const int x0l = cx[0].x & 0xFFFF, x0h = (cx[0].x >> 16) & 0x7FFF;
                const int x1l = cx[1].x & 0xFFFF, x1h = (cx[1].x >> 16) & 0x7FFF;
                v1 = tab[y0l + x0l];
                v2 = tab[y0l + x0h];
                v3 = tab[y0h + x0l];
                v4 = tab[y0h + x0h];
                add(interp(ry0, cx[0]));
                if (!(x1l == x0l && x1h == x0h)) {  // columns x0 -> x1 on rows y0
                    if (x1l == x0h) {
                        v1 = v2;
                        v3 = v4;
                    } else {
                        v1 = tab[y0l + x1l];
                        v3 = tab[y0h + x1l];
                    }
                    v2 = tab[y0l + x1h];
                    v4 = tab[y0h + x1h];
                }
                add(interp(ry0, cx[1]));
                if (!(y1l == y0l && y1h == y0h)) {  // rows y0 -> y1 on columns x1
                    if (y1l == y0h) {
                        v1 = v3;
                        v2 = v4;
                    } else {
                        v1 = tab[y1l + x1l];
                        v2 = tab[y1l + x1h];
                    }
                    v3 = tab[y1h + x1l];
                    v4 = tab[y1h + x1h];
                }
                const float4 t11 = interp(ry1, cx[1]);
                if (!(x1l == x0l && x1h == x0h)) {  // columns x1 -> x0 on rows y1
                    if (x0h == x1l) {
                        v2 = v1;
                        v4 = v3;
                    } else {
                        v2 = tab[y1l + x0h];
                        v4 = tab[y1h + x0h];
                    }
                    v1 = tab[y1l + x0l];
                    v3 = tab[y1h + x0l];
                }
                add(interp(ry1, cx[0]));
                add(t11);
            } else {
                // the four taps of the previous sample stay in registers; a sample reloads them only if its
                // pixel quad differs (per lane: idle lanes cost no shared-memory wavefronts)
                int qlo = -1, qhi = -1;  // pixel indices (ylo+xlo, yhi+xhi) of the cached quad
#pragma unroll
                for (int iy = 0; iy < SR; ++iy) {
                    const int2 ry = rows[iy];
#pragma unroll
                    for (int ix = 0; ix < SR; ++ix) {
                        const int2 rx = cx[ix];
                        const int ylo = ry.x & 0xFFFF, yhi = (ry.x >> 16) & 0x7FFF;
                        const int xlo = rx.x & 0xFFFF, xhi = (rx.x >> 16) & 0x7FFF;
                        if (ylo + xlo != qlo || yhi + xhi != qhi) {
                            v1 = tab[ylo + xlo];
                            v2 = tab[ylo + xhi];
                            v3 = tab[yhi + xlo];
                            v4 = tab[yhi + xhi];
                            qlo = ylo + xlo;
                            qhi = yhi + xhi;
                        }
                        add(interp(ry, rx));
                    }
                }
            }
            if (valid) {
                const float cnt = (float)(SR * SR);
                float* o = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.out) + s_ob[cur][j]) + e;
                o[0] = acc.x / cnt;
                if (cs > 1) o[BINS] = acc.y / cnt;
                if (cs > 2) o[2 * BINS] = acc.z / cnt;
                if (cs > 3) o[3 * BINS] = acc.w / cnt;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RoIAlign forward, FAST variant (the default; north_star asks 1e-5 relative for RoIAlign, not bit-exactness).
//
// Same skeleton and thread mapping as roi_align_tab_kernel -- TMA-staged slab, channel-interleaved float4 table,
// one thread per output bin for good, geometry records handed over through three mbarrier-guarded buffers -- but
// the arithmetic is reorganised, which the bit-exact variant cannot do:
//   * a bin's 2 x 2 samples x 4 bilinear taps are SEPARABLE: out = sum_y Wy[y] * (sum_x Wx[x] * f[y][x]) / 4 with
//     per-axis weights.  The two samples of an axis are bin/2 apart, so their tap pairs coincide (RoI narrower
//     than ~7 px), overlap in one pixel, or are disjoint: the per-axis footprint is 2, 3 or 4 pixels with MERGED
//     weights, computed once per RoI by roi_align_fast_entries_kernel.  A bin reads ny x nx in {4..16} pixels
//     (one LDS.128 = 4 channels) instead of 16 taps, and needs no per-sample weight products;
//   * FMA, two channels per instruction (fma.rn.f32x2, SASS FFMA2): ny*nx*2 + ny*2 FFMA2 per bin and 4 channels
//     instead of ~150 unfused multiplies / adds;
//   * dead slots (weight 0: merged away, out of range, map border) are skipped per lane (no shared-memory
//     wavefronts) and warp-uniformly when no lane needs them (no issue slots);
//   * odd row pitch, so vertically adjacent taps do not alias banks.
// Result: within 1e-5 of the largest tap magnitude of the reference's value (fp32 rounding of a different
// association); not bit-identical.  frcnn_roi_align_forward(exact = 1) selects the reference-order kernel.
// ---------------------------------------------------------------------------------------------
struct FastAxis {     // one bin of one axis: up to four pixels = two pairs (o0, o0 + unit), (o1, o1 + unit)
    float w[4];       // merged weights of the four slots (rows: already divided by the sample count); 0 = dead slot
    uint32_t o0, o1;  // BYTE offsets into the float4 table along this axis (rows: y * pitch * 16, columns: x * 16)
};

// record of RoI row r: float4 w[2P] (rows, then columns), then uint2 o[2P]
__global__ void __launch_bounds__(256) roi_align_fast_entries_kernel(RoiArgs a, unsigned char* __restrict__ rec, int P) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.K * 2 * P) return;
    const int r = idx / (2 * P), e = idx - r * 2 * P;
    const float* rp = a.rois5 + (size_t)roi_at(a, r) * 5;
    const bool is_row = e < P;
    const int p = is_row ? e : e - P;
    const int limit = is_row ? a.H : a.W;
    const float c1 = __ldg(rp + (is_row ? 2 : 1)), c2 = __ldg(rp + (is_row ? 4 : 3));
    const AlignEntry s0 = align_entry(p, 0, P, 2, c1, c2, a.scale, a.aligned, limit, 1);
    const AlignEntry s1 = align_entry(p, 1, P, 2, c1, c2, a.scale, a.aligned, limit, 1);
    const bool v0 = s0.lohi >= 0, v1 = s1.lohi >= 0;
    int x0 = s0.lohi & 0xFFFF, x1 = s1.lohi & 0xFFFF;
    float h0 = v0 ? 1.f - s0.l : 0.f, l0 = v0 ? s0.l : 0.f;
    float h1 = v1 ? 1.f - s1.l : 0.f, l1 = v1 ? s1.l : 0.f;
    bool second = v1;
    if (!v0) {  // only the second sample is in range: it takes the first pair
        x0 = x1;
        h0 = h1;
        l0 = l1;
        second = false;
    }
    float w0 = h0, w1 = l0, w2 = 0.f, w3 = 0.f;
    int xb = 0;
    if (second) {
        if (x1 == x0) {          // same pixel pair
            w0 = w0 + h1;
            w1 = w1 + l1;
        } else if (x1 == x0 + 1) {  // pairs overlap in one pixel
            w1 = w1 + h1;
            w2 = l1;
            xb = min(x1 + 1, limit - 1);
        } else {                 // disjoint pairs
            w2 = h1;
            w3 = l1;
            xb = x1;
        }
    }
    if (!(v0 || v1)) x0 = 0;
    const float norm = is_row ? 0.25f : 1.f;  // 1 / (2 x 2 samples), exact
    const uint32_t unit = is_row ? (uint32_t)a.pitch * 16u : 16u;
    unsigned char* base = rec + (size_t)r * (2 * P * 24);
    reinterpret_cast<float4*>(base)[e] = make_float4(w0 * norm, w1 * norm, w2 * norm, w3 * norm);
    reinterpret_cast<uint2*>(base + 2 * P * 16)[e] = make_uint2((uint32_t)x0 * unit, (uint32_t)xb * unit);
}

// d += a * b for two channels at once (SASS FFMA2; b is one weight for both)
__device__ __forceinline__ void fma2(float2& d, const float2& a, float b) {
    const float2 bb = make_float2(b, b);
    uint64_t& dd = reinterpret_cast<uint64_t&>(d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(dd)
        : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(bb)));
}

template <int P, int AL_THREADS, int MINB>
__global__ void __launch_bounds__(AL_THREADS, MINB) roi_align_fast_kernel(RoiArgs a) {
    constexpr int BINS = P * P;
    constexpr int RPI = AL_THREADS / BINS;   // RoIs per iteration
    constexpr int NB = P <= 7 ? 16 : 8;      // RoIs per batch
    constexpr int ITERS = NB / RPI;
    constexpr int REC = 2 * P * 24;          // bytes per RoI record
    constexpr int WORDS = NB * REC / 16;     // 16-byte words per batch: one per thread
    static_assert(RPI * BINS == AL_THREADS && ITERS * RPI == NB && REC % 16 == 0 && WORDS <= AL_THREADS, "thread mapping");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    constexpr int NWARPS = (AL_THREADS + 31) / 32;
    __shared__ __align__(16) uint4 s_rec[3][WORDS];
    __shared__ size_t s_ob[3][NB];
    __shared__ __align__(8) uint64_t s_full[3];
    float4* tab = reinterpret_cast<float4*>(smem_raw);
    const int H = a.H, W = a.W, HW = H * W, WP = a.pitch, HWp = (H * WP + 3) & ~3;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 4;
    const int cs = min(4, a.C - c0);
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int stride = a.groups * NB;
    int r0 = r_begin + blockIdx.x * NB;
    if (r0 >= r_end) return;
    const int tid = threadIdx.x;
    const uint4* grec = reinterpret_cast<const uint4*>(a.ent);
    uint4 pre;
    int prek;
    auto load_entries = [&](int rr) {  // records of the batch starting at RoI row rr are contiguous
        const int lim = rr < r_end ? min(NB, r_end - rr) * (REC / 16) : 0;
        pre = tid < lim ? __ldg(grec + (size_t)rr * (REC / 16) + tid) : make_uint4(0u, 0u, 0u, 0u);
        prek = (tid < NB && rr + tid < r_end) ? roi_at(a, rr + tid) : 0;
    };
    load_entries(r0);
    float* raw = reinterpret_cast<float*>(tab + HWp);
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], NWARPS);
    }
    stage_slab(raw, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);
    {
        const int step_y = AL_THREADS / W, step_x = AL_THREADS - step_y * W;
        int y = tid / W, x = tid - y * W;
        for (int p = tid; p < HW; p += AL_THREADS) {
            float4 v;
            v.x = raw[p];
            v.y = cs > 1 ? raw[HW + p] : 0.f;
            v.z = cs > 2 ? raw[2 * HW + p] : 0.f;
            v.w = cs > 3 ? raw[3 * HW + p] : 0.f;
            tab[y * WP + x] = v;
            x += step_x;
            y += step_y;
            if (x >= W) {
                x -= W;
                ++y;
            }
        }
    }
    __syncthreads();
    const int e = tid % BINS, ej = tid / BINS;
    const int ph = e / P, pw = e % P;
    const uint32_t row_step = (uint32_t)WP * 16u;
    auto fill = [&](int buf) {
        if (tid < WORDS) s_rec[buf][tid] = pre;
        if (tid < NB) s_ob[buf][tid] = (((size_t)prek * a.C + c0) * BINS) * sizeof(float);
    };
    fill(0);
    load_entries(r0 + stride);
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&s_full[0]);
    uint32_t batch = 0;
    for (; r0 < r_end; r0 += stride, ++batch) {
        const int cur = (int)(batch % 3u), nbuf = (int)((batch + 1u) % 3u);
        mbar_wait(&s_full[cur], (batch / 3u) & 1u);
        fill(nbuf);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&s_full[nbuf]);
        load_entries(r0 + 2 * stride);
        const int nb = min(NB, r_end - r0);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int j = it * RPI + ej;
            if (it * RPI >= nb) break;
            const bool valid = j < nb;
            const unsigned char* rec = reinterpret_cast<const unsigned char*>(&s_rec[cur][0]) + (valid ? j : 0) * REC;
            const float4 rw = reinterpret_cast<const float4*>(rec)[ph];
            const float4 cw = reinterpret_cast<const float4*>(rec)[P + pw];
            const uint2 ro = reinterpret_cast<const uint2*>(rec + 2 * P * 16)[ph];
            const uint2 co = reinterpret_cast<const uint2*>(rec + 2 * P * 16)[P + pw];
            float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
            const bool c1 = cw.y != 0.f, c2 = cw.z != 0.f, c3 = cw.w != 0.f;
            const bool any_c2 = __any_sync(0xFFFFFFFFu, c2), any_c3 = __any_sync(0xFFFFFFFFu, c3);
            auto row = [&](uint32_t rbase, float wy) {
                const unsigned char* pa = smem_raw + rbase + co.x;
                const unsigned char* pb = smem_raw + rbase + co.y;
                float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
                if (cw.x != 0.f) {
                    const float4 v = *reinterpret_cast<const float4*>(pa);
                    fma2(s01, make_float2(v.x, v.y), cw.x);
                    fma2(s23, make_float2(v.z, v.w), cw.x);
                }
                if (c1) {
                    const float4 v = *reinterpret_cast<const float4*>(pa + 16);
                    fma2(s01, make_float2(v.x, v.y), cw.y);
                    fma2(s23, make_float2(v.z, v.w), cw.y);
                }
                if (any_c2) {
                    if (c2) {
                        const float4 v = *reinterpret_cast<const float4*>(pb);
                        fma2(s01, make_float2(v.x, v.y), cw.z);
                        fma2(s23, make_float2(v.z, v.w), cw.z);
                    }
                    if (any_c3) {
                        if (c3) {
                            const float4 v = *reinterpret_cast<const float4*>(pb + 16);
                            fma2(s01, make_float2(v.x, v.y), cw.w);
                            fma2(s23, make_float2(v.z, v.w), cw.w);
                        }
                    }
                }
                fma2(acc01, s01, wy);
                fma2(acc23, s23, wy);
            };
            if (rw.x != 0.f) row(ro.x, rw.x);
            if (__any_sync(0xFFFFFFFFu, rw.y != 0.f)) {
                if (rw.y != 0.f) row(ro.x + row_step, rw.y);
            }
            if (__any_sync(0xFFFFFFFFu, rw.z != 0.f)) {
                if (rw.z != 0.f) row(ro.y, rw.z);
                if (__any_sync(0xFFFFFFFFu, rw.w != 0.f)) {
                    if (rw.w != 0.f) row(ro.y + row_step, rw.w);
                }
            }
            if (valid) {
                float* o = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.out) + s_ob[cur][j]) + e;
                o[0] = acc01.x;
                if (cs > 1) o[BINS] = acc01.y;
                if (cs > 2) o[2 * BINS] = acc23.x;
                if (cs > 3) o[3 * BINS] = acc23.y;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RoIAlign forward, STREAMING variant (7x7 bins, sampling_ratio 2; the default for that shape).
//
// roi_align_fast_kernel is bound by the shared-memory data pipe (ncu: l1tex data-pipe wavefronts 93 % of peak):
// a thread per bin re-reads every pixel row for each of the ~2.6 bin rows it feeds, and every bin fetches its
// own geometry.  Here the bilinear sums are reorganised so that a pixel row is read ONCE per bin column:
//   out[ph][pw] = sum_y Wy[y][ph] * T[y][pw],   T[y][pw] = sum_x Wx[pw][x] * f[y][x]
// A thread owns one bin COLUMN pw of one RoI (8 lanes = one RoI: 7 columns + 1 helper lane; a warp = 4 RoIs),
// walks the RoI's pixel rows top to bottom, computes T once per row (<= 4 LDS.128 + FFMA2) and adds it to its
// seven row accumulators with that row's dense weight vector Wy[y][0..6] (zeros where the row does not reach a
// bin): straight-line FFMA2, no per-bin geometry, no data-dependent register index.  The per-RoI "row program"
// {row offset, 7 weights} comes from roi_align_stream_entries_kernel (once per RoI for all channel slabs) and
// reaches the warp through a shared-memory ring (roi_align_stream_pack_kernel / roi_align_stream2_kernel below).
// Stores: the 7 lanes drop their 7 x 4 results into a per-RoI staging block in shared memory, and the helper
// lane hands the finished [4][7][7] block (784 contiguous,
// 16-byte aligned bytes of the output tensor) to the TMA engine: cp.async.bulk.global.shared (SASS UBLKCP).
// The LSU data pipe sees 28 conflict-free STS.32 per RoI instead of 196 scattered STG.32.
// Loops are warp-uniform (trip counts are maxima over the four RoIs of a warp, dead iterations predicated), so
// the four RoIs of a warp never serialise.  Same numerical contract as roi_align_fast_kernel (1e-5 of the
// largest tap magnitude; fixed evaluation order, run-to-run identical).
// ---------------------------------------------------------------------------------------------
constexpr int AS_P = 7;
constexpr int AS_REC = 1152;     // bytes per RoI record
constexpr int AS_ROWS_OFF = 192; // row program starts here: 32 bytes per row {u32 row byte offset, float w[7]}
constexpr int AS_MAX_ROWS = 30;
constexpr int AS_STAGE = 800;    // bytes between staging blocks (784 used; 800: the four RoIs of a warp on distinct banks)

// One WARP per RoI: lanes 0..6 compute the merged per-axis taps of the seven bin rows, lanes 8..14 those of the
// seven bin columns (their taps are dealt into the record's four tap slots by warp 0, see "Tap slots" below); the row
// lanes publish theirs through shared memory, and then the warp's lanes are candidate pixel rows y_min + lane (32 at a time): a
// lane whose row reaches a bin writes that row's {offset, 7 weights} at the position a ballot gives it.
__global__ void __launch_bounds__(128) roi_align_stream_entries_kernel(RoiArgs a, unsigned char* __restrict__ rec) {
    constexpr int P = AS_P;
    __shared__ int s_py[4][P][4];
    __shared__ float s_pw[4][P][4];
    __shared__ __align__(16) int s_cx[4][P][4];
    __shared__ __align__(16) float s_cw[4][P][4];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool live = blockIdx.x * 4 + wid < a.K;  // a warp past the end recomputes the last RoI and writes nothing
    const int r = live ? blockIdx.x * 4 + wid : a.K - 1;
    const float* rp = a.rois5 + (size_t)roi_at(a, r) * 5;
    unsigned char* base = rec + (size_t)r * AS_REC;
    const bool is_row = lane < P, is_col = lane >= 8 && lane < 8 + P;
    int y_lo = 0x7FFFFFFF, y_hi = -1;
    if (is_row || is_col) {
        const int p = is_row ? lane : lane - 8;
        const int limit = is_row ? a.H : a.W;
        const float c1 = __ldg(rp + (is_row ? 2 : 1)), c2 = __ldg(rp + (is_row ? 4 : 3));
        const AlignEntry s0 = align_entry(p, 0, P, 2, c1, c2, a.scale, a.aligned, limit, 1);
        const AlignEntry s1 = align_entry(p, 1, P, 2, c1, c2, a.scale, a.aligned, limit, 1);
        const bool v0 = s0.lohi >= 0, v1 = s1.lohi >= 0;
        int x0 = s0.lohi & 0xFFFF, x1 = s1.lohi & 0xFFFF;
        float h0 = v0 ? 1.f - s0.l : 0.f, l0 = v0 ? s0.l : 0.f;
        float h1 = v1 ? 1.f - s1.l : 0.f, l1 = v1 ? s1.l : 0.f;
        bool second = v1;
        if (!v0) {
            x0 = x1;
            h0 = h1;
            l0 = l1;
            second = false;
        }
        float w0 = h0, w1 = l0, w2 = 0.f, w3 = 0.f;
        int xb = 0;
        if (second) {
            if (x1 == x0) {
                w0 = w0 + h1;
                w1 = w1 + l1;
            } else if (x1 == x0 + 1) {
                w1 = w1 + h1;
                w2 = l1;
                xb = min(x1 + 1, limit - 1);
            } else {
                w2 = h1;
                w3 = l1;
                xb = x1;
            }
        }
        if (!(v0 || v1)) x0 = 0;
        if (is_row) {  // pixel rows x0, x0+1, xb, xb+1 with weights / 4
            const int py[4] = {x0, x0 + 1, xb, xb + 1};
            const float pw[4] = {w0 * 0.25f, w1 * 0.25f, w2 * 0.25f, w3 * 0.25f};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                s_py[wid][p][q] = py[q];
                s_pw[wid][p][q] = pw[q];
                if (pw[q] != 0.f) {
                    y_lo = min(y_lo, py[q]);
                    y_hi = max(y_hi, py[q]);
                }
            }
        } else {
            s_cx[wid][p][0] = x0;
            s_cx[wid][p][1] = x0 + 1;
            s_cx[wid][p][2] = xb;
            s_cx[wid][p][3] = xb + 1;
            s_cw[wid][p][0] = w0;
            s_cw[wid][p][1] = w1;
            s_cw[wid][p][2] = w2;
            s_cw[wid][p][3] = w3;
        }
    }
    // Tap slots.  The streaming kernel reads a bin column's pixels with four LDS.128 "slots"; the seven column lanes
    // of a RoI share a quarter-warp, i.e. ONE shared-memory wavefront per slot as long as their pixels sit in distinct
    // 16-byte bank groups (pixel column mod 8).  With the taps in their natural order (x0, x0+1, xb, xb+1) two lanes
    // collide whenever their pixel columns differ by 8: 4.4 wavefronts per pixel row where 3.1 would do (ncu,
    // profiles/r2_cfg4_roialign_stream2.md).  Which tap travels in which slot is free (a weighted sum), so the taps of
    // columns 0..6 are dealt into slots greedily.  Warp 0 does it for the four RoIs of the CTA at once: a quarter-warp
    // per RoI whose lanes 0..3 ARE the slots (occupant per bank group in a register, one nibble each); every tap is
    // offered to all four and goes to the cheapest (minimum of cost * 4 + slot over the four slot lanes): 0 = a slot
    // that is already open where it meets no other pixel of its bank group (the same pixel is a broadcast, not a
    // conflict), 2 = a slot nobody uses yet (one more wavefront), 3 = a bank-group collision.  The occupant is
    // remembered modulo 64 pixel columns: dealing only decides speed, never the result.
    __syncthreads();
    if (wid == 0) {
        const int g = lane >> 3, sl = lane & 7;
        const bool g_live = blockIdx.x * 4 + g < a.K;
        unsigned char* gb = rec + (size_t)(blockIdx.x * 4 + g) * AS_REC;
        unsigned occ = 0u;
        bool open = false;
#pragma unroll 1
        for (int p = 0; p < P; ++p) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_cw[g][p]);
            const int4 x4 = *reinterpret_cast<const int4*>(s_cx[g][p]);
            const float wq[4] = {w4.x, w4.y, w4.z, w4.w};
            const int xq[4] = {x4.x, x4.y, x4.z, x4.w};
            float ow = 0.f;
            unsigned ox = 0u;
            bool used = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned x = (unsigned)xq[q], sh = (x & 7u) * 4u, id = 8u | ((x >> 3) & 7u);
                const unsigned o = (occ >> sh) & 15u;
                const unsigned cost = used ? 8u : (o != 0u && o != id) ? 3u : open ? 0u : 2u;
                const unsigned key = (sl < 4 && wq[q] != 0.f) ? cost * 4u + (unsigned)sl : 0xFFu;
                unsigned best = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, 1));  // (REDUX with a quarter-warp mask
                best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, 2));         // serialises the four quarters)
                if (key == best && key != 0xFFu) {
                    used = open = true;
                    ow = wq[q];
                    ox = x;
                    if (o == 0u) occ |= id << sh;
                }
            }
            if (sl < 4 && g_live) {
                reinterpret_cast<float*>(gb)[p * 4 + sl] = ow;
                reinterpret_cast<unsigned short*>(gb + P * 16)[p * 4 + sl] = (unsigned short)(ox * 16u);
            }
        }
    }
    const int y_min = __reduce_min_sync(0xFFFFFFFFu, y_lo), y_max = __reduce_max_sync(0xFFFFFFFFu, y_hi);
    __syncwarp();
    // row program: the pixel rows that reach at least one bin, top to bottom, each with its weight per bin row
    int n = 0;
    for (int y0 = y_min; y0 <= y_max && n < AS_MAX_ROWS; y0 += 32) {
        const int y = y0 + lane;
        float w[P];
        bool any = false;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (s_py[wid][p][q] == y) t = t + s_pw[wid][p][q];  // a row appears at most twice in a bin (inverted RoIs)
            w[p] = t;
            any = any || t != 0.f;
        }
        any = any && y <= y_max;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, any);
        const int idx = n + __popc(bal & ((1u << lane) - 1u));
        if (any && idx < AS_MAX_ROWS && live) {
            uint4* dst = reinterpret_cast<uint4*>(base + AS_ROWS_OFF + idx * 32);
            dst[0] = make_uint4((uint32_t)y * (uint32_t)a.pitch * 16u, __float_as_uint(w[0]), __float_as_uint(w[1]),
                                __float_as_uint(w[2]));
            dst[1] = make_uint4(__float_as_uint(w[3]), __float_as_uint(w[4]), __float_as_uint(w[5]), __float_as_uint(w[6]));
        }
        n = min(n + __popc(bal), AS_MAX_ROWS);
    }
    if (lane == 0 && live) *reinterpret_cast<int*>(base + P * 24) = n;
}

// RoI rows of every image ordered by the length of their row program, longest first (counting sort, one CTA per
// image): the four RoIs that share a warp of roi_align_stream2_kernel then have (almost) the same trip count, so the
// warp-uniform row loop wastes no iterations on the shortest of them (unsorted: 16.3 iterations per warp for a mean of
// 10.2 rows per RoI on the 800x800 configuration).  Order inside a bucket is arbitrary; results do not depend on it.
__global__ void __launch_bounds__(256) roi_align_stream_sort_kernel(RoiArgs a, const unsigned char* __restrict__ rec,
                                                                    int* __restrict__ sorted) {
    __shared__ int cnt[AS_MAX_ROWS + 2];
    const int b = blockIdx.x;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    if (threadIdx.x < AS_MAX_ROWS + 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int r = r_begin + threadIdx.x; r < r_end; r += 256)
        atomicAdd(&cnt[AS_MAX_ROWS - *reinterpret_cast<const int*>(rec + (size_t)r * AS_REC + AS_P * 24)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = r_begin;
        for (int i = 0; i <= AS_MAX_ROWS; ++i) {
            const int c = cnt[i];
            cnt[i] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int r = r_begin + threadIdx.x; r < r_end; r += 256)
        sorted[atomicAdd(&cnt[AS_MAX_ROWS - *reinterpret_cast<const int*>(rec + (size_t)r * AS_REC + AS_P * 24)], 1)] = r;
}

// ---------------------------------------------------------------------------------------------
// The streaming kernel proper: the row programs travel through shared memory.
//
// The first form of this kernel read the row records from L2 with __ldg, three rows ahead in 24 registers; ncu
// (profiles/r2_cfg4_roialign_fast_and_stream.md) showed 44 % of a warp's time as "long scoreboard" -- with 16 warps
// per SM three rows are not far enough.  Here roi_align_stream_pack_kernel lays the geometry of a whole PASS of a warp
// (its 4 RoIs x up to 30 pixel rows) out as the warp will consume it -- 512-byte chunks: two header chunks
// (column weights; column offsets, row count, output row) and one chunk per 4 pixel rows, each row
// [4 RoIs][offset, w0..w6] -- and the warp copies chunk p + D into a private shared-memory ring with ONE cp.async
// (32 lanes x 16 bytes, SASS LDGSTS) while it works on chunk p: no registers held for records in flight, no
// L2 latency on the dependency chain, no mbarrier (cp.async.wait_group + __syncwarp; the ring is private to the
// warp).  Rows a RoI does not have point at an all-zero pixel row appended to the table, so the row body needs
// no per-RoI predicate.  The input planes are read straight from global memory into the channel-interleaved
// table (coalesced 128-byte runs per plane) -- without the 40 KB TMA landing zone three CTAs fit on an SM.
// ---------------------------------------------------------------------------------------------
constexpr int AS2_CHUNK = 512;
constexpr int AS2_HDR = 2;
constexpr int AS2_MAXC = AS2_HDR + (AS_MAX_ROWS + 3) / 4;
constexpr int AS2_SLOT = AS2_MAXC * AS2_CHUNK;  // bytes per (pass, warp)

__device__ __forceinline__ int as2_pass_base(int r_begin, int b, int NQ) { return r_begin / NQ + b; }

template <int THREADS>
__global__ void __launch_bounds__(THREADS) roi_align_stream_pack_kernel(RoiArgs a, const unsigned char* __restrict__ rec,
                                                                        const int* __restrict__ sorted,
                                                                        unsigned char* __restrict__ prog) {
    constexpr int P = AS_P, NQ = THREADS / 8, NW = THREADS / 32;
    const int b = blockIdx.y;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, ql = lane & 7;
    for (int pi = blockIdx.x; r_begin + pi * NQ < r_end; pi += gridDim.x) {
    const int pos0 = r_begin + pi * NQ;
    const int pos = pos0 + (tid >> 3);
    const bool valid = pos < r_end;
    const int r = valid ? __ldg(sorted + pos) : r_begin;
    const unsigned char* rb = rec + (size_t)r * AS_REC;
    const int n = valid ? __ldg(reinterpret_cast<const int*>(rb + P * 24)) : 0;
    float4 cw = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 co = make_uint2(0u, 0u);
    if (valid && ql < P) {
        cw = __ldg(reinterpret_cast<const float4*>(rb) + ql);
        co = __ldg(reinterpret_cast<const uint2*>(rb + P * 16) + ql);
    }
    const int nmax = __reduce_max_sync(0xFFFFFFFFu, n);
    const int k = max(1, (nmax + 3) >> 2);
    unsigned char* slot = prog + ((size_t)(as2_pass_base(r_begin, b, NQ) + pi) * NW + w) * AS2_SLOT;
    if (ql == 7) cw = make_float4(__int_as_float(k), __int_as_float(nmax), 0.f, 0.f);  // helper lanes carry the pass shape
    reinterpret_cast<float4*>(slot)[lane] = cw;
    reinterpret_cast<uint4*>(slot + AS2_CHUNK)[lane] =
        make_uint4(co.x, co.y, (uint32_t)n, valid ? (uint32_t)roi_at(a, r) : 0xFFFFFFFFu);
    const uint32_t zero_row = (uint32_t)a.H * (uint32_t)a.pitch * 16u;
    for (int c = 0; c < k; ++c) {
        const int i = 4 * c + (lane >> 3), gg = (lane & 7) >> 1, half = lane & 1;
        const int r_g = __shfl_sync(0xFFFFFFFFu, r, gg * 8);
        const int n_g = __shfl_sync(0xFFFFFFFFu, n, gg * 8);
        uint4 v = half ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(zero_row, 0u, 0u, 0u);
        if (i < n_g) v = __ldg(reinterpret_cast<const uint4*>(rec + (size_t)r_g * AS_REC + AS_ROWS_OFF) + 2 * i + half);
        reinterpret_cast<uint4*>(slot + (AS2_HDR + c) * AS2_CHUNK)[lane] = v;
    }
    }
}

template <int THREADS, int MINB, int RING>
__global__ void __launch_bounds__(THREADS, MINB) roi_align_stream2_kernel(RoiArgs a, const unsigned char* __restrict__ prog) {
    constexpr int P = AS_P, BINS = P * P;
    constexpr int NQ = THREADS / 8, NW = THREADS / 32, D = RING - 1;
    static_assert(D >= 1 && D <= 3, "look-ahead must stay inside the next pass's header + first row chunk");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tab = reinterpret_cast<float4*>(smem_raw);
    const int H = a.H, W = a.W, HW = H * W, WP = a.pitch;
    const uint32_t tab_bytes = ((uint32_t)(H + 1) * WP * 16u + 127u) & ~127u;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 4;
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int n_pass_img = (r_end - r_begin + NQ - 1) / NQ;
    const int pi0 = blockIdx.x;
    if (pi0 >= n_pass_img) return;
    const int n_pass = (n_pass_img - pi0 + a.groups - 1) / a.groups;  // passes of this CTA: pi0, pi0 + groups, ...
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 3, ql = lane & 7;
    unsigned char* stg_b = smem_raw + tab_bytes + (size_t)(tid >> 3) * AS_STAGE;
    float* stg = reinterpret_cast<float*>(stg_b);
    const uint32_t ring = smem_u32(smem_raw + tab_bytes + (size_t)NQ * AS_STAGE + (size_t)w * (RING * AS2_CHUNK));
    // position in the program buffer in 16-byte units (32 bits: two registers less than a pointer and a stride)
    uint32_t slot = (uint32_t)(((size_t)(as2_pass_base(r_begin, b, NQ) + pi0) * NW + w) * (AS2_SLOT / 16)) + lane;
    const uint32_t slot_stride = (uint32_t)a.groups * NW * (AS2_SLOT / 16);

    // chunk `pos` of the warp's stream lives in ring slot pos % RING
    auto fetch = [&](uint32_t src16, int pos) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (uint32_t)(pos % RING) * AS2_CHUNK + lane * 16),
                     "l"(prog + (size_t)src16 * 16)
                     : "memory");
    };
    auto commit = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };
    auto wait_chunk = [&]() {
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        __syncwarp();
    };
#pragma unroll
    for (int j = 0; j < D; ++j) {  // header (+ first row chunk): on their way while the table is built
        fetch(slot + j * (AS2_CHUNK / 16), j);
        commit();
    }
    {
        const float* src = a.feat + ((size_t)b * a.C + c0) * HW;
#pragma unroll 2
        for (int p = tid; p < HW; p += THREADS) {
            const int y = p / W, x = p - y * W;
            tab[y * WP + x] = make_float4(__ldg(src + p), __ldg(src + HW + p), __ldg(src + 2 * HW + p), __ldg(src + 3 * HW + p));
        }
        for (int x = tid; x < WP; x += THREADS) tab[H * WP + x] = make_float4(0.f, 0.f, 0.f, 0.f);  // the row of no RoI
    }
    __syncthreads();

    const bool col_lane = ql < P;
    int pos = 0;  // chunks consumed so far
    bool store_pending = false;
    // after consuming chunk j of a pass with kk row chunks: start the copy of the chunk D positions further on
    auto prefetch = [&](int j, int kk, bool last_pass) {
        int jj = j + D;
        uint32_t base = slot;
        bool ok = true;
        if (jj >= AS2_HDR + kk) {
            jj -= AS2_HDR + kk;
            base += slot_stride;
            ok = !last_pass;
        }
        if (ok) fetch(base + jj * (AS2_CHUNK / 16), pos + D);
        commit();
    };
    for (int ip = 0; ip < n_pass; ++ip, slot += slot_stride) {
        const bool last_pass = ip + 1 == n_pass;
        wait_chunk();
        float4 cw;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cw.x), "=f"(cw.y), "=f"(cw.z), "=f"(cw.w)
                     : "r"(ring + (uint32_t)(pos % RING) * AS2_CHUNK + lane * 16));
        const int kk = __shfl_sync(0xFFFFFFFFu, __float_as_int(cw.x), 7);
        const int nmax = __shfl_sync(0xFFFFFFFFu, __float_as_int(cw.y), 7);
        if (!col_lane) cw = make_float4(0.f, 0.f, 0.f, 0.f);
        prefetch(0, kk, last_pass);
        ++pos;
        wait_chunk();
        uint4 hd;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hd.x), "=r"(hd.y), "=r"(hd.z), "=r"(hd.w)
                     : "r"(ring + (uint32_t)(pos % RING) * AS2_CHUNK + lane * 16));
        prefetch(1, kk, last_pass);
        ++pos;
        const bool cl0 = cw.x != 0.f, c1 = cw.y != 0.f, c2 = cw.z != 0.f, c3 = cw.w != 0.f;
        // four independent tap slots (byte offsets of their pixel columns; dealt by roi_align_stream_entries_kernel)
        const unsigned char* ca = smem_raw + (hd.x & 0xFFFFu);
        const unsigned char* ca1 = smem_raw + (hd.x >> 16);
        const unsigned char* cb = smem_raw + (hd.y & 0xFFFFu);
        const unsigned char* cb1 = smem_raw + (hd.y >> 16);
        const uint32_t orow = hd.w;
        float2 acc[P][2];
#pragma unroll
        for (int k = 0; k < P; ++k) acc[k][0] = acc[k][1] = make_float2(0.f, 0.f);
        auto process = [&](uint32_t raddr) {
            uint4 lo, hi;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(raddr));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                         : "r"(raddr + 16));
            const unsigned char* pa = ca + lo.x;
            const unsigned char* pa1 = ca1 + lo.x;
            const unsigned char* pb = cb + lo.x;
            const unsigned char* pb1 = cb1 + lo.x;
            float2 t01 = make_float2(0.f, 0.f), t23 = make_float2(0.f, 0.f);
            if (cl0) {
                const float4 v = *reinterpret_cast<const float4*>(pa);
                fma2(t01, make_float2(v.x, v.y), cw.x);
                fma2(t23, make_float2(v.z, v.w), cw.x);
            }
            if (c1) {
                const float4 v = *reinterpret_cast<const float4*>(pa1);
                fma2(t01, make_float2(v.x, v.y), cw.y);
                fma2(t23, make_float2(v.z, v.w), cw.y);
            }
            if (c2) {
                const float4 v = *reinterpret_cast<const float4*>(pb);
                fma2(t01, make_float2(v.x, v.y), cw.z);
                fma2(t23, make_float2(v.z, v.w), cw.z);
            }
            if (c3) {
                const float4 v = *reinterpret_cast<const float4*>(pb1);
                fma2(t01, make_float2(v.x, v.y), cw.w);
                fma2(t23, make_float2(v.z, v.w), cw.w);
            }
            const float wk[P] = {__uint_as_float(lo.y), __uint_as_float(lo.z), __uint_as_float(lo.w), __uint_as_float(hi.x),
                                 __uint_as_float(hi.y), __uint_as_float(hi.z), __uint_as_float(hi.w)};
#pragma unroll
            for (int k = 0; k < P; ++k) {
                fma2(acc[k][0], t01, wk[k]);
                fma2(acc[k][1], t23, wk[k]);
            }
        };
        for (int c = 0; c < kk; ++c) {
            wait_chunk();
            const uint32_t cbase = ring + (uint32_t)(pos % RING) * AS2_CHUNK + g * 32;
            const int left = nmax - 4 * c;  // rows of this chunk some RoI of the warp really has (warp-uniform)
            if (left >= 4) {
                process(cbase);
                process(cbase + 128);
                process(cbase + 256);
                process(cbase + 384);
            } else {
                if (left > 0) process(cbase);
                if (left > 1) process(cbase + 128);
                if (left > 2) process(cbase + 256);
            }
            prefetch(AS2_HDR + c, kk, last_pass);
            ++pos;
        }
        // the staging block is free once the previous bulk store has READ it
        if (store_pending && ql == P) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        const bool valid = orow != 0xFFFFFFFFu;
        if (col_lane && valid) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
                stg[0 * BINS + k * P + ql] = acc[k][0].x;
                stg[1 * BINS + k * P + ql] = acc[k][0].y;
                stg[2 * BINS + k * P + ql] = acc[k][1].x;
                stg[3 * BINS + k * P + ql] = acc[k][1].y;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (ql == P && valid) {
            float* dst = a.out + ((size_t)orow * a.C + c0) * BINS;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stg)),
                         "n"(4 * BINS * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            store_pending = true;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (store_pending && ql == P) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // before the CTA's smem goes away
}

// ---------------------------------------------------------------------------------------------
// RoIAlign + global average pool (HarDNet head with the RoIAlign configuration).
//
// mean over bins of RoIAlign is LINEAR in the features and, with a fixed sampling grid, SEPARABLE:
//   out[k,c] = 1/(P*P*SR*SR) * sum_{sy,sx} sum_{taps} wy(sy,y) * wx(sx,x) * feat[c,y,x]
//            = 1/(P*P*SR*SR) * sum_y Wy[y] * sum_x Wx[x] * feat[c,y,x],   Wy[y] = sum_sy wy(sy,y), Wx likewise
// (a sample outside [-1, limit] has zero weight on its axis, so the reference's `continue` factorises too).
// So instead of 4*SR*SR*P*P bilinear taps per (RoI, channel) -- 784 for 7x7 / 2x2 -- the kernel reads each
// pixel of the RoI's window once, weighted: ~100 reads for a typical RoI, and writes [K,C].
// Step 1 (roi_align_weights_kernel, one warp per RoI, once for all channels): per-axis window + weights into
// a 528-byte record.  Step 2 (roi_align_mean_kernel, CTA = image x 4-channel slab like the other gather
// kernels): a warp owns a RoI, its lanes tile the window, weighted sum, warp tree reduction.
// Agreement with roi_align().mean((2,3)): fp32 rounding (different association), 1e-5 relative.
// ---------------------------------------------------------------------------------------------
constexpr int AW_MAX = 64;  // widest / tallest window a record holds (maps up to 64 pixels per side)
struct AlignWeights {
    int x0, nx, y0, ny;  // window origin and extent in pixels (nx, ny = 0: nothing in range)
    float wx[AW_MAX], wy[AW_MAX];
};

__device__ __forceinline__ void axis_weights(int s, int NS, int P, int SR, float c1, float c2, float scale, int aligned,
                                             int limit, int lane, int& origin, int& extent, float& w0, float& w1) {
    AlignEntry e;
    e.lohi = (int)0x80000000;
    e.l = 0.f;
    if (s < NS) e = align_entry(s / SR, s % SR, P, SR, c1, c2, scale, aligned, limit, 1);
    const bool ok = s < NS && e.lohi >= 0;
    const int lo = e.lohi & 0xFFFF, hi = (e.lohi >> 16) & 0x7FFF;
    const int mn = __reduce_min_sync(0xFFFFFFFFu, ok ? lo : 0x7FFFFFFF);
    const int mx = __reduce_max_sync(0xFFFFFFFFu, ok ? hi : -1);
    origin = mn;
    extent = mx >= mn ? min(mx - mn + 1, AW_MAX) : 0;
    w0 = w1 = 0.f;
    for (int t = 0; t < NS; ++t) {  // lane accumulates the weights landing on pixels origin+lane, origin+lane+32
        const int tl = __shfl_sync(0xFFFFFFFFu, e.lohi, t);
        const float l = __shfl_sync(0xFFFFFFFFu, e.l, t);
        if (tl < 0) continue;  // uniform
        const int plo = (tl & 0xFFFF) - mn, phi = ((tl >> 16) & 0x7FFF) - mn;
        const float h = 1.f - l;
        if (plo == lane) w0 += h;
        if (phi == lane) w0 += l;
        if (plo == lane + 32) w1 += h;
        if (phi == lane + 32) w1 += l;
    }
}

__global__ void __launch_bounds__(256) roi_align_weights_kernel(RoiArgs a, AlignWeights* __restrict__ rec) {
    const int k = blockIdx.x * (256 / 32) + (threadIdx.x >> 5);
    if (k >= a.K) return;
    const int lane = threadIdx.x & 31;
    const float* r = a.rois5 + (size_t)k * 5;
    const float x1 = __ldg(r + 1), y1 = __ldg(r + 2), x2 = __ldg(r + 3), y2 = __ldg(r + 4);
    const int SR = a.sampling_ratio;
    int x0, nx, y0, ny;
    float wx0, wx1, wy0, wy1;
    axis_weights(lane, a.PW * SR, a.PW, SR, x1, x2, a.scale, a.aligned, a.W, lane, x0, nx, wx0, wx1);
    axis_weights(lane, a.PH * SR, a.PH, SR, y1, y2, a.scale, a.aligned, a.H, lane, y0, ny, wy0, wy1);
    AlignWeights* o = rec + k;
    if (lane == 0) {
        o->x0 = nx ? x0 : 0;
        o->nx = ny ? nx : 0;  // either axis empty: nothing to add
        o->y0 = ny ? y0 : 0;
        o->ny = nx ? ny : 0;
    }
    o->wx[lane] = wx0;
    o->wx[lane + 32] = wx1;
    o->wy[lane] = wy0;
    o->wy[lane + 32] = wy1;
}

constexpr int AM_THREADS = 512;

__global__ void __launch_bounds__(AM_THREADS, 2) roi_align_mean_kernel(RoiArgs a, const AlignWeights* __restrict__ rec) {
    constexpr int NW = AM_THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_w[NW][2][AW_MAX];
    float4* tab = reinterpret_cast<float4*>(smem_raw);
    const int H = a.H, W = a.W, HW = H * W;
    const int WP = a.pitch, HWp = (H * WP + 3) & ~3;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 4;
    const int cs = min(4, a.C - c0);
    int r_begin, r_end;
    roi_range(a, b, r_begin, r_end);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (r_begin + blockIdx.x * NW >= r_end) return;
    float* raw = reinterpret_cast<float*>(tab + HWp);
    stage_slab(raw, a.feat + ((size_t)b * a.C + c0) * HW, cs * HW, &bar);
    {
        const int step_y = AM_THREADS / W, step_x = AM_THREADS - step_y * W;
        int y = tid / W, x = tid - y * W;
        for (int p = tid; p < HW; p += AM_THREADS) {
            float4 v;
            v.x = raw[p];
            v.y = cs > 1 ? raw[HW + p] : 0.f;
            v.z = cs > 2 ? raw[2 * HW + p] : 0.f;
            v.w = cs > 3 ? raw[3 * HW + p] : 0.f;
            tab[y * WP + x] = v;
            x += step_x;
            y += step_y;
            if (x >= W) {
                x -= W;
                ++y;
            }
        }
    }
    __syncthreads();
    const float norm = (float)(a.PH * a.PW) * (float)(a.sampling_ratio * a.sampling_ratio);
    // a RoI's record (header + 2 x 64 weights) is fetched one RoI ahead: its L2 round trip hides behind the
    // window loop of the current one
    struct Rec {
        int k;
        int4 hdr;
        float wx0, wx1, wy0, wy1;
    };
    auto fetch = [&](int r) {
        Rec q;
        q.k = -1;
        q.hdr = make_int4(0, 0, 0, 0);
        q.wx0 = q.wx1 = q.wy0 = q.wy1 = 0.f;
        if (r < r_end) {
            q.k = roi_at(a, r);
            const AlignWeights* w = rec + q.k;
            q.hdr = __ldg(reinterpret_cast<const int4*>(w));
            q.wx0 = __ldg(w->wx + lane);
            q.wx1 = __ldg(w->wx + lane + 32);
            q.wy0 = __ldg(w->wy + lane);
            q.wy1 = __ldg(w->wy + lane + 32);
        }
        return q;
    };
    const int rstep = a.groups * NW;
    Rec nxt = fetch(r_begin + blockIdx.x * NW + warp);
    for (int r = r_begin + blockIdx.x * NW + warp; r < r_end; r += rstep) {
        const Rec cur = nxt;
        nxt = fetch(r + rstep);
        const int k = cur.k;
        const int x0 = cur.hdr.x, nx = cur.hdr.y, y0 = cur.hdr.z, ny = cur.hdr.w;
        __syncwarp();  // previous RoI's weights no longer read
        s_w[warp][0][lane] = cur.wx0;
        s_w[warp][0][lane + 32] = cur.wx1;
        s_w[warp][1][lane] = cur.wy0;
        s_w[warp][1][lane + 32] = cur.wy1;
        __syncwarp();
        // lanes tile the window TY x TX with TX the smallest of 8 / 16 / 32 that covers its width
        const int sh = nx <= 8 ? 3 : (nx <= 16 ? 4 : 5);
        const int tx = 1 << sh, ty = 32 >> sh;
        const int lx = lane & (tx - 1), ly = lane >> sh;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int yb = 0; yb < ny; yb += ty) {
            const int yy = yb + ly;
            const float wy = yy < ny ? s_w[warp][1][yy] : 0.f;
            const float4* row = tab + (y0 + min(yy, ny - 1)) * WP + x0;
            for (int xb = 0; xb < nx; xb += tx) {
                const int xx = xb + lx;
                const float wgt = xx < nx ? wy * s_w[warp][0][xx] : 0.f;
                const float4 v = row[min(xx, nx - 1)];
                acc.x = __fmaf_rn(wgt, v.x, acc.x);  // explicit FMA (the library is built with -fmad=false):
                acc.y = __fmaf_rn(wgt, v.y, acc.y);  // this kernel's contract is 1e-5, not the reference's
                acc.z = __fmaf_rn(wgt, v.z, acc.z);  // operation order
                acc.w = __fmaf_rn(wgt, v.w, acc.w);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vxor_add(acc, d);
        if (lane == 0) vmean_store(a.out + (size_t)k * a.C + c0, acc, norm, cs);
    }
}

__global__ void roi_align_direct_kernel(RoiArgs a) {
    size_t total = (size_t)a.K * a.C * a.PH * a.PW;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        int pw = o % a.PW, ph = (o / a.PW) % a.PH;
        int c = (o / ((size_t)a.PW * a.PH)) % a.C;
        int k = o / ((size_t)a.PW * a.PH * a.C);
        const float* r = a.rois5 + (size_t)k * 5;
        int b = (int)r[0];
        float v = 0.f;
        if (b >= 0 && b < a.B) {
            AlignRoi q = make_align_roi(r, k, a.scale, a.PH, a.PW, a.sampling_ratio, a.aligned);
            v = align_one(a.feat + ((size_t)b * a.C + c) * a.H * a.W, q, ph, pw, a.H, a.W);
        }
        a.out[o] = v;
    }
}

__global__ void roi_align_backward_kernel(const float* __restrict__ go, RoiArgs a, float* __restrict__ gi) {
    size_t total = (size_t)a.K * a.C * a.PH * a.PW;
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    int pw = o % a.PW, ph = (o / a.PW) % a.PH;
    int c = (o / ((size_t)a.PW * a.PH)) % a.C;
    int k = o / ((size_t)a.PW * a.PH * a.C);
    const float* r = a.rois5 + (size_t)k * 5;
    int b = (int)r[0];
    if (b < 0 || b >= a.B) return;
    AlignRoi q = make_align_roi(r, k, a.scale, a.PH, a.PW, a.sampling_ratio, a.aligned);
    float g = __ldg(go + o) / q.count;
    float* plane = gi + ((size_t)b * a.C + c) * a.H * a.W;
    const int H = a.H, W = a.W;
    for (int iy = 0; iy < q.gh; ++iy) {
        float yy = q.sy + (float)ph * q.bin_h;
        yy = yy + ((float)iy + .5f) * q.bin_h / (float)q.gh;
        for (int ix = 0; ix < q.gw; ++ix) {
            float xx = q.sx + (float)pw * q.bin_w;
            xx = xx + ((float)ix + .5f) * q.bin_w / (float)q.gw;
            float y = yy, x = xx;
            if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
            if (y <= 0.f) y = 0.f;
            if (x <= 0.f) x = 0.f;
            int yl = (int)y, xl = (int)x, yh, xh;
            if (yl >= H - 1) {
                yh = yl = H - 1;
                y = (float)yl;
            } else {
                yh = yl + 1;
            }
            if (xl >= W - 1) {
                xh = xl = W - 1;
                x = (float)xl;
            } else {
                xh = xl + 1;
            }
            float ly = y - (float)yl, lx = x - (float)xl;
            float hy = 1.f - ly, hx = 1.f - lx;
            atomicAdd(plane + yl * W + xl, g * (hy * hx));
            atomicAdd(plane + yl * W + xh, g * (hy * lx));
            atomicAdd(plane + yh * W + xl, g * (ly * hx));
            atomicAdd(plane + yh * W + xh, g * (ly * lx));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RoIPool / RoIAlign backward, SLAB form (maps whose 4-channel gradient slab fits in shared memory).
//
// roi_pool_backward_kernel / roi_align_backward_kernel are torchvision's design: one thread per output element, one
// (RoIPool) or 4 * samples (RoIAlign) global atomicAdd each -- 963 M L2 atomics for a cfg2-sized RoIPool gradient.
// The forward kernels' decomposition works backwards too: a CTA owns the gradient of ONE image's 4-channel slab
// [4][H*W], keeps it in shared memory, walks the RoIs of its image (compacted from the batch-index column in chunks --
// no workspace, no bucketing pass), reads grad_out / argmax in 784-byte-contiguous runs and accumulates with
// shared-memory atomics; the slab is added to grad_in once at the end by the only CTA that ever touches those planes:
// no global atomic, 16 bytes of grad_in traffic per pixel instead of one L2 round trip per output element.
// (Summation order inside a pixel is still unordered, as with any atomic accumulation.)
// ---------------------------------------------------------------------------------------------
constexpr int BW_THREADS = 512;
constexpr int BW_LIST = 1024;  // RoIs compacted per pass

// appends to s_list the RoI rows k in [k0, k0 + BW_LIST) whose batch index is b; returns their number
__device__ __forceinline__ int bw_collect(const float* __restrict__ rois5, int K, int k0, int b, int* s_list, int* s_n) {
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    for (int k = k0 + threadIdx.x; k < min(k0 + BW_LIST, K); k += BW_THREADS)
        if ((int)__ldg(rois5 + (size_t)k * 5) == b) s_list[atomicAdd(s_n, 1)] = k;
    __syncthreads();
    return *s_n;
}

__global__ void __launch_bounds__(BW_THREADS) roi_pool_backward_slab_kernel(const float* __restrict__ go,
                                                                            const int* __restrict__ argmax,
                                                                            const float* __restrict__ rois5, int K, int C,
                                                                            int HW, int PP, float* __restrict__ gi) {
    extern __shared__ __align__(16) float s_acc[];  // [4][HW]
    __shared__ int s_list[BW_LIST];
    __shared__ int s_n;
    const int b = blockIdx.y, c0 = blockIdx.x * 4, cs = min(4, C - c0);
    for (int i = threadIdx.x; i < 4 * HW; i += BW_THREADS) s_acc[i] = 0.f;
    const int per = cs * PP;  // contiguous elements of one RoI that belong to this slab
    for (int k0 = 0; k0 < K; k0 += BW_LIST) {
        const int n = bw_collect(rois5, K, k0, b, s_list, &s_n);
        for (int idx = threadIdx.x; idx < n * per; idx += BW_THREADS) {
            const int j = idx / per, e = idx - j * per;
            const size_t o = ((size_t)s_list[j] * C + c0) * PP + e;
            const int am = __ldg(argmax + o);
            if (am >= 0 && am < HW) atomicAdd(&s_acc[(e / PP) * HW + am], __ldg(go + o));
        }
        __syncthreads();  // s_list is rewritten by the next pass
    }
    float* dst = gi + ((size_t)b * C + c0) * HW;
    for (int i = threadIdx.x; i < cs * HW; i += BW_THREADS) dst[i] += s_acc[i];
}

__global__ void __launch_bounds__(BW_THREADS) roi_align_backward_slab_kernel(const float* __restrict__ go, RoiArgs a,
                                                                             float* __restrict__ gi) {
    extern __shared__ __align__(16) float s_acc[];  // [4][HW]
    __shared__ int s_list[BW_LIST];
    __shared__ int s_n;
    const int H = a.H, W = a.W, HW = H * W, PP = a.PH * a.PW;
    const int b = blockIdx.y, c0 = blockIdx.x * 4, cs = min(4, a.C - c0);
    for (int i = threadIdx.x; i < 4 * HW; i += BW_THREADS) s_acc[i] = 0.f;
    for (int k0 = 0; k0 < a.K; k0 += BW_LIST) {
        const int n = bw_collect(a.rois5, a.K, k0, b, s_list, &s_n);
        // a thread owns one bin of one RoI for the slab's channels: the sample geometry is computed once for all of them
        for (int idx = threadIdx.x; idx < n * PP; idx += BW_THREADS) {
            const int j = idx / PP, e = idx - j * PP;
            const int ph = e / a.PW, pw = e - ph * a.PW;
            const int k = s_list[j];
            const AlignRoi q = make_align_roi(a.rois5 + (size_t)k * 5, k, a.scale, a.PH, a.PW, a.sampling_ratio, a.aligned);
            const float* gp = go + ((size_t)k * a.C + c0) * PP + e;
            float g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) g[c] = c < cs ? __ldg(gp + (size_t)c * PP) / q.count : 0.f;
            for (int iy = 0; iy < q.gh; ++iy) {
                float yy = q.sy + (float)ph * q.bin_h;
                yy = yy + ((float)iy + .5f) * q.bin_h / (float)q.gh;
                for (int ix = 0; ix < q.gw; ++ix) {
                    float xx = q.sx + (float)pw * q.bin_w;
                    xx = xx + ((float)ix + .5f) * q.bin_w / (float)q.gw;
                    float y = yy, x = xx;
                    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
                    if (y <= 0.f) y = 0.f;
                    if (x <= 0.f) x = 0.f;
                    int yl = (int)y, xl = (int)x, yh, xh;
                    if (yl >= H - 1) {
                        yh = yl = H - 1;
                        y = (float)yl;
                    } else {
                        yh = yl + 1;
                    }
                    if (xl >= W - 1) {
                        xh = xl = W - 1;
                        x = (float)xl;
                    } else {
                        xh = xl + 1;
                    }
                    const float ly = y - (float)yl, lx = x - (float)xl;
                    const float hy = 1.f - ly, hx = 1.f - lx;
                    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < cs) {
                            float* plane = s_acc + c * HW;
                            atomicAdd(plane + yl * W + xl, g[c] * w1);
                            atomicAdd(plane + yl * W + xh, g[c] * w2);
                            atomicAdd(plane + yh * W + xl, g[c] * w3);
                            atomicAdd(plane + yh * W + xh, g[c] * w4);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    float* dst = gi + ((size_t)b * a.C + c0) * HW;
    for (int i = threadIdx.x; i < cs * HW; i += BW_THREADS) dst[i] += s_acc[i];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct RoiWs {
    int* perm;
    int* offs;
    int* sorted;  // streaming RoIAlign: rows of every image, longest row program first
    unsigned char* prog;  // streaming RoIAlign: the records re-laid per (pass, warp) by roi_align_stream_pack_kernel
    int2* ent;  // [num_rois][<= 144] bin / sample geometry of the inference table kernels (roi_pool_entries_kernel: 2*P
                // words per RoI, roi_align_entries_kernel: 2*P*SR, roi_align_fast_entries_kernel: 2*P*3,
                // roi_align_stream_entries_kernel: AS_REC bytes)
};

static size_t roi_layout(Workspace& ws, int batch, int num_rois, RoiWs* out) {
    RoiWs w;
    w.perm = ws.take<int>(num_rois > 0 ? num_rois : 1);
    w.offs = ws.take<int>(batch + 2);
    w.sorted = ws.take<int>(num_rois > 0 ? num_rois : 1);
    // up to 1152 bytes per RoI for the streaming RoIAlign records, 1568 for the 7x7 RoIPool lookup lists
    w.ent = ws.take<int2>((size_t)(num_rois > 0 ? num_rois : 1) * 200);
    // streaming RoIAlign, packed per-warp programs: one AS2_SLOT per 4 RoIs, passes of >= 24 RoIs, one ragged pass per image
    w.prog = ws.take<unsigned char>(((size_t)(num_rois > 0 ? num_rois : 1) / 24 + batch + 2) * 8 * AS2_SLOT);
    if (out) *out = w;
    return ws.off;
}

constexpr size_t ROI_SMEM_TARGET = 100 * 1024;  // two CTAs per SM
constexpr size_t ROI_SMEM_MAX = 200 * 1024;

// channels per slab: largest power of two whose planes fit the shared-memory target
static int pick_slab(int C, int HW) {
    size_t plane = (size_t)HW * 4;
    if (plane > ROI_SMEM_MAX) return 0;  // cannot stage
    int cs = 1;
    while (cs * 2 <= C && (size_t)(cs * 2) * plane <= ROI_SMEM_TARGET && cs * 2 <= 64) cs *= 2;
    return cs;
}

static int pick_groups(int K, int B, int slabs) {
    // enough CTAs for >= 4 waves of 2 CTAs/SM, but at least ~2*ROI_NB RoIs per CTA
    int per_image = (K + B - 1) / B;
    int g = (per_image + 2 * ROI_NB - 1) / (2 * ROI_NB);
    int want = (8 * sm_count() + B * slabs - 1) / (B * slabs);
    if (g > want) g = want;
    if (g < 1) g = 1;
    return g;
}

static int check_roi_common(const float* feat, int B, int C, int H, int W, const float* rois5, int K, int PH,
                            int PW, const float* out, const char* who) {
    FRCNN_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && K >= 0 && PH > 0 && PW > 0, "%s: bad shape", who);
    FRCNN_CHECK_ARG(H < 32768 && W < 32768, "%s: feature map too large", who);
    FRCNN_CHECK_ARG(K == 0 || (feat && rois5 && out), "%s: null pointer", who);
    return FRCNN_OK;
}

template <typename KernelT>
static int launch_staged(KernelT kernel, const RoiArgs& a, size_t smem, cudaStream_t stream) {
    FRCNN_SMEM(kernel, smem);
    int slabs = cdiv(a.C, a.CS);
    FRCNN_CHECK_ARG(slabs <= 65535 && a.B <= 65535, "roi op: too many channel slabs / images");
    dim3 grid(a.groups, slabs, a.B);
    kernel<<<grid, ROI_THREADS, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    note_roi_kernel("staged kernel, %d channels per CTA, %d threads", a.CS, ROI_THREADS);
    return FRCNN_OK;
}

template <typename KernelT>
static int launch_align(const char* name, KernelT kernel, const RoiArgs& a, size_t smem, cudaStream_t stream) {
    FRCNN_SMEM(kernel, smem);
    int slabs = cdiv(a.C, 4);
    FRCNN_CHECK_ARG(slabs <= 65535 && a.B <= 65535, "roi op: too many channel slabs / images");
    dim3 grid(a.groups, slabs, a.B);
    kernel<<<grid, 392, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    note_roi_kernel("%s", name);
    return FRCNN_OK;
}

#define FRCNN_STR2(...) #__VA_ARGS__
#define FRCNN_STR(...) FRCNN_STR2(__VA_ARGS__)

template <typename KernelT>
static int launch_tab(const char* name, KernelT kernel, const RoiArgs& a, size_t smem, int threads, cudaStream_t stream) {
    FRCNN_SMEM(kernel, smem);
    int slabs = cdiv(a.C, a.CS);
    FRCNN_CHECK_ARG(slabs <= 65535 && a.B <= 65535, "roi op: too many channel slabs / images");
    dim3 grid(a.groups, slabs, a.B);
    kernel<<<grid, threads, smem, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    note_roi_kernel("%s", name);
    return FRCNN_OK;
}

// geometry entries for an inference table kernel with CS_ channels per table element (see roi_pool_tab_kernel)
static int launch_entries(RoiArgs& a, int P, bool diag, int cs, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream, const char* who) {
    Workspace ws(workspace, workspace_bytes);
    RoiWs w;
    roi_layout(ws, a.B, a.K, &w);
    if (!ws.ok()) {
        set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ws.off, workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    const int HWp = (a.H * a.pitch + 3) & ~3, esz = cs * 4, blocks = cdiv(a.K * 2 * P, 256);
    if (diag) {
        if (P == 7) roi_pool_entries_kernel<2, true, true><<<blocks, 256, 0, stream>>>(a, w.ent, P, esz, HWp);
        else roi_pool_entries_kernel<2, false, true><<<blocks, 256, 0, stream>>>(a, w.ent, P, esz, HWp);
    } else {
        if (P == 7) roi_pool_entries_kernel<2, true, false><<<blocks, 256, 0, stream>>>(a, w.ent, P, esz, HWp);
        else roi_pool_entries_kernel<2, false, false><<<blocks, 256, 0, stream>>>(a, w.ent, P, esz, HWp);
    }
    FRCNN_LAUNCH_CHECK();
    a.ent = w.ent;
    return FRCNN_OK;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_roi_head_coords(const float* rois, const int32_t* roi_indices, int32_t n_images, int32_t per_image,
                          float d0, float d1, int32_t fh, int32_t fw, float* rois5, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(n_images >= 0 && per_image >= 0, "frcnn_roi_head_coords: bad shape");
    int total = n_images * per_image;
    if (total == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(rois && roi_indices && rois5, "frcnn_roi_head_coords: null pointer");
    roi_head_coords_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)rois, roi_indices, total, per_image, d0, d1, (float)fh, (float)fw, rois5);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

size_t frcnn_roi_workspace_bytes(int32_t batch, int32_t num_rois) {
    Workspace ws(nullptr, 0);
    return roi_layout(ws, batch > 0 ? batch : 1, num_rois, nullptr);
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int roi_forward_common(bool align, const float* feat, int B, int C, int H, int W, const float* rois5,
                              int K, int per_image, int PH, int PW, float scale, int sampling_ratio, int aligned, float* out,
                              int32_t* argmax, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                              const char* who, bool mean = false, bool exact = true) {
    int rc = check_roi_common(feat, B, C, H, W, rois5, K, PH, PW, out, who);
    if (rc) return rc;
    if (K == 0) return FRCNN_OK;
    RoiArgs a;
    memset(&a, 0, sizeof(a));
    a.feat = feat;
    a.rois5 = rois5;
    a.B = B;
    a.C = C;
    a.H = H;
    a.W = W;
    a.K = K;
    a.PH = PH;
    a.PW = PW;
    a.scale = scale;
    a.out = out;
    a.argmax = argmax;
    a.sampling_ratio = sampling_ratio;
    a.aligned = aligned;
    int cs = pick_slab(C, H * W);
    bool staged = cs > 0 && PH <= ROI_MAX_P && PW <= ROI_MAX_P && B <= 4096;
    if (mean && !(staged && !align && PH == PW && (PH == 7 || PH == 14))) {
        set_error("%s: fused pool + mean supports 7x7 / 14x14 RoIPool on maps that fit in shared memory", who);
        return FRCNN_ERR_UNSUPPORTED;
    }
    if (!staged) {
        size_t total = (size_t)K * C * PH * PW;
        int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 32);
        if (align) roi_align_direct_kernel<<<blocks, 256, 0, stream>>>(a);
        else if (argmax) roi_pool_direct_kernel<true><<<blocks, 256, 0, stream>>>(a);
        else roi_pool_direct_kernel<false><<<blocks, 256, 0, stream>>>(a);
        FRCNN_LAUNCH_CHECK();
        return FRCNN_OK;
    }
    if (per_image > 0) {
        FRCNN_CHECK_ARG((int64_t)per_image * B == K, "%s: rois_per_image * batch != num_rois", who);
        a.per_image = per_image;  // caller guarantees grouping: no bucketing pass
    } else {
        Workspace ws(workspace, workspace_bytes);
        RoiWs w;
        roi_layout(ws, B, K, &w);
        if (!ws.ok()) {
            set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ws.off, workspace_bytes);
            return FRCNN_ERR_WORKSPACE;
        }
        roi_bucket_kernel<<<1, BUCKET_THREADS, 2 * B * sizeof(int), stream>>>(rois5, K, B, w.perm, w.offs);
        FRCNN_LAUNCH_CHECK();
        roi_fill_dropped_kernel<<<sm_count(), 256, 0, stream>>>(w.perm, w.offs, B, mean ? (size_t)C : (size_t)C * PH * PW,
                                                               out, argmax);
        FRCNN_LAUNCH_CHECK();
        a.perm = w.perm;
        a.offs = w.offs;
    }
    a.CS = cs;
    a.groups = pick_groups(K, B, cdiv(C, cs));
    size_t smem = (size_t)cs * H * W * 4;
    static const int align_impl = env_int("FRCNN_ALIGN_IMPL", 0);  // experiments only: 1 = thread-per-bin fast kernel
    if (align && !exact && sampling_ratio == 2 && PH == 7 && PW == 7 && C % 4 == 0 && align_impl != 1 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // streaming variant: column threads, every pixel row read once per bin column, row programs through a
        // per-warp shared-memory ring, TMA bulk stores (see roi_align_stream2_kernel)
        static const int pitch_override = env_int("FRCNN_ALIGN_PITCH", 0);
        a.pitch = pitch_override >= W ? pitch_override : (W | 1);
        static const int th_override = env_int("FRCNN_ALIGN_THREADS", 0);  // experiments only
        const int T2 = th_override == 256 || th_override == 224 ? th_override : 192;  // cfg4: 0.432 / 0.444 / 0.464 ms
        const int NQ2 = T2 / 8, RING2 = T2 == 224 ? 3 : 4;
        const size_t tab2 = ((size_t)(H + 1) * a.pitch * 16 + 127) & ~(size_t)127;
        const size_t smem2 = tab2 + (size_t)NQ2 * AS_STAGE + (size_t)(T2 / 32) * RING2 * AS2_CHUNK;
        if (smem2 <= 110 * 1024 && W < 4096) {  // at least two CTAs per SM (three on maps up to 50 x 50); 16-bit tap offsets
            a.CS = 4;
            const int slabs = C / 4;
            const int passes = cdiv(cdiv(K, B), NQ2);
            a.groups = std::max(1, std::min(cdiv(passes, 4), cdiv(8 * sm_count(), B * slabs)));
            Workspace ews(workspace, workspace_bytes);
            RoiWs w;
            roi_layout(ews, B, K, &w);
            if (!ews.ok()) {
                set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ews.off, workspace_bytes);
                return FRCNN_ERR_WORKSPACE;
            }
            FRCNN_CHECK_ARG(slabs <= 65535 && B <= 65535, "roi op: too many channel slabs / images");
            FRCNN_CHECK_ARG(((size_t)K / 24 + B + 2) * 8 * (AS2_SLOT / 16) < 0xFFFFFFFFull, "roi op: too many RoIs");
            roi_align_stream_entries_kernel<<<cdiv(K, 4), 128, 0, stream>>>(a, (unsigned char*)w.ent);
            FRCNN_LAUNCH_CHECK();
            roi_align_stream_sort_kernel<<<B, 256, 0, stream>>>(a, (const unsigned char*)w.ent, w.sorted);
            FRCNN_LAUNCH_CHECK();
            a.ent = w.ent;
            a.perm2 = w.sorted;
            const dim3 pgrid(passes, B), grid2(a.groups, slabs, B);
#define FRCNN_STREAM2(TH_, MB_, RG_)                                                                              \
    do {                                                                                                         \
        roi_align_stream_pack_kernel<TH_><<<pgrid, TH_, 0, stream>>>(a, (const unsigned char*)w.ent, w.sorted, w.prog); \
        FRCNN_LAUNCH_CHECK();                                                                                    \
        FRCNN_SMEM((roi_align_stream2_kernel<TH_, MB_, RG_>), smem2);                                            \
        roi_align_stream2_kernel<TH_, MB_, RG_><<<grid2, TH_, smem2, stream>>>(a, w.prog);                       \
        FRCNN_LAUNCH_CHECK();                                                                                    \
        note_roi_kernel("roi_align_stream2_kernel<%d,%d,%d> pitch %d", TH_, MB_, RG_, a.pitch);                  \
        return FRCNN_OK;                                                                                         \
    } while (0)
            if (T2 == 256) FRCNN_STREAM2(256, 2, 4);
            if (T2 == 224) FRCNN_STREAM2(224, 3, 3);
            FRCNN_STREAM2(192, 3, 4);
#undef FRCNN_STREAM2
        }
    }
    if (align && !exact && sampling_ratio == 2 && PH == PW && (PH == 7 || PH == 14)) {
        // fast variant (FMA, merged separable weights): table with an odd pitch + the staging area must fit
        static const int pitch_override = env_int("FRCNN_ALIGN_PITCH", 0);  // experiments only
        a.pitch = pitch_override >= W ? pitch_override : (W | 1);
        const size_t fsmem = (size_t)((H * a.pitch + 3) & ~3) * sizeof(float4) + (size_t)4 * H * W * sizeof(float);
        if (fsmem <= 200 * 1024) {
            a.CS = 4;
            const bool two = fsmem <= 97 * 1024;  // two CTAs per SM (static part: ~17 KB)
            const int slabs = cdiv(C, 4);
            a.groups = std::max(1, std::min(cdiv(cdiv(K, B), 4 * (PH == 7 ? 16 : 8)), cdiv(8 * sm_count(), B * slabs)));
            Workspace ews(workspace, workspace_bytes);
            RoiWs w;
            roi_layout(ews, B, K, &w);
            if (!ews.ok()) {
                set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ews.off, workspace_bytes);
                return FRCNN_ERR_WORKSPACE;
            }
            roi_align_fast_entries_kernel<<<cdiv(K * 2 * PH, 256), 256, 0, stream>>>(a, (unsigned char*)w.ent, PH);
            FRCNN_LAUNCH_CHECK();
            a.ent = w.ent;
            FRCNN_CHECK_ARG(slabs <= 65535 && B <= 65535, "roi op: too many channel slabs / images");
            const dim3 grid(a.groups, slabs, B);
#define FRCNN_FAST(PP_, TH_, MB_)                                                              \
    do {                                                                                      \
        FRCNN_SMEM((roi_align_fast_kernel<PP_, TH_, MB_>), fsmem);                            \
        roi_align_fast_kernel<PP_, TH_, MB_><<<grid, TH_, fsmem, stream>>>(a);               \
        FRCNN_LAUNCH_CHECK();                                                                 \
        note_roi_kernel("roi_align_fast_kernel<%d,%d,%d> pitch %d", PP_, TH_, MB_, a.pitch); \
        return FRCNN_OK;                                                                      \
    } while (0)
            if (PH == 7) {
                if (two) FRCNN_FAST(7, 392, 2);
                FRCNN_FAST(7, 784, 1);
            }
            if (two) FRCNN_FAST(14, 392, 2);
            FRCNN_FAST(14, 784, 1);
#undef FRCNN_FAST
        }
    }
    if (align) {
        // interleaved kernel for the common fixed 2x2 sampling grid; plane + staging area must fit
        const size_t al_smem = (size_t)2 * ((H * W + 3) & ~3) * sizeof(float4);
        if (sampling_ratio == 2 && PH == PW && (PH == 7 || PH == 14) && al_smem <= 200 * 1024) {
            a.CS = 4;
            const int minb = al_smem <= 100 * 1024 ? 2 : 1;
            const int nbatch = PH == 7 ? 40 : 20;  // NB of the kernel template
            int slabs = cdiv(C, 4);
            int g = cdiv(cdiv(K, B), 4 * nbatch);
            int want = cdiv(8 * sm_count(), B * slabs);
            a.groups = std::max(1, std::min(g, want));
            {
                Workspace ews(workspace, workspace_bytes);
                RoiWs w;
                roi_layout(ews, B, K, &w);
                if (!ews.ok()) {
                    set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ews.off, workspace_bytes);
                    return FRCNN_ERR_WORKSPACE;
                }
                roi_align_entries_kernel<<<cdiv(K * 2 * PH * 2, 256), 256, 0, stream>>>(a, (AlignEntry*)w.ent, PH, 2);
                FRCNN_LAUNCH_CHECK();
                a.ent = w.ent;
            }
            if (PH == 7)
                return minb == 2 ? launch_align(FRCNN_STR(roi_align_tab_kernel<7, 2, 392, 2>), roi_align_tab_kernel<7, 2, 392, 2>, a, al_smem, stream)
                                 : launch_align(FRCNN_STR(roi_align_tab_kernel<7, 2, 392, 1>), roi_align_tab_kernel<7, 2, 392, 1>, a, al_smem, stream);
            return minb == 2 ? launch_align(FRCNN_STR(roi_align_tab_kernel<14, 2, 392, 2>), roi_align_tab_kernel<14, 2, 392, 2>, a, al_smem, stream)
                             : launch_align(FRCNN_STR(roi_align_tab_kernel<14, 2, 392, 1>), roi_align_tab_kernel<14, 2, 392, 1>, a, al_smem, stream);
        }
        return launch_staged(roi_align_staged_kernel, a, smem, stream);
    }
    // RoIPool 7x7 / 14x14: thread-per-bin kernel over shared-memory max tables (roi_pool_tab_kernel).
    //   LV = 2 (windows 1,2: bins up to 4 long; 5..8 through four 2-windows on the 7x7 grid)  inference
    //   LV = 1 (pixels only, every bin scanned)  training: value + argmax in one scan (small tables, several
    //          CTAs per SM: the plain two-buffer / CTA-barrier hand-off, a third buffer would cost a CTA)
    // (a three-level form -- windows 1,2,4, nine tables -- was measured and dropped: it only fits with one or
    //  two channels per CTA, and the per-bin instruction overhead is paid per channel group)
    if (PH == PW && (PH == 7 || PH == 14)) {
        const int per_image_rois = cdiv(K, B);
        const size_t budget2 = 92 * 1024, budget1 = 180 * 1024;  // dynamic part for 2 / 1 CTAs per SM
        auto table_bytes = [&](int lv, int tcs, int pitch) {
            return (size_t)(lv == 1 ? 2 : lv * lv) * (size_t)((H * pitch + 3) & ~3) * 4 * tcs;
        };
        auto set_groups = [&](int tcs, int threads) {
            const int per_batch = threads / PH;
            const int slabs = cdiv(C, tcs);
            const int g = cdiv(per_image_rois, 4 * per_batch);  // >= 4 batches per CTA amortise the table build
            const int want = cdiv(8 * sm_count(), B * slabs);
            a.CS = tcs;
            a.groups = std::max(1, std::min(g, want));
        };
#define FRCNN_TAB(PP_, TH_, CS_, MB_, AM_, LV_)                                                             \
    do {                                                                                                    \
        set_groups(CS_, TH_);                                                                               \
        if (!(AM_)) {                                                                                       \
            const int rc_ = launch_entries(a, PP_, false, CS_, workspace, workspace_bytes, stream, who);   \
            if (rc_) return rc_;                                                                            \
        }                                                                                                   \
        return launch_tab(FRCNN_STR(roi_pool_tab_kernel<PP_, TH_, CS_, MB_, AM_, LV_, 1, !(AM_)>), roi_pool_tab_kernel<PP_, TH_, CS_, MB_, AM_, LV_, 1, !(AM_)>, a, table_bytes(LV_, CS_, a.pitch), \
                          TH_, stream);                                                                     \
    } while (0)
        // rows whose byte length is a multiple of 64 would put vertically adjacent bins on the same banks
        // (64-wide maps: 45 % of the shared-memory wavefronts were conflicts): pad the pitch by one pixel
        auto pitch_for = [&](int tcs) { return (W * 4 * tcs) % 64 == 0 ? W + 1 : W; };
        // 7x7: the gather driven by per-bin lookup lists, computed once per RoI for all slabs (roi_pool_gather_kernel).
        // Window tables: pixels, 2 x 2 and -- when that does not cost a CTA per SM -- 3 x 3, with which a typical
        // 2..3 pixel bin is one or two lookups instead of four (64 x 64 map, one CTA per SM either way: 0.562 vs
        // 0.616 ms, table kernel 0.595; 50 x 50 map: two tables and two CTAs per SM 0.362 ms, three tables and one
        // CTA 0.384, table kernel 0.383).
        constexpr int FRCNN_LIST_NA = 1 << 20;
        auto launch_list = [&](int pitch, bool mean_out) -> int {
            static const int pool_impl = env_int("FRCNN_POOL_IMPL", 0);  // experiments only: 1 = table kernel
            const int HWp_d = (H * pitch + 3) & ~3;
            if (!(PH == 7 && C % 4 == 0 && 2 * HWp_d + 2 <= (int)PD_MASK && pool_impl != 1 &&
                  (int64_t)C * 49 * 4 < ((int64_t)1 << 32)))
                return FRCNN_LIST_NA;
            static const int pool_tables = env_int("FRCNN_POOL_TABLES", 0);  // experiments only: force 2 / 3
            const size_t smem2t = (size_t)2 * HWp_d * 16 + 32, smem3t = (size_t)3 * HWp_d * 16 + 32;
            auto ctas_per_sm = [](size_t bytes) { return bytes + 2048 <= 113 * 1024 ? 2 : 1; };
            if (smem2t > 220 * 1024) return FRCNN_LIST_NA;
            bool three = smem3t <= 220 * 1024 && 3 * HWp_d + 2 <= (int)PD_MASK && ctas_per_sm(smem3t) == ctas_per_sm(smem2t);
            if (pool_tables == 2) three = false;
            if (pool_tables == 3 && smem3t <= 220 * 1024 && 3 * HWp_d + 2 <= (int)PD_MASK) three = true;
            const size_t gsmem = three ? smem3t : smem2t;
            const int per_sm = ctas_per_sm(gsmem);
            // fused pool + mean: a warp per RoI leaves 15 of 64 lane slots idle, which only pays where the table kernel is
            // down to one CTA per SM (64 x 64: 0.573 vs 0.641 ms; 50 x 50: 0.409 vs 0.404; 38 x 38: 0.375 vs 0.280)
            if (mean_out && per_sm == 2) return FRCNN_LIST_NA;
            Workspace ews(workspace, workspace_bytes);
            RoiWs w;
            roi_layout(ews, B, K, &w);
            if (!ews.ok()) {
                set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ews.off, workspace_bytes);
                return FRCNN_ERR_WORKSPACE;
            }
            uint2* const desc = reinterpret_cast<uint2*>(w.ent);
            uint2* const extra = desc + (size_t)K * 49;
            a.pitch = pitch;
            a.CS = 4;
            const int slabs = C / 4, th = per_sm == 2 ? 512 : 1024;
            FRCNN_CHECK_ARG(slabs <= 65535 && B <= 65535, "roi op: too many channel slabs / images");
            // every CTA builds the slab's tables: split an image's RoIs over several only to reach ~4 waves
            a.groups = std::max(1, std::min(cdiv(4 * per_sm * sm_count(), B * slabs),
                                            mean_out ? cdiv(per_image_rois, 4 * (th / 32)) : cdiv((int64_t)per_image_rois * 49, 8 * th)));
            static const int groups_override = env_int("FRCNN_POOL_GROUPS", 0);  // experiments only
            if (groups_override > 0) a.groups = groups_override;
            roi_pool_desc_kernel<<<cdiv((int64_t)K * 49, 256), 256, 0, stream>>>(a, desc, extra, 7, HWp_d, three ? 3 : 2);
            FRCNN_LAUNCH_CHECK();
            const dim3 grid(a.groups, slabs, B);
#define FRCNN_GATHER(TH_, MB_, NT_)                                                                      \
do {                                                                                                \
    FRCNN_SMEM((roi_pool_gather_kernel<7, TH_, MB_, NT_>), gsmem);                                  \
    roi_pool_gather_kernel<7, TH_, MB_, NT_><<<grid, TH_, gsmem, stream>>>(a, desc, extra);         \
    FRCNN_LAUNCH_CHECK();                                                                           \
    note_roi_kernel("roi_pool_gather_kernel<7,%d,%d,%d>", TH_, MB_, NT_);                           \
    return FRCNN_OK;                                                                                \
} while (0)
#define FRCNN_MEANL(TH_, MB_, NT_)                                                                       \
do {                                                                                                \
    FRCNN_SMEM((roi_pool_mean_list_kernel<7, TH_, MB_, NT_>), gsmem);                               \
    roi_pool_mean_list_kernel<7, TH_, MB_, NT_><<<grid, TH_, gsmem, stream>>>(a, desc, extra);      \
    FRCNN_LAUNCH_CHECK();                                                                           \
    note_roi_kernel("roi_pool_mean_list_kernel<7,%d,%d,%d>", TH_, MB_, NT_);                        \
    return FRCNN_OK;                                                                                \
} while (0)
            if (mean_out) {
                if (three) {
                    if (per_sm == 2) FRCNN_MEANL(512, 2, 3);
                    FRCNN_MEANL(1024, 1, 3);
                }
                if (per_sm == 2) FRCNN_MEANL(512, 2, 2);
                FRCNN_MEANL(1024, 1, 2);
            }
            if (three) {
                if (per_sm == 2) FRCNN_GATHER(512, 2, 3);
                FRCNN_GATHER(1024, 1, 3);
            }
            if (per_sm == 2) FRCNN_GATHER(512, 2, 2);
            FRCNN_GATHER(1024, 1, 2);
#undef FRCNN_GATHER
#undef FRCNN_MEANL
        };
        if (mean && !argmax) {
            static const int mean_impl = env_int("FRCNN_MEAN_IMPL", 0);  // experiments only: 1 = table kernel
            if (mean_impl != 1) {
                const int rc_ = launch_list(pitch_for(4), true);
                if (rc_ != FRCNN_LIST_NA) return rc_;
            }
        }
        if (mean) {  // [K,C] = mean over the bins of RoIPool, never materialising [K,C,P,P]
            // four 4-channel tables when they fit, else two (pixels, 2 x 2 windows: same bytes as four 2-channel
            // tables, twice the channels per lookup); one CTA per SM when the tables are that large -> 1024 threads
            a.pitch = pitch_for(4);
            const size_t smem4 = table_bytes(2, 4, a.pitch);
            const bool diag = smem4 > 200 * 1024;
            const size_t smem = diag ? smem4 / 2 : smem4;
            if (smem > 200 * 1024) {
                set_error("%s: feature map too large for the fused pool + mean kernel", who);
                return FRCNN_ERR_UNSUPPORTED;
            }
            const int threads = smem > 100 * 1024 ? 1024 : PM_THREADS;
            const int per_iter = (threads / 32) * (PH <= 8 ? 4 : 2);  // RoIs per CTA pass
            a.CS = 4;
            a.groups = std::max(1, std::min(cdiv(per_image_rois, 4 * per_iter), cdiv(8 * sm_count(), B * cdiv(C, 4))));
            {
                const int rc_ = launch_entries(a, PH, diag, 4, workspace, workspace_bytes, stream, who);
                if (rc_) return rc_;
            }
#define FRCNN_MEAN(PP_, DG_, TH_) return launch_tab(FRCNN_STR(roi_pool_mean_kernel<PP_, 4, 2, DG_, TH_>), roi_pool_mean_kernel<PP_, 4, 2, DG_, TH_>, a, smem, TH_, stream)
            if (PH == 7) {
                if (diag) { if (threads == 1024) FRCNN_MEAN(7, true, 1024); FRCNN_MEAN(7, true, PM_THREADS); }
                if (threads == 1024) FRCNN_MEAN(7, false, 1024);
                FRCNN_MEAN(7, false, PM_THREADS);
            }
            if (diag) { if (threads == 1024) FRCNN_MEAN(14, true, 1024); FRCNN_MEAN(14, true, PM_THREADS); }
            if (threads == 1024) FRCNN_MEAN(14, false, 1024);
            FRCNN_MEAN(14, false, PM_THREADS);
#undef FRCNN_MEAN
        }
        if (argmax) {
            a.pitch = W | 1;
            // persistent kernel over (image, slab) items with precomputed bin ranges: interleaved pixels + two
            // staging buffers (the next item's planes arrive while this one is scanned)
            static const int train_impl = env_int("FRCNN_TRAIN_IMPL", 0);  // experiments only: 1 = table-kernel scan
            const size_t tr_smem = (size_t)((H * a.pitch + 3) & ~3) * 16 + (size_t)2 * ((4 * H * W + 3) & ~3) * 4;
            const size_t tr_static = (size_t)TR_CHUNK * (2 * PH + 1) * 4 + 64;
            if (train_impl != 1 && tr_smem + tr_static <= 220 * 1024) {
                Workspace ews(workspace, workspace_bytes);
                RoiWs w;
                roi_layout(ews, B, K, &w);
                if (!ews.ok()) {
                    set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ews.off, workspace_bytes);
                    return FRCNN_ERR_WORKSPACE;
                }
                FRCNN_CHECK_ARG((int64_t)C * PH * PW * 4 < ((int64_t)1 << 32), "%s: too many channels", who);
                a.CS = 4;
                const int items = B * cdiv(C, 4);
                const int per_sm = tr_smem + tr_static + 1024 <= 113 * 1024 ? 2 : 1;
                const int ctas = std::min(items, per_sm * sm_count());
                int* const rng = reinterpret_cast<int*>(w.ent);
                roi_pool_ranges_kernel<<<cdiv((int64_t)K * 2 * PH, 256), 256, 0, stream>>>(a, rng, PH, w.sorted, ctas);
                FRCNN_LAUNCH_CHECK();
                static const int tr_threads = env_int("FRCNN_TRAIN_THREADS", 512);  // experiments only
#define FRCNN_TRAIN(PP_, TH_)                                                                   \
    do {                                                                                       \
        FRCNN_SMEM((roi_pool_train_kernel<PP_, TH_>), tr_smem);                                \
        roi_pool_train_kernel<PP_, TH_><<<ctas, TH_, tr_smem, stream>>>(a, rng, w.sorted);     \
        FRCNN_LAUNCH_CHECK();                                                                  \
        note_roi_kernel("roi_pool_train_kernel<%d,%d>", PP_, TH_);                             \
        return FRCNN_OK;                                                                       \
    } while (0)
                if (PH == 7) {
                    if (tr_threads == 384) FRCNN_TRAIN(7, 384);
                    FRCNN_TRAIN(7, 512);
                }
                if (tr_threads == 384) FRCNN_TRAIN(14, 384);
                FRCNN_TRAIN(14, 512);
#undef FRCNN_TRAIN
            }
            if (table_bytes(1, 4, a.pitch) <= budget2) {
                if (PH == 7) FRCNN_TAB(7, 392, 4, 2, true, 1);
                FRCNN_TAB(14, 392, 4, 2, true, 1);
            }
            if (table_bytes(1, 1, a.pitch) <= budget1) {
                if (PH == 7) FRCNN_TAB(7, 392, 1, 2, true, 1);
                FRCNN_TAB(14, 392, 1, 2, true, 1);
            }
        }
        if (!argmax) {
            const size_t smem4 = table_bytes(2, 4, pitch_for(4)), smem2 = table_bytes(2, 2, pitch_for(2));
            int tcs = 0, minb = 0;
            if (smem4 <= budget2) tcs = 4, minb = 2;
            else if (smem2 <= budget2) tcs = 2, minb = 2;
            else if (smem4 <= 200 * 1024) tcs = 4, minb = 1;
            else if (smem2 <= 200 * 1024) tcs = 2, minb = 1;
            a.pitch = pitch_for(tcs ? tcs : 4);
            // maps whose four 4-channel tables do not fit: two tables (pixels, 2 x 2 windows) with four
            // channels per lookup instead of four tables with two
            const int pitch_d = pitch_for(4);  // (a sweep over 65 ... 71 pixels on the 64-wide map moved the time by < 0.5 %)
            const size_t smemd = (size_t)2 * (size_t)((H * pitch_d + 3) & ~3) * 16;
            if (tcs == 2 && smemd + 40 * 1024 <= 220 * 1024) {
                a.pitch = pitch_d;
                const bool two = smemd <= budget2;
                {
                    const int rc_ = launch_list(pitch_d, false);
                    if (rc_ != FRCNN_LIST_NA) return rc_;
                }
#define FRCNN_TABD(PP_, TH_, MB_)                                                                            \
    do {                                                                                                    \
        set_groups(4, TH_);                                                                                 \
        const int rc_ = launch_entries(a, PP_, true, 4, workspace, workspace_bytes, stream, who);          \
        if (rc_) return rc_;                                                                                \
        return launch_tab(FRCNN_STR(roi_pool_tab_kernel<PP_, TH_, 4, MB_, false, 2, 1, true, true>), roi_pool_tab_kernel<PP_, TH_, 4, MB_, false, 2, 1, true, true>, a, smemd, TH_, stream); \
    } while (0)
                if (PH == 7) {
                    if (two) FRCNN_TABD(7, 392, 2);
                    FRCNN_TABD(7, 784, 1);
                }
                if (two) FRCNN_TABD(14, 392, 2);
                FRCNN_TABD(14, 392, 1);
#undef FRCNN_TABD
            }
            if (tcs == 4 && minb == 2) {
                // maps small enough for the four-table kernel at two CTAs per SM take the list gather too (38 x 38 x 1024,
                // 4 800 RoIs: 0.308 vs 0.336 ms)
                if (PH == 7) {
                    const int rc_ = launch_list(pitch_for(4), false);
                    if (rc_ != FRCNN_LIST_NA) return rc_;
                }
                if (PH == 7) FRCNN_TAB(7, 392, 4, 2, false, 2);
                set_groups(4, 392);  // 14x14: two adjacent bins per thread
                return launch_tab(FRCNN_STR(roi_pool_tab_kernel<14, 392, 4, 2, false, 2, 2, true>), roi_pool_tab_kernel<14, 392, 4, 2, false, 2, 2, true>, a, table_bytes(2, 4, a.pitch), 392,
                                  stream);
            } else if (tcs == 2 && minb == 2) {
                if (PH == 7) FRCNN_TAB(7, 392, 2, 2, false, 2);
                FRCNN_TAB(14, 392, 2, 2, false, 2);
            } else if (tcs == 4) {
                if (PH == 7) FRCNN_TAB(7, 392, 4, 1, false, 2);
                FRCNN_TAB(14, 392, 4, 1, false, 2);
            } else if (tcs == 2) {
                // one CTA per SM: twice the threads to keep the SM's latency hidden
                if (PH == 7 && smem2 + 40 * 1024 <= 220 * 1024) FRCNN_TAB(7, 784, 2, 1, false, 2);
                if (PH == 7) FRCNN_TAB(7, 392, 2, 1, false, 2);
                FRCNN_TAB(14, 392, 2, 1, false, 2);
            }
        }
#undef FRCNN_TAB
    }
    if (PH == PW && PH == 7)
        return argmax ? launch_staged(roi_pool_staged_kernel<7, true>, a, smem, stream)
                      : launch_staged(roi_pool_staged_kernel<7, false>, a, smem, stream);
    if (PH == PW && PH == 14)
        return argmax ? launch_staged(roi_pool_staged_kernel<14, true>, a, smem, stream)
                      : launch_staged(roi_pool_staged_kernel<14, false>, a, smem, stream);
    return argmax ? launch_staged(roi_pool_staged_kernel<0, true>, a, smem, stream)
                  : launch_staged(roi_pool_staged_kernel<0, false>, a, smem, stream);
}

int frcnn_roi_pool_forward(const float* feat, int32_t B, int32_t C, int32_t H, int32_t W, const float* rois5,
                           int32_t K, int32_t per_image, int32_t PH, int32_t PW, float scale, float* out,
                           int32_t* argmax, void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    return roi_forward_common(false, feat, B, C, H, W, rois5, K, per_image, PH, PW, scale, 0, 0, out, argmax, workspace,
                              workspace_bytes, (cudaStream_t)stream, "frcnn_roi_pool_forward");
}

int frcnn_roi_pool_mean_forward(const float* feat, int32_t B, int32_t C, int32_t H, int32_t W, const float* rois5,
                                int32_t K, int32_t per_image, int32_t PH, int32_t PW, float scale, float* out,
                                void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    return roi_forward_common(false, feat, B, C, H, W, rois5, K, per_image, PH, PW, scale, 0, 0, out, nullptr, workspace,
                              workspace_bytes, (cudaStream_t)stream, "frcnn_roi_pool_mean_forward", true);
}

int frcnn_roi_align_forward(const float* feat, int32_t B, int32_t C, int32_t H, int32_t W, const float* rois5,
                            int32_t K, int32_t per_image, int32_t PH, int32_t PW, float scale,
                            int32_t sampling_ratio, int32_t aligned, int32_t exact, float* out, void* workspace,
                            size_t workspace_bytes, frcnn_stream_t stream) {
    return roi_forward_common(true, feat, B, C, H, W, rois5, K, per_image, PH, PW, scale, sampling_ratio, aligned, out,
                              nullptr, workspace, workspace_bytes, (cudaStream_t)stream, "frcnn_roi_align_forward", false,
                              exact != 0);
}

size_t frcnn_roi_align_mean_workspace_bytes(int32_t batch, int32_t num_rois) {
    Workspace ws(nullptr, 0);
    roi_layout(ws, batch > 0 ? batch : 1, num_rois, nullptr);
    ws.take<AlignWeights>(num_rois > 0 ? num_rois : 1);
    return ws.off;
}

int frcnn_roi_align_mean_forward(const float* feat, int32_t B, int32_t C, int32_t H, int32_t W, const float* rois5,
                                 int32_t K, int32_t per_image, int32_t PH, int32_t PW, float scale,
                                 int32_t sampling_ratio, int32_t aligned, float* out, void* workspace,
                                 size_t workspace_bytes, frcnn_stream_t stream_) {
    const char* who = "frcnn_roi_align_mean_forward";
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_roi_common(feat, B, C, H, W, rois5, K, PH, PW, out, who);
    if (rc) return rc;
    if (K == 0) return FRCNN_OK;
    const int pitch = (W % 2 == 0) ? W + 1 : W;  // odd pitch: the rows of a window start on different banks
    const size_t smem = (size_t)2 * ((H * pitch + 3) & ~3) * sizeof(float4);
    if (sampling_ratio <= 0 || PH * sampling_ratio > 32 || PW * sampling_ratio > 32 || H > AW_MAX || W > AW_MAX ||
        smem > 200 * 1024 || B > 65535) {
        set_error("%s: needs a fixed sampling grid with P*sampling_ratio <= 32 and a map of at most %dx%d", who, AW_MAX,
                  AW_MAX);
        return FRCNN_ERR_UNSUPPORTED;
    }
    RoiArgs a;
    memset(&a, 0, sizeof(a));
    a.feat = feat;
    a.rois5 = rois5;
    a.B = B;
    a.C = C;
    a.H = H;
    a.W = W;
    a.K = K;
    a.PH = PH;
    a.PW = PW;
    a.scale = scale;
    a.out = out;
    a.sampling_ratio = sampling_ratio;
    a.aligned = aligned;
    a.pitch = pitch;
    Workspace ws(workspace, workspace_bytes);
    RoiWs w;
    roi_layout(ws, B, K, &w);
    AlignWeights* rec = ws.take<AlignWeights>(K);
    if (!ws.ok()) {
        set_error("%s: workspace too small or misaligned (%zu needed, %zu given)", who, ws.off, workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    if (per_image > 0) {
        FRCNN_CHECK_ARG((int64_t)per_image * B == K, "%s: rois_per_image * batch != num_rois", who);
        a.per_image = per_image;
    } else {
        roi_bucket_kernel<<<1, BUCKET_THREADS, 2 * B * sizeof(int), stream>>>(rois5, K, B, w.perm, w.offs);
        FRCNN_LAUNCH_CHECK();
        roi_fill_dropped_kernel<<<sm_count(), 256, 0, stream>>>(w.perm, w.offs, B, (size_t)C, out, nullptr);
        FRCNN_LAUNCH_CHECK();
        a.perm = w.perm;
        a.offs = w.offs;
    }
    roi_align_weights_kernel<<<cdiv(K, 8), 256, 0, stream>>>(a, rec);
    FRCNN_LAUNCH_CHECK();
    const int slabs = cdiv(C, 4);
    FRCNN_CHECK_ARG(slabs <= 65535, "%s: too many channel slabs", who);
    const int nw = AM_THREADS / 32;
    a.groups = std::max(1, std::min(cdiv(cdiv(K, B), 8 * nw), cdiv(8 * sm_count(), B * slabs)));
    FRCNN_SMEM(roi_align_mean_kernel, smem);
    roi_align_mean_kernel<<<dim3(a.groups, slabs, B), AM_THREADS, smem, stream>>>(a, rec);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_roi_pool_backward(const float* grad_out, const int32_t* argmax, const float* rois5, int32_t K,
                            int32_t B, int32_t C, int32_t H, int32_t W, int32_t PH, int32_t PW, float* grad_in,
                            frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(K >= 0 && B > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0,
                    "frcnn_roi_pool_backward: bad shape");
    if (K == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(grad_out && argmax && rois5 && grad_in, "frcnn_roi_pool_backward: null pointer");
    size_t total = (size_t)K * C * PH * PW;
    static const int bw_direct = env_int("FRCNN_BACKWARD_DIRECT", 0);  // experiments only: 1 = one global atomic per element
    const size_t slab = (size_t)4 * H * W * sizeof(float);
    if (slab <= 160 * 1024 && B <= 65535 && cdiv(C, 4) <= 65535 && !bw_direct) {
        FRCNN_SMEM(roi_pool_backward_slab_kernel, slab);
        roi_pool_backward_slab_kernel<<<dim3(cdiv(C, 4), B), BW_THREADS, slab, (cudaStream_t)stream>>>(
            grad_out, argmax, rois5, K, C, H * W, PH * PW, grad_in);
        FRCNN_LAUNCH_CHECK();
        return FRCNN_OK;
    }
    roi_pool_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad_out, argmax, rois5, total, B, C, H * W, PH * PW, grad_in);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

int frcnn_roi_align_backward(const float* grad_out, const float* rois5, int32_t K, int32_t B, int32_t C, int32_t H,
                             int32_t W, int32_t PH, int32_t PW, float scale, int32_t sampling_ratio,
                             int32_t aligned, float* grad_in, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(K >= 0 && B > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0,
                    "frcnn_roi_align_backward: bad shape");
    if (K == 0) return FRCNN_OK;
    FRCNN_CHECK_ARG(grad_out && rois5 && grad_in, "frcnn_roi_align_backward: null pointer");
    RoiArgs a;
    memset(&a, 0, sizeof(a));
    a.rois5 = rois5;
    a.B = B;
    a.C = C;
    a.H = H;
    a.W = W;
    a.K = K;
    a.PH = PH;
    a.PW = PW;
    a.scale = scale;
    a.sampling_ratio = sampling_ratio;
    a.aligned = aligned;
    size_t total = (size_t)K * C * PH * PW;
    static const int bw_direct = env_int("FRCNN_BACKWARD_DIRECT", 0);  // experiments only
    const size_t slab = (size_t)4 * H * W * sizeof(float);
    // (grids above 8 x 8: 16 shared-memory CAS loops per bin and channel lose against the L2 atomics -- 26.6 vs 23.5 ms
    //  for a cfg2-sized 14 x 14 gradient; 7 x 7: 4.8 vs 6.4 ms)
    if (slab <= 160 * 1024 && PH * PW <= 64 && B <= 65535 && cdiv(C, 4) <= 65535 && !bw_direct) {
        FRCNN_SMEM(roi_align_backward_slab_kernel, slab);
        roi_align_backward_slab_kernel<<<dim3(cdiv(C, 4), B), BW_THREADS, slab, (cudaStream_t)stream>>>(grad_out, a, grad_in);
        FRCNN_LAUNCH_CHECK();
        return FRCNN_OK;
    }
    roi_align_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_out, a,
                                                                                                  grad_in);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

}  // extern "C"
