// proposals.cu -- the RPN proposal layer (nets/rpn.py:36-70 + the per-image loop :129-139 +
// torchvision.ops.nms), batched over images, no host synchronisation anywhere.
//
//   decode_clip_key   one thread per anchor: loc2bbox, clamp, min-size test, score -> sortable key
//   topk_sort         one 8-CTA cluster per image: stable LSD radix sort of (~key), histograms
//                     exchanged over distributed shared memory, warp-match ranking; emits anchor
//                     indices in (score desc, index asc) order + gathered boxes
//   nms_mask          super-block S of sorted candidates: warp-ballot IoU>thr bitmask tiles
//                     (upper triangle only) + suppression by boxes kept in earlier super-blocks
//   nms_scan          one CTA per image: resolves 32 candidates per step, early exit at keep_cap
//   finalize          pad-with-arange / truncate / gather (nets/rpn.py:65-69)
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace frcnn {

// ---------------------------------------------------------------------------------------------
// decode + clip + min-size + key
// ---------------------------------------------------------------------------------------------
struct DecodeArgs {
    const float4* loc;
    const float* score;
    AnchorGen gen;
    int batch, n;
    float xmax, ymax, min_size;
    int score_mode, decoded;
    float4* boxes;
    uint32_t* keys;
    float* fg_out;
};

__device__ __forceinline__ float clamp_torch(float v, float hi) {
    v = v < 0.f ? 0.f : v;  // NaN propagates like torch.clamp
    v = v > hi ? hi : v;
    return v;
}

__global__ void __launch_bounds__(256) decode_clip_key_kernel(DecodeArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // anchor; blockIdx.y = image
    if (i >= a.n) return;
    const int64_t t = (int64_t)blockIdx.y * a.n + i;
    float4 l = __ldg(a.loc + t);
    float4 r;
    if (a.decoded) {
        r = l;
    } else {
        r = decode_box(load_anchor(a.gen, i), l);
    }
    float s;
    if (a.score_mode == 0) {
        s = __ldg(a.score + t);
    } else {
        float2 lg = __ldg(reinterpret_cast<const float2*>(a.score) + t);
        // softmax([l0, l1])[1] = e1 / (e0 + e1) with e = exp(l - max): the larger logit's term is exp(0) = 1
        // exactly, so one expf is enough and the bits are those of the two-exponential form
        // (1 + (big - big) keeps the NaN an infinite logit produces there)
        const bool fg_larger = lg.y >= lg.x;
        const float big = fg_larger ? lg.y : lg.x, small = fg_larger ? lg.x : lg.y;
        const float eb = 1.f + (big - big), es = expf(small - big);
        const float e0 = fg_larger ? es : eb, e1 = fg_larger ? eb : es;
        s = e1 / (e0 + e1);
    }
    r.x = clamp_torch(r.x, a.xmax);
    r.z = clamp_torch(r.z, a.xmax);
    r.y = clamp_torch(r.y, a.ymax);
    r.w = clamp_torch(r.w, a.ymax);
    bool ok = ((r.z - r.x) >= a.min_size) && ((r.w - r.y) >= a.min_size);
    a.boxes[t] = r;
    a.keys[t] = ok ? score_key(s) : 0u;
    if (a.fg_out) a.fg_out[t] = s;
}

// ---------------------------------------------------------------------------------------------
// The same, reading the RPN's 1x1-conv outputs IN PLACE (score_mode 2): loc [B,4A,H,W], logits [B,2A,H,W],
// NCHW as cuDNN writes them.  nets/rpn.py:107-118 permutes both to NHWC and makes them contiguous before the
// proposal layer -- two full passes over [B,54,H,W] that exist only to change the layout.  Here a CTA takes a
// tile of DEC_TILE consecutive pixels: every one of the 6A planes contributes one contiguous run (coalesced along
// W), the tile is transposed through shared memory, and each thread then owns anchors in OUTPUT order
// (i = pixel*A + a), so boxes (16 B per anchor) and keys are written fully coalesced.  Anchor a of a pixel reads
// loc channels 4a..4a+3 and logit channels 2a, 2a+1 -- exactly the elements view(n,-1,4) / view(n,-1,2) of the
// permuted tensors would hand to the NHWC kernel, in the same arithmetic: boxes and keys are bit-identical.
// ---------------------------------------------------------------------------------------------
constexpr int DEC_TILE = 64;
constexpr int DEC_PITCH = DEC_TILE + 1;

__global__ void __launch_bounds__(256) decode_clip_key_nchw_kernel(DecodeArgs a) {
    extern __shared__ float dsm[];  // [6A][DEC_PITCH]: planes 0..4A-1 = loc, 4A..6A-1 = logits
    const int A = a.gen.num_base, HW = a.gen.height * a.gen.width;
    const int p0 = blockIdx.x * DEC_TILE, b = blockIdx.y;
    const int np = min(DEC_TILE, HW - p0);
    const float* loc = reinterpret_cast<const float*>(a.loc) + (size_t)b * 4 * A * HW + p0;
    const float* sc = a.score + (size_t)b * 2 * A * HW + p0;
    for (int e = threadIdx.x; e < 6 * A * DEC_TILE; e += blockDim.x) {
        const int pl = e / DEC_TILE, q = e - pl * DEC_TILE;
        if (q < np) dsm[pl * DEC_PITCH + q] = pl < 4 * A ? __ldg(loc + (size_t)pl * HW + q) : __ldg(sc + (size_t)(pl - 4 * A) * HW + q);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < np * A; j += blockDim.x) {
        const int q = j / A, an = j - q * A;
        const int i = p0 * A + j;  // anchor index within the image
        const int64_t t = (int64_t)b * a.n + i;
        const float* lp = dsm + (an * 4) * DEC_PITCH + q;
        const float4 l = make_float4(lp[0], lp[DEC_PITCH], lp[2 * DEC_PITCH], lp[3 * DEC_PITCH]);
        float4 r = decode_box(load_anchor(a.gen, i), l);
        const float* sp = dsm + (4 * A + an * 2) * DEC_PITCH + q;
        const float2 lg = make_float2(sp[0], sp[DEC_PITCH]);
        const bool fg_larger = lg.y >= lg.x;  // same softmax form as decode_clip_key_kernel
        const float big = fg_larger ? lg.y : lg.x, small = fg_larger ? lg.x : lg.y;
        const float eb = 1.f + (big - big), es = expf(small - big);
        const float e0 = fg_larger ? es : eb, e1 = fg_larger ? eb : es;
        const float s = e1 / (e0 + e1);
        r.x = clamp_torch(r.x, a.xmax);
        r.z = clamp_torch(r.z, a.xmax);
        r.y = clamp_torch(r.y, a.ymax);
        r.w = clamp_torch(r.w, a.ymax);
        const bool ok = ((r.z - r.x) >= a.min_size) && ((r.w - r.y) >= a.min_size);
        a.boxes[t] = r;
        a.keys[t] = ok ? score_key(s) : 0u;
        if (a.fg_out) a.fg_out[t] = s;
    }
}

__global__ void scores_to_keys_kernel(const float* __restrict__ s, int n, uint32_t* __restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = score_key(__ldg(s + i));
}

// ---------------------------------------------------------------------------------------------
// per-image stable radix sort (descending key, ascending index on ties)
//
// One thread-block CLUSTER of 8 CTAs per image (16 images -> 128 SMs instead of 16).  CTA r owns the
// r-th contiguous eighth of the keys in their current order.  Per 8-bit pass:
//   1. every CTA histograms its segment (warp-match aggregated shared-memory atomics);
//   2. cluster barrier; each CTA reads the eight histograms through distributed shared memory and
//      derives, per digit, the global base + the count in lower-ranked CTAs = its own running offset;
//   3. stable scatter of the segment into the other global ping-pong buffer (chunks in order, ranks
//      from __match_any_sync + a scan of per-warp digit counts);
//   4. cluster barrier (release/acquire: the scatter is visible to the whole cluster).
// Sorting ~key ascending gives key descending; LSD stability gives index-ascending ties; filtered
// anchors (key 0) sort last.  Passes whose digit is identical for all keys are skipped.
// ---------------------------------------------------------------------------------------------
constexpr int TK_CL = 8;
constexpr int TK_THREADS = 512;
constexpr int TK_WARPS = TK_THREADS / 32;
constexpr int TK_E = 4;                          // keys per thread per round
constexpr int TK_ROUND = TK_E * TK_THREADS;      // keys per CTA per round (2048)
constexpr size_t TK_SMEM = sizeof(uint32_t) * TK_E * TK_WARPS * 256;  // per-(chunk,warp) digit counters

__global__ void __cluster_dims__(TK_CL, 1, 1) __launch_bounds__(TK_THREADS)
topk_sort_kernel(const uint32_t* __restrict__ keys_all, const float4* __restrict__ boxes_all, int n,
                 int k_cap, uint32_t* __restrict__ ws_keys, uint32_t* __restrict__ ws_idx,
                 int* __restrict__ order_all, int* __restrict__ n_sel_all,
                 float4* __restrict__ sorted_all) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint32_t cnt[];  // [TK_E][TK_WARPS][256]: digit counts, then exclusive prefixes
    __shared__ uint32_t hist[256];     // this CTA's digit totals (read by the whole cluster)
    __shared__ uint32_t offs[256];     // running destination offsets of this CTA
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t s_nzero;       // filtered (key == 0) anchors in this CTA's segment

    const int crank = (int)cluster.block_rank();
    const int b = blockIdx.x / TK_CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* keys = keys_all + (size_t)b * n;
    uint32_t* bufK0 = ws_keys + (size_t)b * 2 * n;
    uint32_t* bufI0 = ws_idx + (size_t)b * 2 * n;
    const uint32_t lt = lanemask_lt();
    const int seg = (n + TK_CL - 1) / TK_CL;
    const int lo = min(crank * seg, n), hi = min(lo + seg, n);
    const int rounds = (seg + TK_ROUND - 1) / TK_ROUND;  // same in every CTA of the cluster

    if (tid == 0) s_nzero = 0;
    int cur = -1;  // -1: data still in `keys` (identity permutation)

    uint32_t k[TK_E], id[TK_E], rank[TK_E];
    bool valid[TK_E];
    // one round = TK_E chunks of TK_THREADS consecutive keys; order inside the CTA = (chunk, warp, lane)
    auto load_round = [&](int r0, const uint32_t* srcK, const uint32_t* srcI) {
#pragma unroll
        for (int e = 0; e < TK_E; ++e) {
            const int i = r0 + e * TK_THREADS + tid;
            valid[e] = i < hi;
            k[e] = 0;
            id[e] = (uint32_t)i;
            if (valid[e]) {
                if (srcK) {  // written by other SMs in the previous pass: read through L2
                    k[e] = __ldcg(srcK + i);
                    id[e] = __ldcg(srcI + i);
                } else {
                    k[e] = ~__ldg(keys + i);
                }
            }
        }
    };
    // per-(chunk,warp) digit counts into cnt, per-key rank among equal digits of its warp
    auto count_round = [&](int shift) {
        for (int j = lane; j < TK_E * 256; j += 32) cnt[((j >> 8) * TK_WARPS + warp) * 256 + (j & 255)] = 0;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < TK_E; ++e) {
            const uint32_t d = valid[e] ? ((k[e] >> shift) & 255u) : 256u;
            const uint32_t m = __match_any_sync(0xFFFFFFFFu, d);
            rank[e] = __popc(m & lt);
            if (valid[e] && rank[e] == 0) cnt[(e * TK_WARPS + warp) * 256 + d] = __popc(m);
        }
    };
    // thread d < 256: turn the counts of digit d into exclusive prefixes, return their sum
    auto scan_round = [&]() -> uint32_t {
        uint32_t run = 0;
#pragma unroll 16
        for (int j = 0; j < TK_E * TK_WARPS; ++j) {
            const uint32_t c = cnt[j * 256 + tid];
            cnt[j * 256 + tid] = run;
            run += c;
        }
        return run;
    };
    auto scatter_round = [&](int shift, uint32_t* dstK, uint32_t* dstI) {
#pragma unroll
        for (int e = 0; e < TK_E; ++e) {
            if (valid[e]) {
                const uint32_t d = (k[e] >> shift) & 255u;
                const uint32_t pos = offs[d] + cnt[(e * TK_WARPS + warp) * 256 + d] + rank[e];
                dstK[pos] = k[e];
                dstI[pos] = id[e];
            }
        }
    };

    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        const uint32_t* srcK = cur < 0 ? nullptr : bufK0 + (size_t)cur * n;
        const uint32_t* srcI = cur < 0 ? nullptr : bufI0 + (size_t)cur * n;
        // ---- this CTA's digit histogram ----
        if (rounds == 1) {
            load_round(lo, srcK, srcI);
            count_round(shift);
            if (pass == 0) {
#pragma unroll
                for (int e = 0; e < TK_E; ++e) {
                    uint32_t z = __ballot_sync(0xFFFFFFFFu, valid[e] && k[e] == 0xFFFFFFFFu);
                    if (lane == 0 && z) atomicAdd(&s_nzero, (uint32_t)__popc(z));
                }
            }
            __syncthreads();
            if (tid < 256) hist[tid] = scan_round();
        } else {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for (int r = 0; r < rounds; ++r) {
                load_round(lo + r * TK_ROUND, srcK, srcI);
#pragma unroll
                for (int e = 0; e < TK_E; ++e) {
                    const uint32_t d = valid[e] ? ((k[e] >> shift) & 255u) : 256u;
                    const uint32_t m = __match_any_sync(0xFFFFFFFFu, d);
                    if (valid[e] && (m & lt) == 0) atomicAdd(&hist[d], (uint32_t)__popc(m));
                    if (pass == 0) {
                        uint32_t z = __ballot_sync(0xFFFFFFFFu, valid[e] && k[e] == 0xFFFFFFFFu);
                        if (lane == 0 && z) atomicAdd(&s_nzero, (uint32_t)__popc(z));
                    }
                }
            }
        }
        cluster.sync();  // all eight histograms complete
        uint32_t total = 0, pre = 0;
        if (tid < 256) {
#pragma unroll
            for (int c = 0; c < TK_CL; ++c) {
                uint32_t v = cluster.map_shared_rank(hist, c)[tid];
                pre += c < crank ? v : 0u;
                total += v;
            }
        }
        // identical in every CTA of the cluster, so the whole cluster takes the same branch
        int trivial = __syncthreads_or(tid < 256 && total == (uint32_t)n);
        if (trivial) {
            cluster.sync();  // remote reads of hist are done before the next pass rewrites it
            continue;
        }
        uint32_t incl = total;
        if (tid < 256) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += warp_tot[w];
            offs[tid] = before + incl - total + pre;
        }
        __syncthreads();
        // ---- stable scatter into the other ping-pong buffer ----
        const int dst = cur < 0 ? 0 : (cur ^ 1);
        uint32_t* dstK = bufK0 + (size_t)dst * n;
        uint32_t* dstI = bufI0 + (size_t)dst * n;
        if (rounds == 1) {
            scatter_round(shift, dstK, dstI);  // keys, ranks and prefixes are still live
        } else {
            for (int r = 0; r < rounds; ++r) {
                load_round(lo + r * TK_ROUND, srcK, srcI);
                __syncthreads();  // previous round's scatter has read cnt / offs
                count_round(shift);
                __syncthreads();
                uint32_t run = 0;
                if (tid < 256) run = scan_round();
                __syncthreads();
                scatter_round(shift, dstK, dstI);
                __syncthreads();
                if (tid < 256) offs[tid] += run;
            }
        }
        cluster.sync();  // scatter visible cluster-wide; hist may be rewritten
        cur = dst;
    }
    uint32_t nzero = 0;
#pragma unroll
    for (int c = 0; c < TK_CL; ++c) nzero += *cluster.map_shared_rank(&s_nzero, c);
    const int n_valid = n - (int)nzero;
    const int n_sel = n_valid < k_cap ? n_valid : k_cap;
    if (crank == 0 && tid == 0) n_sel_all[b] = n_sel;
    int* order = order_all + (size_t)b * k_cap;
    const uint32_t* finI = cur < 0 ? nullptr : bufI0 + (size_t)cur * n;
    for (int j = crank * TK_THREADS + tid; j < k_cap; j += TK_CL * TK_THREADS) {
        int idx = -1;
        if (j < n_sel) idx = finI ? (int)__ldcg(finI + j) : j;
        order[j] = idx;
        if (sorted_all) {
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx >= 0) bx = __ldg(boxes_all + (size_t)b * n + idx);
            sorted_all[(size_t)b * k_cap + j] = bx;
        }
    }
    cluster.sync();  // nobody exits while its shared memory may still be read remotely
}

// ---------------------------------------------------------------------------------------------
// The same sort with the keys resident in DISTRIBUTED SHARED MEMORY (used whenever a segment fits):
// CTA r holds positions [r*seg, (r+1)*seg) of the current ordering as (key, index) pairs in its own
// shared memory, double-buffered; the scatter of a pass writes each pair straight into the shared
// memory of the CTA that owns its destination position (st.shared::cluster through
// cluster.map_shared_rank), so no pass touches global memory and the uncoalesced 4-byte global
// scatter of the global-memory variant (its dominant cost, measured) disappears.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
    // lanes holding the same 9-bit digit (8 bits + "invalid" bit), by 9 ballots; __match_any_sync
    // measured ~900 cycles per call here
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int bit = 0; bit < 9; ++bit) {
        const uint32_t v = __ballot_sync(0xFFFFFFFFu, (d >> bit) & 1u);
        m &= ((d >> bit) & 1u) ? v : ~v;
    }
    return m;
}

// CL = CTAs per cluster = per image (4 or 8; the launch sets the cluster dimension, see run_topk).
// E = keys per thread per round (4; 9 where that makes a 4 608-key segment a single round, see run_topk).
template <int CL, int E = TK_E>
__global__ void __launch_bounds__(TK_THREADS)
topk_sort_dsmem_kernel(const uint32_t* __restrict__ keys_all, const float4* __restrict__ boxes_all, int n,
                       int k_cap, int seg, int presel, int* __restrict__ order_all, int* __restrict__ n_sel_all,
                       float4* __restrict__ sorted_all) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) uint32_t dyn[];
    uint32_t* cnt = dyn;                                                     // [E][TK_WARPS][256]
    uint2* buf0 = reinterpret_cast<uint2*>(dyn + E * TK_WARPS * 256);     // [2][seg] (key, index)
    __shared__ uint32_t hist[256];
    __shared__ uint32_t offs[256];
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t s_nzero;
    __shared__ uint32_t s_wsum[TK_WARPS];
    __shared__ uint32_t s_sel[3];  // pre-selection: threshold digit, keys at or above it, this CTA's share of them

    const int crank = (int)cluster.block_rank();
    const int b = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* keys = keys_all + (size_t)b * n;
    const uint32_t lt = lanemask_lt();
    int lo = min(crank * seg, n), hi = min(lo + seg, n);
    int cnt_local = hi - lo;
    int rounds = (seg + (E * TK_THREADS) - 1) / (E * TK_THREADS);
    int n_eff = n, seg_eff = seg;  // what is being sorted: everything, or the pre-selected keys (below)
    // rows of TK_THREADS keys a round really has (a pre-selected segment is often a single row): the per-row work of
    // counting, scanning and scattering is skipped for the others
    int e_lim = rounds == 1 ? (seg + TK_THREADS - 1) / TK_THREADS : E;

    // initial ordering: key = ~sortable key (ascending sort), index = anchor id
    if (tid == 0) s_nzero = 0;
    __syncthreads();
    {
        uint32_t nz = 0;
        for (int i = tid; i < cnt_local; i += TK_THREADS) {
            const uint32_t k = ~__ldg(keys + lo + i);
            buf0[i] = make_uint2(k, (uint32_t)(lo + i));
            nz += k == 0xFFFFFFFFu;
        }
        nz = __reduce_add_sync(0xFFFFFFFFu, nz);
        if (lane == 0 && nz) atomicAdd(&s_nzero, nz);
    }
    __syncthreads();
    int cur = 0;

    // ---- pre-selection (k_cap <= n / 2) --------------------------------------------------------------------
    // Only the k_cap best keys are wanted in order, so sorting all n is mostly wasted work (3 000 of 12 996 or of
    // 22 500 on the inference configurations).  One histogram of the top 12 bits over the whole cluster finds the
    // digit D that holds the k_cap-th key; the M >= k_cap keys with digit <= D are compacted IN INDEX ORDER across
    // the cluster (so the stable LSD passes below still break ties by index) and only they are sorted.  The first
    // k_cap positions are exactly those of the full sort.  A distribution that leaves M > 3n/4 skips the compaction.
    if (presel) {
        uint32_t* h12 = cnt;  // [4096], aliases the per-warp digit counters (not in use yet)
        uint32_t* rowcnt = cnt + 4096;  // [rows][TK_WARPS] selected keys per (row of TK_THREADS keys, warp)
        for (int j = tid; j < 4096; j += TK_THREADS) h12[j] = 0;
        __syncthreads();
        for (int i = tid; i < cnt_local; i += TK_THREADS) atomicAdd(&h12[buf0[i].x >> 20], 1u);
        cluster.sync();
        // every CTA: totals of bins 8 tid .. 8 tid + 7 over the cluster, block scan, the thread whose bins hold the
        // k_cap-th key publishes (D, M)
        uint32_t tot[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tot[j] = 0;
#pragma unroll
        for (int c = 0; c < CL; ++c) {
            const uint4* rh = reinterpret_cast<const uint4*>(cluster.map_shared_rank(h12, c));
            const uint4 u = rh[2 * tid], v = rh[2 * tid + 1];
            tot[0] += u.x; tot[1] += u.y; tot[2] += u.z; tot[3] += u.w;
            tot[4] += v.x; tot[5] += v.y; tot[6] += v.z; tot[7] += v.w;
        }
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += tot[j];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t excl = incl - sum;
        for (int w = 0; w < warp; ++w) excl += s_wsum[w];
        if (excl < (uint32_t)k_cap && (uint32_t)k_cap <= excl + sum) {
            uint32_t run = excl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (run < (uint32_t)k_cap && (uint32_t)k_cap <= run + tot[j]) {
                    s_sel[0] = (uint32_t)(8 * tid + j);
                    s_sel[1] = run + tot[j];
                }
                run += tot[j];
            }
        }
        __syncthreads();
        const uint32_t D = s_sel[0];
        const int M = (int)s_sel[1];
        if (4ll * M <= 3ll * n) {  // cluster-uniform: every CTA derived (D, M) from the same totals
            const int nrows = (cnt_local + TK_THREADS - 1) / TK_THREADS;
            for (int row = 0; row < nrows; ++row) {
                const int i = row * TK_THREADS + tid;
                const bool sel = i < cnt_local && (buf0[i].x >> 20) <= D;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, sel);
                if (lane == 0) rowcnt[row * TK_WARPS + warp] = __popc(bal);
            }
            __syncthreads();
            {  // exclusive scan of the (row, warp) counts in index order; nrows * TK_WARPS <= TK_THREADS (host)
                const uint32_t v = tid < nrows * TK_WARPS ? rowcnt[tid] : 0u;
                uint32_t inc2 = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc2, o);
                    if (lane >= o) inc2 += t;
                }
                __syncthreads();  // s_wsum was read above by every thread
                if (lane == 31) s_wsum[warp] = inc2;
                __syncthreads();
                uint32_t ex2 = inc2 - v;
                for (int w = 0; w < warp; ++w) ex2 += s_wsum[w];
                if (tid < nrows * TK_WARPS) rowcnt[tid] = ex2;
                if (tid == TK_THREADS - 1) s_sel[2] = ex2 + v;
            }
            cluster.sync();  // shares of every CTA visible; nobody reads a remote h12 any more
            uint32_t base = 0;
#pragma unroll
            for (int c = 0; c < CL; ++c) {
                const uint32_t v = cluster.map_shared_rank(s_sel, c)[2];
                base += c < crank ? v : 0u;
            }
            seg_eff = (M + CL - 1) / CL;
            for (int row = 0; row < nrows; ++row) {
                const int i = row * TK_THREADS + tid;
                uint2 kv0 = make_uint2(0u, 0u);
                if (i < cnt_local) kv0 = buf0[i];
                const bool sel = i < cnt_local && (kv0.x >> 20) <= D;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, sel);
                if (sel) {
                    const uint32_t pos = base + rowcnt[row * TK_WARPS + warp] + __popc(bal & lt);
                    uint32_t dc = 0;
#pragma unroll
                    for (int c = 1; c < CL; ++c) dc += pos >= (uint32_t)(c * seg_eff);
                    uint2* remote = cluster.map_shared_rank(buf0 + (size_t)seg, dc);
                    remote[pos - dc * seg_eff] = kv0;
                }
            }
            cluster.sync();  // compacted keys in place
            cur = 1;
            n_eff = M;
            lo = min(crank * seg_eff, M);
            hi = min(lo + seg_eff, M);
            cnt_local = hi - lo;
            rounds = (seg_eff + (E * TK_THREADS) - 1) / (E * TK_THREADS);
            if (rounds == 1) e_lim = (seg_eff + TK_THREADS - 1) / TK_THREADS;
        } else {
            cluster.sync();  // the counters h12 aliases are rewritten below: wait for the remote readers
        }
    }

    uint2 kv[E];
    uint32_t rank[E];
    bool valid[E];
    auto load_round = [&](int r, const uint2* src) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = r * (E * TK_THREADS) + e * TK_THREADS + tid;
            valid[e] = i < cnt_local;
            kv[e] = valid[e] ? src[i] : make_uint2(0u, 0u);
        }
    };
    auto count_round = [&](int shift) {
        for (int j = lane; j < e_lim * 256; j += 32) cnt[((j >> 8) * TK_WARPS + warp) * 256 + (j & 255)] = 0;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (e >= e_lim) break;
            const uint32_t d = valid[e] ? ((kv[e].x >> shift) & 255u) : 256u;
            const uint32_t m = match_digit(d);
            rank[e] = __popc(m & lt);
            if (valid[e] && rank[e] == 0) cnt[(e * TK_WARPS + warp) * 256 + d] = __popc(m);
        }
    };
    auto scan_round = [&]() -> uint32_t {  // thread d < 256: counts of digit d -> exclusive prefixes
        uint32_t run = 0;
#pragma unroll
        for (int j0 = 0; j0 < E * TK_WARPS; j0 += 16) {
            if (j0 >= e_lim * TK_WARPS) break;
            uint32_t c[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = cnt[(j0 + j) * 256 + tid];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                cnt[(j0 + j) * 256 + tid] = run;
                run += c[j];
            }
        }
        return run;
    };
    auto scatter_round = [&](int shift, int dst) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (e >= e_lim) break;
            if (valid[e]) {
                const uint32_t d = (kv[e].x >> shift) & 255u;
                const uint32_t pos = offs[d] + cnt[(e * TK_WARPS + warp) * 256 + d] + rank[e];
                uint32_t dc = 0;  // owner of position pos
#pragma unroll
                for (int c = 1; c < CL; ++c) dc += pos >= (uint32_t)(c * seg_eff);
                uint2* remote = cluster.map_shared_rank(buf0 + (size_t)dst * seg, dc);
                remote[pos - dc * seg_eff] = kv[e];
            }
        }
    };

    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        const uint2* src = buf0 + (size_t)cur * seg;
        if (rounds == 1) {
            load_round(0, src);
            count_round(shift);
            __syncthreads();
            if (tid < 256) hist[tid] = scan_round();
        } else {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for (int r = 0; r < rounds; ++r) {
                load_round(r, src);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint32_t d = valid[e] ? ((kv[e].x >> shift) & 255u) : 256u;
                    const uint32_t m = match_digit(d);
                    if (valid[e] && (m & lt) == 0) atomicAdd(&hist[d], (uint32_t)__popc(m));
                }
            }
        }
        cluster.sync();  // all eight histograms complete
        uint32_t total = 0, pre = 0;
        if (tid < 256) {
#pragma unroll
            for (int c = 0; c < CL; ++c) {
                uint32_t v = cluster.map_shared_rank(hist, c)[tid];
                pre += c < crank ? v : 0u;
                total += v;
            }
        }
        int trivial = __syncthreads_or(tid < 256 && total == (uint32_t)n_eff);
        if (trivial) {
            cluster.sync();
            continue;
        }
        uint32_t incl = total;
        if (tid < 256) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += warp_tot[w];
            offs[tid] = before + incl - total + pre;
        }
        __syncthreads();
        const int dst = cur ^ 1;
        if (rounds == 1) {
            scatter_round(shift, dst);
        } else {
            for (int r = 0; r < rounds; ++r) {
                load_round(r, src);
                __syncthreads();
                count_round(shift);
                __syncthreads();
                uint32_t run = 0;
                if (tid < 256) run = scan_round();
                __syncthreads();
                scatter_round(shift, dst);
                __syncthreads();
                if (tid < 256) offs[tid] += run;
            }
        }
        cluster.sync();  // remote shared-memory writes visible; hist may be rewritten
        cur = dst;
    }
    uint32_t nzero = 0;
#pragma unroll
    for (int c = 0; c < CL; ++c) nzero += *cluster.map_shared_rank(&s_nzero, c);
    const int n_valid = n - (int)nzero;
    const int n_sel = n_valid < k_cap ? n_valid : k_cap;
    if (crank == 0 && tid == 0) n_sel_all[b] = n_sel;
    int* order = order_all + (size_t)b * k_cap;
    const uint2* fin = buf0 + (size_t)cur * seg;
    for (int i = tid; i < cnt_local; i += TK_THREADS) {
        const int j = lo + i;
        if (j >= k_cap) break;
        const int idx = j < n_sel ? (int)fin[i].y : -1;
        order[j] = idx;
        if (sorted_all) {
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx >= 0) bx = __ldg(boxes_all + (size_t)b * n + idx);
            sorted_all[(size_t)b * k_cap + j] = bx;
        }
    }
    // rows past n (k_cap > n never happens: k_cap <= n) need nothing; keep every CTA alive until all
    // remote reads of its shared memory are done
    cluster.sync();
}

// ---------------------------------------------------------------------------------------------
// NMS over score-sorted boxes
// ---------------------------------------------------------------------------------------------
struct NmsState {
    int n_kept;
    int done;
    int c_next;  // first sorted position not yet scanned (super-block boundaries are per image, see nms_block_len)
    int pad;
};

constexpr int NMS_RDY_PREV = 128;    // index of the kept-list tile counter (row blocks: 8192 / 64 = 128 at most)
constexpr int NMS_RDY_STRIDE = 132;  // counters per (launch, image)
constexpr int NMS_CB = 256;   // columns per CTA tile (8 warps x 32 lanes)
constexpr int NMS_RB = 64;    // rows staged per in-block tile
constexpr int NMS_KC = 256;   // kept boxes staged per prev tile

struct NmsArgs {
    const float4* boxes;   // [B,row_stride]
    const int* n_sel;      // [B]
    int row_stride, keep_cap;
    int S;                 // mask row stride = largest super-block
    int len;               // this launch's super-block: at most `len` sorted positions from the image's c_next
    int cover_after;       // sum of `len` over the launches that follow this one
    int adaptive;          // later super-blocks may be cut to what the image still needs (nms_block_len)
    float thr;
    uint32_t* mask;        // [B][S/32][S]
    uint32_t* removed;     // [B][S/32]
    float4* kept_box;      // [B][keep_cap]
    NmsState* state;       // [B] read by this launch (and updated in place by nms_tail_kernel, whose tiles and scan
                           // are separated by cluster barriers)
    NmsState* state_out;   // [B] written by the scan.  nms_block_kernel: the OTHER buffer of the pair, because its scan
                           // runs beside tiles that may not have read the image's state yet
    unsigned* tile_count;  // [B] CTAs of the current launch that finished their mask tile
    int* keep;             // [B][keep_cap]
    int* n_keep;           // [B]
    int tri_tiles;
    int prev_tiles;        // tiles of this launch that suppress against the kept list of earlier super-blocks
    int batch;
    int* ready;            // [B][NMS_RDY_STRIDE] tiles finished per 64-row block (+ the kept-list tiles), this launch
    // the proposal layer's last step (nets/rpn.py:65-69), done by nms_tail_kernel when fin_rois is set
    const int* fin_order;  // [B][row_stride] sorted position -> anchor index
    float4* fin_rois;      // [B][keep_cap]
    int* fin_src;          // [B][keep_cap] or null
    int* fin_n_keep;       // [B] or null
    int* fin_status;       // [B]
};

struct NmsFinalize {
    const int* order;
    float* rois;
    int* roi_src;
    int* n_keep;
    int* status;
    bool fused;  // out: the NMS launches did it (otherwise the caller launches finalize_kernel)
};

// Length of an image's super-block in this launch.  The first block is sized by the host for keep_cap; a
// later one only has to supply the keep_cap - n_kept boxes still missing, so it is cut to twice what the
// keep rate so far predicts for them (a 2048-wide mask of which the scan reads 300 columns is the single
// largest item of the proposal-stress configuration) -- but never below what the remaining launches need
// this one to cover, so every candidate is still seen in the worst case.  Same value in every CTA of the
// image: the state is only written by the scan, after all tiles of the launch.
__device__ __forceinline__ int nms_block_len(const NmsArgs& a, const NmsState& st, int n) {
    int L = a.len;
    if (a.adaptive && st.c_next > 0) {
        const int need = a.keep_cap - st.n_kept;
        const long long want = 2ll * need * st.c_next / max(st.n_kept, 1);
        const int must = n - st.c_next - a.cover_after;
        const long long lo = max((long long)must, max(want, (long long)NMS_CB));
        if (lo < L) L = (int)((lo + NMS_CB - 1) / NMS_CB * NMS_CB);
    }
    return L;
}

// one mask tile (in-block triangle tile, or suppression of a column block by a chunk of the kept list)
// Returns the readiness counter the tile belongs to: its 64-row block, or NMS_RDY_PREV.
__device__ __forceinline__ int nms_mask_tile(const NmsArgs& a, int b, const NmsState& st, float4* srow,
                                             float* sarea, int t, int tiles_total) {
    // in-block tiles in ROW-block-major order (the scan consumes the mask row block by row block, see
    // nms_block_kernel): t -> (rb, cb), rb < 4*(cb+1); row blocks 4q..4q+3 have ncb - q column blocks each
    int rb = NMS_RDY_PREV, cb = 0;
    if (t < a.tri_tiles) {
        const int ncb_l = a.len / NMS_CB;
        int q = 0, tt = t;
        while (tt >= 4 * (ncb_l - q)) {
            tt -= 4 * (ncb_l - q);
            ++q;
        }
        const int per = ncb_l - q;
        rb = 4 * q + tt / per;
        cb = q + tt % per;
    }
    const int n = a.n_sel[b];
    const int c0 = st.c_next;
    if (c0 >= n) return rb;
    const int c1 = min(c0 + nms_block_len(a, st, n), n);
    const float4* boxes = a.boxes + (size_t)b * a.row_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (t < a.tri_tiles) {
        int col0 = c0 + cb * NMS_CB, row0 = c0 + rb * NMS_RB;
        if (col0 >= c1 || row0 >= c1) return rb;
        if (threadIdx.x < NMS_RB) {
            int r = row0 + threadIdx.x;
            float4 v = r < c1 ? __ldg(boxes + r) : make_float4(0.f, 0.f, 0.f, 0.f);
            srow[threadIdx.x] = v;
            sarea[threadIdx.x] = box_area(v);
        }
        __syncthreads();
        int c = col0 + warp * 32 + lane;
        bool cvalid = c < c1;
        float4 cbx = cvalid ? __ldg(boxes + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ca = box_area(cbx);
        int cw = cb * (NMS_CB / 32) + warp;
        if (col0 + warp * 32 >= c1) return rb;
        uint32_t* mrow = a.mask + ((size_t)b * (a.S / 32) + cw) * a.S + (row0 - c0);
        uint32_t mine = 0;
        if (row0 + NMS_RB <= c1) {  // every row of the tile exists (all but the last row block): no per-row test
#pragma unroll 8
            for (int r = 0; r < NMS_RB; ++r) {
                const bool p = nms_suppresses(srow[r], sarea[r], cbx, ca, a.thr) && cvalid;
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, p);
                if ((r & 31) == lane) mine = w;
                if ((r & 31) == 31) mrow[(r - 31) + lane] = mine;
            }
        } else {
#pragma unroll 4
            for (int r = 0; r < NMS_RB; ++r) {
                bool p = cvalid && (row0 + r < c1) && nms_suppresses(srow[r], sarea[r], cbx, ca, a.thr);
                uint32_t w = __ballot_sync(0xFFFFFFFFu, p);
                if ((r & 31) == lane) mine = w;
                if ((r & 31) == 31) mrow[(r - 31) + lane] = mine;
            }
        }
    } else {
        // suppression by boxes kept in earlier super-blocks
        // The launch has (len / NMS_CB) * ceil(keep_cap / NMS_KC) of these tiles per image.  They are dealt over
        // the column blocks the image really has (a cut super-block has fewer): the kept list is then split into
        // more, shorter chunks, and a chunk is a serial loop of exact IoUs per column thread -- the latency of
        // the launch when only a few column blocks are live.
        t -= a.tri_tiles;
        const int ntile = tiles_total - a.tri_tiles;
        const int ncb = (c1 - c0 + NMS_CB - 1) / NMS_CB;
        const int nchunk = ntile / ncb;  // >= ceil(keep_cap / NMS_KC), so a chunk has at most NMS_KC rows
        const int cb = t % ncb, kc = t / ncb;
        if (kc >= nchunk) return rb;
        const int per = (st.n_kept + nchunk - 1) / nchunk;
        const int k0 = kc * per;
        if (k0 >= st.n_kept) return rb;
        const int col0 = c0 + cb * NMS_CB;
        const int kn = min(per, st.n_kept - k0);
        if (threadIdx.x < kn) {
            // L2 read: inside nms_tail_kernel the list grows between two reads of the same line
            float4 v = __ldcg(a.kept_box + (size_t)b * a.keep_cap + k0 + threadIdx.x);
            srow[threadIdx.x] = v;
            sarea[threadIdx.x] = box_area(v);
        }
        __syncthreads();
        int c = col0 + warp * 32 + lane;
        bool cvalid = c < c1;
        float4 cbx = cvalid ? __ldg(boxes + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ca = box_area(cbx);
        bool sup = false;
        for (int r = 0; r < kn; ++r) {
            sup = sup || nms_suppresses(srow[r], sarea[r], cbx, ca, a.thr);
            if ((r & 15) == 15 && __all_sync(0xFFFFFFFFu, sup || !cvalid)) break;
        }
        uint32_t w = __ballot_sync(0xFFFFFFFFu, sup && cvalid);
        if (lane == 0 && w) atomicOr(a.removed + (size_t)b * (a.S / 32) + cb * (NMS_CB / 32) + warp, w);
    }
    return rb;
}

constexpr int NMS_MAX_WORDS = 256;  // S <= 8192

// One CTA per image, one thread per 32-candidate column word t of the super-block.  Step u resolves
// the 32 candidates of block u in warp 0 (serial over the 32 bits, diagonal mask word per lane), then
// every thread t > u ORs the mask rows of the newly kept candidates into its removed word R[t].
//
// The mask rows of step u (for every column word t >= u: 32 words, 128 contiguous bytes of the
// column-word-major mask) do not depend on the outcome of earlier steps, so the whole CTA streams them
// from L2 into a shared-memory ring with cp.async, NMS_RING_DEPTH steps ahead: a step then costs the
// resolve chain plus two barriers instead of an L2 round trip (measured on the 2048-wide super-blocks of
// the proposal-stress configuration: ~1 us per step with register prefetch one step ahead).
constexpr int NMS_RING_WORDS = 64;   // ring path: super-blocks up to 2048 candidates
constexpr int NMS_RING_DEPTH = 3;    // steps in flight
constexpr int NMS_RING_STAGES = NMS_RING_DEPTH + 1;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// named barriers (ids 1..15; 0 is __syncthreads): `count` threads in all, arriving or waiting
__device__ __forceinline__ void nms_bar_sync(int id, int count) {
    asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nms_bar_arrive(int id, int count) {
    asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Greedy resolve of 32 consecutive candidates: `open` = candidates no earlier box removed, D (per lane) =
// which later candidates of the word lane's candidate suppresses (strictly upper triangular).  Candidate i
// is kept iff it is open and no KEPT j < i suppresses it.  The serial form is a 32-long dependent chain
// (~25 cycles per bit through a predicate); this is its fixed-point form: start from "every open candidate
// kept", let the kept ones vote their rows together (one REDUX.OR), drop what they suppress, repeat.  Bit i
// depends only on bits < i, so after t rounds the first t bits are final and an unchanged word is the
// greedy answer; typical suppression chains are 2-4 deep.
__device__ __forceinline__ uint32_t nms_resolve_word(uint32_t open, uint32_t D, int lane) {
    uint32_t kept = open;
    for (;;) {
        const uint32_t hit = __reduce_or_sync(0xFFFFFFFFu, ((kept >> lane) & 1u) ? D : 0u);
        const uint32_t next = open & ~hit;
        if (next == kept) return kept;
        kept = next;
    }
}

// Boxes of the candidates kept in this super-block, for the "suppress against the kept list" tiles of the
// next ones.  Done once after the scan: a load -> store of the box inside a step made every step wait for
// an L2 round trip (the store needs the loaded registers before the warp can reach the step's barrier).
__device__ __forceinline__ void nms_copy_kept_boxes(const NmsArgs& a, int b, const float4* boxes, int first,
                                                    const int* s_nkept) {
    __syncthreads();  // keep[] of this scan written (same CTA), *s_nkept final
    const int last = *s_nkept;
    for (int i = first + (int)threadIdx.x; i < last; i += (int)blockDim.x)
        a.kept_box[(size_t)b * a.keep_cap + i] = __ldg(boxes + __ldcg(a.keep + (size_t)b * a.keep_cap + i));
    __syncthreads();  // thread 0 reads *s_nkept again below
}

// `ready` (nms_block_kernel only): the launch's per-row-block tile counters of this image.  The scan runs BESIDE the
// mask tiles of its launch and waits, row block by row block, for the part of the mask it is about to read: every
// thread polls for itself (an acquire load, then its own reads), remembering how far it has seen the mask complete.
__device__ __forceinline__ int nms_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void nms_wait_counter(const int* ready, int idx, int expected) {
    if (ready == nullptr) return;
    while (nms_ld_acquire(ready + idx) < expected) __nanosleep(40);
}
// row blocks [0, known) are complete as far as this thread knows; rb advances by at most one per call site step
__device__ __forceinline__ void nms_wait_rows(const int* ready, int rb, int ncb_launch, int& known) {
    if (ready == nullptr || rb < known) return;
    for (int r = known; r <= rb; ++r) nms_wait_counter(ready, r, ncb_launch - (r >> 2));
    known = rb + 1;
}

__device__ __forceinline__ void nms_scan_image(const NmsArgs& a, int b, const NmsState& st, uint32_t* R,
                                               uint32_t* ring, const int* ready = nullptr) {
    __shared__ uint32_t s_kb;
    __shared__ int s_nkept, s_done;
    const int n = a.n_sel[b];
    const int c0 = st.c_next;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ncb_launch = a.len / NMS_CB;
    int known = 0;
    if (c0 >= n) {
        if (t == 0) {
            NmsState o = st;
            o.done = 1;
            a.state_out[b] = o;
            a.n_keep[b] = st.n_kept;
        }
        return;
    }
    const int c1 = min(c0 + nms_block_len(a, st, n), n);
    const int ncol = c1 - c0, nw = (ncol + 31) / 32;
    const float4* boxes = a.boxes + (size_t)b * a.row_stride;
    uint32_t* removed = a.removed + (size_t)b * (a.S / 32);
    const uint32_t* mask = a.mask + (size_t)b * (a.S / 32) * a.S;
    if (a.prev_tiles > 0) nms_wait_counter(ready, NMS_RDY_PREV, a.prev_tiles);  // removed[] is final
    if (t < a.S / 32) {
        uint32_t v = __ldcg(removed + t);  // written by other CTAs of this launch
        removed[t] = 0;  // ready for the next super-block
        if (t == nw - 1 && (ncol & 31)) v |= ~0u << (ncol & 31);
        R[t] = v;
    }
    if (t == 0) {
        s_nkept = st.n_kept;
        s_done = 0;
    }
    const bool use_ring = nw <= NMS_RING_WORDS;
    // ---- ring path: resolver warp + helper warps ----------------------------------------------------
    // Warp 0 walks the words; all it needs of the mask are two words per lane and step -- the diagonal word
    // (rows 32u.., column word u) and the same rows' word of column u+1 -- which do not depend on earlier
    // outcomes and are prefetched into registers four steps ahead.  The removed word of step u is
    //   R[u]  (everything kept up to word u-2, ORed in by the helper warps, see below)
    //   | the rows of the candidates kept in word u-1 (one REDUX.OR in warp 0, no round trip through the helpers).
    // Warps 1..7 stream the mask rows through the shared-memory ring and OR the rows of word v's kept candidates
    // into R[v+2..]; they have the whole of step v+1 to do it.  Hand-offs are named barriers (ids by step
    // parity): kb_v ready (resolver arrives, helpers sync) and R[..] of step v done (helpers arrive, the
    // resolver syncs at step v+2).  A step of the serial chain is then resolve + bookkeeping (~400 cycles)
    // instead of resolve, barrier, OR phase, barrier (~1 500).
    if (use_ring) {
        __shared__ uint32_t s_kbs[4], s_dns[4];  // per step (mod 4): kept bits, "keep_cap reached"
        // warps 1..6 help; warp 7 follows the mask's readiness counters when the scan runs beside its launch's tiles
        constexpr int NH = NMS_MAX_WORDS - 64;  // helper threads (192: as many loop trips as 224 for up to 64 words)
        constexpr int NSYNC = 32 + NH;          // resolver + helpers: the parties of the step barriers
        __shared__ volatile int s_known, s_over;  // row blocks [0, s_known) of the mask are complete; the scan is over
        if (t == 0) {
            s_known = ready ? 0 : 1 << 30;
            s_over = 0;
        }
        __syncthreads();                        // R[], s_nkept, s_done, s_known written
        // wait until row block rb of the mask is complete.  The L2 polling is warp 7's job (an acquire load per
        // poll would otherwise sit on the step chains of the resolver and the helpers); everybody else spins on the
        // shared-memory copy, which is normally ahead
        auto wait_rows = [&](int rb) {
            if (rb < known) return;
            while (s_known <= rb) {}
            __threadfence_block();
            known = rb + 1;
        };
        if (t >= 32 + NH) {
            if (ready) {
                const int nrb = (nw + 1) >> 1;
                for (int r = 0; r < nrb && !s_over; ++r) {
                    if (lane == 0) {
                        while (nms_ld_acquire(ready + r) < ncb_launch - (r >> 2) && !s_over) __nanosleep(32);
                        __threadfence_block();
                        s_known = r + 1;
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 0) {
            auto ld_diag = [&](int u) -> uint32_t {  // rows 32u + lane of column word u, strictly upper triangle
                const int rowi = 32 * u + lane;
                const uint32_t d = (u < nw && rowi < ncol) ? __ldcg(mask + (size_t)u * a.S + rowi) : 0u;
                return d & ~((2u << lane) - 1u);
            };
            auto ld_next = [&](int u) -> uint32_t {  // the same rows in column word u + 1
                const int rowi = 32 * u + lane;
                return (u + 1 < nw && rowi < ncol) ? __ldcg(mask + (size_t)(u + 1) * a.S + rowi) : 0u;
            };
            uint32_t Dp[4], Np[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < nw) wait_rows(i >> 1);
                Dp[i] = ld_diag(i);
                Np[i] = ld_next(i);
            }
            int nk = st.n_kept;
            uint32_t carry = 0u;  // rows of word u-1's kept candidates in column word u
            bool stop = false;
#ifdef FRCNN_NMS_TIMING
            long long q0 = clock64(), q_wait = 0, q_res = 0, qa;
#endif
            for (int u0 = 0; u0 < nw && !stop; u0 += 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int u = u0 + i;
                    if (u < nw && !stop) {
#ifdef FRCNN_NMS_TIMING
                        qa = clock64();
#endif
                        if (u >= 2) nms_bar_sync(3 + (u & 1), NSYNC);  // helpers finished word u-2
#ifdef FRCNN_NMS_TIMING
                        q_wait += clock64() - qa;
                        qa = clock64();
#endif
                        const uint32_t D = Dp[i], Nw = Np[i];
                        if (u + 4 < nw) wait_rows((u + 4) >> 1);
                        Dp[i] = ld_diag(u + 4);
                        Np[i] = ld_next(u + 4);
                        uint32_t kb = nms_resolve_word(~(R[u] | carry), D, lane);
                        const int room = a.keep_cap - nk;
                        int cnt = __popc(kb);
                        int done = 0;
                        if (cnt >= room) {
                            while (__popc(kb) > room) kb &= ~(0x80000000u >> __clz(kb));
                            cnt = __popc(kb);
                            done = 1;
                        }
                        if (lane == 0) {
                            s_kbs[u & 3] = kb;
                            s_dns[u & 3] = (uint32_t)done;  // read by the helpers for THIS step: all of them leave together
                            if (done) s_done = 1;
                        }
                        __threadfence_block();
                        nms_bar_arrive(1 + (u & 1), NSYNC);  // kb of word u published
                        if ((kb >> lane) & 1u) {
                            const int pos = nk + __popc(kb & ((1u << lane) - 1u));
                            a.keep[(size_t)b * a.keep_cap + pos] = c0 + 32 * u + lane;  // its box is copied after the loop
                        }
                        carry = __reduce_or_sync(0xFFFFFFFFu, ((kb >> lane) & 1u) ? Nw : 0u);
                        nk += cnt;
                        stop = done != 0;
#ifdef FRCNN_NMS_TIMING
                        q_res += clock64() - qa;
#endif
                    }
                }
            }
            if (lane == 0) {
                s_nkept = nk;
                s_over = 1;
            }
#ifdef FRCNN_NMS_TIMING
            if (lane == 0 && b == 0 && c0 == 0)
                printf("nms resolver nw=%d total=%lld wait_helpers=%lld resolve=%lld kept=%d\n", nw, clock64() - q0, q_wait, q_res, nk);
#endif
        } else {
            const int ht = t - 32;
            // stage of step u: [tw][32 words] for tw in [u, nw); 16-byte chunk c -> column word u + c/8, part c%8
            auto issue = [&](int u) {
                if (u < nw) {
                    wait_rows(u >> 1);
                    uint32_t* dst = ring + (u % NMS_RING_STAGES) * (NMS_RING_WORDS * 32);
                    const int chunks = (nw - u) * 8;
                    for (int c = ht; c < chunks; c += NH) {
                        const int tw = u + (c >> 3), part = c & 7;
                        cp_async_16(dst + tw * 32 + part * 4, mask + (size_t)tw * a.S + 32 * u + part * 4);
                    }
                }
                cp_async_commit();  // one group per step, empty past the end: the wait count stays uniform
            };
#pragma unroll
            for (int u = 0; u < NMS_RING_DEPTH; ++u) issue(u);
#ifdef FRCNN_NMS_TIMING
            long long h_stage = 0, h_kb = 0, h_iss = 0, h_or = 0, ha;
#endif
            for (int v = 0; v < nw; ++v) {
#ifdef FRCNN_NMS_TIMING
                ha = clock64();
#endif
                cp_async_wait<NMS_RING_DEPTH - 1>();
                nms_bar_sync(5, NH);                             // stage v landed for every helper; R[] of step v-1 written
#ifdef FRCNN_NMS_TIMING
                h_stage += clock64() - ha; ha = clock64();
#endif
                nms_bar_sync(1 + (v & 1), NSYNC);        // kb of word v
#ifdef FRCNN_NMS_TIMING
                h_kb += clock64() - ha; ha = clock64();
#endif
                const uint32_t kb = s_kbs[v & 3];
                if (s_dns[v & 3]) break;
                // next stage first: it overwrites the stage step v-1 used (every helper is past this step's barrier)
                issue(v + NMS_RING_DEPTH);
#ifdef FRCNN_NMS_TIMING
                h_iss += clock64() - ha; ha = clock64();
#endif
                const uint32_t* stage = ring + (v % NMS_RING_STAGES) * (NMS_RING_WORDS * 32);
                // rows of the kept candidates ORed into the removed words of column words >= v + 2 (v + 1 is the
                // resolver's own carry): four threads per column word (two 16-byte quarters of its 32 rows each,
                // selected without predicates), combined by two xor shuffles
                const int part = ht & 3;
                const uint32_t ka = kb >> (8 * part), kc = ka >> 4;
                const uint32_t a0 = 0u - (ka & 1u), a1 = 0u - ((ka >> 1) & 1u), a2 = 0u - ((ka >> 2) & 1u),
                               a3 = 0u - ((ka >> 3) & 1u);
                const uint32_t c0_ = 0u - (kc & 1u), c1_ = 0u - ((kc >> 1) & 1u), c2_ = 0u - ((kc >> 2) & 1u),
                               c3_ = 0u - ((kc >> 3) & 1u);
                for (int tw = v + 2 + (ht >> 2); tw - (ht >> 2) < nw; tw += NH / 4) {  // warp-uniform trip count
                    uint32_t acc = 0u;
                    if (tw < nw) {
                        const uint4* m = reinterpret_cast<const uint4*>(stage + tw * 32 + part * 8);
                        const uint4 x = m[0], w = m[1];
                        acc = (x.x & a0) | (x.y & a1) | (x.z & a2) | (x.w & a3) | (w.x & c0_) | (w.y & c1_) | (w.z & c2_) |
                              (w.w & c3_);
                    }
                    acc |= __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
                    acc |= __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
                    if (part == 0 && tw < nw) R[tw] |= acc;
                }
                if (v + 2 < nw) {
                    __threadfence_block();
                    nms_bar_arrive(3 + (v & 1), NSYNC);  // R[v+2..] has word v's rows
                }
#ifdef FRCNN_NMS_TIMING
                h_or += clock64() - ha;
#endif
            }
#ifdef FRCNN_NMS_TIMING
            if (t == 32 && b == 0 && c0 == 0)
                printf("nms helpers stage_wait=%lld kb_wait=%lld issue=%lld or=%lld\n", h_stage, h_kb, h_iss, h_or);
#endif
            cp_async_wait<0>();
        }
        nms_copy_kept_boxes(a, b, boxes, st.n_kept, &s_nkept);
        if (t == 0) {
            int done = s_done || (c1 >= n);
            NmsState o;
            o.n_kept = s_nkept;
            o.done = done;
            o.c_next = c1;
            o.pad = 0;
            a.state_out[b] = o;
            a.n_keep[b] = s_nkept;
        }
        return;
    }
    // ---- wide super-blocks (only on request): rows prefetched into registers one step ahead ----------
    if (nw > 0) nms_wait_rows(ready, (nw - 1) >> 1, ncb_launch, known);  // the whole mask (no overlap on this path)
    const uint32_t* mine = mask + (size_t)t * a.S;  // rows of my column word
    uint4 m[8];
    auto prefetch = [&](int u) {
        if (t > u && t < nw) {
#pragma unroll
            for (int q = 0; q < 8; ++q) m[q] = __ldcg(reinterpret_cast<const uint4*>(mine + 32 * u) + q);
        }
    };
    auto diag = [&](int u) -> uint32_t {  // warp 0: which later candidates of block u does row 32u+lane suppress
        int row = 32 * u + lane;
        uint32_t d = (u < nw && row < ncol) ? __ldcg(mask + (size_t)u * a.S + row) : 0u;
        return d & ~((2u << lane) - 1u);
    };
    prefetch(0);
    uint32_t D = warp == 0 ? diag(0) : 0u;
    __syncthreads();
    for (int u = 0; u < nw; ++u) {
        if (warp == 0) {
            const uint32_t Dn = diag(u + 1);  // in flight during the resolve
            uint32_t kb = nms_resolve_word(~R[u], D, lane);
            D = Dn;
            int nk = s_nkept;
            int room = a.keep_cap - nk;
            int cnt = __popc(kb);
            int done = 0;
            if (cnt >= room) {
                while (__popc(kb) > room) kb &= ~(0x80000000u >> __clz(kb));
                cnt = __popc(kb);
                done = 1;
            }
            if ((kb >> lane) & 1u) {
                int pos = nk + __popc(kb & ((1u << lane) - 1u));
                int row = c0 + 32 * u + lane;
                a.keep[(size_t)b * a.keep_cap + pos] = row;  // its box is copied after the loop
            }
            if (lane == 0) {
                s_kb = kb;
                s_nkept = nk + cnt;
                if (done) s_done = 1;
            }
        }
        __syncthreads();
        if (s_done) break;
        if (t > u && t < nw) {
            const uint32_t kb = s_kb;
            uint32_t acc = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc |= ((kb >> (4 * q)) & 1u) ? m[q].x : 0u;
                acc |= ((kb >> (4 * q + 1)) & 1u) ? m[q].y : 0u;
                acc |= ((kb >> (4 * q + 2)) & 1u) ? m[q].z : 0u;
                acc |= ((kb >> (4 * q + 3)) & 1u) ? m[q].w : 0u;
            }
            R[t] |= acc;
        }
        prefetch(u + 1);
        __syncthreads();
    }
    nms_copy_kept_boxes(a, b, boxes, st.n_kept, &s_nkept);
    if (t == 0) {
        int done = s_done || (c1 >= n);
        NmsState o;
        o.n_kept = s_nkept;
        o.done = done;
        o.c_next = c1;
        o.pad = 0;
        a.state_out[b] = o;
        a.n_keep[b] = s_nkept;
    }
}

// One launch per super-block: mask tiles plus one scan CTA per image that consumes the mask WHILE it is being
// computed.  The scan reads the mask one 64-row block at a time, top to bottom (resolver: the diagonal words; helpers:
// the rows of the step's 32 candidates in every later column word), so the tiles are issued row block by row block
// and count themselves done per row block (`ready`, after a __threadfence); the scan's threads wait for exactly the
// row block they are about to load.  Before: the CTA that finished an image's last tile ran the scan, i.e. mask and
// scan in series -- and the scan is the longer half (0.5 us per 32 candidates).  A scan that reaches keep_cap closes
// the launch for its image: tiles that have not started yet return at once.
__global__ void __launch_bounds__(NMS_MAX_WORDS) nms_block_kernel(NmsArgs a) {
    __shared__ float4 srow[NMS_KC];
    __shared__ float sarea[NMS_KC];
    __shared__ uint32_t R[NMS_MAX_WORDS];
    __shared__ __align__(16) uint32_t ring[NMS_RING_STAGES * NMS_RING_WORDS * 32];  // 32 KB
    // 1-D grid: [one scan CTA per image][kept-list tiles][in-block tiles, row block by row block]; inside each
    // group the IMAGE is the fastest index, so the masks of all images advance together and every image's scan can
    // follow its mask from the first row block on.  A tile CTA never waits for anything, so the scan CTAs (first in
    // the grid, resident from the start) always see their counters arrive.
    const int B = a.batch;
    if (a.ready == nullptr) {
        // Very large batches (more scan CTAs than the GPU should hold spinning): no dedicated scan CTAs; the CTA that
        // finishes an image's last tile (counted with an atomic after a __threadfence, nobody waits) runs its scan.
        __shared__ int s_last;
        const int b = (int)blockIdx.x % B, t = (int)blockIdx.x / B, tiles = a.tri_tiles + a.prev_tiles;
        const NmsState st = a.state[b];
        if (st.done) {
            if (t == 0 && threadIdx.x == 0) a.state_out[b] = st;
            return;
        }
        nms_mask_tile(a, b, st, srow, sarea, t, tiles);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned prev = atomicAdd(a.tile_count + b, 1u);
            s_last = prev == (unsigned)tiles - 1u;
            if (s_last) a.tile_count[b] = 0;  // ready for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        nms_scan_image(a, b, st, R, ring);
        return;
    }
    const int idx = (int)blockIdx.x;
    const bool scan = idx < B;
    const int b = scan ? idx : (idx - B) % B;
    const NmsState st = a.state[b];
    if (st.done) {  // set by an earlier launch: identical for all CTAs of this image
        if (scan && threadIdx.x == 0) a.state_out[b] = st;  // carried into the buffer the next launch reads
        return;
    }
    int* ready = a.ready + (size_t)b * NMS_RDY_STRIDE;
    if (scan) {
        nms_scan_image(a, b, st, R, ring, ready);
        // the image's scan is over (often early: keep_cap reached): tiles that have not started yet have no reader
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(ready + NMS_RDY_PREV + 1, 1);
        return;
    }
    if (nms_ld_acquire(ready + NMS_RDY_PREV + 1)) return;
    const int sched = (idx - B) / B;
    const int t = sched < a.prev_tiles ? a.tri_tiles + sched : sched - a.prev_tiles;
    const int counter = nms_mask_tile(a, b, st, srow, sarea, t, a.tri_tiles + a.prev_tiles);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(ready + counter, 1);
}

// Whatever the sized launches leave over, in ONE launch.  The default schedule sizes its first super-block to finish
// the job and lets one or two adaptive ones follow; covering the worst case (every candidate) used to take a further
// row_stride / 2048 launches that found `done` set and returned -- 13 of them, ~2.5 us each, on the 30 000-candidate
// configuration.  Here a cluster of 8 CTAs per image (co-resident by construction, so a cluster barrier is a safe
// image-wide barrier) loops over super-blocks itself: the mask tiles are dealt over the cluster's CTAs, cluster
// barrier, CTA 0 scans, cluster barrier, until keep_cap boxes are kept or the candidates run out.  An image that is
// already done costs one state read.  Everything that crosses CTAs inside the loop is read from L2 (__ldcg /
// cp.async.cg): state, removed words, mask, kept boxes.
constexpr int NMS_TAIL_CL = 8;
__global__ void __cluster_dims__(NMS_TAIL_CL, 1, 1) __launch_bounds__(NMS_MAX_WORDS) nms_tail_kernel(NmsArgs a) {
    __shared__ float4 srow[NMS_KC];
    __shared__ float sarea[NMS_KC];
    __shared__ uint32_t R[NMS_MAX_WORDS];
    __shared__ __align__(16) uint32_t ring[NMS_RING_STAGES * NMS_RING_WORDS * 32];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.x / NMS_TAIL_CL;
    const int rank = (int)cluster.block_rank();
    const int n = a.n_sel[b];
    const int tiles_total = (int)a.tri_tiles + (a.len / NMS_CB) * ((a.keep_cap + NMS_KC - 1) / NMS_KC);
    for (;;) {
        NmsState st;
        {
            const int4 v = __ldcg(reinterpret_cast<const int4*>(a.state + b));
            st.n_kept = v.x;
            st.done = v.y;
            st.c_next = v.z;
            st.pad = v.w;
        }
        if (st.done) break;  // identical in every CTA of the cluster
        if (st.c_next < n) {
            for (int t = rank; t < tiles_total; t += NMS_TAIL_CL) {
                nms_mask_tile(a, b, st, srow, sarea, t, tiles_total);
                __syncthreads();  // srow / sarea are reused by the next tile
            }
        }
        __threadfence();
        cluster.sync();
        if (rank == 0) nms_scan_image(a, b, st, R, ring);  // sets done when the candidates are exhausted
        __threadfence();
        cluster.sync();
    }
    // Pad with arange / truncate / gather (finalize_kernel's job) by the cluster that has just seen the image done:
    // one launch less per proposal layer.  keep_cap == n_post here.  A padded row r >= k takes sorted position
    // r - k; the reference raises IndexError when that runs past the candidate list, i.e. iff n_post - 1 - k >= n.
    if (a.fin_rois) {
        const int k = __ldcg(a.n_keep + b), n_post = a.keep_cap;
        for (int r = rank * NMS_CB + (int)threadIdx.x; r < n_post; r += NMS_TAIL_CL * NMS_CB) {
            const int p = r < k ? __ldcg(a.keep + (size_t)b * n_post + r) : r - k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int src = -1;
            if (p < n) {
                v = __ldg(a.boxes + (size_t)b * a.row_stride + p);
                src = __ldg(a.fin_order + (size_t)b * a.row_stride + p);
            }
            a.fin_rois[(size_t)b * n_post + r] = v;
            if (a.fin_src) a.fin_src[(size_t)b * n_post + r] = src;
        }
        if (rank == 0 && threadIdx.x == 0) {
            a.fin_status[b] = (k < n_post && n_post - 1 - k >= n) ? FRCNN_IMG_PAD_INDEX_ERROR : FRCNN_IMG_OK;
            if (a.fin_n_keep) a.fin_n_keep[b] = k;
        }
    }
}

// nets/rpn.py:65-69: pad with arange, truncate, gather
__global__ void finalize_kernel(const float4* __restrict__ sorted, const int* __restrict__ order,
                                const int* __restrict__ n_sel_all, const int* __restrict__ keep,
                                const int* __restrict__ n_keep_all, int row_stride, int n_post,
                                float4* __restrict__ rois, int* __restrict__ roi_src,
                                int* __restrict__ n_keep_out, int* __restrict__ status) {
    const int b = blockIdx.x;
    const int n_sel = n_sel_all[b], k = n_keep_all[b];
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < n_post; r += blockDim.x) {
        int p = r < k ? keep[(size_t)b * n_post + r] : r - k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int src = -1;
        if (p < n_sel) {
            v = sorted[(size_t)b * row_stride + p];
            src = order[(size_t)b * row_stride + p];
        } else {
            bad = 1;
        }
        rois[(size_t)b * n_post + r] = v;
        if (roi_src) roi_src[(size_t)b * n_post + r] = src;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        status[b] = bad ? FRCNN_IMG_PAD_INDEX_ERROR : FRCNN_IMG_OK;
        if (n_keep_out) n_keep_out[b] = k;
    }
}

__global__ void keep_to_index_kernel(const int* __restrict__ keep, const int* __restrict__ n_keep,
                                     const int* __restrict__ order, int64_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_keep[0]) out[i] = order[keep[i]];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int pick_superblock(int requested, int n_rows, int keep_cap) {
    // default: about twice the number of boxes wanted, so that the first super-block usually
    // finishes the job and little of the IoU mask is computed for nothing (measured on cfg2: a
    // 2048-wide block spent 97 us on a mask of which ~430 columns were needed)
    int S = requested > 0 ? requested : 2 * keep_cap;
    S = (S + NMS_CB - 1) / NMS_CB * NMS_CB;
    if (S < NMS_CB) S = NMS_CB;
    if (S > 2048 && requested <= 0) S = 2048;
    if (S > NMS_MAX_WORDS * 32) S = NMS_MAX_WORDS * 32;
    int need = (n_rows + NMS_CB - 1) / NMS_CB * NMS_CB;
    if (need < NMS_CB) need = NMS_CB;
    if (S > need) S = need;
    return S;
}

struct NmsLayout {
    int S;
    size_t mask_words, removed_words;
};

static int max_superblock(int requested, int n_rows, int keep_cap) {
    int s0 = pick_superblock(requested, n_rows, keep_cap);
    if (requested > 0) return s0;
    int need = (n_rows + NMS_CB - 1) / NMS_CB * NMS_CB;
    return std::max(s0, std::min(2048, need));
}

// upper bound of the nms_block_kernel launches of one run_nms_sorted call (each has its own readiness counters)
static int nms_max_launches(int n_rows, int keep_cap, int superblock) {
    return cdiv(std::max(n_rows, 1), pick_superblock(superblock, n_rows, keep_cap)) + 3;
}

static size_t nms_ws_layout(Workspace& ws, int batch, int n_rows, int keep_cap, int superblock,
                            NmsArgs* a) {
    int S = max_superblock(superblock, n_rows, keep_cap);
    uint32_t* mask = ws.take<uint32_t>((size_t)batch * (S / 32) * S);
    // state (two buffers, see NmsArgs::state_out) + removed + the per-launch readiness counters are cleared together
    // by one memset
    const int max_launch = nms_max_launches(n_rows, keep_cap, superblock);
    size_t clear_bytes = align_up((size_t)2 * batch * sizeof(NmsState)) + align_up((size_t)batch * sizeof(unsigned)) +
                         align_up((size_t)batch * (S / 32) * 4) +
                         align_up((size_t)max_launch * batch * NMS_RDY_STRIDE * sizeof(int));
    NmsState* state = ws.take<NmsState>((size_t)2 * batch);
    unsigned* tile_count = ws.take<unsigned>(batch);
    uint32_t* removed = ws.take<uint32_t>((size_t)batch * (S / 32));
    int* ready = ws.take<int>((size_t)max_launch * batch * NMS_RDY_STRIDE);
    float4* kept_box = ws.take<float4>((size_t)batch * (keep_cap > 0 ? keep_cap : 1));
    if (a) {
        a->S = S;
        a->mask = mask;
        a->state = state;
        a->ready = ready;
        a->tile_count = tile_count;
        a->removed = removed;
        a->kept_box = kept_box;
    }
    return clear_bytes;
}

static int run_nms_sorted(const float* sorted_boxes, const int32_t* n_sel, int batch, int row_stride,
                          double thresh, int keep_cap, int superblock, int32_t* keep, int32_t* n_keep,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream, NmsFinalize* fin = nullptr) {
    Workspace ws(workspace, workspace_bytes);
    NmsArgs a;
    memset(&a, 0, sizeof(a));
    size_t clear_bytes = nms_ws_layout(ws, batch, row_stride, keep_cap, superblock, &a);
    if (!ws.ok()) {
        set_error("nms: workspace too small or misaligned (%zu needed, %zu given)", ws.off, workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    a.boxes = (const float4*)sorted_boxes;
    a.n_sel = n_sel;
    a.row_stride = row_stride;
    a.keep_cap = keep_cap;
    a.thr = float_threshold(thresh);
    a.keep = keep;
    a.n_keep = n_keep;
    FRCNN_CUDA(cudaMemsetAsync(a.state, 0, clear_bytes, stream));
    a.batch = batch;
    NmsState* const state2 = a.state;  // [2][B]
    int* const ready_all = a.ready;
    int n_block_launches = 0;
    static const int early_off = []() { const char* v = getenv("FRCNN_NMS_NOEARLY"); return v && *v ? atoi(v) : 0; }();
    // launch i: reads state buffer i & 1, its scan writes the other one; own readiness counters
    auto launch_block = [&](int prev_tiles) -> cudaError_t {
        const int i = n_block_launches++;
        a.state = state2 + (size_t)(i & 1) * batch;
        a.state_out = state2 + (size_t)((i + 1) & 1) * batch;
        // a scan CTA per image that follows the mask while it is being computed -- unless that would park more
        // spinning CTAs on the GPU than it has SMs (they must never keep the tiles from running)
        // It pays only when the tiles do not all fit on the GPU at once (4 CTAs per SM): in a single wave every row
        // block completes at about the same time and there is nothing to overlap (16 images x 24 tiles: 26.6 vs
        // 24.8 us; 8 x 144 tiles: 59.4 vs 72.4 us).
        const bool early = batch <= sm_count() && !early_off &&
                           (int64_t)batch * (a.tri_tiles + prev_tiles) > (int64_t)4 * sm_count();
        a.ready = early ? ready_all + (size_t)i * batch * NMS_RDY_STRIDE : nullptr;
        a.prev_tiles = prev_tiles;
        nms_block_kernel<<<batch * ((early ? 1 : 0) + a.tri_tiles + prev_tiles), NMS_CB, 0, stream>>>(a);
        return cudaGetLastError();
    };
    const int max_launch = nms_max_launches(row_stride, keep_cap, superblock);
    // super-block schedule: the first block is sized for keep_cap (pick_superblock); later blocks double
    // up to the mask stride, so a run that finishes early launches few no-op kernels
    const int s0 = pick_superblock(superblock, row_stride, keep_cap);
    std::vector<int> lens;
    int n_launch = 0, total = 0;
    for (int len = s0; total < row_stride; ++n_launch) {
        lens.push_back(len);
        total += len;
        if (superblock <= 0) len = std::min(2 * len, a.S);
    }
    if (superblock <= 0 && s0 < row_stride && 2 * keep_cap <= row_stride) {  // (keep_cap ~ n: plain NMS, all rounds wide)
        // default schedule: the first block (sized for keep_cap), one adaptive block -- two when the first cannot
        // finish the job by itself -- and nms_tail_kernel for whatever is left (usually nothing)
        const int n_adaptive = s0 < 2 * keep_cap ? 2 : 1;
        a.adaptive = 1;
        a.cover_after = 1 << 30;  // the tail covers every candidate: later blocks may be cut freely
        for (int i = 0; i <= n_adaptive; ++i) {
            a.len = i == 0 ? s0 : a.S;
            const int ncb = a.len / NMS_CB;
            a.tri_tiles = 2 * ncb * (ncb + 1);
            const int prev_tiles = i > 0 ? ncb * cdiv(keep_cap, NMS_KC) : 0;
            FRCNN_CUDA(launch_block(prev_tiles));
            count_launch();
        }
        a.state = a.state_out = state2 + (size_t)(n_block_launches & 1) * batch;  // the tail updates it in place
        a.ready = nullptr;
        a.prev_tiles = 0;
        a.len = std::min(a.S, 1024);  // 8 CTAs per image: 40 + 4 * ceil(keep_cap / 256) tiles per round
        const int ncb = a.len / NMS_CB;
        a.tri_tiles = 2 * ncb * (ncb + 1);
        if (fin) {
            a.fin_order = fin->order;
            a.fin_rois = (float4*)fin->rois;
            a.fin_src = fin->roi_src;
            a.fin_n_keep = fin->n_keep;
            a.fin_status = fin->status;
            fin->fused = true;
        }
        nms_tail_kernel<<<batch * NMS_TAIL_CL, NMS_CB, 0, stream>>>(a);
        FRCNN_LAUNCH_CHECK();
        return FRCNN_OK;
    }
    // Long schedules get one spare launch: its coverage is the slack that lets the device cut later super-blocks
    // to what an image still needs (nms_block_len); a no-op launch costs ~2 us, a needless 2048-wide mask ~100
    // (only where the first block is too narrow to finish the job: otherwise launch 2 onwards are no-ops anyway)
    a.adaptive = superblock <= 0 && n_launch >= 4 && s0 < 2 * keep_cap;
    if (a.adaptive) {
        lens.push_back(a.S);
        ++n_launch;
        total += a.S;
    }
    for (int i = 0, after = total; i < n_launch; ++i) {
        const int len = lens[i];
        after -= len;
        a.len = len;
        a.cover_after = after;
        int ncb = len / NMS_CB;
        a.tri_tiles = 2 * ncb * (ncb + 1);
        int prev_tiles = i > 0 ? ncb * cdiv(keep_cap, NMS_KC) : 0;
        if (n_block_launches >= max_launch) {
            set_error("nms: internal error, more launches than readiness counters");
            return FRCNN_ERR_UNSUPPORTED;
        }
        FRCNN_CUDA(launch_block(prev_tiles));
        count_launch();
    }
    return FRCNN_OK;
}

static int run_topk(const uint32_t* keys, const float* boxes, int batch, int n, int k_cap, int32_t* order,
                    int32_t* n_sel, float* sorted_boxes, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
    Workspace ws(workspace, workspace_bytes);
    uint32_t* wk = ws.take<uint32_t>((size_t)batch * 2 * n);
    uint32_t* wi = ws.take<uint32_t>((size_t)batch * 2 * n);
    if (!ws.ok()) {
        set_error("topk: workspace too small or misaligned (%zu needed, %zu given)", ws.off, workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    // cluster size per image: 8 CTAs, or 4 when 8 per image would need well over one wave of the GPU and 4 fit in one
    // (B = 32 images of 22 500 keys: 158 -> 131 us for the whole proposal stage).  16 (non-portable) was measured
    // slower than 8 everywhere (B = 8 x 36 864 keys: 236 vs 205 us): the cluster barriers cost more than the shorter
    // segments save.
    static const int cl_override = []() { const char* v = getenv("FRCNN_TOPK_CL"); return v && *v ? atoi(v) : 0; }();
    static const int presel_off = []() { const char* v = getenv("FRCNN_TOPK_NOPRESEL"); return v && *v ? atoi(v) : 0; }();
    // pre-select the k_cap best before sorting when that is at most half of the keys (see the kernel).  The sort
    // that follows handles a few thousand keys and is bound by its cluster barriers, which are cheaper among 4
    // CTAs than among 8 (16 x 12 996 keys, 3 000 wanted: 38 us full sort on 8-CTA clusters, 43 us pre-selected on 8,
    // 26 us pre-selected on 4; 32 x 22 500: 72 -> 29 us).
    const bool want_presel = 2ll * k_cap <= n && !presel_off;
    int cl = 8;
    if (batch * 8 > sm_count() + sm_count() / 2 && batch * 4 <= sm_count()) cl = 4;
    if (want_presel) cl = 4;
    if (cl_override == 4 || cl_override == 8) cl = cl_override;
    static const int e9_off = []() { const char* v = getenv("FRCNN_TOPK_NOE9"); return v && *v ? atoi(v) : 0; }();
    for (; k_cap <= n; cl = 8) {
        const int seg = (n + cl - 1) / cl;
        size_t dsmem = TK_SMEM + (size_t)2 * seg * sizeof(uint2);
        // Nine keys per thread instead of four where that turns a multi-round segment into a single round (the
        // multi-round passes re-load and re-synchronise per round): 8 x 36 864 keys, 30 000 wanted.  Needs 144 KB of
        // per-(row, warp) digit counters, so only where the segment still fits beside them.
        const size_t dsmem9 = (size_t)9 * TK_WARPS * 256 * sizeof(uint32_t) + (size_t)2 * seg * sizeof(uint2);
        const bool e9 = cl == 8 && !want_presel && !e9_off && seg > TK_ROUND && seg <= 9 * TK_THREADS && dsmem9 <= 224 * 1024;
        if (e9) dsmem = dsmem9;
        if (!e9 && dsmem > 200 * 1024) {
            if (cl == 8) break;
            continue;
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(batch * cl);
        cfg.blockDim = dim3(TK_THREADS);
        cfg.dynamicSmemBytes = dsmem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cl;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaSuccess;
        const float4* bx = (const float4*)boxes;
        float4* sb = (float4*)sorted_boxes;
        // the (row, warp) count table of the compaction is scanned by one pass of the CTA: seg <= 32 rows of TK_THREADS
        const int presel = (want_presel && seg <= 32 * TK_THREADS) ? 1 : 0;
        if (cl == 4) {
            FRCNN_SMEM(topk_sort_dsmem_kernel<4>, dsmem);
            e = cudaLaunchKernelEx(&cfg, topk_sort_dsmem_kernel<4>, keys, bx, n, k_cap, seg, presel, order, n_sel, sb);
        } else if (e9) {
            FRCNN_SMEM((topk_sort_dsmem_kernel<8, 9>), dsmem);
            e = cudaLaunchKernelEx(&cfg, topk_sort_dsmem_kernel<8, 9>, keys, bx, n, k_cap, seg, presel, order, n_sel, sb);
        } else {
            FRCNN_SMEM(topk_sort_dsmem_kernel<8>, dsmem);
            e = cudaLaunchKernelEx(&cfg, topk_sort_dsmem_kernel<8>, keys, bx, n, k_cap, seg, presel, order, n_sel, sb);
        }
        FRCNN_CUDA(e);
        count_launch();
        return FRCNN_OK;
    }
    {  // global-memory ping-pong buffers
        FRCNN_SMEM(topk_sort_kernel, TK_SMEM);
        topk_sort_kernel<<<batch * TK_CL, TK_THREADS, TK_SMEM, stream>>>(keys, (const float4*)boxes, n, k_cap, wk, wi,
                                                                         order, n_sel, (float4*)sorted_boxes);
    }
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

static AnchorGen make_gen(const frcnn_anchor_spec* s) {
    return make_anchor_gen(s->anchors, s->base, s->num_base, s->feat_stride, s->height, s->width);
}

static int check_anchor_spec(const frcnn_anchor_spec* s, int n, const char* who) {
    if (!s) {
        set_error("%s: anchor spec is null", who);
        return FRCNN_ERR_INVALID_ARG;
    }
    if (s->anchors) return FRCNN_OK;
    if (!s->base || s->num_base <= 0 || s->num_base > FRCNN_MAX_BASE_ANCHORS || s->width <= 0 ||
        (int64_t)s->num_base * s->height * s->width != n) {
        set_error("%s: generated anchors need base/num_base and H*W*A == N (%d)", who, n);
        return FRCNN_ERR_INVALID_ARG;
    }
    return FRCNN_OK;
}

static int run_decode(const frcnn_proposal_params* p, const frcnn_anchor_spec* anchors, const float* loc,
                      const float* score, float* boxes, uint32_t* keys, float* fg_out, cudaStream_t stream) {
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.loc = (const float4*)loc;
    a.score = score;
    if (!p->boxes_are_decoded) a.gen = make_gen(anchors);
    a.batch = p->batch;
    a.n = p->num_anchors;
    a.xmax = p->clip_x_max;
    a.ymax = p->clip_y_max;
    a.min_size = p->min_size;
    a.score_mode = p->score_mode;
    a.decoded = p->boxes_are_decoded;
    a.boxes = (float4*)boxes;
    a.keys = keys;
    a.fg_out = fg_out;
    FRCNN_CHECK_ARG(p->batch <= 65535, "frcnn_decode_clip_score: batch %d > 65535", p->batch);
    if (p->score_mode == 2) {
        const size_t smem = (size_t)6 * a.gen.num_base * DEC_PITCH * sizeof(float);
        FRCNN_SMEM(decode_clip_key_nchw_kernel, smem);
        decode_clip_key_nchw_kernel<<<dim3(cdiv(a.gen.height * a.gen.width, DEC_TILE), p->batch), 256, smem, stream>>>(a);
        FRCNN_LAUNCH_CHECK();
        return FRCNN_OK;
    }
    decode_clip_key_kernel<<<dim3(cdiv(p->num_anchors, 256), p->batch), 256, 0, stream>>>(a);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

static int check_params(const frcnn_proposal_params* p, const frcnn_anchor_spec* anchors, const char* who) {
    FRCNN_CHECK_ARG(p, "%s: params is null", who);
    FRCNN_CHECK_ARG(p->batch > 0 && p->num_anchors > 0, "%s: batch and num_anchors must be positive", who);
    FRCNN_CHECK_ARG(p->n_post_nms >= 0, "%s: n_post_nms must be >= 0", who);
    FRCNN_CHECK_ARG((int64_t)p->batch * p->num_anchors < (1ll << 31), "%s: batch*num_anchors too large", who);
    FRCNN_CHECK_ARG(p->score_mode >= 0 && p->score_mode <= 2, "%s: bad score_mode", who);
    if (!p->boxes_are_decoded) {
        int rc = check_anchor_spec(anchors, p->num_anchors, who);
        if (rc) return rc;
    }
    if (p->score_mode == 2) {  // NCHW conv outputs: the (A, H, W) factorisation of N must be known
        FRCNN_CHECK_ARG(!p->boxes_are_decoded && anchors && anchors->num_base > 0 &&
                            anchors->num_base <= FRCNN_MAX_BASE_ANCHORS && anchors->height > 0 && anchors->width > 0 &&
                            (int64_t)anchors->num_base * anchors->height * anchors->width == p->num_anchors,
                        "%s: score_mode 2 needs num_base, height, width in the anchor spec with A*H*W == N", who);
    }
    return FRCNN_OK;
}

static int n_pre_rows(const frcnn_proposal_params* p) {
    return (p->n_pre_nms > 0 && p->n_pre_nms < p->num_anchors) ? p->n_pre_nms : p->num_anchors;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

size_t frcnn_topk_workspace_bytes(int32_t batch, int32_t n) {
    Workspace ws(nullptr, 0);
    ws.take<uint32_t>((size_t)batch * 2 * n);
    ws.take<uint32_t>((size_t)batch * 2 * n);
    return ws.off;
}

size_t frcnn_nms_sorted_workspace_bytes(int32_t batch, int32_t n_rows, int32_t keep_cap, int32_t superblock) {
    Workspace ws(nullptr, 0);
    nms_ws_layout(ws, batch, n_rows, keep_cap, superblock, nullptr);
    return ws.off;
}

int frcnn_decode_clip_score(const frcnn_proposal_params* p, const frcnn_anchor_spec* anchors,
                            const float* loc, const float* score, float* boxes, uint32_t* keys,
                            float* fg_out, frcnn_stream_t stream) {
    int rc = check_params(p, anchors, "frcnn_decode_clip_score");
    if (rc) return rc;
    FRCNN_CHECK_ARG(loc && score && boxes && keys, "frcnn_decode_clip_score: null pointer");
    return run_decode(p, anchors, loc, score, boxes, keys, fg_out, (cudaStream_t)stream);
}

int frcnn_topk_sorted(const uint32_t* keys, const float* boxes, int32_t batch, int32_t n, int32_t k_cap,
                      int32_t* order, int32_t* n_sel, float* sorted_boxes, void* workspace,
                      size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(batch > 0 && n > 0 && k_cap > 0, "frcnn_topk_sorted: bad shape");
    FRCNN_CHECK_ARG(keys && order && n_sel, "frcnn_topk_sorted: null pointer");
    FRCNN_CHECK_ARG((sorted_boxes == nullptr) || (boxes != nullptr), "frcnn_topk_sorted: sorted_boxes needs boxes");
    return run_topk(keys, boxes, batch, n, k_cap, order, n_sel, sorted_boxes, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

int frcnn_nms_sorted(const float* sorted_boxes, const int32_t* n_sel, int32_t batch, int32_t row_stride,
                     double thresh, int32_t keep_cap, int32_t superblock, int32_t* keep, int32_t* n_keep,
                     void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_CHECK_ARG(batch > 0 && row_stride > 0 && keep_cap > 0, "frcnn_nms_sorted: bad shape");
    FRCNN_CHECK_ARG(sorted_boxes && n_sel && keep && n_keep, "frcnn_nms_sorted: null pointer");
    return run_nms_sorted(sorted_boxes, n_sel, batch, row_stride, thresh, keep_cap, superblock, keep, n_keep,
                          workspace, workspace_bytes, (cudaStream_t)stream);
}

// workspace of the fused pipeline: boxes, keys, order, n_sel, sorted, keep, n_keep, topk ws, nms ws
struct ProposalWs {
    float* boxes;
    uint32_t* keys;
    int32_t* order;
    int32_t* n_sel;
    float* sorted;
    int32_t* keep;
    int32_t* n_keep;
    void* topk_ws;
    size_t topk_bytes;
    void* nms_ws;
    size_t nms_bytes;
};

static size_t proposal_layout(Workspace& ws, const frcnn_proposal_params* p, ProposalWs* out) {
    int B = p->batch, N = p->num_anchors, rows = n_pre_rows(p);
    int cap = p->n_post_nms > 0 ? p->n_post_nms : 1;
    ProposalWs w;
    w.boxes = ws.take<float>((size_t)B * N * 4);
    w.keys = ws.take<uint32_t>((size_t)B * N);
    w.order = ws.take<int32_t>((size_t)B * rows);
    w.n_sel = ws.take<int32_t>(B);
    w.sorted = ws.take<float>((size_t)B * rows * 4);
    w.keep = ws.take<int32_t>((size_t)B * cap);
    w.n_keep = ws.take<int32_t>(B);
    w.topk_bytes = frcnn_topk_workspace_bytes(B, N);
    w.topk_ws = ws.take<char>(w.topk_bytes);
    w.nms_bytes = frcnn_nms_sorted_workspace_bytes(B, rows, cap, p->nms_superblock);
    w.nms_ws = ws.take<char>(w.nms_bytes);
    if (out) *out = w;
    return ws.off;
}

size_t frcnn_proposals_workspace_bytes(const frcnn_proposal_params* p) {
    if (!p || p->batch <= 0 || p->num_anchors <= 0) return 0;
    Workspace ws(nullptr, 0);
    return proposal_layout(ws, p, nullptr);
}

int frcnn_proposals(const frcnn_proposal_params* p, const frcnn_anchor_spec* anchors, const float* loc,
                    const float* score, float* rois, int32_t* roi_src, int32_t* n_keep, int32_t* status,
                    void* workspace, size_t workspace_bytes, frcnn_stream_t stream_) {
    int rc = check_params(p, anchors, "frcnn_proposals");
    if (rc) return rc;
    FRCNN_CHECK_ARG(loc && score && status, "frcnn_proposals: null pointer");
    FRCNN_CHECK_ARG(p->n_post_nms == 0 || rois, "frcnn_proposals: rois is null");
    cudaStream_t stream = (cudaStream_t)stream_;
    Workspace ws(workspace, workspace_bytes);
    ProposalWs w;
    proposal_layout(ws, p, &w);
    if (!ws.ok()) {
        set_error("frcnn_proposals: workspace too small or misaligned (%zu needed, %zu given)", ws.off,
                  workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    int B = p->batch, N = p->num_anchors, rows = n_pre_rows(p);
    if (p->n_post_nms == 0) {
        FRCNN_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, stream));
        if (n_keep) FRCNN_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t) * B, stream));
        return FRCNN_OK;
    }
    rc = run_decode(p, anchors, loc, score, w.boxes, w.keys, nullptr, stream);
    if (rc) return rc;
    rc = run_topk(w.keys, w.boxes, B, N, rows, w.order, w.n_sel, w.sorted, w.topk_ws, w.topk_bytes, stream);
    if (rc) return rc;
    NmsFinalize fin = {w.order, rois, roi_src, n_keep, status, false};
    rc = run_nms_sorted(w.sorted, w.n_sel, B, rows, p->nms_thresh, p->n_post_nms, p->nms_superblock, w.keep,
                        w.n_keep, w.nms_ws, w.nms_bytes, stream, &fin);
    if (rc) return rc;
    if (fin.fused) return FRCNN_OK;  // nms_tail_kernel wrote rois / roi_src / n_keep / status
    // one CTA per image (the status flag is a CTA-wide OR); two dependent L2 round trips per row, so as many rows
    // in flight as the CTA can hold
    finalize_kernel<<<B, p->n_post_nms > 512 ? 1024 : (p->n_post_nms > 256 ? 512 : 256), 0, stream>>>((const float4*)w.sorted, w.order, w.n_sel, w.keep, w.n_keep, rows,
                                           p->n_post_nms, (float4*)rois, roi_src, n_keep, status);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

struct NmsWs {
    uint32_t* keys;
    int32_t* order;
    int32_t* n_sel;
    float* sorted;
    int32_t* keep;
    void* topk_ws;
    size_t topk_bytes;
    void* nms_ws;
    size_t nms_bytes;
};

static size_t nms_layout(Workspace& ws, int n, NmsWs* out) {
    NmsWs w;
    w.keys = ws.take<uint32_t>(n);
    w.order = ws.take<int32_t>(n);
    w.n_sel = ws.take<int32_t>(1);
    w.sorted = ws.take<float>((size_t)n * 4);
    w.keep = ws.take<int32_t>(n);
    w.topk_bytes = frcnn_topk_workspace_bytes(1, n);
    w.topk_ws = ws.take<char>(w.topk_bytes);
    w.nms_bytes = frcnn_nms_sorted_workspace_bytes(1, n, n, 0);
    w.nms_ws = ws.take<char>(w.nms_bytes);
    if (out) *out = w;
    return ws.off;
}

size_t frcnn_nms_workspace_bytes(int32_t n) {
    if (n <= 0) return 256;
    Workspace ws(nullptr, 0);
    return nms_layout(ws, n, nullptr);
}

int frcnn_nms(const float* boxes, const float* scores, int32_t n, double thresh, int64_t* keep,
              int32_t* n_keep, void* workspace, size_t workspace_bytes, frcnn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FRCNN_CHECK_ARG(n >= 0 && n_keep, "frcnn_nms: bad arguments");
    if (n == 0) {
        FRCNN_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), stream));
        return FRCNN_OK;
    }
    FRCNN_CHECK_ARG(boxes && scores && keep, "frcnn_nms: null pointer");
    Workspace ws(workspace, workspace_bytes);
    NmsWs w;
    nms_layout(ws, n, &w);
    if (!ws.ok()) {
        set_error("frcnn_nms: workspace too small or misaligned (%zu needed, %zu given)", ws.off, workspace_bytes);
        return FRCNN_ERR_WORKSPACE;
    }
    scores_to_keys_kernel<<<cdiv(n, 256), 256, 0, stream>>>(scores, n, w.keys);
    FRCNN_LAUNCH_CHECK();
    int rc = run_topk(w.keys, boxes, 1, n, n, w.order, w.n_sel, w.sorted, w.topk_ws, w.topk_bytes, stream);
    if (rc) return rc;
    rc = run_nms_sorted(w.sorted, w.n_sel, 1, n, thresh, n, 0, w.keep, n_keep, w.nms_ws, w.nms_bytes, stream);
    if (rc) return rc;
    keep_to_index_kernel<<<cdiv(n, 256), 256, 0, stream>>>(w.keep, n_keep, w.order, keep);
    FRCNN_LAUNCH_CHECK();
    return FRCNN_OK;
}

}  // extern "C"
