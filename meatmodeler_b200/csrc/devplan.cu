// Device-side tile plan and reduced-camera-matrix pattern: see devplan.h.
#include "devplan.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/mmba.h"

namespace mmba {

using u64 = unsigned long long;
using u32 = unsigned;

cudaError_t DevBuf::ensure(size_t bytes) {
    if (bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 4096;   // a little headroom: successive problems of similar size reuse it
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        return e;
    }
    cap = want;
    return cudaSuccess;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
void DevPlanner::release() {
    in.release();
    in2.release();
    tmp.release();
    if (d_counts) cudaFree(d_counts);
    d_counts = nullptr;
    work.release();
    work2.release();
    out.release();
    pat.release();
    if (d_info) cudaFree(d_info);
    if (h_info) cudaFreeHost(h_info);
    d_info = nullptr;
    h_info = nullptr;
}

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
inline int cdiv64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int bits_for(int64_t max_value) {   // bits needed to represent 0..max_value
    int b = 0;
    while (b < 63 && ((int64_t)1 << b) <= max_value) ++b;
    return b;
}

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan of n ints -> out[0..n] (out[n] = total); in == out is allowed (out then has n + 1 entries)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 512, kScanItems = 8, kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(kFullMask, v, off);
        if (lane >= off) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_kernel(const int* __restrict__ in, int* __restrict__ out, int n,
                                                                 int* __restrict__ bsum) {
    __shared__ int s_w[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        sum += v[k];
    }
    const int inc = warp_incl_scan(sum, lane);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < kScanThreads / 32 ? s_w[lane] : 0;
        const int wi = warp_incl_scan(w, lane);
        if (lane < kScanThreads / 32) s_w[lane] = wi - w;
        if (lane == kScanThreads / 32 - 1) bsum[blockIdx.x] = wi;
    }
    __syncthreads();
    int ex = inc - sum + s_w[warp];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

// one block: exclusive scan of bsum[0..nb) in place, bsum[nb] = total
__global__ void __launch_bounds__(1024) scan_bsum_kernel(int* __restrict__ bsum, int nb) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? bsum[i] : 0;
        const int inc = warp_incl_scan(v, lane);
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            const int wi = warp_incl_scan(w, lane);
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int ex = carry + s_w[warp] + inc - v;
        if (i < nb) bsum[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = s_carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ out, int n, const int* __restrict__ bsum, int nb) {
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    const int add = bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

constexpr int kBsumCap = 1 << 17;   // scan block sums: n up to 2^17 * 4096

void exclusive_scan(const int* in, int* out, int64_t n, int* bsum, cudaStream_t s) {
    const int nb = std::max(1, cdiv64(n, kScanChunk));
    scan_block_kernel<<<nb, kScanThreads, 0, s>>>(in, out, (int)n, bsum);
    scan_bsum_kernel<<<1, 1024, 0, s>>>(bsum, nb);
    scan_add_kernel<<<nb, kScanThreads, 0, s>>>(out, (int)n, bsum, nb);
}

// ---------------------------------------------------------------------------------------------------------------
// stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.  Every WARP owns a contiguous chunk of the
// input; a pass is histogram (per warp) -> exclusive scan over (digit, warp) -> scatter, in which a warp walks its
// chunk row by row (32 items) and ranks equal digits with match.any: positions inside a digit follow the input order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256, kSortWarps = kSortThreads / 32, kSortMaxBlocks = 592;

__global__ void __launch_bounds__(kSortThreads) rs_hist_kernel(const u64* __restrict__ keys, int64_t n, int shift, int64_t chunk,
                                                              int* __restrict__ hist, int W) {
    __shared__ int h[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const int64_t gw = (int64_t)blockIdx.x * kSortWarps + warp;
    const int64_t b = gw * chunk, e = min(n, b + chunk);
    for (int64_t i = b + lane; i < e; i += 32) atomicAdd(&h[warp][(int)((keys[i] >> shift) & 255)], 1);
    __syncwarp();
    for (int d = lane; d < 256; d += 32) hist[(int64_t)d * W + gw] = h[warp][d];
}

__global__ void __launch_bounds__(kSortThreads) rs_scatter_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals,
                                                                 u64* __restrict__ keys_out, u32* __restrict__ vals_out, int64_t n,
                                                                 int shift, int64_t chunk, const int* __restrict__ offs, int W) {
    __shared__ int cnt[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * kSortWarps + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = offs[(int64_t)d * W + gw];
    __syncwarp();
    const int64_t b = gw * chunk, e = min(n, b + chunk);
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t base = b; base < e; base += 32) {
        const int64_t i = base + lane;
        const bool valid = i < e;
        u64 k = 0;
        u32 v = 0;
        if (valid) {
            k = keys[i];
            v = vals[i];
        }
        const int d = valid ? (int)((k >> shift) & 255) : 256 + lane;
        const unsigned peers = __match_any_sync(kFullMask, d);
        const int rank = __popc(peers & lt);
        int pos = 0;
        if (valid) pos = cnt[warp][d] + rank;
        __syncwarp();
        if (valid && rank == 0) cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) {
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
    }
}

struct Sorted {
    const u64* keys;
    const u32* vals;
};

// hist: 256 * 8 * kSortMaxBlocks + 1 ints
Sorted radix_sort(u64* k0, u64* k1, u32* v0, u32* v1, int64_t n, int bits, int* hist, int* bsum, cudaStream_t s) {
    u64 *ka = k0, *kb = k1;
    u32 *va = v0, *vb = v1;
    if (n > 1) {
        const int G = (int)std::max<int64_t>(1, std::min<int64_t>(kSortMaxBlocks, (n + 2047) / 2048));
        const int W = G * kSortWarps;
        int64_t chunk = (n + W - 1) / W;
        chunk = (chunk + 31) / 32 * 32;
        for (int shift = 0; shift < bits; shift += 8) {
            rs_hist_kernel<<<G, kSortThreads, 0, s>>>(ka, n, shift, chunk, hist, W);
            exclusive_scan(hist, hist, (int64_t)256 * W, bsum, s);
            rs_scatter_kernel<<<G, kSortThreads, 0, s>>>(ka, va, kb, vb, n, shift, chunk, hist, W);
            std::swap(ka, kb);
            std::swap(va, vb);
        }
    }
    return Sorted{ka, va};
}
constexpr size_t kHistInts = (size_t)256 * kSortWarps * kSortMaxBlocks + 1;

// ---------------------------------------------------------------------------------------------------------------
// stage A: per-point statistics
// ---------------------------------------------------------------------------------------------------------------
__global__ void stats_init_kernel(int* count, int* first, int* last, int* first_hi, int n_points, int n_cams) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_points) {
        count[p] = 0;
        first[p] = n_cams;
        last[p] = 0;       // the smallest camera id is a neutral start; unobserved points are recognised by their count
        first_hi[p] = n_cams;
    }
}

// observations usually arrive grouped by point: lanes with the same point combine (match.any + warp reductions)
// before touching the per-point counters
__global__ void stats_kernel(const int* __restrict__ cam, const int* __restrict__ pt, int64_t n, int half, int* count, int* first,
                             int* last, int* first_hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    const int p = valid ? pt[i] : -1 - lane, c = valid ? cam[i] : 0;
    const unsigned peers = __match_any_sync(kFullMask, p);
    const int cmin = __reduce_min_sync(peers, c), cmax = __reduce_max_sync(peers, c);
    const int chi = __reduce_min_sync(peers, c >= half ? c : 0x7fffffff);
    if (valid && (peers & ((1u << lane) - 1u)) == 0) {
        atomicAdd(&count[p], __popc(peers));
        atomicMin(&first[p], cmin);
        atomicMax(&last[p], cmax);
        if (chi != 0x7fffffff) atomicMin(&first_hi[p], chi);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stage B: point order
// ---------------------------------------------------------------------------------------------------------------
// ring-aware key camera (plan.cpp step 1): a track spanning more than half of the camera ids is keyed by its
// smallest camera in the upper half; unobserved points get key n_cams (they sort last)
__global__ void point_key_kernel(const int* __restrict__ count, const int* __restrict__ first, const int* __restrict__ last,
                                 const int* __restrict__ first_hi, int n_points, int n_cams, int half, u64* __restrict__ keys,
                                 u32* __restrict__ vals, long long* total_pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    long long pairs = 0;
    if (p < n_points) {
        int k = n_cams;
        const int L = count[p];
        if (L) {
            k = first[p];
            if (last[p] - first[p] > half && first_hi[p] < n_cams) k = first_hi[p];
        }
        keys[p] = (u64)k;
        vals[p] = (u32)p;
        pairs = (long long)L * (L + 1) / 2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) pairs += __shfl_xor_sync(kFullMask, pairs, off);
    if ((threadIdx.x & 31) == 0 && pairs) atomicAdd((u64*)total_pairs, (u64)pairs);
}

__global__ void point_perm_kernel(const u32* __restrict__ sorted_vals, const int* __restrict__ count, int n_points,
                                  int* __restrict__ perm, int* __restrict__ inv, int* __restrict__ cnt_sorted) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_points) {
        const int p = (int)sorted_vals[q];
        perm[q] = p;
        inv[p] = q;
        cnt_sorted[q] = count[p];
    }
}

// shard cuts (plan.cpp step 3): shard_begin[r] = first internal point q with start_all[q] >= ceil(n_obs r / nranks)
__global__ void cuts_kernel(const int* __restrict__ start_all, const u64* __restrict__ sorted_keys, int n_points, int n_cams,
                            long long n_obs, int rank, int nranks, PlanInfo* info) {
    const int r = threadIdx.x;
    if (r <= nranks) {
        int cut = n_points;
        if (r == 0) cut = 0;
        else if (r < nranks) {
            const long long target = (n_obs * r + nranks - 1) / nranks;
            int lo = 0, hi = n_points;   // first q in [0, n_points) with start_all[q] >= target, else n_points
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (start_all[mid] >= target) hi = mid;
                else lo = mid + 1;
            }
            cut = lo;
        }
        info->shard_begin[r] = cut;
    }
    __syncthreads();
    if (r == 0) {
        const int b = info->shard_begin[rank], e = info->shard_begin[rank + 1];
        info->pt_begin = b;
        info->pt_end = e;
        // observed points form a prefix of the internal order
        int lo = 0, hi = n_points;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sorted_keys[mid] >= (u64)n_cams) hi = mid;
            else lo = mid + 1;
        }
        info->npo_local = max(0, min(lo, e) - b);
        info->n_obs_local = start_all[e] - start_all[b];
        info->err_track = 0;
        info->n_tiles = 0;
        info->max_tile_cams = 0;
        info->max_tile_pts = 0;
        info->nnz_up = info->nnz_full = info->nblk_max = info->nh_max = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sharded set-up: destination rank of every staged observation = owner of its point
// ---------------------------------------------------------------------------------------------------------------
__global__ void dest_key_kernel(const int* __restrict__ pt, int64_t n, const int* __restrict__ inv, const PlanInfo* __restrict__ info,
                                int nranks, u64* __restrict__ keys, u32* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q = inv[pt[i]];
    int r = 0;
    while (r + 1 < nranks && q >= info->shard_begin[r + 1]) ++r;
    keys[i] = (u64)r;
    vals[i] = (u32)i;
}
__global__ void dest_counts_kernel(const u64* __restrict__ sorted_keys, int64_t n, int nranks, int* __restrict__ counts) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    auto lower = [&](u64 v) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_keys[mid] >= v) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    counts[r] = (int)(lower((u64)r + 1) - lower((u64)r));
}
__global__ void dispatch_pack_kernel(const u32* __restrict__ sorted_vals, int64_t n, const int* __restrict__ cam, const int* __restrict__ pt,
                                     const double* __restrict__ uv, int64_t o0, int* __restrict__ cam_s, int* __restrict__ pt_s,
                                     double* __restrict__ uv_s, int* __restrict__ gidx_s) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = sorted_vals[k];
    cam_s[k] = cam[i];
    pt_s[k] = pt[i];
    uv_s[2 * k] = uv[2 * i];
    uv_s[2 * k + 1] = uv[2 * i + 1];
    gidx_s[k] = (int)(o0 + i);
}
__global__ void or_bitmaps_kernel(u64* __restrict__ bits, const u64* __restrict__ gathered, size_t n_words, int nranks) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    u64 v = bits[w];
    for (int r = 0; r < nranks; ++r) v |= gathered[(size_t)r * n_words + w];
    bits[w] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// stage C: observations -> tiles
// ---------------------------------------------------------------------------------------------------------------
__global__ void obs_key_kernel(const int* __restrict__ cam, const int* __restrict__ pt, int64_t n, const int* __restrict__ inv,
                               const PlanInfo* __restrict__ info, int cam_bits, u64* __restrict__ keys, u32* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int q = inv[pt[i]] - info->pt_begin;
        keys[i] = ((u64)(unsigned)q << cam_bits) | (u64)(unsigned)cam[i];
        vals[i] = (u32)i;
    }
}

__global__ void local_start_kernel(const int* __restrict__ start_all, const PlanInfo* __restrict__ info, int* __restrict__ start,
                                   int n_alloc) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = info->pt_begin, npl = info->pt_end - b;
    if (q < n_alloc) start[q] = q <= npl ? start_all[b + q] - start_all[b] : start_all[b + npl] - start_all[b];
}

// jump[p] = first point of the tile after the one that starts at p: the largest q <= npo with
// start[q] - start[p] <= 256 (plan.cpp step 5).  jump[npo] = npo.
__global__ void jump_init_kernel(const int* __restrict__ start, const int* __restrict__ perm, PlanInfo* info, int* __restrict__ jump,
                                 int* __restrict__ mark, int n_alloc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_alloc) return;
    const int npo = info->npo_local;
    mark[p] = (p == 0 && npo > 0) ? 1 : 0;
    if (p >= npo) {
        jump[p] = npo;
        return;
    }
    const int s0 = start[p];
    if (start[p + 1] - s0 > kTileObs) {
        atomicMax(&info->err_track, perm[info->pt_begin + p] + 1);     // point index + 1: 0 = no error, max-reduced over ranks
        jump[p] = p + 1;
        return;
    }
    int lo = p + 1, hi = npo;   // largest q in [p + 1, npo] with start[q] - s0 <= 256
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (start[mid] - s0 <= kTileObs) lo = mid;
        else hi = mid - 1;
    }
    jump[p] = lo;
}

// one pointer-doubling round: every marked point marks its 2^k-th successor; jump <- jump o jump
__global__ void jump_round_kernel(const int* __restrict__ jin, int* __restrict__ jout, int* __restrict__ mark,
                                  const PlanInfo* __restrict__ info, int n_alloc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_alloc) return;
    const int npo = info->npo_local;
    const int j = jin[p];
    if (p < npo && mark[p] && j < npo) mark[j] = 1;
    jout[p] = j >= npo ? npo : jin[j];
}

__global__ void tile_list_kernel(const int* __restrict__ mark, const int* __restrict__ tidx, PlanInfo* info, int* __restrict__ tile_p0,
                                 int n_alloc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_alloc) return;
    const int npo = info->npo_local;
    if (p < npo && mark[p]) tile_p0[tidx[p]] = p;
    if (p == npo) {
        tile_p0[tidx[p]] = npo;
        info->n_tiles = tidx[p];
    }
}

__device__ __forceinline__ int block_incl_scan_256(int v, int* s_w /* 8 */, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int off = 0, t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) off += s_w[w];
        t += s_w[w];
    }
    total = t;
    return inc + off;
}
__device__ __forceinline__ int block_incl_max_256(int v, int* s_w /* 8 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(kFullMask, v, off);
        if (lane >= off) v = max(v, t);
    }
    __syncthreads();
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    int m = v;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (w < warp) m = max(m, s_w[w]);
    return m;
}

// One CTA per tile (grid-stride): thread j = slot j.  Everything plan.cpp step 6 derives per tile.
__global__ void __launch_bounds__(kTileObs) tile_build_kernel(const int* __restrict__ tile_p0, const int* __restrict__ start,
                                                             const u64* __restrict__ keys, const u32* __restrict__ vals,
                                                             const int* __restrict__ gidx, const double* __restrict__ uv, int cam_bits,
                                                             int cam_stride, PlanInfo* info, TileMeta* __restrict__ meta,
                                                             int* __restrict__ tile_cams, double* __restrict__ uvt,
                                                             int* __restrict__ slot_obs) {
    __shared__ unsigned s_key[kTileObs];
    __shared__ int s_w[8];
    __shared__ int s_flag[2];
    const int j = threadIdx.x;
    const int n_tiles = info->n_tiles;
    const u64 cam_mask = ((u64)1 << cam_bits) - 1;
    int max_cams = 0, max_pts = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int p0 = tile_p0[t], p1 = tile_p0[t + 1];
        const int o0 = start[p0], nobs = start[p1] - o0, npts = p1 - p0;
        const bool valid = j < nobs;
        u64 key = 0;
        int obs = -1;
        if (valid) {
            key = keys[o0 + j];
            obs = (int)vals[o0 + j];
        }
        const int cam = (int)(key & cam_mask), q = (int)(key >> cam_bits);
        slot_obs[(int64_t)t * kTileObs + j] = valid ? (gidx ? gidx[obs] : obs) : -1;
        uvt[((int64_t)t * 2) * kTileObs + j] = valid ? uv[2 * (int64_t)obs] : 0.0;
        uvt[((int64_t)t * 2 + 1) * kTileObs + j] = valid ? uv[2 * (int64_t)obs + 1] : 0.0;
        if (j == 0) s_flag[0] = 0;
        // camera-sorted order: bitonic sort of (camera, slot)
        s_key[j] = valid ? ((unsigned)cam << 8) | (unsigned)j : 0xFFFFFFFFu;
        __syncthreads();
        for (int k = 2; k <= kTileObs; k <<= 1)
            for (int st = k >> 1; st > 0; st >>= 1) {
                const int other = j ^ st;
                if (other > j) {
                    const unsigned a = s_key[j], b = s_key[other];
                    const bool asc = (j & k) == 0;
                    if ((a > b) == asc) {
                        s_key[j] = b;
                        s_key[other] = a;
                    }
                }
                __syncthreads();
            }
        const unsigned sk = s_key[j];
        const bool svalid = j < nobs;
        const int scam = (int)(sk >> 8), sslot = (int)(sk & 255u);
        const bool head = svalid && (j == 0 || (int)(s_key[j - 1] >> 8) != scam);
        int ncams, nruns;
        const int lc = block_incl_scan_256(head ? 1 : 0, s_w, ncams) - 1;
        const int gs = block_incl_max_256(head ? j : 0, s_w);
        const bool runhead = svalid && ((j - gs) % kMaxRun == 0);
        const int ri = block_incl_scan_256(runhead ? 1 : 0, s_w, nruns) - 1;
        TileMeta& m = meta[t];
        // defaults of the unused entries (as plan.cpp): slot_cam 0, slot_pt 0xFFFF, sort_src j, runs 0
        m.slot_pt[j] = valid ? (uint16_t)(q - p0) : (uint16_t)0xFFFF;
        if (!valid) m.slot_cam[j] = 0;
        m.sort_src[j] = svalid ? (uint16_t)sslot : (uint16_t)j;
        if (svalid) m.slot_cam[sslot] = (uint16_t)lc;
        if (j >= nruns) {
            m.run_start[j] = 0;
            m.run_cam[j] = 0;
        }
        if (runhead) {
            m.run_start[ri] = (uint16_t)j;
            m.run_cam[ri] = (uint16_t)lc;
        }
        if (head) tile_cams[(int64_t)t * cam_stride + lc] = scam;
        for (int c = ncams + j; c < cam_stride; c += kTileObs) tile_cams[(int64_t)t * cam_stride + c] = -1;
        // duplicates (one camera observing a point twice) and the pair count sum L (L + 1) / 2
        int pairs = 0;
        if (valid) {
            const u64 prev = j > 0 ? keys[o0 + j - 1] : ~(u64)0;
            if (j > 0 && prev == key) s_flag[0] = 1;
            if (j == 0 || (int)(prev >> cam_bits) != q) {
                const int L = start[q + 1] - start[q];
                pairs = L * (L + 1) / 2;
            }
        }
        int npairs;
        block_incl_scan_256(pairs, s_w, npairs);
        __syncthreads();
        if (j == 0) {
            const bool dup = s_flag[0] != 0;
            const long long campairs = (long long)ncams * (ncams + 1) / 2;
            int mode = (!dup && (long long)npts * ncams <= kRcmTabCap && campairs * npts <= 4ll * npairs) ? 1 : 0;
            if (mode == 1 && campairs <= kTileObs) mode = 2;
            m.pt0 = p0;
            m.npts = npts;
            m.ncams = ncams;
            m.nobs = nobs;
            m.nruns = nruns;
            m.pair_mode = mode;
            m.npairs = npairs;
            m.pad = 0;
        }
        max_cams = max(max_cams, ncams);
        max_pts = max(max_pts, npts);
        __syncthreads();
    }
    if (j == 0 && max_cams) {
        atomicMax(&info->max_tile_cams, max_cams);
        atomicMax(&info->max_tile_pts, max_pts);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// co-visibility pattern
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void set_bit(u64* bits, size_t words, int i, int j) {
    u64* w = bits + (size_t)i * words + (j >> 6);
    const u64 m = (u64)1 << (j & 63);
    if (!(*reinterpret_cast<volatile u64*>(w) & m)) atomicOr(w, m);
}

// thread k = sorted observation k of a local point: marks (camera(k), camera(k')) for the later observations k' of
// the same point, both triangles
__global__ void pattern_mark_kernel(const u64* __restrict__ keys, const int* __restrict__ start, const PlanInfo* __restrict__ info,
                                    int cam_bits, u64* bits, size_t words) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= info->n_obs_local) return;
    const u64 cam_mask = ((u64)1 << cam_bits) - 1;
    const u64 key = keys[k];
    const int ci = (int)(key & cam_mask), q = (int)(key >> cam_bits);
    const int end = start[q + 1];
    for (int k2 = (int)k + 1; k2 < end; ++k2) {
        const int cj = (int)(keys[k2] & cam_mask);
        if (cj == ci) continue;
        set_bit(bits, words, ci, cj);
        set_bit(bits, words, cj, ci);
    }
}

// every camera owns its diagonal block; per row: block counts (full / upper) and per-word prefix popcounts
__global__ void pattern_count_kernel(u64* bits, size_t words, int n_cams, int* __restrict__ pc, int* __restrict__ cnt_up,
                                     int* __restrict__ cnt_full) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cams) return;
    u64* row = bits + (size_t)i * words;
    row[i >> 6] |= (u64)1 << (i & 63);
    int run = 0, below = 0;
    for (size_t w = 0; w < words; ++w) {
        pc[(size_t)i * words + w] = run;
        const u64 v = row[w];
        if ((int)w == (i >> 6)) below = run + __popcll(v & (((u64)1 << (i & 63)) - 1));
        run += __popcll(v);
    }
    cnt_full[i] = run;
    cnt_up[i] = run - below;
}

__global__ void pattern_sizes_kernel(const int* __restrict__ up_rowptr, const int* __restrict__ rowptr, int n_cams, PlanInfo* info) {
    info->nnz_up = up_rowptr[n_cams];
    info->nnz_full = rowptr[n_cams];
}

__device__ __forceinline__ int row_rank(const u64* __restrict__ row, const int* __restrict__ pcrow, int j) {
    return pcrow[j >> 6] + __popcll(row[j >> 6] & (((u64)1 << (j & 63)) - 1));
}

// one warp per row i of the full pattern (rcm.cpp: ascending columns; blocks below the diagonal read their source
// block (j, i) transposed)
__global__ void pattern_fill_kernel(const u64* __restrict__ bits, const int* __restrict__ pc, size_t words, int n_cams,
                                    const int* __restrict__ up_rowptr, const int* __restrict__ rowptr, int* __restrict__ up_cols,
                                    int* __restrict__ cols, int* __restrict__ rows, int* __restrict__ src, int* __restrict__ diag) {
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= n_cams) return;
    const u64* row = bits + (size_t)i * words;
    const int* pcrow = pc + (size_t)i * words;
    const int below = row_rank(row, pcrow, i);
    for (size_t w = lane; w < words; w += 32) {
        u64 v = row[w];
        int e = rowptr[i] + pcrow[w];
        while (v) {
            const int j = (int)(w * 64) + __ffsll((long long)v) - 1;
            v &= v - 1;
            cols[e] = j;
            rows[e] = i;
            if (j >= i) {
                const int u = up_rowptr[i] + (e - rowptr[i]) - below;
                src[e] = u;
                up_cols[u] = j;
                if (j == i) diag[i] = e;
            } else {
                const u64* rj = bits + (size_t)j * words;
                const int* pj = pc + (size_t)j * words;
                const int u = up_rowptr[j] + row_rank(rj, pj, i) - row_rank(rj, pj, j);
                src[e] = (int)((unsigned)u | 0x80000000u);
            }
            ++e;
        }
    }
}

// PCG partition: CTA b owns cameras [b cpc, (b + 1) cpc); its halo = union of the columns of its rows
__global__ void halo_or_kernel(const u64* __restrict__ bits, size_t words, int n_cams, int cpc, u64* __restrict__ hbits,
                               int* __restrict__ hcnt, const int* __restrict__ rowptr, PlanInfo* info) {
    const int b = blockIdx.x;
    const int c0 = b * cpc, c1 = min(n_cams, c0 + cpc);
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int local = 0;
    for (size_t w = threadIdx.x; w < words; w += blockDim.x) {
        u64 v = 0;
        for (int c = c0; c < c1; ++c) v |= bits[(size_t)c * words + w];
        hbits[(size_t)b * words + w] = v;
        local += __popcll(v);
    }
    atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) {
        hcnt[b] = s_cnt;
        atomicMax(&info->nh_max, s_cnt);
        atomicMax(&info->nblk_max, rowptr[c1] - rowptr[c0]);
    }
}

__global__ void halo_ptr_kernel(const int* __restrict__ hcnt, int g, int* __restrict__ halo_ptr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < g; ++b) {
            halo_ptr[b] = acc;
            acc += hcnt[b];
        }
        halo_ptr[g] = acc;
    }
}

__global__ void halo_emit_kernel(const u64* __restrict__ hbits, size_t words, int n_cams, int cpc, const int* __restrict__ halo_ptr,
                                 const int* __restrict__ rowptr, const int* __restrict__ cols, int* __restrict__ halo_cols,
                                 uint16_t* __restrict__ lcol, int* __restrict__ own_l) {
    extern __shared__ int s_pre[];   // [words] prefix popcounts of the CTA's halo bitmap
    const int b = blockIdx.x;
    const int c0 = b * cpc, c1 = min(n_cams, c0 + cpc);
    const u64* hb = hbits + (size_t)b * words;
    if (threadIdx.x == 0) {
        int run = 0;
        for (size_t w = 0; w < words; ++w) {
            s_pre[w] = run;
            run += __popcll(hb[w]);
        }
    }
    __syncthreads();
    const int h0 = halo_ptr[b];
    for (size_t w = threadIdx.x; w < words; w += blockDim.x) {
        u64 v = hb[w];
        int o = h0 + s_pre[w];
        while (v) {
            halo_cols[o++] = (int)(w * 64) + __ffsll((long long)v) - 1;
            v &= v - 1;
        }
    }
    auto rank = [&](int j) { return s_pre[j >> 6] + __popcll(hb[j >> 6] & (((u64)1 << (j & 63)) - 1)); };
    for (int e = rowptr[c0] + threadIdx.x; e < rowptr[c1]; e += blockDim.x) lcol[e] = (uint16_t)rank(cols[e]);
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) own_l[c] = rank(c);
}

// ---------------------------------------------------------------------------------------------------------------
// parameter / residual permutations between the caller's order and the internal layout
// ---------------------------------------------------------------------------------------------------------------
__global__ void gather_x_kernel(const double* __restrict__ xc, double* __restrict__ xi, const int* __restrict__ perm, int64_t ncam6,
                                int64_t pt_begin, int64_t npl) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ncam6) xi[e] = xc[e];
    else if (e < ncam6 + 3 * npl) {
        const int64_t q = (e - ncam6) / 3, k = (e - ncam6) - 3 * q;
        xi[e] = xc[ncam6 + 3 * (int64_t)perm[pt_begin + q] + k];
    }
}
__global__ void scatter_x_kernel(const double* __restrict__ xi, double* __restrict__ xc, const int* __restrict__ perm, int64_t ncam6,
                                 int64_t pt_begin, int64_t npl, int with_cams) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ncam6) {
        if (with_cams) xc[e] = xi[e];
    } else if (e < ncam6 + 3 * npl) {
        const int64_t q = (e - ncam6) / 3, k = (e - ncam6) - 3 * q;
        xc[ncam6 + 3 * (int64_t)perm[pt_begin + q] + k] = xi[e];
    }
}
__global__ void scatter_slots_kernel(const double* __restrict__ src, int src_rows, int row0, int rows,
                                     const int* __restrict__ slot_obs, int64_t n_slots, double* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int o = slot_obs[s];
    if (o < 0) return;
    const int64_t t = s / kTileObs, j = s - t * kTileObs;
    for (int r = 0; r < rows; ++r) out[(int64_t)o * rows + r] = src[(t * src_rows + row0 + r) * kTileObs + j];
}

#define DP_CU(call)                                                             \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);           \
            return MMBA_ERR_CUDA;                                               \
        }                                                                       \
    } while (0)

int ensure_info(DevPlanner& P, std::string& err) {
    if (!P.d_info) DP_CU(cudaMalloc(&P.d_info, sizeof(PlanInfo)));
    if (!P.h_info) DP_CU(cudaMallocHost(&P.h_info, sizeof(PlanInfo)));
    return MMBA_OK;
}

}  // namespace

// ===============================================================================================================
// stages
// ===============================================================================================================
int devplan_stats(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err) {
    if (D.n_cams <= 0 || D.n_points <= 0 || D.n_obs <= 0) {
        err = "set_problem: sizes must be positive";
        return MMBA_ERR_ARG;
    }
    if (D.n_cams >= (1 << 24) || D.n_points > INT32_MAX / 16 || D.n_obs > INT32_MAX - 4096) {
        err = "set_problem: problem too large for 32-bit device indices";
        return MMBA_ERR_ARG;
    }
    if (D.nranks > 16) {
        err = "set_problem: at most 16 ranks";
        return MMBA_ERR_ARG;
    }
    int rc = ensure_info(P, err);
    if (rc != MMBA_OK) return rc;
    const size_t np = (size_t)D.n_points;
    for (int pass = 0; pass < 2; ++pass) {
        Carver c;
        c.base = pass ? P.work.p : nullptr;
        // one block [count | first | last | first_hi], each np_pad ints: a sharded set-up all-gathers it as a whole
        const size_t np_pad = (np + 63) / 64 * 64;
        P.a.count = c.take<int>(4 * np_pad);
        P.a.first = P.a.count + np_pad;
        P.a.last = P.a.first + np_pad;
        P.a.first_hi = P.a.last + np_pad;
        P.a.key = c.take<int>(np + 1);
        P.a.start_all = c.take<int>(np + 2);
        P.a.hist = c.take<int>(kHistInts);
        P.a.bsum = c.take<int>(kBsumCap + 2);
        P.a.k0 = c.take<u64>(np);
        P.a.k1 = c.take<u64>(np);
        P.a.v0 = c.take<u32>(np);
        P.a.v1 = c.take<u32>(np);
        D.point_perm = c.take<int32_t>(np);
        D.point_inv = c.take<int32_t>(np);
        if (!pass) DP_CU(P.work.ensure(c.off + 256));
    }
    const int nb = cdiv64(D.n_points, 256);
    stats_init_kernel<<<nb, 256, 0, s>>>(P.a.count, P.a.first, P.a.last, P.a.first_hi, (int)D.n_points, (int)D.n_cams);
    if (P.n_in > 0)
        stats_kernel<<<cdiv64(P.n_in, 256), 256, 0, s>>>(P.cam, P.pt, P.n_in, (int)(D.n_cams / 2), P.a.count, P.a.first, P.a.last,
                                                        P.a.first_hi);
    DP_CU(cudaGetLastError());
    return MMBA_OK;
}

// combine the all-gathered per-rank statistics blocks: counts add, first / first_hi take the minimum, last the maximum
__global__ void combine_stats_kernel(const int* __restrict__ gathered, int nranks, size_t np_pad, int n_points, int* __restrict__ count,
                                     int* __restrict__ first, int* __restrict__ last, int* __restrict__ first_hi) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    int c = 0, f = 0x7fffffff, l = 0, fh = 0x7fffffff;
    for (int r = 0; r < nranks; ++r) {
        const int* blk = gathered + (size_t)r * 4 * np_pad;
        c += blk[p];
        f = min(f, blk[np_pad + p]);
        l = max(l, blk[2 * np_pad + p]);
        fh = min(fh, blk[3 * np_pad + p]);
    }
    count[p] = c;
    first[p] = f;
    last[p] = l;
    first_hi[p] = fh;
}

void devplan_stat_block(DevPlanner& P, const DevPlan& D, int** block, size_t* n_ints) {
    *block = P.a.count;
    *n_ints = 4 * (((size_t)D.n_points + 63) / 64 * 64);
}

void devplan_combine_stats(DevPlanner& P, const DevPlan& D, const int* gathered, int nranks, cudaStream_t s) {
    const size_t np_pad = ((size_t)D.n_points + 63) / 64 * 64;
    combine_stats_kernel<<<cdiv64(D.n_points, 256), 256, 0, s>>>(gathered, nranks, np_pad, (int)D.n_points, P.a.count, P.a.first,
                                                                 P.a.last, P.a.first_hi);
}

void devplan_stat_arrays(DevPlanner& P, const DevPlan&, int** count, int** first, int** last, int** first_hi) {
    *count = P.a.count;
    *first = P.a.first;
    *last = P.a.last;
    *first_hi = P.a.first_hi;
}

int devplan_order(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err) {
    const int np = (int)D.n_points, nb = cdiv64(np, 256);
    DP_CU(cudaMemsetAsync(&P.d_info->total_pairs, 0, sizeof(long long), s));
    point_key_kernel<<<nb, 256, 0, s>>>(P.a.count, P.a.first, P.a.last, P.a.first_hi, np, (int)D.n_cams, (int)(D.n_cams / 2), P.a.k0,
                                        P.a.v0, &P.d_info->total_pairs);
    const Sorted so = radix_sort(P.a.k0, P.a.k1, P.a.v0, P.a.v1, np, bits_for(D.n_cams), P.a.hist, P.a.bsum, s);
    point_perm_kernel<<<nb, 256, 0, s>>>(so.vals, P.a.count, np, D.point_perm, D.point_inv, P.a.key);
    exclusive_scan(P.a.key, P.a.start_all, np, P.a.bsum, s);
    cuts_kernel<<<1, 32, 0, s>>>(P.a.start_all, so.keys, np, (int)D.n_cams, (long long)D.n_obs, D.rank, D.nranks, P.d_info);
    DP_CU(cudaGetLastError());
    return MMBA_OK;
}

int devplan_dispatch_pack(DevPlanner& P, DevPlan& D, int64_t o0, cudaStream_t s, std::string& err) {
    const int64_t n = P.n_in;
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    if (!P.d_counts) DP_CU(cudaMalloc(&P.d_counts, (16 + 16 * 16) * sizeof(int)));
    u64 *k0 = nullptr, *k1 = nullptr;
    u32 *v0 = nullptr, *v1 = nullptr;
    int *hist = nullptr, *bsum = nullptr;
    for (int pass = 0; pass < 2; ++pass) {   // scratch of this step; devplan_tiles carves work2 afresh afterwards
        Carver c;
        c.base = pass ? P.work2.p : nullptr;
        k0 = c.take<u64>(nn);
        k1 = c.take<u64>(nn);
        v0 = c.take<u32>(nn);
        v1 = c.take<u32>(nn);
        hist = c.take<int>(kHistInts);
        bsum = c.take<int>(kBsumCap + 2);
        if (!pass) DP_CU(P.work2.ensure(c.off + 256));
    }
    DP_CU(cudaMemsetAsync(P.d_counts, 0, 16 * sizeof(int), s));
    if (n > 0) {
        dest_key_kernel<<<cdiv64(n, 256), 256, 0, s>>>(P.pt, n, D.point_inv, P.d_info, D.nranks, k0, v0);
        const Sorted so = radix_sort(k0, k1, v0, v1, n, std::max(1, bits_for(D.nranks - 1)), hist, bsum, s);
        dest_counts_kernel<<<1, 32, 0, s>>>(so.keys, n, D.nranks, P.d_counts);
        dispatch_pack_kernel<<<cdiv64(n, 256), 256, 0, s>>>(so.vals, n, P.cam, P.pt, P.uv, o0, P.cam_s, P.pt_s, P.uv_s, P.gidx_s);
    }
    DP_CU(cudaGetLastError());
    return MMBA_OK;
}

int devplan_dispatch_recv(DevPlanner& P, int64_t n_recv, std::string& err) {
    const size_t nn = (size_t)std::max<int64_t>(n_recv, 1);
    for (int pass = 0; pass < 2; ++pass) {
        Carver c;
        c.base = pass ? P.in2.p : nullptr;
        P.cam_r = c.take<int32_t>(nn);
        P.pt_r = c.take<int32_t>(nn);
        P.gidx_r = c.take<int32_t>(nn);
        P.uv_r = c.take<double>(2 * nn);
        if (!pass) DP_CU(P.in2.ensure(c.off + 256));
    }
    P.cam = P.cam_r;
    P.pt = P.pt_r;
    P.uv = P.uv_r;
    P.gidx = P.gidx_r;
    P.n_in = n_recv;
    return MMBA_OK;
}

void devplan_or_bitmaps(unsigned long long* bits, const unsigned long long* gathered, size_t n_words, int nranks, cudaStream_t s) {
    or_bitmaps_kernel<<<cdiv64((int64_t)n_words, 256), 256, 0, s>>>(bits, gathered, n_words, nranks);
}

int devplan_tiles(DevPlanner& P, DevPlan& D, bool want_pattern, cudaStream_t s, std::string& err) {
    const int64_t n = P.n_in;                      // local observations
    const size_t np = (size_t)D.n_points;
    const int cam_bits = std::max(1, bits_for(D.n_cams - 1));
    const int key_bits = cam_bits + bits_for(D.n_points - 1);
    const size_t words = (size_t)(D.n_cams + 63) / 64;
    const int64_t tiles_max = 2 * n / kTileObs + 2;
    D.cam_stride = (int)std::min<int64_t>(kTileObs, std::max<int64_t>(4, (D.n_cams + 3) / 4 * 4));
    P.c.words = words;
    P.c.tiles_max = tiles_max;
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    for (int pass = 0; pass < 2; ++pass) {
        Carver c;
        c.base = pass ? P.work2.p : nullptr;
        P.c.k0 = c.take<u64>(nn);
        P.c.k1 = c.take<u64>(nn);
        P.c.v0 = c.take<u32>(nn);
        P.c.v1 = c.take<u32>(nn);
        P.c.hist = c.take<int>(kHistInts);
        P.c.bsum = c.take<int>(kBsumCap + 2);
        P.c.start = c.take<int>(np + 2);
        P.c.jump0 = c.take<int>(np + 2);
        P.c.jump1 = c.take<int>(np + 2);
        P.c.mark = c.take<int>(np + 2);
        P.c.tidx = c.take<int>(np + 3);
        P.c.tile_p0 = c.take<int>((size_t)tiles_max + 2);
        if (want_pattern) {
            P.c.bits = c.take<u64>((size_t)D.n_cams * words);
            P.c.pc = c.take<int>((size_t)D.n_cams * words);
            P.c.cnt_up = c.take<int>((size_t)D.n_cams + 2);
            P.c.cnt_full = c.take<int>((size_t)D.n_cams + 2);
            P.c.hbits = c.take<u64>((size_t)256 * words);
            P.c.hbits_cnt = c.take<int>(256);
        } else {
            P.c.bits = nullptr;
        }
        if (!pass) DP_CU(P.work2.ensure(c.off + 256));
    }
    for (int pass = 0; pass < 2; ++pass) {
        Carver c;
        c.base = pass ? P.out.p : nullptr;
        D.meta = c.take<TileMeta>((size_t)tiles_max);
        D.tile_cams = c.take<int32_t>((size_t)tiles_max * D.cam_stride);
        D.uvt = c.take<double>((size_t)tiles_max * 2 * kTileObs);
        D.slot_obs = c.take<int32_t>((size_t)tiles_max * kTileObs);
        if (!pass) DP_CU(P.out.ensure(c.off + 256));
    }
    const int n_alloc = (int)np + 1;
    Sorted so{P.c.k0, P.c.v0};
    if (n > 0) {
        obs_key_kernel<<<cdiv64(n, 256), 256, 0, s>>>(P.cam, P.pt, n, D.point_inv, P.d_info, cam_bits, P.c.k0, P.c.v0);
        so = radix_sort(P.c.k0, P.c.k1, P.c.v0, P.c.v1, n, key_bits, P.c.hist, P.c.bsum, s);
    }
    P.c.keys_sorted = so.keys;
    P.c.vals_sorted = so.vals;
    local_start_kernel<<<cdiv64(n_alloc + 1, 256), 256, 0, s>>>(P.a.start_all, P.d_info, P.c.start, n_alloc + 1);
    jump_init_kernel<<<cdiv64(n_alloc, 256), 256, 0, s>>>(P.c.start, D.point_perm, P.d_info, P.c.jump0, P.c.mark, n_alloc);
    int* jin = P.c.jump0;
    int* jout = P.c.jump1;
    const int rounds = bits_for(tiles_max) + 1;
    for (int r = 0; r < rounds; ++r) {
        jump_round_kernel<<<cdiv64(n_alloc, 256), 256, 0, s>>>(jin, jout, P.c.mark, P.d_info, n_alloc);
        std::swap(jin, jout);
    }
    exclusive_scan(P.c.mark, P.c.tidx, n_alloc, P.c.bsum, s);
    tile_list_kernel<<<cdiv64(n_alloc, 256), 256, 0, s>>>(P.c.mark, P.c.tidx, P.d_info, P.c.tile_p0, n_alloc);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles_max, 148 * 8));
    tile_build_kernel<<<grid, kTileObs, 0, s>>>(P.c.tile_p0, P.c.start, so.keys, so.vals, P.gidx, P.uv, cam_bits, D.cam_stride, P.d_info,
                                                D.meta, D.tile_cams, D.uvt, D.slot_obs);
    if (want_pattern) {
        DP_CU(cudaMemsetAsync(P.c.bits, 0, (size_t)D.n_cams * words * sizeof(u64), s));
        if (n > 0) pattern_mark_kernel<<<cdiv64(n, 256), 256, 0, s>>>(so.keys, P.c.start, P.d_info, cam_bits, P.c.bits, words);
    }
    DP_CU(cudaGetLastError());
    return MMBA_OK;
}

void devplan_bitmap(DevPlanner& P, const DevPlan& D, unsigned long long** bits, size_t* n_words) {
    *bits = P.c.bits;
    *n_words = (size_t)D.n_cams * P.c.words;
}

int devplan_pattern_sizes(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err) {
    const int nc = (int)D.n_cams;
    pattern_count_kernel<<<cdiv64(nc, 128), 128, 0, s>>>(P.c.bits, P.c.words, nc, P.c.pc, P.c.cnt_up, P.c.cnt_full);
    exclusive_scan(P.c.cnt_up, P.c.cnt_up, nc, P.c.bsum, s);
    exclusive_scan(P.c.cnt_full, P.c.cnt_full, nc, P.c.bsum, s);
    pattern_sizes_kernel<<<1, 1, 0, s>>>(P.c.cnt_up, P.c.cnt_full, nc, P.d_info);
    DP_CU(cudaGetLastError());
    return MMBA_OK;
}

int devplan_sync_sizes(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err) {
    DP_CU(cudaMemcpyAsync(P.h_info, P.d_info, sizeof(PlanInfo), cudaMemcpyDeviceToHost, s));
    DP_CU(cudaStreamSynchronize(s));
    const PlanInfo& I = *P.h_info;
    if (I.err_track > 0) {
        err = "set_problem: point " + std::to_string(I.err_track - 1) + " has more observations than one tile holds (" +
              std::to_string(kTileObs) + ")";
        return MMBA_ERR_TRACK;
    }
    D.pt_begin = I.pt_begin;
    D.pt_end = I.pt_end;
    D.n_obs_local = I.n_obs_local;
    D.n_tiles = I.n_tiles;
    D.n_slots = (int64_t)I.n_tiles * kTileObs;
    D.max_tile_cams = I.max_tile_cams;
    D.max_tile_pts = I.max_tile_pts;
    D.nnz_up = I.nnz_up;
    D.nnz_full = I.nnz_full;
    D.total_pairs = I.total_pairs;
    D.nblk_max = I.nblk_max;
    D.nh_max = I.nh_max;
    if (D.n_tiles > P.c.tiles_max) {
        err = "set_problem: internal error, tile bound exceeded";
        return MMBA_ERR_STATE;
    }
    return MMBA_OK;
}

int devplan_pattern_fill(DevPlanner& P, DevPlan& D, int max_ctas, cudaStream_t s, std::string& err) {
    const int nc = (int)D.n_cams;
    const size_t words = P.c.words;
    // partition (rcm.cpp build_rcm_partition): a few cameras per CTA even for small problems
    int64_t g = std::max<int64_t>(1, std::min<int64_t>(max_ctas, (nc + 3) / 4));
    const int64_t cpc = (nc + g - 1) / g;
    g = (nc + cpc - 1) / cpc;
    D.n_ctas = (int)g;
    D.cpc = (int)cpc;
    for (int pass = 0; pass < 2; ++pass) {
        Carver c;
        c.base = pass ? P.pat.p : nullptr;
        D.up_rowptr = c.take<int>((size_t)nc + 1);
        D.up_cols = c.take<int>((size_t)D.nnz_up);
        D.rowptr = c.take<int>((size_t)nc + 1);
        D.cols = c.take<int>((size_t)D.nnz_full);
        D.rows = c.take<int>((size_t)D.nnz_full);
        D.src = c.take<int>((size_t)D.nnz_full);
        D.diag = c.take<int>((size_t)nc);
        D.halo_ptr = c.take<int>((size_t)g + 1);
        D.halo_cols = c.take<int>((size_t)D.nnz_full + 1);
        D.own_l = c.take<int>((size_t)nc);
        D.lcol = c.take<uint16_t>((size_t)D.nnz_full);
        if (!pass) DP_CU(P.pat.ensure(c.off + 256));
    }
    DP_CU(cudaMemcpyAsync(D.up_rowptr, P.c.cnt_up, ((size_t)nc + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s));
    DP_CU(cudaMemcpyAsync(D.rowptr, P.c.cnt_full, ((size_t)nc + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s));
    pattern_fill_kernel<<<cdiv64((int64_t)nc * 32, 256), 256, 0, s>>>(P.c.bits, P.c.pc, words, nc, D.up_rowptr, D.rowptr, D.up_cols, D.cols,
                                                                     D.rows, D.src, D.diag);
    halo_or_kernel<<<(int)g, 256, 0, s>>>(P.c.bits, words, nc, (int)cpc, P.c.hbits, P.c.hbits_cnt, D.rowptr, P.d_info);
    halo_ptr_kernel<<<1, 32, 0, s>>>(P.c.hbits_cnt, (int)g, D.halo_ptr);
    halo_emit_kernel<<<(int)g, 256, words * sizeof(int), s>>>(P.c.hbits, words, nc, (int)cpc, D.halo_ptr, D.rowptr, D.cols, D.halo_cols,
                                                             D.lcol, D.own_l);
    DP_CU(cudaGetLastError());
    DP_CU(cudaMemcpyAsync(P.h_info, P.d_info, sizeof(PlanInfo), cudaMemcpyDeviceToHost, s));
    DP_CU(cudaStreamSynchronize(s));
    D.nblk_max = P.h_info->nblk_max;
    D.nh_max = P.h_info->nh_max;
    return MMBA_OK;
}

void devplan_gather_x(const double* x_caller, double* x_internal, const int32_t* point_perm, int64_t n_cams, int64_t pt_begin, int64_t npl,
                      cudaStream_t s) {
    const int64_t n = 6 * n_cams + 3 * npl;
    gather_x_kernel<<<cdiv64(n, 256), 256, 0, s>>>(x_caller, x_internal, point_perm, 6 * n_cams, pt_begin, npl);
}
void devplan_scatter_x(const double* x_internal, double* x_caller, const int32_t* point_perm, int64_t n_cams, int64_t pt_begin,
                       int64_t npl, bool with_cams, cudaStream_t s) {
    const int64_t n = 6 * n_cams + 3 * npl;
    scatter_x_kernel<<<cdiv64(n, 256), 256, 0, s>>>(x_internal, x_caller, point_perm, 6 * n_cams, pt_begin, npl, with_cams ? 1 : 0);
}
void devplan_scatter_slots(const double* src, int src_rows, int row0, int rows, const int32_t* slot_obs, int64_t n_slots, double* out,
                           cudaStream_t s) {
    if (n_slots <= 0) return;
    scatter_slots_kernel<<<cdiv64(n_slots, 256), 256, 0, s>>>(src, src_rows, row0, rows, slot_obs, n_slots, out);
}

}  // namespace mmba
