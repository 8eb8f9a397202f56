// Device code of libmmba.so — float64 CUDA kernels for sm_100a (B200): the observation-streaming
// kernels.  (Small-vector kernels: veckernels.cuh.)
//
// Data layout in HBM, per shard (built on the device by devplan.cu; plan.cpp is the host statement of the same plan)
//   * observations are reordered into point-aligned TILES of kT = 256 slots; a point's observations
//     are contiguous and never straddle a tile.  Everything per-observation is tile-major, so one
//     tile is one contiguous block that a single TMA bulk copy moves:
//       Jt   [tile][18][256] f64   36 864 B/tile   rows 0..11 = 2x6 camera block (row-major, columns
//                                                  w0 w1 w2 t0 t1 t2), rows 12..17 = 2x3 point block
//       uv   [tile][2][256]  f64    4 096 B/tile   observed pixel
//       res  [tile][2][256]  f64                   residual (du, dv)
//       ju1  [tile][2][256]  f64                   J u1 of the JV1 pass (read back by BACKSUB for the subspace Gram sums)
//       meta [tile] TileMeta        2 592 B/tile   header + 5 x u16 per slot: local camera slot, local point, the
//                                                  tile's camera-sorted order and its camera runs (start, camera)
//       tile_cams [tile][cam_stride] i32           global camera ids of the tile's local camera slots
//   * cameras: camtab[Nc][24] = R (9) | t (3) | Q (9) from cam_prep_kernel (rotate's trigonometry
//     hoisted from per-observation to per-camera); camera vectors are [Nc][6].
//   * points (internal order, tile-contiguous): x_p[3], V[6], g_p[3], M[6] (damped inverse), zg[3] = M g_p
//
// Kernel structure: ONE persistent, warp-specialised kernel template (tile_kernel<MODE>).  Each CTA
// owns a contiguous range of tiles and runs a 2..4-stage mbarrier pipeline:
//   * producer warps (one per stage): cp.async.bulk (TMA bulk copy, SASS UBLKCP) of the tile's J block,
//     metadata (and uv / ju1) into shared memory, plus the gather of the cameras the tile touches (rotation rows
//     or vector entries) and of the tile's point payloads, completing on the stage's "full" barrier;
//   * 8 consumer warps: one thread per observation slot; per-point sums by warp-shuffle segmented
//     reduction over the contiguous point runs, per-camera sums by staging the values in shared memory and
//     summing the tile's precomputed camera runs, one f64 RED per (run, component).  The S-build (SBUILD)
//     instead accumulates whole 6x6 blocks of the reduced camera matrix in registers across tiles.
// Every pass is therefore a streaming pass at 24..184 B/observation with all latency (tile header ->
// camera list -> camera rows) hidden behind the previous tile's arithmetic.
//
// Reference sites replaced: rotate/project/pointFun (bundleAdjuster.py:7-52, 81-102), scipy's
// finite-difference Jacobian (_numdiff.py:770-893), J^T J / J^T f (common.py:590-610) and the
// matvec/rmatvec pair inside LSMR (lsmr.py:337,343).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pcg.cuh"
#include "plan.h"

namespace mmba {

constexpr int kT = kTileObs;
constexpr int kConsumers = kT;            // consumer threads per CTA (one per slot)
constexpr int kMaxStages = 4;             // pipeline depth is per MODE (Traits::kStages); one producer warp per stage
constexpr int kCamTab = 24;               // doubles per camera-table row in HBM
constexpr unsigned kFull = 0xffffffffu;
constexpr uint16_t kPadPt = 0xFFFF;
constexpr int kProducerShare = 1024;      // gathered values per tile the producer warp stages itself
constexpr int kJRows = 18;
constexpr int kJTileBytes = kJRows * kT * 8;
constexpr int kUVTileBytes = 2 * kT * 8;

static_assert(sizeof(TileMeta) == 2592 && sizeof(TileMeta) % 16 == 0, "TileMeta is bulk-copied: 16-byte multiple");
static_assert(9 * (kTileObs + 1) * 8 <= 18 * kTileObs * 8, "scatter staging rows must fit inside a J block");
constexpr int kBufStride = kT + 1;         // row stride of the scatter staging buffer (bank spread)
// S-build, register-accumulation strategy (pair mode 2): the (point, camera) -> slot table has two parities of
// kRcmTab2Cap entries {generation << 8 | slot}; a tile whose table would not fit is handled by strategy 1.
constexpr int kRcmTab2Cap = kRcmTabCap / 2;
// ints of the accumulation-run scratch: cameras [0..31], per parity {flags, missing} [32..35], per parity the
// run index -> local camera slot map [40..103], block index of every pair of the run (flush) [112..239], per parity
// the camera list of the tile [240..303]
constexpr int kRunInts = 304;
constexpr int kRunMap = 40, kRunBlk = 112, kRunPrev = 240;

// scalar slots in device memory (doubles).  Groups that are reduced across ranks together are
// contiguous: [S_COST] sum, [S_GH2..S_X2] sum, S_GINF max, S_COST_NEW sum, [S_JV00..S_JV11] sum,
// [S_DOT0..S_DOT9] sum.
enum Scal {
    S_COST = 0,      // sum r^2 (build)
    S_GH2,           // ||g_h||^2
    S_XSI2,          // sum (x * scale_inv)^2
    S_X2,            // ||x||^2
    S_GINF,          // ||g||_inf as the bit pattern of a non-negative double (atomicMax on u64)
    S_COST_NEW,      // sum r^2 (trial)
    S_JV00, S_JV01, S_JV11,   // Gram of J*v products
    S_DOT0, S_DOT1, S_DOT2, S_DOT3, S_DOT4, S_DOT5, S_DOT6, S_DOT7, S_DOT8, S_DOT9,   // subspace dots
    S_PEER_ERR = 23,  // set by peer_allreduce_small_kernel when a rank did not arrive
    S_COUNT = 24
};

enum Mode { M_BUILD = 0, M_RESID, M_RESID_STORE, M_MATVEC, M_RHS, M_BACKSUB, M_JV1, M_JV2, M_BUILD_FULL, M_SBUILD, M_COUNT };
__host__ __device__ constexpr bool is_build(int m) { return m == M_BUILD || m == M_BUILD_FULL; }
__host__ __device__ constexpr bool is_project(int m) { return is_build(m) || m == M_RESID || m == M_RESID_STORE; }

template <int MODE>
struct Traits {
    static constexpr bool kLoadJ = !is_project(MODE);
    // J-streaming tiles are 39 KB: two stages per CTA, two CTAs per SM.  The projection passes (BUILD, RESID)
    // stage only 7-13 KB per tile but gather 12-21 doubles per camera, so their per-tile latency chains need
    // more tiles (and producer warps) in flight.
    static constexpr int kStages = is_project(MODE) ? 4 : MODE == M_SBUILD ? 3 : 2;   // SBUILD: one CTA per SM
    static constexpr int kThreads = kConsumers + 32 * kStages;
    static constexpr bool kLoadUV = !kLoadJ;
    // BACKSUB also takes the Gram sums of J [u1 u2] (u2 = the step it completes): J u1 per observation, stored by the
    // JV1 pass, comes with the tile (16 B/observation instead of a third pass over J)
    static constexpr bool kLoadAux = MODE == M_BACKSUB;
    static constexpr int kPtBufs = MODE == M_BACKSUB ? 3 : 2;     // rotating per-point accumulators
    // doubles gathered per camera into shared memory, and the (odd, conflict-free) smem stride
    static constexpr int kCamRows = is_build(MODE) ? 21 : (MODE == M_RESID || MODE == M_RESID_STORE) ? 12
                                  : (MODE == M_RHS || MODE == M_SBUILD) ? 0 : MODE == M_JV2 ? 12 : 6;
    static constexpr int kCamStride = kCamRows | 1;
    // per-point payloads staged by the producer
    static constexpr int kPA = (MODE == M_MATVEC || MODE == M_RHS || MODE == M_BACKSUB || MODE == M_SBUILD) ? 6 : 3;
    static constexpr int kPB = (MODE == M_RHS || MODE == M_BACKSUB || MODE == M_JV2 || MODE == M_SBUILD) ? 3 : 0;
    static constexpr bool kScatter = is_build(MODE) || MODE == M_MATVEC || MODE == M_RHS;
    static constexpr int kPtAcc = (MODE == M_MATVEC || MODE == M_BACKSUB) ? 3 : 0;
    // scatter staging rows: BUILD stages 12 camera + 9 point rows + 1 row of point-run starts in its
    // own buffer.  The Schur passes stage 6 / 9 rows INSIDE the current pipeline stage's J block:
    // every consumer holds its 18 J values in registers by then, the block is dead until the stage
    // is released, and consecutive tiles use different stages (free double buffering).
    // SBUILD keeps the J block intact for its pair phase and stages two parities of 8 rows per observation
    // (6 rows Jp M, 2 rows v = Jp M g_p), so that consecutive tiles need ONE consumer barrier each; the per-tile
    // strategies (pair modes 0 / 1) use rows 0..5 for the right-hand-side scatter and rows 8..13 for Jp M.
    static constexpr int kStageRows = is_build(MODE) ? 22 : MODE == M_SBUILD ? 16 : 0;
    static constexpr int kStageBufs = 1;
    static constexpr int kMinBlocks = MODE == M_SBUILD ? 1 : 2;   // SBUILD: one CTA per SM, up to 204 registers per thread
};

struct TileArgs {
    const TileMeta* meta;
    const int32_t* tile_cams;
    const double* uv;
    int n_tiles, cam_stride, max_cams, max_pts;
    int n_cams, ytab_cams;   // ytab_cams = n_cams when MATVEC keeps a per-CTA camera table in shared memory, else 0
    int sb_stages;           // SBUILD: pipeline stages in use (3, or 2 when the tiles' point payloads are large)
    long long* dbg;          // optional phase cycle counters (diagnostics, options.profile bit 2), else nullptr
    double K[9];
};

// operands of one launch; which ones are read depends on MODE
struct ModeArgs {
    const double* J;       // Jt (read)                      MATVEC RHS BACKSUB JV
    double* Jw;            // Jt (written)                   BUILD
    double* res;           // residuals                      BUILD RESID_STORE
    const double* cam0;    // camtab / xt / vc0              gathered per tile camera
    const double* cam1;    // vc1                            JV2
    const double* ptA;     // x_p / M / vp0
    const double* ptB;     // zg / g_p / vp1
    double* y;             // [Nc][6]  scatter target        MATVEC RHS
    double* Sd;            // [Nc][21]                       RHS
    double* U;             // [Nc][21] full camera blocks     BUILD_FULL (evaluation hook only)
    double* Ud;            // [Nc][6]  diag(J_c^T J_c)        BUILD
    double* gc;            // [Nc][6]                        BUILD
    double* V;             // [Np][6]                        BUILD
    double* gp;            // [Np][3]                        BUILD
    double* dp;            // [Np][3]                        BACKSUB
    double* Tup;           // [nnz_up][36] reduced camera matrix, upper blocks, unscaled   SBUILD
    const int* up_rowptr;  // [Nc + 1]  block pattern of Tup (rcm.h)                       SBUILD
    const int* up_cols;    // [nnz_up]
    double* scal;          // scalar slots                   BUILD (S_COST) JV (S_JV*)
    double* cost;          // trial cost slot                RESID
    const int* done;       // PCG converged flag             MATVEC
    double* aux_w;         // [tile][2][256] J u1 per observation (written)      JV1
    const double* aux;     // the same, read with the tile                       BACKSUB
};

// ---------------------------------------------------------------------------------------------
// shared-memory layout (identical on host and device)
// ---------------------------------------------------------------------------------------------
struct SmemLayout {
    int off_J, off_meta, off_uv, off_camid, off_camvec, off_pa, off_pb, stage_bytes;
    int off_stages, off_pt, off_z, off_buf, off_red, off_ids, off_ytab, off_tab, off_pstart, off_run, off_yacc, total;
};

__host__ __device__ constexpr int align_up(int v, int a) { return (v + a - 1) / a * a; }

template <int MODE>
__host__ __device__ inline SmemLayout smem_layout(int max_cams, int max_pts, int ytab_cams = 0, int n_stages = 0) {
    using T = Traits<MODE>;
    if (n_stages <= 0) n_stages = T::kStages;
    SmemLayout L{};
    int o = 0;
    L.off_J = o;
    if (T::kLoadJ) o += kJTileBytes;
    L.off_meta = o;
    o += align_up((int)sizeof(TileMeta), 128);
    L.off_uv = o;
    if (T::kLoadUV || T::kLoadAux) o += kUVTileBytes;
    L.off_camid = o;
    o += align_up(max_cams * 4, 16);
    L.off_camvec = o;
    o += align_up(max_cams * T::kCamStride * 8, 16);
    L.off_pa = o;
    o += align_up(max_pts * T::kPA * 8, 16);
    L.off_pb = o;
    o += align_up(max_pts * T::kPB * 8, 16);
    L.stage_bytes = align_up(o, 128);
    L.off_stages = 128;                                   // mbarriers live in the first 128 bytes
    o = L.off_stages + n_stages * L.stage_bytes;
    L.off_pt = o;
    o += T::kPtBufs * align_up(max_pts * T::kPtAcc * 8, 16);   // rotating buffers (see the MATVEC / BACKSUB flow)
    L.off_z = o;
    L.off_buf = o;
    o += T::kStageBufs * T::kStageRows * kBufStride * 8;
    L.off_red = o;
    o += 64 * 8;
    L.off_ids = o;
    o += align_up(T::kStages * max_cams * 4, 16);
    L.off_ytab = o;
    if (MODE == M_MATVEC) o += align_up(ytab_cams * 6 * 8, 16);
    L.off_tab = o;
    if (MODE == M_SBUILD) o += kRcmTabCap * 2;                     // (point, camera) -> slot table, u16
    L.off_pstart = o;
    if (MODE == M_SBUILD) o += 2 * align_up((max_pts + 2) * 4, 16);  // first slot / pair offset of every point
    L.off_run = o;
    if (MODE == M_SBUILD) o += kRunInts * 4;                       // accumulation run (see kRunInts)
    L.off_yacc = o;
    if (MODE == M_SBUILD) o += 6 * kT * 8;                         // right-hand-side sums of the run, one column per thread
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy + named barrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA unit, no register staging); completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// barrier among the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(double* addr, double v) { atomicAdd(addr, v); }  // -> REDG.E.ADD.F64

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// Sum of NV values per consumer thread over the CTA's consumers, added to out[i] with one RED each.
template <int NV>
__device__ __forceinline__ void consumer_accumulate(const double (&v)[NV], double* s_red /* >= 8*NV */, double* const* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double w = warp_sum(v[i]);
        if (lane == 0) s_red[warp * NV + i] = w;
    }
    consumer_sync();
    if (threadIdx.x < NV) {
        double t = 0;
        for (int w = 0; w < kConsumers / 32; ++w) t += s_red[w * NV + threadIdx.x];
        red_add(out[threadIdx.x], t);
    }
}

// Block-wide sum for the plain (non-persistent) vector kernels.
template <int NV>
__device__ __forceinline__ void block_accumulate(const double (&v)[NV], double* s_red /* >= 8*NV */, double* const* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double w = warp_sum(v[i]);
        if (lane == 0) s_red[warp * NV + i] = w;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w * NV + threadIdx.x];
        red_add(out[threadIdx.x], t);
    }
    __syncthreads();
}

// Segmented reduction inside a warp over non-decreasing keys: afterwards the first lane of every
// run of equal keys holds the sum of the run (the part of it that lies in this warp).
template <int NV>
__device__ __forceinline__ void warp_seg_reduce(double (&v)[NV], unsigned key, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned k2 = __shfl_down_sync(kFull, key, off);
        const bool take = (lane + off < 32) && (k2 == key);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double o = __shfl_down_sync(kFull, v[i], off);
            if (take) v[i] += o;
        }
    }
}

__device__ __forceinline__ bool run_head(unsigned key, int lane) {
    const unsigned prev = __shfl_up_sync(kFull, key, 1);
    return lane == 0 || prev != key;
}

// Per-point sums inside a tile: observations of a point are contiguous slots, so a warp-level
// segmented reduction leaves one partial per (warp, point); partials are combined in shared
// memory (a point of L observations spans at most ceil(L/32)+1 warps).  s_pt is zero on entry.
template <int NV>
__device__ __forceinline__ void tile_point_reduce(double (&v)[NV], unsigned lp, double* s_pt /* [npts][NV] */) {
    const int lane = threadIdx.x & 31;
    warp_seg_reduce<NV>(v, lp, lane);
    if (run_head(lp, lane) && lp != kPadPt) {
#pragma unroll
        for (int i = 0; i < NV; ++i) atomicAdd(&s_pt[lp * NV + i], v[i]);
    }
}

// Per-camera scatter-add of NV values per observation (one "round"): every consumer stages its
// values in shared memory; after one barrier, thread (run r, component k) sums the run's values
// sequentially — the tile's camera-sorted order and its runs (<= kMaxRun observations of one camera)
// come precomputed in the tile metadata — and issues one f64 RED to out[cam*stride + offset + k].
// No shuffles, conflict-free staging, 6..9 adjacent lanes RED adjacent doubles of one camera.
// Rounds alternate between two staging buffers, so one consumer barrier per round suffices.
template <int NV>
__device__ __forceinline__ void camera_scatter_round(const double (&v)[NV], double* s_buf /* [NV][kBufStride] */,
                                                     const TileMeta* mt, const int* s_camid, double* out, int stride,
                                                     int offset, int n0 = NV, double* out1 = nullptr, int stride1 = 0,
                                                     int offset1 = 0, double* s_tab = nullptr) {
    // components k < n0 go to out[cam*stride + offset + k], the rest to out1[cam*stride1 + offset1 + k - n0]
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < NV; ++i) s_buf[i * kBufStride + tid] = v[i];
    consumer_sync();
    const int nruns = mt->nruns, nobs = mt->nobs;
    for (int idx = tid; idx < nruns * NV; idx += kConsumers) {
        const int r = idx / NV, k = idx - r * NV;
        const int j0 = mt->run_start[r];
        const int j1 = r + 1 < nruns ? (int)mt->run_start[r + 1] : nobs;
        const double* row = s_buf + k * kBufStride;
        double sum = 0.0;
        for (int j = j0; j < j1; ++j) sum += row[mt->sort_src[j]];
        const int64_t cam = s_camid[mt->run_cam[r]];
        if (s_tab) atomicAdd(s_tab + cam * stride + offset + k, sum);   // per-CTA table, flushed once at the end
        else if (k < n0) red_add(out + cam * stride + offset + k, sum);
        else red_add(out1 + cam * stride1 + offset1 + (k - n0), sum);
    }
}

// ---------------------------------------------------------------------------------------------
// K0: per-camera rotation tables  (rotate's trigonometry, bundleAdjuster.py:16-26, hoisted from
// per-observation to per-camera)
//   R = I + a [w]x + b [w]x^2,  Q = R (a I - b [w]x + c w w^T),  d(R X)/dw = -[R X]x Q
//   a = sin t / t, b = (1 - cos t)/t^2, c = (t - sin t)/t^3, Taylor series below t = 0.05.
// ---------------------------------------------------------------------------------------------
__global__ void cam_prep_kernel(const double* __restrict__ xc, double* __restrict__ camtab, int n_cams) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cams) return;
    const double w0 = xc[c * 6 + 0], w1 = xc[c * 6 + 1], w2 = xc[c * 6 + 2];
    const double t2 = w0 * w0 + w1 * w1 + w2 * w2;
    double a, b, cc;
    if (t2 < 0.05 * 0.05) {
        a = 1.0 - t2 / 6 * (1.0 - t2 / 20 * (1.0 - t2 / 42 * (1.0 - t2 / 72)));
        b = 0.5 * (1.0 - t2 / 12 * (1.0 - t2 / 30 * (1.0 - t2 / 56 * (1.0 - t2 / 90))));
        cc = (1.0 / 6) * (1.0 - t2 / 20 * (1.0 - t2 / 42 * (1.0 - t2 / 72 * (1.0 - t2 / 110))));
    } else {
        const double t = sqrt(t2);
        double s, co, sh, ch;
        sincos(t, &s, &co);
        sincos(0.5 * t, &sh, &ch);
        a = s / t;
        b = 2.0 * sh * sh / t2;
        cc = (t - s) / (t2 * t);
    }
    double R[9], G[9];
    // [w]x^2 = w w^T - t2 I
    R[0] = 1.0 + b * (w0 * w0 - t2);
    R[1] = -a * w2 + b * w0 * w1;
    R[2] = a * w1 + b * w0 * w2;
    R[3] = a * w2 + b * w0 * w1;
    R[4] = 1.0 + b * (w1 * w1 - t2);
    R[5] = -a * w0 + b * w1 * w2;
    R[6] = -a * w1 + b * w0 * w2;
    R[7] = a * w0 + b * w1 * w2;
    R[8] = 1.0 + b * (w2 * w2 - t2);
    G[0] = a + cc * w0 * w0;
    G[1] = b * w2 + cc * w0 * w1;
    G[2] = -b * w1 + cc * w0 * w2;
    G[3] = -b * w2 + cc * w0 * w1;
    G[4] = a + cc * w1 * w1;
    G[5] = b * w0 + cc * w1 * w2;
    G[6] = b * w1 + cc * w0 * w2;
    G[7] = -b * w0 + cc * w1 * w2;
    G[8] = a + cc * w2 * w2;
    double* row = camtab + (int64_t)c * kCamTab;
#pragma unroll
    for (int i = 0; i < 9; ++i) row[i] = R[i];
    row[9] = xc[c * 6 + 3];
    row[10] = xc[c * 6 + 4];
    row[11] = xc[c * 6 + 5];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            row[12 + i * 3 + j] = R[i * 3 + 0] * G[0 * 3 + j] + R[i * 3 + 1] * G[1 * 3 + j] + R[i * 3 + 2] * G[2 * 3 + j];
    row[21] = row[22] = row[23] = 0.0;
}

// one observation: residual and (optionally) the analytic 2x6 / 2x3 blocks
template <bool WITH_JAC>
__device__ __forceinline__ void project_obs(const double* __restrict__ cam /* smem row: R t [Q] */, const double X0,
                                            const double X1, const double X2, const double* __restrict__ K,
                                            const double u_obs, const double v_obs, double (&r)[2], double (&jc)[12],
                                            double (&jp)[6]) {
    const double Y0 = cam[0] * X0 + cam[1] * X1 + cam[2] * X2;
    const double Y1 = cam[3] * X0 + cam[4] * X1 + cam[5] * X2;
    const double Y2 = cam[6] * X0 + cam[7] * X1 + cam[8] * X2;
    const double C0 = Y0 + cam[9], C1 = Y1 + cam[10], C2 = Y2 + cam[11];
    const double q0 = K[0] * C0 + K[1] * C1 + K[2] * C2;
    const double q1 = K[3] * C0 + K[4] * C1 + K[5] * C2;
    const double q2 = K[6] * C0 + K[7] * C1 + K[8] * C2;
    const double inv = 1.0 / q2;
    const double u = q0 * inv, v = q1 * inv;
    r[0] = u - u_obs;
    r[1] = v - v_obs;
    if (WITH_JAC) {
        // A = (1/q2) [[1,0,-u],[0,1,-v]] K
        double a[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            a[k] = (K[k] - u * K[6 + k]) * inv;
            a[3 + k] = (K[3 + k] - v * K[6 + k]) * inv;
        }
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            const double a0 = a[row * 3], a1 = a[row * 3 + 1], a2 = a[row * 3 + 2];
            // d/dt = A ; d/dX = A R
            jc[row * 6 + 3] = a0;
            jc[row * 6 + 4] = a1;
            jc[row * 6 + 5] = a2;
#pragma unroll
            for (int k = 0; k < 3; ++k) jp[row * 3 + k] = a0 * cam[k] + a1 * cam[3 + k] + a2 * cam[6 + k];
            // d/dw = -A [Y]x Q = (Y x a)^T Q
            const double b0 = Y1 * a2 - Y2 * a1, b1 = Y2 * a0 - Y0 * a2, b2 = Y0 * a1 - Y1 * a0;
#pragma unroll
            for (int k = 0; k < 3; ++k) jc[row * 6 + k] = b0 * cam[12 + k] + b1 * cam[15 + k] + b2 * cam[18 + k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// S-build helpers (explicit reduced camera matrix, rcm.h)
// ---------------------------------------------------------------------------------------------
// pr-th pair (a <= b) of the upper triangle of an n x n matrix in row-major order
__device__ __forceinline__ void tri_decode(int pr, int n, int& a, int& b) {
    const double t = 2.0 * n + 1.0;
    int aa = (int)((t - sqrt(t * t - 8.0 * pr)) * 0.5);
    aa = max(0, min(aa, n - 1));
    while (aa > 0 && aa * n - aa * (aa - 1) / 2 > pr) --aa;
    while (aa + 1 < n && (aa + 1) * n - (aa + 1) * aa / 2 <= pr) ++aa;
    a = aa;
    b = aa + (pr - (aa * n - aa * (aa - 1) / 2));
}

// index of block (ci, cj), ci <= cj, in the upper-triangle pattern (video-like visibility: a band, so the
// first guess is usually right)
__device__ __forceinline__ int rcm_lookup(const int* __restrict__ rowptr, const int* __restrict__ cols, int ci, int cj) {
    int lo = __ldg(rowptr + ci), hi = __ldg(rowptr + ci + 1);
    const int guess = lo + (cj - ci);
    if (guess < hi && __ldg(cols + guess) == cj) return guess;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cols + mid) <= cj) lo = mid;
        else hi = mid;
    }
    return lo;
}

// acc[b] += row `row` of  Jc_i^T (delta_ij I - Jp_i M Jp_j^T) Jc_j  for the observations in slots i, j of one point:
// the (camera(i), camera(j)) block of  J_c^T J_c - W V'^-1 W^T  restricted to this point.  sJ: the tile's J block,
// s_pm: rows k of Jp M (2x3, row-major) at stride kBufStride.
__device__ __forceinline__ void sbuild_row(const double* __restrict__ sJ, const double* __restrict__ s_pm, int i, int j,
                                           int row, double (&acc)[6]) {
    double g00 = i == j ? 1.0 : 0.0, g01 = 0.0, g10 = 0.0, g11 = g00;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double pm0 = s_pm[k * kBufStride + i], pm1 = s_pm[(3 + k) * kBufStride + i];
        const double q0 = sJ[(12 + k) * kT + j], q1 = sJ[(15 + k) * kT + j];
        g00 -= pm0 * q0;
        g01 -= pm0 * q1;
        g10 -= pm1 * q0;
        g11 -= pm1 * q1;
    }
    const double c0 = sJ[row * kT + i], c1 = sJ[(6 + row) * kT + i];
    const double t0 = c0 * g00 + c1 * g10, t1 = c0 * g01 + c1 * g11;
#pragma unroll
    for (int b = 0; b < 6; ++b) acc[b] = fma(t0, sJ[b * kT + j], fma(t1, sJ[(6 + b) * kT + j], acc[b]));
}

// acc (6x6, row-major) += Jc_i^T (delta_ij I - Jp_i M Jp_j^T) Jc_j : the whole block of one observation pair
__device__ __forceinline__ void sbuild_block(const double* __restrict__ sJ, const double* __restrict__ s_pm, int i, int j,
                                             double (&acc)[36]) {
    double g00 = i == j ? 1.0 : 0.0, g01 = 0.0, g10 = 0.0, g11 = g00;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double pm0 = s_pm[k * kBufStride + i], pm1 = s_pm[(3 + k) * kBufStride + i];
        const double q0 = sJ[(12 + k) * kT + j], q1 = sJ[(15 + k) * kT + j];
        g00 -= pm0 * q0;
        g01 -= pm0 * q1;
        g10 -= pm1 * q0;
        g11 -= pm1 * q1;
    }
    double h0[6], h1[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const double c0 = sJ[b * kT + j], c1 = sJ[(6 + b) * kT + j];
        h0[b] = g00 * c0 + g01 * c1;
        h1[b] = g10 * c0 + g11 * c1;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const double c0 = sJ[a * kT + i], c1 = sJ[(6 + a) * kT + i];
#pragma unroll
        for (int b = 0; b < 6; ++b) acc[a * 6 + b] = fma(c0, h0[b], fma(c1, h1[b], acc[a * 6 + b]));   // 2 DFMA per entry
    }
}

// point slices per camera pair of the register-accumulation S-build path: slices x pair capacity <= 256 threads with
// the capacity as tight as possible (whole warps of live lanes: a DFMA costs the same issue time with 15 live lanes
// as with 32); at most 32 slices (the flush sums the slices in shared memory)
__device__ __forceinline__ int sbuild_slices(int npair) { return max(1, min(32, kConsumers / max(1, npair))); }

// ---------------------------------------------------------------------------------------------
// The streaming kernel.  Algorithmic bytes per observation:
//   BUILD   184  (24 in: metadata + uv; 16 residual + 144 Jacobian out) + fused V/g_p, U/g_c, cost
//   RESID    24  trial cost only
//   MATVEC  152  y_c += sum_i Jc_i^T Jp_i M_p (sum_{j in p} Jp_j^T Jc_j xt_c(j))              (PCG)
//   RHS     152  y_c += sum_i Jc_i^T Jp_i zg_p ;  Sd_c += sum_i E_i M_p E_i^T, E_i = Jc_i^T Jp_i
//   BACKSUB 152  dp_p = M_p (g_p - sum_{j in p} Jp_j^T Jc_j xt_c(j))
//   JV1/JV2 152  ||J v||^2 / 2x2 Gram of J [v0 v1]   (build_quadratic_1d, J_h.dot(S): common.py:282-288,
//                trf.py:498-499); only scalars leave the SM
//   SBUILD  152  explicit reduced camera matrix: T_(ci,cj) += Jc_i^T (delta_ij I - Jp_i M_p Jp_j^T) Jc_j over the
//                observation pairs of every point (upper blocks), and the Schur right-hand side y_c (as RHS)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(Traits<MODE>::kThreads, Traits<MODE>::kMinBlocks) tile_kernel(const TileArgs A, const ModeArgs P) {
    using T = Traits<MODE>;
    constexpr int kStages = T::kStages;
    constexpr int kThreads = T::kThreads;
    // programmatic dependent launch: this grid may have been scheduled while its predecessor drains
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (MODE == M_MATVEC && P.done && *P.done) return;
    extern __shared__ __align__(128) unsigned char smem[];
    const int nst = MODE == M_SBUILD ? A.sb_stages : kStages;      // stages in use (<= kStages producer warps)
    const SmemLayout L = smem_layout<MODE>(A.max_cams, A.max_pts, A.ytab_cams, nst);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kStages;
    const int tid = threadIdx.x;

    // contiguous tile range of this CTA
    const int t_begin = (int)(((int64_t)blockIdx.x * A.n_tiles) / gridDim.x);
    const int t_end = (int)(((int64_t)(blockIdx.x + 1) * A.n_tiles) / gridDim.x);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // scratch that must start at zero
    if (MODE == M_MATVEC && A.ytab_cams) {
        double* s_ytab = reinterpret_cast<double*>(smem + L.off_ytab);
        for (int i = tid; i < A.ytab_cams * 6; i += kThreads) s_ytab[i] = 0.0;
    }
    if (MODE == M_SBUILD) {
        double* s_yacc = reinterpret_cast<double*>(smem + L.off_yacc);
        for (int i = tid; i < 6 * kT; i += kThreads) s_yacc[i] = 0.0;
    }
    if (T::kPtAcc) {
        double* s_pt = reinterpret_cast<double*>(smem + L.off_pt);
        for (int i = tid; i < T::kPtBufs * align_up(A.max_pts * T::kPtAcc * 8, 16) / 8; i += kThreads) s_pt[i] = 0.0;
    }
    __syncthreads();

    if (tid >= kConsumers) {
        // ======================= producer warps =======================
        // Producer warp w owns stage w and the CTA's tiles t_begin + w, t_begin + w + kStages, ...
        // Everything a tile needs besides its bulk-copied blocks is fetched AHEAD of the stage
        // becoming free: header + camera list two tiles ahead, the first kBatch*32 gathered values
        // one tile ahead (held in registers while the warp sleeps on the "empty" barrier).  When the
        // consumers release the stage the producer only issues the bulk copies, stores registers to
        // shared memory and arrives, so a stage is almost never without a copy in flight.
        const int pw = (tid - kConsumers) >> 5, lane = tid & 31;
        if (pw >= nst) return;
        constexpr int kCPL = kT / 32;   // camera ids per lane (registers)
        // gathered values prefetched one tile ahead per lane (SBUILD stages 9 values per point: up to 1 152 per tile)
        constexpr int kBatch = MODE == M_SBUILD ? 16 : 8;
        const int stage = pw;
        unsigned char* st = smem + L.off_stages + stage * L.stage_bytes;
        int* s_camid = reinterpret_cast<int*>(st + L.off_camid);
        double* s_cv = reinterpret_cast<double*>(st + L.off_camvec);
        double* s_pa = reinterpret_cast<double*>(st + L.off_pa);
        double* s_pb = reinterpret_cast<double*>(st + L.off_pb);
        int* s_ids = reinterpret_cast<int*>(smem + L.off_ids) + pw * A.max_cams;   // producer-private camera ids

        int4 hdr_far = make_int4(0, 0, 0, 0);   // header / cameras of the tile after next
        int cams_far[kCPL];
        auto fetch_far = [&](int t) {
            if (t < t_end) {
                hdr_far = *reinterpret_cast<const int4*>(&A.meta[t]);
#pragma unroll
                for (int j = 0; j < kCPL; ++j) {
                    const int c = lane + 32 * j;
                    cams_far[j] = c < A.max_cams ? A.tile_cams[(int64_t)t * A.cam_stride + c] : -1;
                }
            }
        };
        int4 hdr = make_int4(0, 0, 0, 0);       // header / cameras / first gathered batch of the next tile
        int cams[kCPL];
        double vals[kBatch];
        int n_cam = 0, n_pa = 0, n_pb = 0, total = 0;
        const double *ptA = nullptr, *ptB = nullptr;
        auto src_of = [&](int i, const int* ids) -> const double* {
            if (i < n_cam) {
                const int c = i / (T::kCamRows ? T::kCamRows : 1), k = i - c * T::kCamRows;
                const int cam = ids[c];
                if (is_project(MODE)) return P.cam0 + (int64_t)cam * kCamTab + k;
                if (MODE == M_JV2) return k < 6 ? P.cam0 + (int64_t)cam * 6 + k : P.cam1 + (int64_t)cam * 6 + (k - 6);
                return P.cam0 + (int64_t)cam * 6 + k;
            }
            if (i < n_cam + n_pa) return ptA + (i - n_cam);
            return ptB + (i - n_cam - n_pa);
        };
        auto dst_of = [&](int i) -> double* {
            if (i < n_cam) {
                const int c = i / (T::kCamRows ? T::kCamRows : 1), k = i - c * T::kCamRows;
                return s_cv + c * T::kCamStride + k;
            }
            if (i < n_cam + n_pa) return s_pa + (i - n_cam);
            return s_pb + (i - n_cam - n_pa);
        };
        // make (hdr_far, cams_far) the current tile and start its first gather batch
        auto advance = [&]() {
            hdr = hdr_far;
#pragma unroll
            for (int j = 0; j < kCPL; ++j) cams[j] = cams_far[j];
            n_cam = hdr.z * T::kCamRows;
            n_pa = hdr.y * T::kPA;
            n_pb = hdr.y * T::kPB;
            total = n_cam + n_pa + n_pb;
            ptA = P.ptA + (int64_t)hdr.x * T::kPA;
            ptB = T::kPB ? P.ptB + (int64_t)hdr.x * T::kPB : nullptr;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kCPL; ++j) {
                const int c = lane + 32 * j;
                if (c < hdr.z) s_ids[c] = cams[j];
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int i = lane + 32 * u;
                vals[u] = i < total ? __ldg(src_of(i, s_ids)) : 0.0;
            }
        };
        int t = t_begin + pw;
        unsigned phase = 0;
        bool first = true;
        for (; t < t_end; t += nst) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                if (T::kLoadJ) bulk_g2s(st + L.off_J, P.J + (int64_t)t * kJRows * kT, kJTileBytes, &full[stage]);
                bulk_g2s(st + L.off_meta, &A.meta[t], (unsigned)sizeof(TileMeta), &full[stage]);
                if (T::kLoadUV) bulk_g2s(st + L.off_uv, A.uv + (int64_t)t * 2 * kT, kUVTileBytes, &full[stage]);
                if (T::kLoadAux) bulk_g2s(st + L.off_uv, P.aux + (int64_t)t * 2 * kT, kUVTileBytes, &full[stage]);
            }
            if (first) {
                // the first tile's bulk copies are already in flight while its header chain resolves
                first = false;
                fetch_far(t);
                advance();
                fetch_far(t + nst);
            }
#pragma unroll
            for (int j = 0; j < kCPL; ++j) {
                const int c = lane + 32 * j;
                if (c < hdr.z) s_camid[c] = cams[j];
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int i = lane + 32 * u;
                if (i < total) *dst_of(i) = vals[u];
            }
            // up to three more batches here (one memory round trip each, overlapped with the other stages'
            // arithmetic); values beyond kProducerShare (tiles touching many cameras) are fetched cooperatively
            // by the 256 consumer threads once the stage is full
            for (int base = 32 * kBatch + lane; base < min(total, kProducerShare); base += 32 * kBatch) {
                double v[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int i = base + 32 * u;
                    v[u] = i < min(total, kProducerShare) ? __ldg(src_of(i, s_ids)) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int i = base + 32 * u;
                    if (i < min(total, kProducerShare)) *dst_of(i) = v[u];
                }
            }
            __syncwarp();
            if (lane == 0) {
                const unsigned tx = (T::kLoadJ ? kJTileBytes : 0) + (unsigned)sizeof(TileMeta) +
                                    ((T::kLoadUV || T::kLoadAux) ? kUVTileBytes : 0);
                mbar_arrive_expect_tx(&full[stage], tx);
            }
            phase ^= 1;
            if (t + nst < t_end) {
                advance();                       // next tile of this warp: ids + first batch in flight
                fetch_far(t + 2 * nst);
            }
        }
        return;
    }

    // ======================= consumer warps =======================
    const int lane = tid & 31;
    double* s_pt = reinterpret_cast<double*>(smem + L.off_pt);
    double* s_buf = reinterpret_cast<double*>(smem + L.off_buf);
    double* s_red = reinterpret_cast<double*>(smem + L.off_red);
    int par = 0;                   // per-point accumulator parity (MATVEC / BACKSUB)
    double acc[3] = {0, 0, 0};     // cost (BUILD / RESID) or Gram (JV)
    // SBUILD: one 6x6 block of the reduced camera matrix per thread, accumulated over a run of tiles with the
    // same camera list (video-like visibility: tens of tiles) and added to HBM once per run
    double sacc[MODE == M_SBUILD ? 36 : 1];
#pragma unroll
    for (int i = 0; i < (MODE == M_SBUILD ? 36 : 1); ++i) sacc[i] = 0.0;
    // An accumulation RUN: a list of cameras (ascending ids, s_run[0..run_n)) whose pair blocks are being accumulated
    // in registers.  Thread (slice my_sl, pair my_pr) = thread my_sl * run_P + my_pr owns the block of cameras
    // (run[my_a], run[my_b]), my_a <= my_b, for the points p = my_sl (mod run_ns) of every tile; pairs are numbered
    // column by column (pr = b (b + 1) / 2 + a) so that appending a camera to the list adds pairs without renumbering
    // the existing ones.  A tile continues the run when its cameras are a subset of the list, extends it when the new
    // cameras are larger than the last one and the pairs still fit the capacity run_P, else the run is flushed and
    // restarted.  Video-like visibility: one flush per 10..40 tiles.
    int run_n = 0, run_ns = 1, run_P = kConsumers;     // cameras, point slices, pair capacity (= 256 / slices)
    int my_sl = 0, my_pr = 0, my_a = 0, my_b = 0;
    int my_blk = 0;                                           // block of (run[my_a], run[my_b]) in the upper pattern (slice 0)
    // right-hand side y_c += Jc^T Jp M g_p: thread (camera y_a of the run, point slice y_sl) = thread y_sl * y_cap + y_a
    // walks the points of its slice that camera y_a observes and keeps six sums in its column of s_yacc
    int y_cap = 1, y_ns = 1, y_sl = 0, y_a = 0;
    int prev_n = -1;                                          // cameras of the previous tile (register-accumulation tiles), else -1
    bool run_touched = false;
    bool prev_mode2 = false;                                  // the previous tile left no closing barrier behind
    int* s_run = reinterpret_cast<int*>(smem + L.off_run);   // layout: kRunInts
    double* s_yacc = reinterpret_cast<double*>(smem + L.off_yacc);
    auto run_pairs = [](int n) { return n * (n + 1) / 2; };
    // diagnostics (-DMMBA_PHASE_TIMING builds only): cycles per phase of the S-build tile loop as seen by thread 0 of CTA 0
#ifdef MMBA_PHASE_TIMING
    long long t_last = 0, phc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_fl = 0, flc[6] = {0, 0, 0, 0, 0, 0};
    const bool timing = MODE == M_SBUILD && A.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    auto lap = [&](int k) {
        if (MODE == M_SBUILD && timing) {
            const long long tt = clock64();
            phc[k] += tt - t_last;
            t_last = tt;
        }
    };
    auto flap = [&](int k) {      // sub-phases of a flush: k < 0 starts the clock
        if (MODE == M_SBUILD && timing) {
            const long long tt = clock64();
            if (k >= 0) flc[k] += tt - t_fl;
            t_fl = tt;
        }
    };
    if (timing) t_last = clock64();
#else
    auto lap = [](int) {};
    auto flap = [](int) {};
#endif
    // Adds the run's blocks and right-hand-side sums to HBM and ends the run.  One slice: every thread adds its own
    // block.  Several slices: the slices are summed through `scratch` (6 staging rows nobody reads at this point),
    // one block row per round, so that a block costs 36 REDs whatever the number of slices.  Every consumer thread
    // must call it after a consumer barrier (all accumulation of the run is complete); barriers inside.
    auto sbuild_flush = [&](double* scratch) {
        if constexpr (MODE == M_SBUILD) {
            if (run_n == 0) return;
            const int np = run_pairs(run_n);
            flap(-1);
            {
                // right-hand side: (camera, component, group of point slices) sums, one RED each
                const int G = max(1, min(8, kConsumers / (run_n * 6)));
                for (int q = tid; q < run_n * 6 * G; q += kConsumers) {
                    const int g = q % G, ak = q / G;
                    const int a = ak / 6, k = ak - a * 6;
                    double* col = s_yacc + k * kT + a;
                    double sum = 0.0;
                    for (int sl = g; sl < y_ns; sl += G) {
                        sum += col[sl * y_cap];
                        col[sl * y_cap] = 0.0;
                    }
                    if (sum != 0.0) red_add(P.y + (int64_t)s_run[a] * 6 + k, sum);
                }
            }
            flap(0);
            if (run_ns == 1) {
                if (my_pr < np && run_touched) {
                    double* dst = P.Tup + (int64_t)my_blk * 36;
#pragma unroll
                    for (int e = 0; e < 36; ++e) red_add(dst + e, sacc[e]);
                }
            } else {
                int* s_blk = s_run + kRunBlk;
                // (a pair of the list that no point sees has no block: its sums are zero and never added)
                if (my_sl == 0 && my_pr < np) s_blk[my_pr] = my_blk;
                flap(1);
#pragma unroll
                for (int rd = 0; rd < 6; ++rd) {      // one block row per round
#pragma unroll
                    for (int e = 0; e < 6; ++e) scratch[e * kBufStride + tid] = sacc[rd * 6 + e];
                    consumer_sync();
                    flap(2);
                    for (int q = tid; q < np * 6; q += kConsumers) {
                        const int e = q / np, pr = q - e * np;      // lanes: consecutive pairs (conflict-free columns)
                        const double* col = scratch + e * kBufStride + pr;
                        const int blk = s_blk[pr];
                        double s0 = 0.0;
                        for (int sl = 0; sl < run_ns; sl += 8) {      // eight independent loads, then a tree
                            double v[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) v[u] = sl + u < run_ns ? col[(sl + u) * run_P] : 0.0;
                            s0 += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                        }
                        if (s0 != 0.0) red_add(P.Tup + (int64_t)blk * 36 + rd * 6 + e, s0);
                    }
                    flap(3);
                    consumer_sync();
                    flap(4);
                }
            }
#pragma unroll
            for (int e = 0; e < 36; ++e) sacc[e] = 0.0;
            run_touched = false;
            run_n = 0;
            flap(5);
#ifdef MMBA_PHASE_TIMING
            if (timing) A.dbg[40] += 1;
#endif
        }
    };
    auto run_start = [&](int ncams) {      // thread mapping of a new run of `ncams` cameras
        run_n = ncams;
        run_ns = sbuild_slices(run_pairs(ncams));
        run_P = kConsumers / run_ns;
        my_sl = tid / run_P;
        my_pr = tid - my_sl * run_P;
        int b = (int)((sqrtf(8.0f * (float)my_pr + 1.0f) - 1.0f) * 0.5f);
        while (b * (b + 1) / 2 > my_pr) --b;
        while ((b + 1) * (b + 2) / 2 <= my_pr) ++b;
        my_b = b;
        my_a = my_pr - b * (b + 1) / 2;
        // cameras the list can grow to within the pair capacity
        y_cap = ncams;
        while (y_cap < 31 && run_pairs(y_cap + 1) <= run_P) ++y_cap;
        y_ns = max(1, kConsumers / y_cap);
        y_sl = tid / y_cap;
        y_a = tid - y_sl * y_cap;
    };
    // block index of this thread's pair (slice 0): the loads are in flight while the run accumulates
    auto run_lookup = [&]() {
        if constexpr (MODE == M_SBUILD) {
            if (my_sl == 0 && my_pr < run_pairs(run_n)) my_blk = rcm_lookup(P.up_rowptr, P.up_cols, s_run[my_a], s_run[my_b]);
        }
    };
    int stage = 0;
    unsigned phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
        unsigned char* st = smem + L.off_stages + stage * L.stage_bytes;
        mbar_wait(&full[stage], phase);
        lap(0);
        const TileMeta* mt = reinterpret_cast<const TileMeta*>(st + L.off_meta);
        const double* sJ = reinterpret_cast<const double*>(st + L.off_J);
        const double* s_cv = reinterpret_cast<const double*>(st + L.off_camvec);
        const double* s_pa = reinterpret_cast<const double*>(st + L.off_pa);
        const double* s_pb = reinterpret_cast<const double*>(st + L.off_pb);
        const int* s_camid = reinterpret_cast<const int*>(st + L.off_camid);
        const int pt0 = mt->pt0, npts = mt->npts;
        {
            // remainder of the tile's gathered operands (see the producer): index space
            // (camera rows | point payload A | point payload B), entries >= kProducerShare
            const int n_cam = mt->ncams * T::kCamRows, n_pa = npts * T::kPA, n_pb = npts * T::kPB;
            const int total = n_cam + n_pa + n_pb;
            if (total > kProducerShare) {
                double* w_cv = reinterpret_cast<double*>(st + L.off_camvec);
                double* w_pa = reinterpret_cast<double*>(st + L.off_pa);
                double* w_pb = reinterpret_cast<double*>(st + L.off_pb);
                for (int i = kProducerShare + tid; i < total; i += kConsumers) {
                    if (i < n_cam) {
                        const int c = i / (T::kCamRows ? T::kCamRows : 1), k = i - c * T::kCamRows;
                        const int cam = s_camid[c];
                        double v;
                        if (is_project(MODE)) v = __ldg(P.cam0 + (int64_t)cam * kCamTab + k);
                        else if (MODE == M_JV2) v = k < 6 ? __ldg(P.cam0 + (int64_t)cam * 6 + k) : __ldg(P.cam1 + (int64_t)cam * 6 + (k - 6));
                        else v = __ldg(P.cam0 + (int64_t)cam * 6 + k);
                        w_cv[c * T::kCamStride + k] = v;
                    } else if (i < n_cam + n_pa) {
                        w_pa[i - n_cam] = __ldg(P.ptA + (int64_t)pt0 * T::kPA + (i - n_cam));
                    } else if (T::kPB) {
                        w_pb[i - n_cam - n_pa] = __ldg(P.ptB + (int64_t)pt0 * T::kPB + (i - n_cam - n_pa));
                    }
                }
                consumer_sync();
            }
        }
        const unsigned lp = mt->slot_pt[tid], lc = mt->slot_cam[tid];
        const bool valid = lp != kPadPt;
        const int lps = valid ? (int)lp : 0;

        if constexpr (is_project(MODE)) {
            const double* s_uv = reinterpret_cast<const double*>(st + L.off_uv);
            double r[2] = {0, 0}, jc[12], jp[6];
#pragma unroll
            for (int i = 0; i < 12; ++i) jc[i] = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = 0;
            if (valid) {
                const double* X = s_pa + lp * 3;
                project_obs<is_build(MODE)>(s_cv + lc * T::kCamStride, X[0], X[1], X[2], A.K, s_uv[tid], s_uv[kT + tid], r, jc, jp);
            }
            acc[0] += r[0] * r[0] + r[1] * r[1];
            if (MODE != M_RESID) {
                double* rt = P.res + (int64_t)t * 2 * kT;
                rt[tid] = r[0];
                rt[kT + tid] = r[1];
            }
            if constexpr (is_build(MODE)) {
                double* Jt = P.Jw + (int64_t)t * kJRows * kT;
#pragma unroll
                for (int i = 0; i < 12; ++i) Jt[i * kT + tid] = jc[i];
#pragma unroll
                for (int i = 0; i < 6; ++i) Jt[(12 + i) * kT + tid] = jp[i];
                // One staging round for everything the normal equations need from this observation:
                //   rows 0..5   diag(Jc^T Jc)  (column norms of the camera part: the Marquardt scale)
                //   rows 6..11  Jc^T r         (camera gradient)
                //   rows 12..17 Jp^T Jp (upper triangle), rows 18..20 Jp^T r   (point block, gradient)
                //   row  21     first slot of every point of the tile (int)
                // then thread (camera run, k) / (point, k) sums one run sequentially: no shuffles, no atomics.
                double* buf = s_buf;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    buf[k * kBufStride + tid] = jc[k] * jc[k] + jc[6 + k] * jc[6 + k];
                    buf[(6 + k) * kBufStride + tid] = jc[k] * r[0] + jc[6 + k] * r[1];
                }
                buf[12 * kBufStride + tid] = jp[0] * jp[0] + jp[3] * jp[3];
                buf[13 * kBufStride + tid] = jp[0] * jp[1] + jp[3] * jp[4];
                buf[14 * kBufStride + tid] = jp[0] * jp[2] + jp[3] * jp[5];
                buf[15 * kBufStride + tid] = jp[1] * jp[1] + jp[4] * jp[4];
                buf[16 * kBufStride + tid] = jp[1] * jp[2] + jp[4] * jp[5];
                buf[17 * kBufStride + tid] = jp[2] * jp[2] + jp[5] * jp[5];
                buf[18 * kBufStride + tid] = jp[0] * r[0] + jp[3] * r[1];
                buf[19 * kBufStride + tid] = jp[1] * r[0] + jp[4] * r[1];
                buf[20 * kBufStride + tid] = jp[2] * r[0] + jp[5] * r[1];
                int* s_pstart = reinterpret_cast<int*>(buf + 21 * kBufStride);
                if (valid && (tid == 0 || mt->slot_pt[tid - 1] != lp)) s_pstart[lp] = tid;
                if (tid == 0) s_pstart[npts] = mt->nobs;
                consumer_sync();
                {
                    const int nruns = mt->nruns, nobs = mt->nobs;
                    const int n_cam = nruns * 12, n_tot = n_cam + npts * 9;
                    for (int idx = tid; idx < n_tot; idx += kConsumers) {
                        if (idx < n_cam) {
                            const int rr = idx / 12, k = idx - rr * 12;
                            const int j0 = mt->run_start[rr];
                            const int j1 = rr + 1 < nruns ? (int)mt->run_start[rr + 1] : nobs;
                            const double* row = buf + k * kBufStride;
                            double sum = 0.0;
                            for (int j = j0; j < j1; ++j) sum += row[mt->sort_src[j]];
                            const int64_t cam = s_camid[mt->run_cam[rr]];
                            red_add((k < 6 ? P.Ud : P.gc - 6) + cam * 6 + k, sum);
                        } else {
                            const int q = idx - n_cam;
                            const int p = q / 9, k = q - p * 9;
                            const double* row = buf + (12 + k) * kBufStride;
                            const int j1 = s_pstart[p + 1];
                            double sum = 0.0;
                            for (int j = s_pstart[p]; j < j1; ++j) sum += row[j];
                            // the tile owns its points: plain stores
                            if (k < 6) P.V[((int64_t)pt0 + p) * 6 + k] = sum;
                            else P.gp[((int64_t)pt0 + p) * 3 + (k - 6)] = sum;
                        }
                    }
                }
                consumer_sync();   // the single staging buffer is rewritten by the next round
                if constexpr (MODE == M_BUILD_FULL) {
                    // evaluation hook only: the full 6x6 camera blocks (upper triangle) in a second round
                    double cv[21];
#pragma unroll
                    for (int i = 0; i < 21; ++i) cv[i] = jc[tri6_row(i)] * jc[tri6_col(i)] + jc[6 + tri6_row(i)] * jc[6 + tri6_col(i)];
                    camera_scatter_round<21>(cv, buf, mt, s_camid, P.U, 21, 0);
                    consumer_sync();
                }
            }
        } else if constexpr (MODE == M_JV1 || MODE == M_JV2) {
            constexpr int NV = MODE == M_JV2 ? 2 : 1;
            double e[NV][2];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double* c = s_cv + lc * T::kCamStride + v * 6;
                const double* p = (v == 0 ? s_pa : s_pb) + lps * 3;
                double e0 = 0, e1 = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    e0 += sJ[k * kT + tid] * c[k];
                    e1 += sJ[(6 + k) * kT + tid] * c[k];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    e0 += sJ[(12 + k) * kT + tid] * p[k];
                    e1 += sJ[(15 + k) * kT + tid] * p[k];
                }
                e[v][0] = valid ? e0 : 0.0;
                e[v][1] = valid ? e1 : 0.0;
            }
            acc[0] += e[0][0] * e[0][0] + e[0][1] * e[0][1];
            if (MODE == M_JV1 && P.aux_w) {
                double* o = P.aux_w + (int64_t)t * 2 * kT;
                o[tid] = e[0][0];
                o[kT + tid] = e[0][1];
            }
            if (NV == 2) {
                acc[1] += e[0][0] * e[NV - 1][0] + e[0][1] * e[NV - 1][1];
                acc[2] += e[NV - 1][0] * e[NV - 1][0] + e[NV - 1][1] * e[NV - 1][1];
            }
        } else if constexpr (MODE == M_SBUILD) {
            // ---- explicit reduced camera matrix (upper blocks) + Schur right-hand side ----
            const int ncams = mt->ncams;
            int pair_mode = mt->pair_mode;
            // the register-accumulation strategy needs the tile's table in one parity half and its camera list in one warp
            if (pair_mode == 2 && (npts * ncams > kRcmTab2Cap || ncams > 32)) pair_mode = 1;
            uint16_t* s_tab = reinterpret_cast<uint16_t*>(smem + L.off_tab);
            double jp[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = valid ? sJ[(12 + i) * kT + tid] : 0.0;
            const double* m = s_pa + lps * 6;
            const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
            if (pair_mode == 2) {
                // One consumer barrier per tile: the staged rows, the table and the run bookkeeping alternate between
                // two parities, so a warp that is done with this tile's pairs stages the next tile while the others
                // still accumulate.  The right-hand side's camera-run sums are taken AFTER the barrier, a few per
                // warp, so that their latency chains hide behind the other warps' pair arithmetic.
                const int k_loc = t - t_begin, q = k_loc & 1;
                const unsigned gen = 1u + (unsigned)((k_loc >> 1) % 255);     // (q, gen) identifies the tile among 510
                if (!prev_mode2) {
                    // first tile of its kind: forget whatever the table region held (the previous tile closed with a barrier)
                    for (int i = tid; i < kRcmTabCap / 4; i += kConsumers) reinterpret_cast<uint2*>(s_tab)[i] = make_uint2(0u, 0u);
                    consumer_sync();
                } else if (k_loc >= 510 && k_loc % 510 < 2) {
                    // (q, gen) repeats: forget this parity's stale entries (its last readers passed the previous barrier)
                    for (int i = tid; i < kRcmTab2Cap / 4; i += kConsumers)
                        reinterpret_cast<uint2*>(s_tab + q * kRcmTab2Cap)[i] = make_uint2(0u, 0u);
                    consumer_sync();
                }
                double* s_pm = s_buf + q * 8 * kBufStride;      // rows 0..5: Jp M (2x3), rows 6..7: v = Jp (M_p g_p)
                uint16_t* tab = s_tab + q * kRcmTab2Cap;
                int* s_flag = s_run + 32 + 2 * q;
                int* s_map = s_run + kRunMap + 32 * q;          // run index -> local camera slot of this tile, -1 = absent
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double a0 = jp[r * 3], a1 = jp[r * 3 + 1], a2 = jp[r * 3 + 2];
                    s_pm[(r * 3 + 0) * kBufStride + tid] = a0 * m0 + a1 * m1 + a2 * m2;
                    s_pm[(r * 3 + 1) * kBufStride + tid] = a0 * m1 + a1 * m3 + a2 * m4;
                    s_pm[(r * 3 + 2) * kBufStride + tid] = a0 * m2 + a1 * m4 + a2 * m5;
                }
                {
                    const double* zg = s_pb + lps * 3;
                    s_pm[6 * kBufStride + tid] = jp[0] * zg[0] + jp[1] * zg[1] + jp[2] * zg[2];
                    s_pm[7 * kBufStride + tid] = jp[3] * zg[0] + jp[4] * zg[1] + jp[5] * zg[2];
                }
                if (valid) tab[lp * ncams + lc] = (uint16_t)((unsigned)tid | (gen << 8));
                if (tid < 32) {
                    // Warp 0: is every camera of the tile in the run's list?  (bit 0: some are missing, bit 1: a missing
                    // one is not larger than the list's last camera, i.e. the list cannot simply be extended), and the
                    // local camera slot of every run camera in this tile.  Usual case: the same cameras as the previous
                    // tile — nothing is missing and the map carries over.
                    int* s_prev = s_run + kRunPrev + 32 * q;
                    const int* s_prev_o = s_run + kRunPrev + 32 * (q ^ 1);
                    const int mycam = tid < ncams ? s_camid[tid] : -1;
                    s_prev[tid] = mycam;
                    const bool same = __all_sync(kFull, prev_n == ncams && s_prev_o[tid] == mycam);
                    if (same) {
                        if (tid < run_n) s_map[tid] = s_run[kRunMap + 32 * (q ^ 1) + tid];
                        if (tid == 0) {
                            s_flag[0] = 4;      // bit 2: the same cameras as the previous tile
                            s_flag[1] = 0;
                        }
                    } else {
                        bool miss = false, low = false;
                        if (tid < ncams) {
                            bool found = false;
                            for (int r = 0; r < run_n; ++r) found |= s_run[r] == mycam;
                            miss = !found;
                            low = miss && run_n > 0 && mycam < s_run[run_n - 1];
                        }
                        const unsigned mm = __ballot_sync(kFull, miss), ml = __ballot_sync(kFull, low);
                        if (tid == 0) {
                            s_flag[0] = (mm ? 1 : 0) | (ml ? 2 : 0);
                            s_flag[1] = __popc(mm);
                        }
                        if (tid < run_n) {
                            const int cam = s_run[tid];
                            int at = -1;
                            for (int c = 0; c < ncams; ++c)
                                if (s_camid[c] == cam) at = c;
                            s_map[tid] = at;
                        }
                    }
                }
                prev_n = ncams;
                lap(1);
                consumer_sync();
                lap(3);
                const int flags = s_flag[0], n_missing = s_flag[1];
                // (Restarting the run on a subset tile's own cameras — tighter capacity, no idle pair lanes — was measured:
                // the extra flushes cost more than the lanes gain, profiles/r2_sbuild_phases_repack_experiment.log.)
                if (flags & 1) {
                    const int new_n = run_n + n_missing;
                    if (run_n > 0 && !(flags & 2) && new_n <= 31 && run_pairs(new_n) <= run_P) {
                        // extend: the missing cameras are the last n_missing of the tile's (ascending) list
                        if (tid >= ncams - n_missing && tid < ncams) {
                            const int pos = run_n + (tid - (ncams - n_missing));
                            s_run[pos] = s_camid[tid];
                            s_map[pos] = tid;
                        }
                        run_n = new_n;
                    } else {
                        const bool had_run = run_n != 0;
                        sbuild_flush(s_buf + (q ^ 1) * 8 * kBufStride);   // the other parity's rows: their readers are done
                        if (had_run) consumer_sync();   // every thread has read the old camera list
                        if (tid < ncams) {
                            s_run[tid] = s_camid[tid];
                            s_map[tid] = tid;
                        }
                        run_start(ncams);
                    }
                    consumer_sync();
#ifdef MMBA_PHASE_TIMING
                    const long long tl0 = timing ? clock64() : 0;
#endif
                    run_lookup();
#ifdef MMBA_PHASE_TIMING
                    if (timing) A.dbg[41] += (my_blk >= 0 ? clock64() : 0) - tl0;     // (reads my_blk: waits for the loads)
#endif
                }
                lap(4);
                if (y_sl < y_ns && y_a < run_n) {
                    // right-hand side: the observations of camera y_a among the points of slice y_sl
                    const int la = s_map[y_a];
                    if (la >= 0) {
                        double ty[6] = {0, 0, 0, 0, 0, 0};
                        bool any = false;
                        for (int p = y_sl; p < npts; p += y_ns) {
                            const unsigned e = tab[p * ncams + la];
                            if ((e >> 8) != gen) continue;
                            const int sl = (int)(e & 255u);
                            const double w0 = s_pm[6 * kBufStride + sl], w1 = s_pm[7 * kBufStride + sl];
                            any = true;
#pragma unroll
                            for (int k = 0; k < 6; ++k) ty[k] = fma(sJ[k * kT + sl], w0, fma(sJ[(6 + k) * kT + sl], w1, ty[k]));
                        }
                        if (any) {
#pragma unroll
                            for (int k = 0; k < 6; ++k) s_yacc[k * kT + tid] += ty[k];
                        }
                    }
                }
                lap(2);
                // unit = (camera pair a <= b of the run, point slice): whole 6x6 block in registers.  Lanes of a warp
                // are consecutive pairs of one or two slices: they walk the same few points, so the J rows they read
                // are a handful of neighbouring slots (broadcast / conflict-free); a lane whose pair is absent from a
                // point skips ahead on its own instead of idling through the other lanes' block.
                if (my_sl < run_ns && my_pr < run_pairs(run_n)) {
                    const int la = s_map[my_a], lb = s_map[my_b];
                    if (la >= 0 && lb >= 0) {
                        // the table entries of the next candidate point are fetched before the current block's
                        // arithmetic (an entry of generation 0 never matches)
                        int p = my_sl;
                        unsigned i = 0u, j = 0u;
                        auto probe = [&](int pp) {
                            if (pp < npts) {
                                i = tab[pp * ncams + la];
                                j = tab[pp * ncams + lb];
                            }
                        };
                        probe(p);
                        while (true) {
                            while (p < npts && !((i >> 8) == gen && (j >> 8) == gen)) {
                                p += run_ns;
                                probe(p);
                            }
                            if (p >= npts) break;
                            run_touched = true;
                            const int si = (int)(i & 255u), sj = (int)(j & 255u);
                            p += run_ns;
                            probe(p);
                            sbuild_block(sJ, s_pm, si, sj, sacc);
                        }
                    }
                }
                lap(5);
                prev_mode2 = true;
                lap(6);
            } else {
                // ---- per-tile strategies (many cameras per tile / duplicate observations): flushed tile by tile ----
                if (prev_mode2) consumer_sync();     // the previous tile's pair phase still reads its staged rows and table
                sbuild_flush(s_buf);                 // a tile of another strategy ends the run
                prev_mode2 = false;
                prev_n = -1;
                double jc[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) jc[i] = valid ? sJ[i * kT + tid] : 0.0;
                int* s_pstart = reinterpret_cast<int*>(smem + L.off_pstart);
                int* s_poff = s_pstart + align_up((A.max_pts + 2) * 4, 16) / 4;
                double* s_pm = s_buf + 8 * kBufStride;
                {
                    // y_c += Jc^T Jp (M_p g_p)
                    const double* zg = s_pb + lps * 3;
                    const double v0 = jp[0] * zg[0] + jp[1] * zg[1] + jp[2] * zg[2];
                    const double v1 = jp[3] * zg[0] + jp[4] * zg[1] + jp[5] * zg[2];
                    double cv[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) cv[k] = jc[k] * v0 + jc[6 + k] * v1;
                    // Jp M (2x3) of this observation
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const double a0 = jp[r * 3], a1 = jp[r * 3 + 1], a2 = jp[r * 3 + 2];
                        s_pm[(r * 3 + 0) * kBufStride + tid] = a0 * m0 + a1 * m1 + a2 * m2;
                        s_pm[(r * 3 + 1) * kBufStride + tid] = a0 * m1 + a1 * m3 + a2 * m4;
                        s_pm[(r * 3 + 2) * kBufStride + tid] = a0 * m2 + a1 * m4 + a2 * m5;
                    }
                    if (pair_mode) {
                        for (int i = tid; i < npts * ncams; i += kConsumers) s_tab[i] = 0xFFFF;
                    } else {
                        if (valid && (tid == 0 || mt->slot_pt[tid - 1] != lp)) s_pstart[lp] = tid;
                        if (tid == 0) s_pstart[npts] = mt->nobs;
                    }
                    lap(1);
                    camera_scatter_round<6>(cv, s_buf, mt, s_camid, P.y, 6, 0);   // one consumer barrier inside
                    lap(2);
                }
                if (pair_mode) {
                    if (valid) s_tab[lp * ncams + lc] = (uint16_t)tid;
                } else if (tid == 0) {
                    int o = 0;
                    for (int p = 0; p < npts; ++p) {
                        s_poff[p] = o;
                        const int len = s_pstart[p + 1] - s_pstart[p];
                        o += len * (len + 1) / 2;
                    }
                    s_poff[npts] = o;
                }
                consumer_sync();
                lap(3);
                if (pair_mode == 1) {
                    // unit = (camera pair a <= b of the tile, block row, point slice): register accumulation over the
                    // tile's points, then one RED per entry
                    const int npair = ncams * (ncams + 1) / 2;
                    int nslice = 1;
                    if (npair * 6 < kConsumers) nslice = max(1, min(npts, kConsumers / (npair * 6)));
                    const int units = npair * 6 * nslice;
                    for (int u = tid; u < units; u += kConsumers) {
                        const int sl = u % nslice, rest = u / nslice;
                        const int row = rest % 6, pr = rest / 6;
                        int a, b;
                        tri_decode(pr, ncams, a, b);
                        double acc[6] = {0, 0, 0, 0, 0, 0};
                        bool hit = false;
                        for (int p = sl; p < npts; p += nslice) {
                            const unsigned i = s_tab[p * ncams + a];
                            if (i == 0xFFFFu) continue;
                            const unsigned j = s_tab[p * ncams + b];
                            if (j == 0xFFFFu) continue;
                            hit = true;
                            sbuild_row(sJ, s_pm, (int)i, (int)j, row, acc);
                        }
                        if (hit) {
                            const int k = rcm_lookup(P.up_rowptr, P.up_cols, s_camid[a], s_camid[b]);
                            double* dst = P.Tup + (int64_t)k * 36 + row * 6;
#pragma unroll
                            for (int bb = 0; bb < 6; ++bb) red_add(dst + bb, acc[bb]);
                        }
                    }
                } else {
                    // unit = (observation pair i <= j of one point, block row): cameras ascend inside a point
                    const int units = mt->npairs * 6;
                    for (int u = tid; u < units; u += kConsumers) {
                        const int row = u % 6, q = u / 6;
                        int lo = 0, hi = npts;
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (s_poff[mid] <= q) lo = mid;
                            else hi = mid;
                        }
                        const int start = s_pstart[lo], len = s_pstart[lo + 1] - start;
                        int ii, jj;
                        tri_decode(q - s_poff[lo], len, ii, jj);
                        const int i = start + ii, j = start + jj;
                        double acc[6] = {0, 0, 0, 0, 0, 0};
                        sbuild_row(sJ, s_pm, i, j, row, acc);
                        const int ci = s_camid[mt->slot_cam[i]], cj = s_camid[mt->slot_cam[j]];
                        const int k = rcm_lookup(P.up_rowptr, P.up_cols, ci, cj);
                        double* dst = P.Tup + (int64_t)k * 36;
#pragma unroll
                        for (int bb = 0; bb < 6; ++bb) red_add(dst + row * 6 + bb, acc[bb]);
                        if (ci == cj && i != j) {
                            // one camera observing the point twice: the mirrored pair lands in the same diagonal block
#pragma unroll
                            for (int bb = 0; bb < 6; ++bb) red_add(dst + bb * 6 + row, acc[bb]);
                        }
                    }
                }
                lap(5);
                consumer_sync();   // staging rows and tables are rewritten by the next tile
                lap(6);
            }
        } else {
            // ---- Schur passes: MATVEC / RHS / BACKSUB ----
            double jc[12], jp[6];
#pragma unroll
            for (int i = 0; i < 12; ++i) jc[i] = valid ? sJ[i * kT + tid] : 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) jp[i] = valid ? sJ[(12 + i) * kT + tid] : 0.0;
            double z0, z1, z2;
            double uu0 = 0, uu1 = 0;
            double cvy[6] = {0, 0, 0, 0, 0, 0};   // RHS: the y contribution rides in the first Sd round
            if constexpr (MODE != M_RHS) {
                // u = Jc xt_c ; w = Jp^T u ; t_p = sum over the point's observations.  The per-point accumulators
                // rotate: tile i's sums are read after its barrier while tile i+1 already accumulates into the next
                // buffer.  MATVEC (two buffers): each is re-zeroed by its readers before the scatter round's barrier.
                // BACKSUB (three buffers, one barrier per tile): after the barrier of tile i the buffer of tile i+2 is
                // zeroed — its last readers (tile i-1) are past that barrier, its next writers wait for the one of tile i+1.
                const int pt_buf = align_up(A.max_pts * T::kPtAcc * 8, 16) / 8;
                double* s_ptp = s_pt + par * pt_buf;
                const double* xc = s_cv + lc * T::kCamStride;
                double u0 = 0, u1 = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    u0 += jc[k] * xc[k];
                    u1 += jc[6 + k] * xc[k];
                }
                uu0 = u0;
                uu1 = u1;
                double w[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) w[k] = jp[k] * u0 + jp[3 + k] * u1;
                tile_point_reduce<3>(w, lp, s_ptp);
                consumer_sync();
                if constexpr (MODE == M_BACKSUB) {
                    // every observation forms its point's step dp = M (g - t) itself; the first slot of a point stores it
                    const double t0 = s_pb[lps * 3] - s_ptp[lps * 3], t1 = s_pb[lps * 3 + 1] - s_ptp[lps * 3 + 1],
                                 t2 = s_pb[lps * 3 + 2] - s_ptp[lps * 3 + 2];
                    const double* m = s_pa + lps * 6;
                    const double d0 = m[0] * t0 + m[1] * t1 + m[2] * t2;
                    const double d1 = m[1] * t0 + m[3] * t1 + m[4] * t2;
                    const double d2 = m[2] * t0 + m[4] * t1 + m[5] * t2;
                    if (valid && (tid == 0 || mt->slot_pt[tid - 1] != lp)) {
                        double* o = P.dp + ((int64_t)pt0 + lps) * 3;
                        o[0] = d0;
                        o[1] = d1;
                        o[2] = d2;
                    }
                    {
                        double* s_ptz = s_pt + ((par + 2) % 3) * pt_buf;
                        for (int i = tid; i < A.max_pts * 3; i += kConsumers) s_ptz[i] = 0.0;
                    }
                    if (P.aux) {
                        // Gram sums of J [u1 u2]: J u2 = Jc xt_c + Jp dp_p (u2 = the unscaled Gauss-Newton step),
                        // J u1 from the JV1 pass (trf.py:498-499)
                        const double* s_ju = reinterpret_cast<const double*>(st + L.off_uv);
                        const double e10 = valid ? s_ju[tid] : 0.0, e11 = valid ? s_ju[kT + tid] : 0.0;
                        const double e20 = valid ? uu0 + jp[0] * d0 + jp[1] * d1 + jp[2] * d2 : 0.0;
                        const double e21 = valid ? uu1 + jp[3] * d0 + jp[4] * d1 + jp[5] * d2 : 0.0;
                        acc[1] += e10 * e20 + e11 * e21;
                        acc[2] += e20 * e20 + e21 * e21;
                    }
                    z0 = z1 = z2 = 0.0;
                } else {
                    // every observation applies its point's damped inverse itself: z = M_p t_p
                    const double t0 = s_ptp[lps * 3], t1 = s_ptp[lps * 3 + 1], t2 = s_ptp[lps * 3 + 2];
                    const double* m = s_pa + lps * 6;
                    z0 = m[0] * t0 + m[1] * t1 + m[2] * t2;
                    z1 = m[1] * t0 + m[3] * t1 + m[4] * t2;
                    z2 = m[2] * t0 + m[4] * t1 + m[5] * t2;
                }
            } else {
                z0 = s_pb[lps * 3];
                z1 = s_pb[lps * 3 + 1];
                z2 = s_pb[lps * 3 + 2];
            }
            if constexpr (MODE != M_BACKSUB) {
                // v = Jp z_p.  MATVEC scatters Jc^T (Jc xt - v): the camera's own block U_c xt is folded into
                // the same pass, so the reduced-system product needs no stored U.  RHS scatters Jc^T v.
                double v0 = jp[0] * z0 + jp[1] * z1 + jp[2] * z2;
                double v1 = jp[3] * z0 + jp[4] * z1 + jp[5] * z2;
                if (MODE == M_MATVEC) {
                    v0 = uu0 - v0;
                    v1 = uu1 - v1;
                }
                double cv[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) cv[k] = jc[k] * v0 + jc[6 + k] * v1;
                if constexpr (MODE == M_RHS) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) cvy[k] = cv[k];
                } else {
                    camera_scatter_round<6>(cv, const_cast<double*>(sJ), mt, s_camid, P.y, 6, 0, 6,
                                            nullptr, 0, 0, A.ytab_cams ? reinterpret_cast<double*>(smem + L.off_ytab) : nullptr);
                }
                if constexpr (MODE == M_MATVEC) {
                    // all reads of this parity's point sums happened before the round's barrier
                    double* s_ptp = s_pt + par * (align_up(A.max_pts * T::kPtAcc * 8, 16) / 8);
                    for (int i = tid; i < npts * 3; i += kConsumers) s_ptp[i] = 0.0;
                }
            }
            if constexpr (MODE == M_RHS) {
                // Diagonal blocks of the reduced system: S_cc = sum_i (Jc^T Jc - F E^T), E = Jc^T Jp (6x3),
                // F = E M (6x3); upper triangle, 21 values, scattered in the y round + two more rounds of 9
                const double* m = s_pa + lps * 6;
                const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
                double E[18], F[18];
#pragma unroll
                for (int a = 0; a < 6; ++a) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) E[a * 3 + k] = jc[a] * jp[k] + jc[6 + a] * jp[3 + k];
                    F[a * 3 + 0] = E[a * 3] * m0 + E[a * 3 + 1] * m1 + E[a * 3 + 2] * m2;
                    F[a * 3 + 1] = E[a * 3] * m1 + E[a * 3 + 1] * m3 + E[a * 3 + 2] * m4;
                    F[a * 3 + 2] = E[a * 3] * m2 + E[a * 3 + 1] * m4 + E[a * 3 + 2] * m5;
                }
                auto sval = [&](int i) {
                    const int a = tri6_row(i), b = tri6_col(i);
                    return jc[a] * jc[b] + jc[6 + a] * jc[6 + b] -
                           (F[a * 3] * E[b * 3] + F[a * 3 + 1] * E[b * 3 + 1] + F[a * 3 + 2] * E[b * 3 + 2]);
                };
                double sv[9];
#pragma unroll
                for (int k = 0; k < 6; ++k) sv[k] = cvy[k];
#pragma unroll
                for (int i = 0; i < 3; ++i) sv[6 + i] = sval(i);
                double* jbuf = const_cast<double*>(sJ);   // 9 staging rows inside the J block
                consumer_sync();                          // every consumer has its J values in registers
                camera_scatter_round<9>(sv, jbuf, mt, s_camid, P.y, 6, 0, 6, P.Sd, 21, 0);
#pragma unroll
                for (int i = 0; i < 9; ++i) sv[i] = sval(3 + i);
                consumer_sync();                          // previous round's sums are done
                camera_scatter_round<9>(sv, jbuf, mt, s_camid, P.Sd, 21, 3);
#pragma unroll
                for (int i = 0; i < 9; ++i) sv[i] = sval(12 + i);
                consumer_sync();
                camera_scatter_round<9>(sv, jbuf, mt, s_camid, P.Sd, 21, 12);
            }
        }
        // all reads of this stage are done: hand it back to the producer
        par = T::kPtBufs == 3 ? (par + 1) % 3 : par ^ 1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == nst) {
            stage = 0;
            phase ^= 1;
        }
    }
    if constexpr (MODE == M_SBUILD) {
        if (prev_mode2) consumer_sync();     // the last tile's pair phase still reads its staged rows
        sbuild_flush(reinterpret_cast<double*>(smem + L.off_buf));
    }
#ifdef MMBA_PHASE_TIMING
    if (MODE == M_SBUILD && timing) {
        lap(7);
        for (int k = 0; k < 8; ++k) A.dbg[16 + k] += phc[k];
        for (int k = 0; k < 6; ++k) A.dbg[32 + k] += flc[k];
        A.dbg[24] += t_end - t_begin;
    }
#endif
    if constexpr (MODE == M_MATVEC) {
        // few cameras (heavy RED contention on few addresses): the CTA's sums were kept in shared memory
        if (A.ytab_cams) {
            consumer_sync();
            const double* s_ytab = reinterpret_cast<const double*>(smem + L.off_ytab);
            for (int i = tid; i < A.ytab_cams * 6; i += kConsumers) {
                const double v = s_ytab[i];
                if (v != 0.0) red_add(P.y + i, v);
            }
        }
    }
    // per-CTA scalar results
    if constexpr (is_build(MODE)) {
        double c[1] = {acc[0]};
        double* outp[1] = {P.scal + S_COST};
        consumer_accumulate<1>(c, s_red, outp);
    } else if constexpr (MODE == M_RESID || MODE == M_RESID_STORE) {
        double c[1] = {acc[0]};
        double* outp[1] = {P.cost};
        consumer_accumulate<1>(c, s_red, outp);
    } else if constexpr (MODE == M_JV1) {
        double c[1] = {acc[0]};
        double* outp[1] = {P.scal + S_JV00};
        consumer_accumulate<1>(c, s_red, outp);
    } else if constexpr (MODE == M_JV2) {
        double* outp[3] = {P.scal + S_JV00, P.scal + S_JV01, P.scal + S_JV11};
        consumer_accumulate<3>(acc, s_red, outp);
    } else if constexpr (MODE == M_BACKSUB) {
        if (P.aux) {
            double c[2] = {acc[1], acc[2]};
            double* outp[2] = {P.scal + S_JV01, P.scal + S_JV11};
            consumer_accumulate<2>(c, s_red, outp);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4: damped 3x3 point-block inversion
//   M_p = D_p (D_p V_p D_p + reg I)^-1 D_p  (symmetric, 6 doubles) and zg_p = M_p g_p
// ---------------------------------------------------------------------------------------------
__global__ void point_invert_kernel(const double* __restrict__ V, const double* __restrict__ gp,
                                    const double* __restrict__ sinv_p, double reg, double* __restrict__ M,
                                    double* __restrict__ zg, int64_t n_pts) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const double d0 = 1.0 / sinv_p[p * 3], d1 = 1.0 / sinv_p[p * 3 + 1], d2 = 1.0 / sinv_p[p * 3 + 2];
    const double* v = V + p * 6;
    const double a00 = v[0] * d0 * d0 + reg, a01 = v[1] * d0 * d1, a02 = v[2] * d0 * d2;
    const double a11 = v[3] * d1 * d1 + reg, a12 = v[4] * d1 * d2, a22 = v[5] * d2 * d2 + reg;
    // adjugate / determinant of a symmetric 3x3
    const double c00 = a11 * a22 - a12 * a12;
    const double c01 = a02 * a12 - a01 * a22;
    const double c02 = a01 * a12 - a02 * a11;
    const double c11 = a00 * a22 - a02 * a02;
    const double c12 = a01 * a02 - a00 * a12;
    const double c22 = a00 * a11 - a01 * a01;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    double m[6];
    if (det > 0.0 && isfinite(det)) {
        const double id = 1.0 / det;
        m[0] = c00 * id * d0 * d0;
        m[1] = c01 * id * d0 * d1;
        m[2] = c02 * id * d0 * d2;
        m[3] = c11 * id * d1 * d1;
        m[4] = c12 * id * d1 * d2;
        m[5] = c22 * id * d2 * d2;
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) m[i] = 0.0;   // unobserved point with reg == 0: no step
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) M[p * 6 + i] = m[i];
    const double g0 = gp[p * 3], g1 = gp[p * 3 + 1], g2 = gp[p * 3 + 2];
    zg[p * 3 + 0] = m[0] * g0 + m[1] * g1 + m[2] * g2;
    zg[p * 3 + 1] = m[1] * g0 + m[3] * g1 + m[4] * g2;
    zg[p * 3 + 2] = m[2] * g0 + m[4] * g1 + m[5] * g2;
}

}  // namespace mmba
