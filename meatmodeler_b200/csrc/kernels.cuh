// Device code of libmmba.so — float64 CUDA kernels for sm_100a (B200).
//
// Data layout (all in HBM, per shard)
//   * observations are reordered by the host plan (plan.h) into point-aligned tiles of kT = 256
//     slots; one CTA of 256 threads owns one tile, one thread one observation slot.
//   * J   : 18 rows x n_slots doubles, structure-of-arrays.  Rows 0..11 = 2x6 camera block
//           (row-major, columns w0 w1 w2 t0 t1 t2), rows 12..17 = 2x3 point block.  Every pass
//           over J is 18 fully coalesced 8-byte streams per warp.
//   * res : 2 rows x n_slots (du, dv);  uv: 2 rows x n_slots.
//   * per-slot metadata, 2 bytes each: local camera slot, local point index, and the tile's
//     camera-sorted order (source slot + key) used by the warp-shuffle segmented scatter.
//   * cameras: camtab[Nc][24] = R (9) | t (3) | Q (9) produced by cam_prep_kernel; each tile
//     stages only the cameras it touches in shared memory.
//   * points: x_p[3*Np] in the plan's internal order; V (6), g_p (3), M (6) per point.
//   * camera accumulators U (21 upper-triangle doubles), g_c (6), y (6), Sd (21) per camera,
//     updated with one native f64 RED per (warp-level camera run, component).
//
// Reference sites replaced: rotate/project/pointFun (bundleAdjuster.py:7-52, 81-102), scipy's
// finite-difference Jacobian (_numdiff.py:770-893), J^T J / J^T f (common.py:590-610) and the
// matvec/rmatvec pair inside LSMR (lsmr.py:337,343).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace mmba {

constexpr int kT = kTileObs;
constexpr int kCamTab = 24;   // doubles per camera-table row in HBM
constexpr int kCamS = 21;     // staged row length in shared memory (odd: conflict-free)
constexpr int kVecS = 7;      // staged stride of a 6-vector per camera (odd)
constexpr unsigned kFull = 0xffffffffu;
constexpr uint16_t kPadPt = 0xFFFF;

struct TileArgs {
    const int4* tiles;
    const int32_t* tile_cams;
    const uint16_t* slot_cam;
    const uint16_t* slot_pt;
    const uint16_t* sort_src;
    const uint16_t* sort_key;
    const double* uv;
    int64_t n_slots;
    double K[9];
};

// scalar slots in device memory (doubles).  Groups that are reduced across ranks together are
// contiguous: [S_COST..S_X2] sum, S_GINF max, [S_JV00..S_JV11] sum, [S_DOT0..S_DOT9] sum, S_COST_NEW sum.
enum Scal {
    S_COST = 0,      // sum r^2 (build)
    S_GH2,           // ||g_h||^2
    S_XSI2,          // sum (x * scale_inv)^2
    S_X2,            // ||x||^2
    S_GINF,          // ||g||_inf as the bit pattern of a non-negative double (atomicMax on u64)
    S_COST_NEW,      // sum r^2 (trial)
    S_JV00, S_JV01, S_JV11,   // Gram of J*v products
    S_DOT0, S_DOT1, S_DOT2, S_DOT3, S_DOT4, S_DOT5, S_DOT6, S_DOT7, S_DOT8, S_DOT9,   // subspace dots
    S_COUNT = 24
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(double* addr, double v) { atomicAdd(addr, v); }  // -> REDG.E.ADD.F64

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// Block-wide sum of NV values per thread, result added to out[i] with one RED per block.
template <int NV>
__device__ __forceinline__ void block_accumulate(const double (&v)[NV], double* s_red /* >= 8*NV */,
                                                 double* const* out) {
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
// memory (a point of L observations spans at most ceil(L/32)+1 warps).
template <int NV>
__device__ __forceinline__ void tile_point_reduce(double (&v)[NV], unsigned lp, double* s_pt /* [npts][NV], zeroed */) {
    const int lane = threadIdx.x & 31;
    warp_seg_reduce<NV>(v, lp, lane);
    if (run_head(lp, lane) && lp != kPadPt) {
#pragma unroll
        for (int i = 0; i < NV; ++i) atomicAdd(&s_pt[lp * NV + i], v[i]);
    }
}

// Per-camera scatter-add of NV values per observation: stage the values in shared memory, re-read
// them in the tile's camera-sorted order, reduce runs of equal cameras with warp shuffles and issue
// one f64 RED per (run, component) to out[cam * stride + offset + i].
template <int NV>
__device__ __forceinline__ void tile_camera_scatter(const double (&v)[NV], double* s_stage /* [NV][kT] */,
                                                    unsigned src, unsigned key, const int* s_camid,
                                                    double* out, int stride, int offset) {
    const int tid = threadIdx.x, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) s_stage[i * kT + tid] = v[i];
    __syncthreads();
    double w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = s_stage[i * kT + src];
    warp_seg_reduce<NV>(w, key, lane);
    if (run_head(key, lane) && key != kPadKey) {
        double* dst = out + (int64_t)s_camid[key] * stride + offset;
#pragma unroll
        for (int i = 0; i < NV; ++i) red_add(dst + i, w[i]);
    }
    __syncthreads();
}

// packed upper-triangle index helpers for 6x6 (21) and 3x3 (6)
__host__ __device__ constexpr int tri6(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }
__host__ __device__ constexpr int tri3(int a, int b) { return a * 3 - a * (a - 1) / 2 + (b - a); }
__host__ __device__ constexpr int tri6_row(int idx) {
    return idx < 6 ? 0 : idx < 11 ? 1 : idx < 15 ? 2 : idx < 18 ? 3 : idx < 20 ? 4 : 5;
}
__host__ __device__ constexpr int tri6_col(int idx) {
    return idx - tri6(tri6_row(idx), tri6_row(idx)) + tri6_row(idx);
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

// ---------------------------------------------------------------------------------------------
// tile prologue shared by the projection kernels: stage the tile's cameras and points
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_cameras(const TileArgs& A, const int4 ti, const double* __restrict__ camtab,
                                              double* s_cam, int* s_camid) {
    for (int i = threadIdx.x; i < ti.w * kCamS; i += blockDim.x) {
        const int c = i / kCamS, k = i - c * kCamS;
        s_cam[i] = camtab[(int64_t)A.tile_cams[ti.z + c] * kCamTab + k];
    }
    if (s_camid)
        for (int i = threadIdx.x; i < ti.w; i += blockDim.x) s_camid[i] = A.tile_cams[ti.z + i];
}

// one observation: residual and (optionally) the analytic 2x6 / 2x3 blocks
template <bool WITH_JAC>
__device__ __forceinline__ void project_obs(const double* __restrict__ cam /* smem row: R t Q */,
                                            const double X0, const double X1, const double X2,
                                            const double* __restrict__ K, const double u_obs, const double v_obs,
                                            double (&r)[2], double (&jc)[12], double (&jp)[6]) {
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
// K1: build — gather, project, residual, Jacobian blocks, and the normal-equation blocks
//   per slot: 24 B in (indices + uv), 16 B residual + 144 B Jacobian out  -> 184 B/observation
//   fused: cost, V_p / g_p (point-segment sums), U_c / g_c (camera scatter)
// dynamic smem: s_cam[max_cams*21] | s_X[max_pts*3] | s_pt[max_pts*9] | s_stage[9*kT] | s_red[64] | s_camid[max_cams]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT)
build_kernel(const TileArgs A, const double* __restrict__ camtab, const double* __restrict__ xp,
             double* __restrict__ J, double* __restrict__ res, double* __restrict__ U, double* __restrict__ gc,
             double* __restrict__ V, double* __restrict__ gp, double* __restrict__ scal, int max_cams, int max_pts) {
    extern __shared__ double smem[];
    double* s_cam = smem;
    double* s_X = s_cam + max_cams * kCamS;
    double* s_pt = s_X + max_pts * 3;
    double* s_stage = s_pt + max_pts * 9;
    double* s_red = s_stage + 9 * kT;
    int* s_camid = reinterpret_cast<int*>(s_red + 64);

    const int tid = threadIdx.x;
    const int4 ti = A.tiles[blockIdx.x];
    const int64_t slot = (int64_t)blockIdx.x * kT + tid;
    stage_cameras(A, ti, camtab, s_cam, s_camid);
    for (int i = tid; i < ti.y * 3; i += kT) s_X[i] = xp[(int64_t)ti.x * 3 + i];
    for (int i = tid; i < ti.y * 9; i += kT) s_pt[i] = 0.0;
    const unsigned lp = A.slot_pt[slot], lc = A.slot_cam[slot];
    const unsigned src = A.sort_src[slot], key = A.sort_key[slot];
    const bool valid = lp != kPadPt;
    const double uo = valid ? A.uv[slot] : 0.0, vo = valid ? A.uv[A.n_slots + slot] : 0.0;
    __syncthreads();

    double r[2] = {0, 0}, jc[12], jp[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) jc[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) jp[i] = 0;
    if (valid) {
        const double* X = s_X + lp * 3;
        project_obs<true>(s_cam + lc * kCamS, X[0], X[1], X[2], A.K, uo, vo, r, jc, jp);
#pragma unroll
        for (int i = 0; i < 12; ++i) J[(int64_t)i * A.n_slots + slot] = jc[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) J[(int64_t)(12 + i) * A.n_slots + slot] = jp[i];
        res[slot] = r[0];
        res[A.n_slots + slot] = r[1];
    }
    // cost
    {
        double c[1] = {r[0] * r[0] + r[1] * r[1]};
        double* outp[1] = {scal + S_COST};
        block_accumulate<1>(c, s_red, outp);
    }
    // point blocks: V (6) and g_p (3)
    {
        double pv[9];
        pv[0] = jp[0] * jp[0] + jp[3] * jp[3];
        pv[1] = jp[0] * jp[1] + jp[3] * jp[4];
        pv[2] = jp[0] * jp[2] + jp[3] * jp[5];
        pv[3] = jp[1] * jp[1] + jp[4] * jp[4];
        pv[4] = jp[1] * jp[2] + jp[4] * jp[5];
        pv[5] = jp[2] * jp[2] + jp[5] * jp[5];
        pv[6] = jp[0] * r[0] + jp[3] * r[1];
        pv[7] = jp[1] * r[0] + jp[4] * r[1];
        pv[8] = jp[2] * r[0] + jp[5] * r[1];
        tile_point_reduce<9>(pv, lp, s_pt);
    }
    // camera blocks: U (21 upper-triangle) + g_c (6) in three rounds of 9
    {
        double cv[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) cv[i] = jc[tri6_row(i)] * jc[tri6_col(i)] + jc[6 + tri6_row(i)] * jc[6 + tri6_col(i)];
        tile_camera_scatter<9>(cv, s_stage, src, key, s_camid, U, 21, 0);
#pragma unroll
        for (int i = 0; i < 9; ++i) cv[i] = jc[tri6_row(9 + i)] * jc[tri6_col(9 + i)] + jc[6 + tri6_row(9 + i)] * jc[6 + tri6_col(9 + i)];
        tile_camera_scatter<9>(cv, s_stage, src, key, s_camid, U, 21, 9);
        double cw[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cw[i] = jc[tri6_row(18 + i)] * jc[tri6_col(18 + i)] + jc[6 + tri6_row(18 + i)] * jc[6 + tri6_col(18 + i)];
        tile_camera_scatter<3>(cw, s_stage, src, key, s_camid, U, 21, 18);
        double cg[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) cg[i] = jc[i] * r[0] + jc[6 + i] * r[1];
        tile_camera_scatter<6>(cg, s_stage, src, key, s_camid, gc, 6, 0);
    }
    // the tile owns its points: plain coalesced stores
    for (int i = tid; i < ti.y * 9; i += kT) {
        const int p = i / 9, k = i - p * 9;
        if (k < 6) V[((int64_t)ti.x + p) * 6 + k] = s_pt[i];
        else gp[((int64_t)ti.x + p) * 3 + (k - 6)] = s_pt[i];
    }
}

// ---------------------------------------------------------------------------------------------
// K1r: residual only (trial point) -> cost; optionally stores the residuals
//   24 B/observation in, one scalar out
// dynamic smem: s_cam[max_cams*21] | s_X[max_pts*3] | s_red[64]
// ---------------------------------------------------------------------------------------------
template <bool STORE>
__global__ void __launch_bounds__(kT)
resid_kernel(const TileArgs A, const double* __restrict__ camtab, const double* __restrict__ xp,
             double* __restrict__ res, double* __restrict__ cost_out, int max_cams, int max_pts) {
    extern __shared__ double smem[];
    double* s_cam = smem;
    double* s_X = s_cam + max_cams * kCamS;
    double* s_red = s_X + max_pts * 3;
    const int tid = threadIdx.x;
    const int4 ti = A.tiles[blockIdx.x];
    const int64_t slot = (int64_t)blockIdx.x * kT + tid;
    stage_cameras(A, ti, camtab, s_cam, nullptr);
    for (int i = tid; i < ti.y * 3; i += kT) s_X[i] = xp[(int64_t)ti.x * 3 + i];
    const unsigned lp = A.slot_pt[slot], lc = A.slot_cam[slot];
    const bool valid = lp != kPadPt;
    const double uo = valid ? A.uv[slot] : 0.0, vo = valid ? A.uv[A.n_slots + slot] : 0.0;
    __syncthreads();
    double r[2] = {0, 0}, jc[12], jp[6];
    if (valid) {
        const double* X = s_X + lp * 3;
        project_obs<false>(s_cam + lc * kCamS, X[0], X[1], X[2], A.K, uo, vo, r, jc, jp);
        if (STORE) {
            res[slot] = r[0];
            res[A.n_slots + slot] = r[1];
        }
    }
    double c[1] = {r[0] * r[0] + r[1] * r[1]};
    double* outp[1] = {cost_out};
    block_accumulate<1>(c, s_red, outp);
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

// ---------------------------------------------------------------------------------------------
// K5 family: one streaming pass over J per launch (144 B Jacobian + 8 B metadata per observation)
//   SCHUR_MATVEC : y_c += sum_i Jc_i^T Jp_i M_p (sum_{j in p} Jp_j^T Jc_j xt_c(j))       (PCG)
//   SCHUR_RHS    : y_c += sum_i Jc_i^T Jp_i zg_p ;  Sd_c += sum_i E_i M_p E_i^T, E_i = Jc_i^T Jp_i
//   SCHUR_BACKSUB: dp_p = M_p (g_p - sum_{j in p} Jp_j^T Jc_j xt_c(j))
// dynamic smem: s_xc[max_cams*7] | s_pt[max_pts*3] | s_stage[9*kT] | s_camid[max_cams]
// ---------------------------------------------------------------------------------------------
enum SchurMode { SCHUR_MATVEC = 0, SCHUR_RHS = 1, SCHUR_BACKSUB = 2 };

template <int MODE>
__global__ void __launch_bounds__(kT)
schur_kernel(const TileArgs A, const double* __restrict__ J, const double* __restrict__ xt,
             const double* __restrict__ M, const double* __restrict__ zg, const double* __restrict__ gp,
             double* __restrict__ y, double* __restrict__ Sd, double* __restrict__ dp,
             const int* __restrict__ done, int max_cams, int max_pts) {
    if (MODE == SCHUR_MATVEC && done && *done) return;
    extern __shared__ double smem[];
    double* s_xc = smem;
    double* s_pt = s_xc + max_cams * kVecS;
    double* s_stage = s_pt + max_pts * 3;
    int* s_camid = reinterpret_cast<int*>(s_stage + 9 * kT);

    const int tid = threadIdx.x;
    const int4 ti = A.tiles[blockIdx.x];
    const int64_t slot = (int64_t)blockIdx.x * kT + tid;
    const int64_t ns = A.n_slots;
    // issue the J loads first: 18 independent coalesced streams
    double jc[12], jp[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) jc[i] = __ldg(J + (int64_t)i * ns + slot);
#pragma unroll
    for (int i = 0; i < 6; ++i) jp[i] = __ldg(J + (int64_t)(12 + i) * ns + slot);
    const unsigned lp = A.slot_pt[slot], lc = A.slot_cam[slot];
    const unsigned src = A.sort_src[slot], key = A.sort_key[slot];
    const bool valid = lp != kPadPt;

    if (MODE != SCHUR_BACKSUB)
        for (int i = tid; i < ti.w; i += kT) s_camid[i] = A.tile_cams[ti.z + i];
    if (MODE != SCHUR_RHS) {
        for (int i = tid; i < ti.w * 6; i += kT) {
            const int c = i / 6, k = i - c * 6;
            s_xc[c * kVecS + k] = xt[(int64_t)A.tile_cams[ti.z + c] * 6 + k];
        }
        for (int i = tid; i < ti.y * 3; i += kT) s_pt[i] = 0.0;
    } else {
        for (int i = tid; i < ti.y * 3; i += kT) s_pt[i] = zg[(int64_t)ti.x * 3 + i];
    }
    __syncthreads();

    if (MODE != SCHUR_RHS) {
        // u = Jc xt_c ; w = Jp^T u ; t_p = sum over the point's observations
        const double* xc = s_xc + lc * kVecS;
        double u0 = 0, u1 = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            u0 += jc[k] * xc[k];
            u1 += jc[6 + k] * xc[k];
        }
        double w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = jp[k] * u0 + jp[3 + k] * u1;
        tile_point_reduce<3>(w, lp, s_pt);
        __syncthreads();
        // one thread per point: z = M t  (MATVEC)  or  dp = M (g - t)  (BACKSUB)
        if (tid < ti.y) {
            const int64_t p = (int64_t)ti.x + tid;
            double t0 = s_pt[tid * 3], t1 = s_pt[tid * 3 + 1], t2 = s_pt[tid * 3 + 2];
            if (MODE == SCHUR_BACKSUB) {
                t0 = gp[p * 3] - t0;
                t1 = gp[p * 3 + 1] - t1;
                t2 = gp[p * 3 + 2] - t2;
            }
            const double* m = M + p * 6;
            const double z0 = m[0] * t0 + m[1] * t1 + m[2] * t2;
            const double z1 = m[1] * t0 + m[3] * t1 + m[4] * t2;
            const double z2 = m[2] * t0 + m[4] * t1 + m[5] * t2;
            if (MODE == SCHUR_BACKSUB) {
                dp[p * 3] = z0;
                dp[p * 3 + 1] = z1;
                dp[p * 3 + 2] = z2;
            } else {
                s_pt[tid * 3] = z0;
                s_pt[tid * 3 + 1] = z1;
                s_pt[tid * 3 + 2] = z2;
            }
        }
        if (MODE == SCHUR_BACKSUB) return;
        __syncthreads();
    }
    // v = Jp z_p ; contribution Jc^T v to the camera
    const int lps = valid ? lp : 0;
    const double z0 = s_pt[lps * 3], z1 = s_pt[lps * 3 + 1], z2 = s_pt[lps * 3 + 2];
    const double v0 = jp[0] * z0 + jp[1] * z1 + jp[2] * z2;
    const double v1 = jp[3] * z0 + jp[4] * z1 + jp[5] * z2;
    double cv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) cv[k] = jc[k] * v0 + jc[6 + k] * v1;
    tile_camera_scatter<6>(cv, s_stage, src, key, s_camid, y, 6, 0);

    if (MODE == SCHUR_RHS) {
        // Schur diagonal: E = Jc^T Jp (6x3), F = E M (6x3), Sd += F E^T (upper triangle)
        const double* m = M + ((int64_t)ti.x + lps) * 6;
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
        double sv[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int a = tri6_row(i), b = tri6_col(i);
            sv[i] = F[a * 3] * E[b * 3] + F[a * 3 + 1] * E[b * 3 + 1] + F[a * 3 + 2] * E[b * 3 + 2];
        }
        tile_camera_scatter<9>(sv, s_stage, src, key, s_camid, Sd, 21, 0);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int a = tri6_row(9 + i), b = tri6_col(9 + i);
            sv[i] = F[a * 3] * E[b * 3] + F[a * 3 + 1] * E[b * 3 + 1] + F[a * 3 + 2] * E[b * 3 + 2];
        }
        tile_camera_scatter<9>(sv, s_stage, src, key, s_camid, Sd, 21, 9);
        double sw[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int a = tri6_row(18 + i), b = tri6_col(18 + i);
            sw[i] = F[a * 3] * E[b * 3] + F[a * 3 + 1] * E[b * 3 + 1] + F[a * 3 + 2] * E[b * 3 + 2];
        }
        tile_camera_scatter<3>(sw, s_stage, src, key, s_camid, Sd, 21, 18);
    }
}

// ---------------------------------------------------------------------------------------------
// K8: J*v products for NVEC (1 or 2) unscaled n-vectors; only the Gram scalars leave the SM
//   (build_quadratic_1d / J_h.dot(S), common.py:282-288, trf.py:498-499).  Optionally stores J*v0.
// dynamic smem: s_vc[NVEC][max_cams*7] | s_vp[NVEC][max_pts*3] | s_red[64]
// ---------------------------------------------------------------------------------------------
template <int NVEC, bool STORE>
__global__ void __launch_bounds__(kT)
jv_kernel(const TileArgs A, const double* __restrict__ J, const double* __restrict__ vc0,
          const double* __restrict__ vp0, const double* __restrict__ vc1, const double* __restrict__ vp1,
          double* __restrict__ scal, double* __restrict__ jv_out, int max_cams, int max_pts) {
    extern __shared__ double smem[];
    double* s_vc = smem;
    double* s_vp = s_vc + NVEC * max_cams * kVecS;
    double* s_red = s_vp + NVEC * max_pts * 3;
    const int tid = threadIdx.x;
    const int4 ti = A.tiles[blockIdx.x];
    const int64_t slot = (int64_t)blockIdx.x * kT + tid;
    const int64_t ns = A.n_slots;
    double jc[12], jp[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) jc[i] = __ldg(J + (int64_t)i * ns + slot);
#pragma unroll
    for (int i = 0; i < 6; ++i) jp[i] = __ldg(J + (int64_t)(12 + i) * ns + slot);
    const unsigned lp = A.slot_pt[slot], lc = A.slot_cam[slot];
    const bool valid = lp != kPadPt;
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
        const double* vc = v == 0 ? vc0 : vc1;
        const double* vp = v == 0 ? vp0 : vp1;
        for (int i = tid; i < ti.w * 6; i += kT) {
            const int c = i / 6, k = i - c * 6;
            s_vc[v * max_cams * kVecS + c * kVecS + k] = vc[(int64_t)A.tile_cams[ti.z + c] * 6 + k];
        }
        for (int i = tid; i < ti.y * 3; i += kT) s_vp[v * max_pts * 3 + i] = vp[(int64_t)ti.x * 3 + i];
    }
    __syncthreads();
    double e[NVEC][2];
    const int lps = valid ? lp : 0;
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
        const double* c = s_vc + v * max_cams * kVecS + lc * kVecS;
        const double* p = s_vp + v * max_pts * 3 + lps * 3;
        double e0 = 0, e1 = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            e0 += jc[k] * c[k];
            e1 += jc[6 + k] * c[k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            e0 += jp[k] * p[k];
            e1 += jp[3 + k] * p[k];
        }
        e[v][0] = e0;
        e[v][1] = e1;
    }
    if (STORE && valid) {
        jv_out[slot] = e[0][0];
        jv_out[ns + slot] = e[0][1];
    }
    if (NVEC == 1) {
        double g[1] = {e[0][0] * e[0][0] + e[0][1] * e[0][1]};
        double* outp[1] = {scal + S_JV00};
        block_accumulate<1>(g, s_red, outp);
    } else {
        double g[3] = {e[0][0] * e[0][0] + e[0][1] * e[0][1],
                       e[0][0] * e[NVEC - 1][0] + e[0][1] * e[NVEC - 1][1],
                       e[NVEC - 1][0] * e[NVEC - 1][0] + e[NVEC - 1][1] * e[NVEC - 1][1]};
        double* outp[3] = {scal + S_JV00, scal + S_JV01, scal + S_JV11};
        block_accumulate<3>(g, s_red, outp);
    }
}

}  // namespace mmba
