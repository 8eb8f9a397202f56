// Explicit reduced camera matrix: scaling / mirroring of the accumulated blocks, block-Jacobi preconditioner,
// and the whole PCG solve in ONE cooperative kernel (grid barriers instead of 2 launches + a streaming pass
// over J per iteration).
//
//   S = D (J_c^T J_c - W V'^-1 W^T) D + reg I      (D = diag(1 / scale_inv) of the camera parameters)
//
// The S-build pass (tile_kernel<M_SBUILD>, kernels.cuh) accumulates the unscaled upper blocks; here they are
// scaled and expanded to the full block-CSR pattern (rcm.h) the PCG multiplies with.  S is a few MB (C2: 2.2 MB,
// C4: 6.6 MB): it stays in L1/L2 for the whole solve, so a PCG iteration costs two grid barriers plus an
// L1-resident block-sparse product instead of 152 B/observation of HBM traffic.
// Replaces: lsmr(J_h, f, damp=sqrt(reg)) (trf.py:494-495) together with kernels.cuh's MATVEC pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pcg.cuh"

namespace mmba {

constexpr int kRcmMaxCtas = 256;   // capacity of the per-CTA partial-sum slots
constexpr int kRcmSlots = 5;       // lanes of a warp: 5 block slots x 6 rows (lanes 30, 31 idle)

// S[k] = D_i T[src(k)]^(T) D_j (+ reg I on the diagonal); one thread per entry
__global__ void __launch_bounds__(256) rcm_finalize_kernel(const double* __restrict__ Tup, const int* __restrict__ rows,
                                                           const int* __restrict__ cols, const int* __restrict__ src,
                                                           const double* __restrict__ sinv, double reg,
                                                           double* __restrict__ S, int64_t n_entries) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_entries) return;
    const int64_t k = idx / 36;
    const int e = (int)(idx - k * 36), a = e / 6, b = e - a * 6;
    const int s = src[k];
    const int64_t sb = (int64_t)(s & 0x7fffffff);
    const int i = rows[k], j = cols[k];
    double v = Tup[sb * 36 + (s < 0 ? b * 6 + a : a * 6 + b)];
    v = v * ((1.0 / sinv[i * 6 + a]) * (1.0 / sinv[j * 6 + b]));
    if (i == j && a == b) v += reg;
    S[idx] = v;
}

// per camera: Pinv = (S_cc)^-1 (block-Jacobi), b = d o (g_c - y); y is cleared for the next pass
__global__ void __launch_bounds__(kCamBlock) rcm_prepare_kernel(const double* __restrict__ S, const int* __restrict__ diag,
                                                                const double* __restrict__ gc, double* __restrict__ y,
                                                                const double* __restrict__ sinv, double* __restrict__ Pinv,
                                                                double* __restrict__ b, int n_cams) {
    const int c = blockIdx.x * kCamBlock + threadIdx.x;
    if (c >= n_cams) return;
    const double* blk = S + (int64_t)diag[c] * 36;
    double s[21], inv[21];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int bb = a; bb < 6; ++bb) s[tri6(a, bb)] = 0.5 * (blk[a * 6 + bb] + blk[bb * 6 + a]);
    sym6_inverse(s, inv);
#pragma unroll
    for (int i = 0; i < 21; ++i) Pinv[c * 21 + i] = inv[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        b[c * 6 + k] = (gc[c * 6 + k] - y[c * 6 + k]) / sinv[c * 6 + k];
        y[c * 6 + k] = 0.0;
    }
}

struct RcmPcgArgs {
    const double* S;        // [nnz_full][36]
    const int* rowptr;      // [Nc + 1]
    const int* cols;        // [nnz_full]
    const double* Pinv;     // [Nc][21]
    const double* b;        // [Nc][6]
    double* x;              // [Nc][6]  result (scaled step)
    double* z;              // [Nc][6]  preconditioned residual, shared through L2
    double* p0;             // [Nc][6]  search direction, two parities
    double* p1;
    double* part;           // [2][kRcmMaxCtas][2] per-CTA partial sums
    unsigned* bar;          // grid-barrier counter, zero at launch
    int* flags;             // [0] stop code (0 = maxit reached, 1 = converged, 2 = breakdown), [1] iterations,
                            // [2] set when a grid barrier timed out
    double* state;          // [1] ||b||^2, [2] ||r||^2
    int n_cams, maxit, kmax;
    double rtol2;
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (co-resident, cooperatively launched) grid.  Returns false when the other CTAs did not
// arrive within ~1 s (never expected; the kernel then gives up instead of hanging the device).
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned target, int* s_dead) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const long long t0 = clock64();
        while (ld_acquire_gpu_u32(bar) < target) {
            if (clock64() - t0 > (1ll << 31)) {
                *s_dead = 1;
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
    return *s_dead == 0;
}

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Preconditioned conjugate gradients on S x = b, zero initial guess, block-Jacobi preconditioner, relative
// residual stop (the same recurrences and stopping rules as pcg_update_kernel).  One warp per camera (strided):
// lane = (block slot 0..4, row 0..5); lanes 0..5 own the camera's x, r, p, z, q entries (kept in shared memory).
// Two grid barriers per iteration: after q = S p (for p.q) and after z = Pinv r (for r.z, ||r||^2).  The
// search direction of OTHER cameras is never waited for: p_j = z_j + beta p_j(old) is recomputed by the reader
// from the two published vectors with the same fma the owner uses.
// Every reduction is a fixed-order sum evaluated identically by all warps of all CTAs (and of all ranks of a
// sharded solve: S and b are all-reduced, the PCG itself is replicated), so all take the same decisions.
__global__ void __launch_bounds__(256, 1) rcm_pcg_kernel(const RcmPcgArgs A) {
    extern __shared__ double s_vec[];   // [warp][k][x r p z q][6]
    __shared__ double s_part[8][2];
    __shared__ int s_dead;
    if (threadIdx.x == 0) s_dead = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int G = gridDim.x, wtot = G * nwarps, gw = blockIdx.x * nwarps + warp;
    const int a = lane % 6, slot = lane / 6;
    const bool rowlane = lane < 6;
    double* my = s_vec + (size_t)warp * A.kmax * 30;
    const int n = A.n_cams;
    unsigned nbar = 0;

    auto reduce2 = [&](double v0, double v1, int phase, double& o0, double& o1) -> bool {
        v0 = warp_sum_all(v0);
        v1 = warp_sum_all(v1);
        if (lane == 0) {
            s_part[warp][0] = v0;
            s_part[warp][1] = v1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t0 = 0, t1 = 0;
            for (int w = 0; w < nwarps; ++w) {
                t0 += s_part[w][0];
                t1 += s_part[w][1];
            }
            double* dst = A.part + ((size_t)phase * kRcmMaxCtas + blockIdx.x) * 2;
            dst[0] = t0;
            dst[1] = t1;
        }
        ++nbar;
        if (!grid_barrier(A.bar, nbar * (unsigned)G, &s_dead)) {
            if (threadIdx.x == 0) A.flags[2] = 1;
            return false;
        }
        double s0 = 0, s1 = 0;
        for (int c = lane; c < G; c += 32) {
            s0 += __ldcg(A.part + ((size_t)phase * kRcmMaxCtas + c) * 2);
            s1 += __ldcg(A.part + ((size_t)phase * kRcmMaxCtas + c) * 2 + 1);
        }
        o0 = warp_sum_all(s0);
        o1 = warp_sum_all(s1);
        return true;
    };
    // z = Pinv_c r for the camera of this warp (r in lanes 0..5); valid in lanes 0..5
    auto precond = [&](int c, double r_a) {
        double z = 0;
        const double* pin = A.Pinv + (int64_t)c * 21;
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) {
            const double rb = __shfl_sync(0xffffffffu, r_a, bb);
            z += __ldg(pin + (a <= bb ? tri6(a, bb) : tri6(bb, a))) * rb;
        }
        return z;
    };

    // x = 0, r = b, z = Pinv r
    double rz = 0, rr = 0;
    for (int k = 0; k < A.kmax; ++k) {
        const int c = gw + k * wtot;
        if (c >= n) break;
        double* v = my + k * 30;
        const double r_a = rowlane ? A.b[c * 6 + a] : 0.0;
        const double z = precond(c, r_a);
        if (rowlane) {
            v[a] = 0.0;
            v[6 + a] = r_a;
            v[12 + a] = 0.0;
            v[18 + a] = z;
            A.z[c * 6 + a] = z;
            rz += r_a * z;
            rr += r_a * r_a;
        }
    }
    double rho, b2;
    if (!reduce2(rz, rr, 1, rho, b2)) return;
    int its = 0, done = 0;
    double beta = 0.0, rr_last = b2;
    if (!(b2 > 0.0)) {
        done = 1;
    } else {
        for (int it = 0; it < A.maxit; ++it) {
            const double* pold = (it & 1) ? A.p0 : A.p1;
            double* pnew = (it & 1) ? A.p1 : A.p0;
            double pq = 0;
            for (int k = 0; k < A.kmax; ++k) {
                const int c = gw + k * wtot;
                if (c >= n) break;
                double* v = my + k * 30;
                double pown = 0;
                if (rowlane) {
                    pown = it == 0 ? v[18 + a] : fma(beta, v[12 + a], v[18 + a]);
                    v[12 + a] = pown;
                    pnew[c * 6 + a] = pown;
                }
                double acc = 0;
                if (slot < kRcmSlots) {
                    const int e1 = __ldg(A.rowptr + c + 1);
                    for (int e = __ldg(A.rowptr + c) + slot; e < e1; e += kRcmSlots) {
                        const int j = __ldg(A.cols + e);
                        const double* srow = A.S + (int64_t)e * 36 + a * 6;
#pragma unroll
                        for (int bb = 0; bb < 6; ++bb) {
                            const double zj = __ldcg(A.z + j * 6 + bb);
                            const double pj = it == 0 ? zj : fma(beta, __ldcg(pold + j * 6 + bb), zj);
                            acc += __ldg(srow + bb) * pj;
                        }
                    }
                }
                double q = acc;
                q += __shfl_down_sync(0xffffffffu, acc, 6);
                q += __shfl_down_sync(0xffffffffu, acc, 12);
                q += __shfl_down_sync(0xffffffffu, acc, 18);
                q += __shfl_down_sync(0xffffffffu, acc, 24);
                if (rowlane) {
                    v[24 + a] = q;
                    pq += pown * q;
                }
            }
            double pq_tot, unused;
            if (!reduce2(pq, 0.0, 0, pq_tot, unused)) return;
            const double alpha = rho / pq_tot;
            rz = 0;
            rr = 0;
            for (int k = 0; k < A.kmax; ++k) {
                const int c = gw + k * wtot;
                if (c >= n) break;
                double* v = my + k * 30;
                double r_a = 0;
                if (rowlane) {
                    v[a] += alpha * v[12 + a];
                    r_a = v[6 + a] - alpha * v[24 + a];
                    v[6 + a] = r_a;
                }
                const double z = precond(c, r_a);
                if (rowlane) {
                    v[18 + a] = z;
                    A.z[c * 6 + a] = z;
                    rz += r_a * z;
                    rr += r_a * r_a;
                }
            }
            double rz_tot, rr_tot;
            if (!reduce2(rz, rr, 1, rz_tot, rr_tot)) return;
            its = it + 1;
            rr_last = rr_tot;
            if (rr_tot <= A.rtol2 * b2) {
                done = 1;
                break;
            }
            if (!(pq_tot > 0.0) || !isfinite(rr_tot) || !(rz_tot > 0.0)) {
                done = 2;
                break;
            }
            beta = rz_tot / rho;
            rho = rz_tot;
        }
    }
    for (int k = 0; k < A.kmax; ++k) {
        const int c = gw + k * wtot;
        if (c >= n) break;
        if (rowlane) A.x[c * 6 + a] = my[k * 30 + a];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.flags[0] = done;
        A.flags[1] = its;
        A.state[1] = b2;
        A.state[2] = rr_last;
    }
}

}  // namespace mmba
