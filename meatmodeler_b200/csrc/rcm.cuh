// Explicit reduced camera matrix: scaling / mirroring of the accumulated blocks, block-Jacobi preconditioner,
// and the whole PCG solve in ONE cooperative kernel (instead of 2 launches + a streaming pass over J per iteration).
//
//   S = D (J_c^T J_c - W V'^-1 W^T) D + reg I      (D = diag(1 / scale_inv) of the camera parameters)
//
// The S-build pass (tile_kernel<M_SBUILD>, kernels.cuh) accumulates the unscaled upper blocks; here they are
// scaled and expanded to the full block-CSR pattern (rcm.h) the PCG multiplies with.  S is a few MB (C2: 2.2 MB,
// C4: 5.6 MB): every PCG CTA keeps its rows in shared memory for the whole solve, so a PCG iteration costs two
// grid-wide exchanges of self-validating 16-byte lines through L2 plus a shared-memory block-sparse product instead
// of 152 B/observation of HBM traffic.
// Replaces: lsmr(J_h, f, damp=sqrt(reg)) (trf.py:494-495) together with kernels.cuh's MATVEC pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pcg.cuh"

namespace mmba {

constexpr int kRcmMaxCtas = 256;   // capacity of the per-CTA reduction slots
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

// ---- flag-in-data exchange (the idea of NCCL's LL protocol) --------------------------------------------------
// A double travels as one 16-byte line of two 64-bit words {lo | seq << 32, hi | seq << 32}: each word is a
// naturally aligned scalar of the vector access (single-copy atomic) and carries the sequence number of the
// exchange, so a reader that finds the expected number in both words holds a valid value.  No fence, no
// separate flag, no second round trip: the datum validates itself.
struct __align__(16) LLLine {
    unsigned long long w0, w1;
};
__device__ __forceinline__ void ll_store(LLLine* line, double v, unsigned seq) {
    const unsigned long long tag = (unsigned long long)seq << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(line), "l"(tag | (unsigned)__double2loint(v)),
                 "l"(tag | (unsigned)__double2hiint(v))
                 : "memory");
}
__device__ __forceinline__ bool ll_try_load(const LLLine* line, unsigned seq, double& v) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(line) : "memory");
    v = __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0);
    return (unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq;
}

// One reduction slot per CTA, 1 KB apart: every CTA polls every slot, so the slots must sit on different L2
// slices (address bits 8 and 10.. select the slice) or a handful of slices serves G^2 requests.
constexpr int kRcmSums = 7;      // partial sums of one PCG iteration: p.q, q.z, q.zq, r.q, q.q, r.z, r.r
struct RcmSlot {
    LLLine v[kRcmSums];
    LLLine pad[64 - kRcmSums];
};
static_assert(sizeof(RcmSlot) == 1024, "RcmSlot stride");

struct RcmPcgArgs {
    const double* S;            // [nnz_full][36]
    const int* rowptr;          // [Nc + 1]
    const uint16_t* lcol;       // [nnz_full] column of each block as an index into its CTA's halo list
    const int* halo_ptr;        // [G + 1]
    const int* halo_cols;       // global camera ids of every CTA's halo
    const int* own_l;           // [Nc] index of camera c in its CTA's halo list
    const double* Pinv;         // [Nc][21]
    const double* b;            // [Nc][6]
    double* x;                  // [Nc][6]  result (scaled step)
    LLLine* z;                  // [2 parities][Nc][6]  the one vector exchanged through L2 (Pinv q; classic kernel: z)
    RcmSlot* slots;             // [2 parities][kRcmMaxCtas] partial sums of the grid-wide reductions
    int* flags;                 // [0] stop code (0 = maxit reached, 1 = converged, 2 = breakdown), [1] iterations,
                                // [2] set when an exchange timed out
    double* state;              // [1] ||b||^2, [2] ||r||^2
    double* hist;               // optional [2 (maxit + 1)]: (||r_k||^2, r_k.z_k) of every iterate, k = 0 first
    int n_cams, maxit, cpc, nblk_max, nh_max, s_in_smem;
    int nsub;                   // warps that share one camera row in the product q = S p
    unsigned seq0;              // sequence numbers of this launch are seq0 + 1 ...: lines of earlier launches never match,
                                // so neither z nor slots are cleared between launches
    double rtol2;
    double atol2f;              // pcg_atol^2 ||f||^2: absolute rule, ||r||^2 <= atol2f
    double ktol2f;              // pcg_ktol^2 ||f||^2: LSMR-like rule on the smoothed residual, nu_k^2 <= ktol2f k
};

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

constexpr int kRcmPcgThreads = 256;

// shared-memory carve-up of rcm_pcg_kernel (same arithmetic on the host)
struct RcmSmem {
    int off_S, off_pinv, off_vec, off_ph, off_zh, off_zqh, off_sums, off_qp, off_hcols, off_rowptr, off_own, off_lcol, total;
};
__host__ __device__ inline RcmSmem rcm_smem(int cpc, int nblk_max, int nh_max, int s_in_smem, int n_ctas) {
    RcmSmem L{};
    int o = 0;
    L.off_S = o;
    if (s_in_smem) o += nblk_max * 36 * 8;
    L.off_pinv = o;
    o += cpc * 21 * 8;
    L.off_vec = o;
    o += cpc * 18 * 8;                       // x, r, q of the CTA's cameras
    L.off_ph = o;
    o += nh_max * 6 * 8;                     // search direction on the CTA's halo
    L.off_zh = o;
    o += nh_max * 6 * 8;                     // z on the halo
    L.off_zqh = o;
    o += nh_max * 6 * 8;                     // Pinv q on the halo, as gathered
    L.off_sums = o;
    o += n_ctas * kRcmSums * 8;              // every CTA's partial sums, as polled
    L.off_qp = o;
    o += (cpc + 8) * 6 * 8;                  // partial products [camera x sub-warp][6]: at most max(cpc, 8) items
    L.off_hcols = o;
    o += nh_max * 4;
    L.off_rowptr = o;
    o += (cpc + 1) * 4;
    L.off_own = o;
    o += cpc * 4;
    L.off_lcol = o;
    o += (nblk_max * 2 + 15) / 16 * 16;
    L.total = o;
    return L;
}

// CLASSIC two-exchange variant, kept for A/B measurements (MMBA_PCG_CLASSIC=1); the solve uses rcm_pcg_kernel below.
// Preconditioned conjugate gradients on S x = b, zero initial guess, block-Jacobi preconditioner, relative
// residual stop (the recurrences and stopping rules of pcg_update_kernel).  CTA `b` owns `cpc` consecutive
// cameras: its rows of S, their preconditioner blocks and x, r, q live in shared memory for the whole solve;
// one warp per camera (strided), lane = (block slot 0..4, row 0..5).
// Only z = Pinv r travels through L2.  Every CTA keeps its own copy of the search direction on its halo (the
// columns its rows touch) and advances it with the owner's recurrence p_j = z_j + beta p_j, so the new
// direction is never waited for: an iteration has two grid-wide reductions (p.q, then r.z and ||r||^2) and
// nothing else crosses CTAs.  Both the z entries and the partial sums travel as self-validating LL lines
// (see ll_store): a reduction is one 16-byte store per CTA and value, polled by warp 0 of every CTA; no fences.
// Partials are added in CTA order by every CTA, so all CTAs (and all ranks of a sharded solve: S and b are
// all-reduced, the PCG is replicated) take identical decisions.
// Buffer reuse is safe without extra synchronisation: a CTA rewrites its z lines / slot only after it has passed
// the next reduction, which every other CTA joins only after it has consumed the previous values.
__global__ void __launch_bounds__(kRcmPcgThreads, 1) rcm_pcg_classic_kernel(const RcmPcgArgs A) {
    extern __shared__ __align__(16) unsigned char rsm[];
    __shared__ double s_part[kRcmPcgThreads / 32][2];
    __shared__ double s_tot[2];
    __shared__ int s_dead;
    const RcmSmem L = rcm_smem(A.cpc, A.nblk_max, A.nh_max, A.s_in_smem, (int)gridDim.x);
    double* S_s = reinterpret_cast<double*>(rsm + L.off_S);
    double* pinv_s = reinterpret_cast<double*>(rsm + L.off_pinv);
    double* vec = reinterpret_cast<double*>(rsm + L.off_vec);
    double* ph = reinterpret_cast<double*>(rsm + L.off_ph);
    double* zh = reinterpret_cast<double*>(rsm + L.off_zh);
    double* qp = reinterpret_cast<double*>(rsm + L.off_qp);
    int* hcols_s = reinterpret_cast<int*>(rsm + L.off_hcols);
    int* rowptr_s = reinterpret_cast<int*>(rsm + L.off_rowptr);
    int* own_s = reinterpret_cast<int*>(rsm + L.off_own);
    uint16_t* lcol_s = reinterpret_cast<uint16_t*>(rsm + L.off_lcol);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int G = gridDim.x;
    const int a = lane % 6, slot = lane / 6;
    const bool rowlane = lane < 6;
    const int c0 = blockIdx.x * A.cpc, ncam = min(A.cpc, A.n_cams - c0);
    const int e0 = A.rowptr[c0], nblk = A.rowptr[c0 + ncam] - e0;
    const int h0 = A.halo_ptr[blockIdx.x], nh = A.halo_ptr[blockIdx.x + 1] - h0;
    if (tid == 0) s_dead = 0;
    if (A.s_in_smem) {
        // blocks are 288 bytes: 16-byte loads, four in flight per thread
        const double2* src = reinterpret_cast<const double2*>(A.S + (int64_t)e0 * 36);
        double2* dst = reinterpret_cast<double2*>(S_s);
        const int n2 = nblk * 18;
        for (int i = tid; i < n2; i += 4 * blockDim.x) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * blockDim.x < n2) v[u] = __ldg(src + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * blockDim.x < n2) dst[i + u * blockDim.x] = v[u];
        }
    }
    for (int i = tid; i < nblk; i += blockDim.x) lcol_s[i] = A.lcol[e0 + i];
    for (int i = tid; i < nh; i += blockDim.x) hcols_s[i] = A.halo_cols[h0 + i];
    for (int i = tid; i < nh * 6; i += blockDim.x) ph[i] = 0.0;
    for (int i = tid; i < ncam * 21; i += blockDim.x) pinv_s[i] = A.Pinv[(int64_t)c0 * 21 + i];
    for (int i = tid; i <= ncam; i += blockDim.x) rowptr_s[i] = A.rowptr[c0 + i] - e0;
    for (int i = tid; i < ncam; i += blockDim.x) own_s[i] = A.own_l[c0 + i];
    const double* S_rows = A.s_in_smem ? S_s : A.S + (int64_t)e0 * 36;
    __syncthreads();
    constexpr long long kSpinLimit = 1ll << 31;   // ~1 s: give up instead of hanging the device

    // grid-wide sums: `which` = 0 -> slot.pq (v0 only), 1 -> slot.rz / slot.rr.  false when the other CTAs did
    // not arrive within ~1 s (never expected)
    // gather_z: while warp 0 polls the partial sums, the other warps fetch the z lines of the same sequence number
    // on the CTA's halo into zh (both were published together, so the two waits overlap)
    auto reduce2 = [&](double v0, double v1, int which, unsigned seq, bool gather_z, double& o0, double& o1) -> bool {
        v0 = warp_sum_all(v0);
        if (which) v1 = warp_sum_all(v1);
        if (lane == 0) {
            s_part[warp][0] = v0;
            s_part[warp][1] = v1;
        }
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) {
                double t0 = 0, t1 = 0;
                for (int w = 0; w < nwarps; ++w) {
                    t0 += s_part[w][0];
                    t1 += s_part[w][1];
                }
                RcmSlot* mine = A.slots + blockIdx.x;
                if (which) {
                    ll_store(&mine->v[1], t0, seq);
                    ll_store(&mine->v[2], t1, seq);
                } else {
                    ll_store(&mine->v[0], t0, seq);
                }
            }
            // lane l polls the slots of CTAs l, l + 32, ...: all of them in flight at once
            constexpr int kPer = kRcmMaxCtas / 32;
            double a0[kPer], a1[kPer];
            unsigned pend = 0;
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                a0[u] = a1[u] = 0.0;
                if (lane + 32 * u < G) pend |= 1u << u;
            }
            const long long t_start = clock64();
            while (pend) {
                bool ok[kPer];
#pragma unroll
                for (int u = 0; u < kPer; ++u) {
                    ok[u] = false;
                    if (pend >> u & 1) {
                        const RcmSlot* sl = A.slots + lane + 32 * u;
                        ok[u] = which ? (ll_try_load(&sl->v[1], seq, a0[u]) & ll_try_load(&sl->v[2], seq, a1[u]))
                                      : ll_try_load(&sl->v[0], seq, a0[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < kPer; ++u)
                    if (ok[u]) pend &= ~(1u << u);
                if (pend && clock64() - t_start > kSpinLimit) {
                    s_dead = 1;
                    break;
                }
            }
            double s0 = 0, s1 = 0;
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                s0 += a0[u];
                s1 += a1[u];
            }
            s0 = warp_sum_all(s0);
            s1 = warp_sum_all(s1);
            if (lane == 0) {
                s_tot[0] = s0;
                s_tot[1] = s1;
            }
        }
        if (gather_z && (warp > 0 || nwarps == 1)) {
            const int first = nwarps == 1 ? tid : tid - 32, step = nwarps == 1 ? 32 : (int)blockDim.x - 32;
            const long long t_start = clock64();
            const int n6 = nh * 6;
            for (int base = first; base < n6; base += 4 * step) {
                double zv[4];
                bool ok[4];
                const LLLine* line[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * step;
                    ok[u] = true;
                    zv[u] = 0.0;
                    if (i < n6) {
                        const int j = i / 6;
                        line[u] = A.z + (int64_t)hcols_s[j] * 6 + (i - j * 6);
                        ok[u] = ll_try_load(line[u], seq, zv[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    while (!ok[u]) {
                        ok[u] = ll_try_load(line[u], seq, zv[u]);
                        if (!ok[u] && clock64() - t_start > kSpinLimit) {
                            s_dead = 1;
                            break;
                        }
                    }
                    const int i = base + u * step;
                    if (i < n6) zh[i] = zv[u];
                }
            }
        }
        __syncthreads();
        if (s_dead) {
            if (tid == 0) A.flags[2] = 1;
            return false;
        }
        o0 = s_tot[0];
        o1 = s_tot[1];
        return true;
    };
    // z = Pinv_k r for camera k of this CTA (r in lanes 0..5); valid in lanes 0..5
    auto precond = [&](int k, double r_a) {
        double z = 0;
        const double* pin = pinv_s + k * 21;
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) {
            const double rb = __shfl_sync(0xffffffffu, r_a, bb);
            z += pin[a <= bb ? tri6(a, bb) : tri6(bb, a)] * rb;
        }
        return z;
    };

    // x = 0, r = b, z = Pinv r
    double rz = 0, rr = 0;
    for (int k = warp; k < ncam; k += nwarps) {
        const int c = c0 + k;
        const double r_a = rowlane ? A.b[c * 6 + a] : 0.0;
        const double z = precond(k, r_a);
        if (rowlane) {
            vec[k * 18 + a] = 0.0;
            vec[k * 18 + 6 + a] = r_a;
            ll_store(A.z + c * 6 + a, z, A.seq0 + 1u);
            rz += r_a * z;
            rr += r_a * r_a;
        }
    }
    double rho, b2;
    if (!reduce2(rz, rr, 1, A.seq0 + 1u, true, rho, b2)) return;
    int its = 0, done = 0;
    double beta = 0.0, rr_last = b2;
    if (!(b2 > 0.0)) {
        done = 1;
    } else {
        for (int it = 0; it < A.maxit; ++it) {
            // p = z + beta p on the halo (z of this iteration was gathered together with the last reduction)
            for (int i = tid; i < nh * 6; i += blockDim.x) ph[i] = it == 0 ? zh[i] : fma(beta, ph[i], zh[i]);
            __syncthreads();
            // q = S p: nsub warps share a camera row (blocks dealt round-robin to nsub x 5 lane groups); two
            // accumulators per lane keep the dependent DFMA chains short
            const int nsub = A.nsub;
            for (int item = warp; item < ncam * nsub; item += nwarps) {
                const int k = item / nsub, sub = item - k * nsub;
                double acc0 = 0, acc1 = 0;
                if (slot < kRcmSlots) {
                    const int e1 = rowptr_s[k + 1];
                    for (int e = rowptr_s[k] + slot + kRcmSlots * sub; e < e1; e += kRcmSlots * nsub) {
                        const double* srow = S_rows + e * 36 + a * 6;
                        const double* pj = ph + lcol_s[e] * 6;
                        if (A.s_in_smem) {
                            acc0 += srow[0] * pj[0];
                            acc1 += srow[1] * pj[1];
                            acc0 += srow[2] * pj[2];
                            acc1 += srow[3] * pj[3];
                            acc0 += srow[4] * pj[4];
                            acc1 += srow[5] * pj[5];
                        } else {
                            acc0 += __ldg(srow + 0) * pj[0];
                            acc1 += __ldg(srow + 1) * pj[1];
                            acc0 += __ldg(srow + 2) * pj[2];
                            acc1 += __ldg(srow + 3) * pj[3];
                            acc0 += __ldg(srow + 4) * pj[4];
                            acc1 += __ldg(srow + 5) * pj[5];
                        }
                    }
                }
                const double acc = acc0 + acc1;
                double q = acc;
                q += __shfl_down_sync(0xffffffffu, acc, 6);
                q += __shfl_down_sync(0xffffffffu, acc, 12);
                q += __shfl_down_sync(0xffffffffu, acc, 18);
                q += __shfl_down_sync(0xffffffffu, acc, 24);
                if (rowlane) qp[item * 6 + a] = q;
            }
            __syncthreads();
            double pq = 0;
            for (int k = warp; k < ncam; k += nwarps) {
                if (rowlane) {
                    double q = 0;
                    for (int sub = 0; sub < nsub; ++sub) q += qp[(k * nsub + sub) * 6 + a];
                    vec[k * 18 + 12 + a] = q;
                    pq += ph[own_s[k] * 6 + a] * q;
                }
            }
            double pq_tot, unused;
            if (!reduce2(pq, 0.0, 0, A.seq0 + (unsigned)it + 1u, false, pq_tot, unused)) return;
            const double alpha = rho / pq_tot;
            rz = 0;
            rr = 0;
            for (int k = warp; k < ncam; k += nwarps) {
                double* v = vec + k * 18;
                double r_a = 0;
                if (rowlane) {
                    v[a] += alpha * ph[own_s[k] * 6 + a];
                    r_a = v[6 + a] - alpha * v[12 + a];
                    v[6 + a] = r_a;
                }
                const double z = precond(k, r_a);
                if (rowlane) {
                    ll_store(A.z + (c0 + k) * 6 + a, z, A.seq0 + (unsigned)it + 2u);
                    rz += r_a * z;
                    rr += r_a * r_a;
                }
            }
            double rz_tot, rr_tot;
            if (!reduce2(rz, rr, 1, A.seq0 + (unsigned)it + 2u, true, rz_tot, rr_tot)) return;
            its = it + 1;
            rr_last = rr_tot;
            if (rr_tot <= A.rtol2 * b2 || rr_tot <= A.atol2f) {
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
    for (int k = warp; k < ncam; k += nwarps)
        if (rowlane) A.x[(c0 + k) * 6 + a] = vec[k * 18 + a];
    if (blockIdx.x == 0 && tid == 0) {
        A.flags[0] = done;
        A.flags[1] = its;
        A.state[1] = b2;
        A.state[2] = rr_last;
    }
}


// ---- the PCG kernel of the solve: ONE grid-wide exchange per iteration -------------------------------------------
// Same method (preconditioned CG, zero initial guess, block-Jacobi), reorganised so that everything that crosses
// CTAs in an iteration travels in one round trip through L2:
//     q = S p                       own rows, p on the CTA's halo in shared memory
//     w = Pinv q                    own rows                                   -> published (LL lines)
//     partial sums over the own rows of  p.q, q.z, q.w, r.q, q.q, r.z, r.r      -> published (one 112-byte slot)
//     ---- exchange: poll every CTA's slot, gather w on the halo ----
//     alpha = (r.z) / (p.q)
//     x += alpha p ;  r -= alpha q ;  z -= alpha w            (z = Pinv r by recurrence, on the whole halo)
//     rho' = r.z - 2 alpha q.z + alpha^2 q.w ,  ||r'||^2 = r.r - 2 alpha r.q + alpha^2 q.q
//     beta = rho' / (r.z) ;  p = z + beta p                  (on the whole halo: no second exchange)
// The inner products of the NEXT residual follow from this iteration's sums (r' = r - alpha q, z' = z - alpha w), so
// beta needs no second reduction; they are one-step predictions from exactly reduced values — r.z and r.r of the
// current iterate ride in the same exchange — so nothing accumulates.  Both the slots and the w lines have two
// parities: a CTA rewrites parity s only after it has passed exchange s + 1, which every other CTA joins only
// after it has consumed exchange s.
// Stopping rules on ||r'||: relative (rtol), absolute (atol ||f||), and the counterpart of LSMR's test 2
// (lsmr.py:430-459: ||A^T res|| <= atol ||A|| ||res|| with the running Frobenius estimate ||A|| ~ sqrt(k)): LSMR is
// a minimal-residual method on the normal equations, and the minimal-residual norm of a CG process is the smoothed
// norm 1 / nu_k^2 = sum_{j<=k} 1 / ||r_j||^2, hence  nu_k <= pcg_ktol sqrt(k) ||f||.
// Breakdown (p.q <= 0 or a non-finite sum): the update is NOT applied; x keeps the last finite iterate, flags[0] = 2.
__global__ void __launch_bounds__(kRcmPcgThreads, 1) rcm_pcg_kernel(const RcmPcgArgs A) {
    extern __shared__ __align__(16) unsigned char rsm[];
    __shared__ double s_part[kRcmPcgThreads / 32][kRcmSums];
    __shared__ double s_tot[kRcmSums];
    __shared__ int s_dead;
    const int G = gridDim.x;
    const RcmSmem L = rcm_smem(A.cpc, A.nblk_max, A.nh_max, A.s_in_smem, G);
    double* S_s = reinterpret_cast<double*>(rsm + L.off_S);
    double* pinv_s = reinterpret_cast<double*>(rsm + L.off_pinv);
    double* vec = reinterpret_cast<double*>(rsm + L.off_vec);      // [camera][x 6 | r 6 | q 6]
    double* ph = reinterpret_cast<double*>(rsm + L.off_ph);
    double* zh = reinterpret_cast<double*>(rsm + L.off_zh);
    double* wh = reinterpret_cast<double*>(rsm + L.off_zqh);
    double* sums_s = reinterpret_cast<double*>(rsm + L.off_sums);
    double* qp = reinterpret_cast<double*>(rsm + L.off_qp);
    int* hcols_s = reinterpret_cast<int*>(rsm + L.off_hcols);
    int* rowptr_s = reinterpret_cast<int*>(rsm + L.off_rowptr);
    int* own_s = reinterpret_cast<int*>(rsm + L.off_own);
    uint16_t* lcol_s = reinterpret_cast<uint16_t*>(rsm + L.off_lcol);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int a = lane % 6, slot = lane / 6;
    const bool rowlane = lane < 6;
    const int c0 = blockIdx.x * A.cpc, ncam = min(A.cpc, A.n_cams - c0);
    const int e0 = A.rowptr[c0], nblk = A.rowptr[c0 + ncam] - e0;
    const int h0 = A.halo_ptr[blockIdx.x], nh = A.halo_ptr[blockIdx.x + 1] - h0;
    if (tid == 0) s_dead = 0;
    if (A.s_in_smem) {
        const double2* src = reinterpret_cast<const double2*>(A.S + (int64_t)e0 * 36);
        double2* dst = reinterpret_cast<double2*>(S_s);
        const int n2 = nblk * 18;
        for (int i = tid; i < n2; i += 4 * blockDim.x) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * blockDim.x < n2) v[u] = __ldg(src + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * blockDim.x < n2) dst[i + u * blockDim.x] = v[u];
        }
    }
    for (int i = tid; i < nblk; i += blockDim.x) lcol_s[i] = A.lcol[e0 + i];
    for (int i = tid; i < nh; i += blockDim.x) hcols_s[i] = A.halo_cols[h0 + i];
    for (int i = tid; i < ncam * 21; i += blockDim.x) pinv_s[i] = A.Pinv[(int64_t)c0 * 21 + i];
    for (int i = tid; i <= ncam; i += blockDim.x) rowptr_s[i] = A.rowptr[c0 + i] - e0;
    for (int i = tid; i < ncam; i += blockDim.x) own_s[i] = A.own_l[c0 + i];
    const double* S_rows = A.s_in_smem ? S_s : A.S + (int64_t)e0 * 36;
    __syncthreads();
    const int own0 = own_s[0];      // the CTA's own cameras are consecutive entries of its (ascending) halo list
    constexpr long long kSpinLimit = 1ll << 31;   // ~1 s: give up instead of hanging the device

    // One exchange: the per-thread partial sums v[0..nv) are reduced over the grid (every CTA adds the published
    // CTA totals in CTA order: identical bits everywhere) and the lines of sequence number `seq` on the CTA's halo
    // are gathered into `gather`.  All threads share the polling: item i < G nv is a partial sum, the rest are
    // halo lines; eight polls in flight per thread.
    auto exchange = [&](const double (&v)[kRcmSums], int nv, unsigned seq, double* gather, double (&tot)[kRcmSums]) -> bool {
        const int par = (int)(seq & 1u);
#pragma unroll
        for (int i = 0; i < kRcmSums; ++i)
            if (i < nv) {
                const double w = warp_sum_all(v[i]);
                if (lane == 0) s_part[warp][i] = w;
            }
        __syncthreads();
        RcmSlot* slots = A.slots + par * kRcmMaxCtas;
        if (tid < nv) {
            double t = 0;
            for (int w = 0; w < nwarps; ++w) t += s_part[w][tid];
            ll_store(&slots[blockIdx.x].v[tid], t, seq);
        }
        const LLLine* zl = A.z + (int64_t)par * 6 * A.n_cams;
        const int n_sum = G * nv, n_items = n_sum + nh * 6;
        const long long t_start = clock64();
        constexpr int kFly = 8;      // polls in flight per thread: G nv + 6 nh <= 8 x 256 items go out in one batch
        for (int base = tid; base < n_items; base += kFly * (int)blockDim.x) {
            const LLLine* line[kFly];
            double val[kFly];
            bool ok[kFly];
#pragma unroll
            for (int u = 0; u < kFly; ++u) {
                const int i = base + u * (int)blockDim.x;
                ok[u] = true;
                val[u] = 0.0;
                line[u] = nullptr;
                if (i < n_items) {
                    if (i < n_sum) {
                        const int cta = i / nv;
                        line[u] = &slots[cta].v[i - cta * nv];
                    } else {
                        const int j = (i - n_sum) / 6;
                        line[u] = zl + (int64_t)hcols_s[j] * 6 + (i - n_sum - j * 6);
                    }
                    ok[u] = ll_try_load(line[u], seq, val[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kFly; ++u) {
                while (!ok[u]) {
                    ok[u] = ll_try_load(line[u], seq, val[u]);
                    if (!ok[u] && clock64() - t_start > kSpinLimit) {
                        s_dead = 1;
                        break;
                    }
                }
                const int i = base + u * (int)blockDim.x;
                if (i < n_items) {
                    if (i < n_sum) sums_s[i] = val[u];
                    else gather[i - n_sum] = val[u];
                }
            }
        }
        __syncthreads();
        if (s_dead) {
            if (tid == 0) A.flags[2] = 1;
            return false;
        }
        for (int vi = warp; vi < nv; vi += nwarps) {
            double s = 0;
            for (int c = lane; c < G; c += 32) s += sums_s[c * nv + vi];
            s = warp_sum_all(s);
            if (lane == 0) s_tot[vi] = s;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kRcmSums; ++i) tot[i] = i < nv ? s_tot[i] : 0.0;
        return true;
    };
    // Pinv_k v for camera k of this CTA (v in lanes 0..5); valid in lanes 0..5
    auto precond = [&](int k, double v_a) {
        double z = 0;
        const double* pin = pinv_s + k * 21;
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) {
            const double vb = __shfl_sync(0xffffffffu, v_a, bb);
            z += pin[a <= bb ? tri6(a, bb) : tri6(bb, a)] * vb;
        }
        return z;
    };

    // x = 0, r = b, z = Pinv r (published with sequence number seq0 + 1 and gathered on the halo), p = z
    double v[kRcmSums], tot[kRcmSums];
#pragma unroll
    for (int i = 0; i < kRcmSums; ++i) v[i] = 0.0;
    {
        const unsigned seq = A.seq0 + 1u;
        LLLine* zl = A.z + (int64_t)(seq & 1u) * 6 * A.n_cams;
        for (int k = warp; k < ncam; k += nwarps) {
            const int c = c0 + k;
            const double r_a = rowlane ? A.b[c * 6 + a] : 0.0;
            const double z = precond(k, r_a);
            if (rowlane) {
                vec[k * 18 + a] = 0.0;
                vec[k * 18 + 6 + a] = r_a;
                ll_store(zl + c * 6 + a, z, seq);
                v[0] += r_a * z;
                v[1] += r_a * r_a;
            }
        }
        if (!exchange(v, 2, seq, zh, tot)) return;
    }
    const double rho0 = tot[0], b2 = tot[1];
    int its = 0, done = 0;
    double rr_last = b2;
    double inv_nu2 = b2 > 0.0 ? 1.0 / b2 : 0.0;     // smoothed (minimal-residual) norm: 1 / nu^2 = sum 1 / ||r_j||^2
    if (A.hist && blockIdx.x == 0 && tid == 0) {
        A.hist[0] = b2;
        A.hist[1] = rho0;
    }
    if (!(b2 > 0.0) || !isfinite(rho0)) {
        done = 1;      // zero (or non-finite) right-hand side: x = 0
    } else {
        for (int i = tid; i < nh * 6; i += blockDim.x) ph[i] = zh[i];
        __syncthreads();
        const int nsub = A.nsub;
        for (int it = 0; it < A.maxit; ++it) {
            const unsigned seq = A.seq0 + (unsigned)it + 2u;
            // q = S p: nsub warps share a camera row (blocks dealt round-robin to nsub x 5 lane groups)
            for (int item = warp; item < ncam * nsub; item += nwarps) {
                const int k = item / nsub, sub = item - k * nsub;
                double acc0 = 0, acc1 = 0;
                if (slot < kRcmSlots) {
                    const int e1 = rowptr_s[k + 1];
                    for (int e = rowptr_s[k] + slot + kRcmSlots * sub; e < e1; e += kRcmSlots * nsub) {
                        const double* srow = S_rows + e * 36 + a * 6;
                        const double* pj = ph + lcol_s[e] * 6;
                        if (A.s_in_smem) {
                            acc0 += srow[0] * pj[0];
                            acc1 += srow[1] * pj[1];
                            acc0 += srow[2] * pj[2];
                            acc1 += srow[3] * pj[3];
                            acc0 += srow[4] * pj[4];
                            acc1 += srow[5] * pj[5];
                        } else {
                            acc0 += __ldg(srow + 0) * pj[0];
                            acc1 += __ldg(srow + 1) * pj[1];
                            acc0 += __ldg(srow + 2) * pj[2];
                            acc1 += __ldg(srow + 3) * pj[3];
                            acc0 += __ldg(srow + 4) * pj[4];
                            acc1 += __ldg(srow + 5) * pj[5];
                        }
                    }
                }
                const double acc = acc0 + acc1;
                double q = acc;
                q += __shfl_down_sync(0xffffffffu, acc, 6);
                q += __shfl_down_sync(0xffffffffu, acc, 12);
                q += __shfl_down_sync(0xffffffffu, acc, 18);
                q += __shfl_down_sync(0xffffffffu, acc, 24);
                if (rowlane) qp[item * 6 + a] = q;
            }
            __syncthreads();
            // w = Pinv q on the own rows (published), and the seven partial sums
#pragma unroll
            for (int i = 0; i < kRcmSums; ++i) v[i] = 0.0;
            LLLine* wl = A.z + (int64_t)(seq & 1u) * 6 * A.n_cams;
            for (int k = warp; k < ncam; k += nwarps) {
                double q = 0, pk = 0, zk = 0, rk = 0;
                if (rowlane) {
                    for (int sub = 0; sub < nsub; ++sub) q += qp[(k * nsub + sub) * 6 + a];
                    vec[k * 18 + 12 + a] = q;
                    pk = ph[(own0 + k) * 6 + a];
                    zk = zh[(own0 + k) * 6 + a];
                    rk = vec[k * 18 + 6 + a];
                }
                const double w = precond(k, q);
                if (rowlane) {
                    ll_store(wl + (c0 + k) * 6 + a, w, seq);
                    v[0] += pk * q;
                    v[1] += q * zk;
                    v[2] += q * w;
                    v[3] += rk * q;
                    v[4] += q * q;
                    v[5] += rk * zk;
                    v[6] += rk * rk;
                }
            }
            if (!exchange(v, kRcmSums, seq, wh, tot)) return;
            const double pq = tot[0], rho = tot[5];
            const double alpha = rho / pq;
            const double rho_n = fma(alpha, fma(alpha, tot[2], -2.0 * tot[1]), rho);
            double rr_n = fma(alpha, fma(alpha, tot[4], -2.0 * tot[3]), tot[6]);
            if (!(pq > 0.0) || !(rho > 0.0) || !isfinite(alpha) || !isfinite(rho_n) || !isfinite(rr_n)) {
                done = 2;      // breakdown: keep the last finite iterate
                break;
            }
            if (rr_n < 0.0) rr_n = 0.0;
            const double beta = rho_n > 0.0 ? rho_n / rho : 0.0;
            // x += alpha p, r -= alpha q (own rows); z -= alpha w, p = z + beta p (whole halo)
            for (int i = tid; i < nh * 6; i += blockDim.x) {
                const int j = i / 6;
                const double p_old = ph[i];
                const double z_new = fma(-alpha, wh[i], zh[i]);
                zh[i] = z_new;
                ph[i] = fma(beta, p_old, z_new);
                const int k = j - own0;
                if (k >= 0 && k < ncam) {
                    double* vk = vec + k * 18 + (i - j * 6);
                    vk[0] = fma(alpha, p_old, vk[0]);
                    vk[6] = fma(-alpha, vk[12], vk[6]);
                }
            }
            its = it + 1;
            rr_last = rr_n;
            if (A.hist && blockIdx.x == 0 && tid == 0) {
                A.hist[2 * its] = rr_n;
                A.hist[2 * its + 1] = rho_n;
            }
            inv_nu2 += rr_n > 0.0 ? 1.0 / rr_n : INFINITY;
            if (rr_n <= A.rtol2 * b2 || rr_n <= A.atol2f || 1.0 <= inv_nu2 * A.ktol2f * (double)its) {
                done = 1;
                break;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int k = warp; k < ncam; k += nwarps)
        if (rowlane) A.x[(c0 + k) * 6 + a] = vec[k * 18 + a];
    if (blockIdx.x == 0 && tid == 0) {
        A.flags[0] = done;
        A.flags[1] = its;
        A.state[1] = b2;
        A.state[2] = rr_last;
    }
}

// Round trip of one self-validating line between two CTAs that sit far apart (first and last CTA of a grid that
// fills the device): the latency unit of the PCG's grid-wide exchange (bench.py: latency model of rcm_pcg_kernel).
__global__ void ll_pingpong_kernel(LLLine* a, LLLine* b, int iters, unsigned seq0) {
    if (threadIdx.x != 0) return;
    double v = 0.0;
    if (blockIdx.x == 0) {
        for (int i = 0; i < iters; ++i) {
            const unsigned seq = seq0 + 1u + (unsigned)i;
            ll_store(a, (double)i, seq);
            const long long t0 = clock64();
            while (!ll_try_load(b, seq, v))
                if (clock64() - t0 > (1ll << 31)) return;
        }
    } else if (blockIdx.x == gridDim.x - 1) {
        for (int i = 0; i < iters; ++i) {
            const unsigned seq = seq0 + 1u + (unsigned)i;
            const long long t0 = clock64();
            while (!ll_try_load(a, seq, v))
                if (clock64() - t0 > (1ll << 31)) return;
            ll_store(b, v, seq);
        }
    }
}

}  // namespace mmba
