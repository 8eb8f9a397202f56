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
    LLLine* z;                  // [2 parities][Nc][6]  the one vector exchanged through L2 (z = Pinv r at the start, then Pinv q)
    RcmSlot* slots;             // [2 parities][kRcmMaxCtas] partial sums of the grid-wide reductions
    int* flags;                 // [0] stop code (0 = maxit reached, 1 = converged, 2 = breakdown), [1] iterations,
                                // [2] set when an exchange timed out
    double* state;              // [1] ||b||^2, [2] ||r||^2
    double* hist;               // optional [2 (maxit + 1)]: (||r_k||^2, r_k.z_k) of every iterate, k = 0 first
    long long* phase;           // optional [8]: cycles CTA 0 spent per phase of the iteration loop (diagnostics)
    int n_cams, maxit, cpc, nblk_max, nh_max, s_in_smem;
    int fault_cta;              // fault injection (tests, MMBA_FAULT_PCG_CTA): this CTA never joins the exchanges; -1 = none
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

constexpr int kRcmRowStride = 7;         // doubles per 6-entry block row of S in shared memory (lanes 7 apart hit 16 distinct banks)
constexpr int kRcmPcgThreads = 512;      // 16 warps, one CTA per SM: every phase of an iteration is latency bound, more warps hide it

// shared-memory carve-up of rcm_pcg_kernel (same arithmetic on the host)
struct RcmSmem {
    int off_S, off_pinv, off_vec, off_ph, off_zh, off_zqh, off_sums, off_qp, off_hcols, off_rowptr, off_own, off_lcol, total;
};
__host__ __device__ inline RcmSmem rcm_smem(int cpc, int nblk_max, int nh_max, int s_in_smem, int n_ctas) {
    RcmSmem L{};
    int o = 0;
    L.off_S = o;
    if (s_in_smem) o += nblk_max * 6 * kRcmRowStride * 8;   // block rows padded to 7 doubles: conflict-free row products
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
    o += (nblk_max > cpc + 8 ? nblk_max : cpc + 8) * 6 * 8;   // row products of q = S p: [block][6]
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
    double* prod = reinterpret_cast<double*>(rsm + L.off_qp);      // [block][6] row products of q = S p
    int* hcols_s = reinterpret_cast<int*>(rsm + L.off_hcols);
    int* rowptr_s = reinterpret_cast<int*>(rsm + L.off_rowptr);
    uint16_t* lcol_s = reinterpret_cast<uint16_t*>(rsm + L.off_lcol);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if ((int)blockIdx.x == A.fault_cta) return;    // (tests: a lost CTA must surface as an error code, not as a hang)
    const int a = lane % 6, slot = lane / 6;       // 6-lane groups: five cameras per warp, lanes 30 and 31 idle
    const int c0 = blockIdx.x * A.cpc, ncam = min(A.cpc, A.n_cams - c0);
    const int e0 = A.rowptr[c0], nblk = A.rowptr[c0 + ncam] - e0;
    const int h0 = A.halo_ptr[blockIdx.x], nh = A.halo_ptr[blockIdx.x + 1] - h0;
    if (tid == 0) s_dead = 0;
    if (A.s_in_smem) {
        // block rows (6 doubles) land kRcmRowStride apart
        const double2* src = reinterpret_cast<const double2*>(A.S + (int64_t)e0 * 36);
        const int n2 = nblk * 18;
        for (int i = tid; i < n2; i += 4 * blockDim.x) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * blockDim.x < n2) v[u] = __ldg(src + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i2 = i + u * (int)blockDim.x;
                if (i2 < n2) {
                    const int row = i2 / 3, k = (i2 - row * 3) * 2;
                    S_s[row * kRcmRowStride + k] = v[u].x;
                    S_s[row * kRcmRowStride + k + 1] = v[u].y;
                }
            }
        }
    }
    for (int i = tid; i < nblk; i += blockDim.x) lcol_s[i] = A.lcol[e0 + i];
    for (int i = tid; i < nh; i += blockDim.x) hcols_s[i] = A.halo_cols[h0 + i];
    for (int i = tid; i < ncam * 21; i += blockDim.x) pinv_s[i] = A.Pinv[(int64_t)c0 * 21 + i];
    for (int i = tid; i <= ncam; i += blockDim.x) rowptr_s[i] = A.rowptr[c0 + i] - e0;
    const double* S_rows = A.s_in_smem ? S_s : A.S + (int64_t)e0 * 36;
    const int own0 = A.own_l[c0];      // the CTA's own cameras are consecutive entries of its (ascending) halo list
    __syncthreads();
    constexpr long long kSpinLimit = 1ll << 31;   // ~1 s: give up instead of hanging the device
    // diagnostics (-DMMBA_PHASE_TIMING builds only): cycles per phase as seen by thread 0 of CTA 0
#ifdef MMBA_PHASE_TIMING
    long long t_last = 0, phc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto lap = [&](int k) {
        if (A.phase) {
            const long long t = clock64();
            phc[k] += t - t_last;
            t_last = t;
        }
    };
#else
    auto lap = [](int) {};
#endif
    // the warps that own camera rows in the "row" phase (five cameras each); the others go straight to polling
    // (measured on C4 / C2: 4.15 / 4.20 us per iteration; one camera per warp with the row's blocks spread over the
    // lanes: 4.72 / 4.04)
    const int row_warps = min(nwarps, (ncam + 4) / 5);
    auto row_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(32 * row_warps) : "memory"); };

    // Publish: the per-thread partial sums v[0..nv) of the row warps are reduced over the CTA and stored to this
    // CTA's slot (every CTA later adds the published CTA totals in CTA order: identical bits everywhere).
    auto publish = [&](const double (&v)[kRcmSums], int nv, unsigned seq) {
        // all (up to eight) sums in one butterfly: every step halves the values a lane carries, 9 shuffles instead
        // of 5 per value; afterwards lane l holds the warp total of value (l >> 2) (bit-reversed over lane bits 4, 3, 2)
        {
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            double a4[4], a2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double lo = v[i], hi = i + 4 < kRcmSums ? v[i + 4] : 0.0;
                a4[i] = (h16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, h16 ? lo : hi, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) a2[i] = (h8 ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, h8 ? a4[i] : a4[i + 2], 8);
            double a1 = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if ((lane & 3) == 0 && idx < kRcmSums) s_part[warp][idx] = a1;
        }
        row_sync();
        if (tid < nv) {
            double t = 0;
            for (int w = 0; w < row_warps; ++w) t += s_part[w][tid];
            ll_store(&(A.slots + (int)(seq & 1u) * kRcmMaxCtas)[blockIdx.x].v[tid], t, seq);
        }
    };
    // Collect: poll every CTA's partial sums and the lines of sequence number `seq` on the CTA's halo (gathered into
    // `gather`).  All threads share the polling: item i < G nv is a partial sum, the rest are halo lines; eight
    // polls in flight per thread.  Returns the grid totals.
    auto collect = [&](int nv, unsigned seq, double* gather, double (&tot)[kRcmSums]) -> bool {
        const int par = (int)(seq & 1u);
        const RcmSlot* slots = A.slots + par * kRcmMaxCtas;
        const LLLine* zl = A.z + (int64_t)par * 6 * A.n_cams;
        const int n_sum = G * nv, n_items = n_sum + nh * 6;
        const long long t_start = clock64();
        constexpr int kFly = 8;
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
        lap(3);
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
    // Pinv_k v for camera k of this CTA, v spread over the six lanes of the group `slot`; valid where slot < 5
    auto precond = [&](int k, double v_a) {
        double z = 0;
        const double* pin = pinv_s + k * 21;
        const int base = slot < 5 ? slot * 6 : 24;
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) {
            const double vb = __shfl_sync(0xffffffffu, v_a, base + bb);
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
        if (warp < row_warps) {
            for (int k0 = 0; k0 < ncam; k0 += 5 * row_warps) {
                const int k = k0 + warp * 5 + slot;
                const bool act = slot < 5 && k < ncam;
                const int kk = act ? k : 0;
                const double r_a = act ? A.b[(c0 + kk) * 6 + a] : 0.0;
                const double z = precond(kk, r_a);
                if (act) {
                    vec[k * 18 + a] = 0.0;
                    vec[k * 18 + 6 + a] = r_a;
                    ll_store(zl + (c0 + k) * 6 + a, z, seq);
                    v[0] += r_a * z;
                    v[1] += r_a * r_a;
                }
            }
            publish(v, 2, seq);
        }
        if (!collect(2, seq, zh, tot)) return;
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
#ifdef MMBA_PHASE_TIMING
        t_last = clock64();
#endif
        for (int it = 0; it < A.maxit; ++it) {
            const unsigned seq = A.seq0 + (unsigned)it + 2u;
            lap(5);
            // q = S p, step 1: every (block, row) pair is one unit of six multiply-adds, dealt to all threads
            const int n_units = nblk * 6;
            for (int u0 = tid; u0 < n_units; u0 += 2 * blockDim.x) {
                // two units per trip: their loads are independent and overlap
                const int u1 = u0 + blockDim.x;
                const bool two = u1 < n_units;
                const int ea = u0 / 6, eb = two ? u1 / 6 : ea;
                const int rs = A.s_in_smem ? kRcmRowStride : 6;
                const double* sa = S_rows + (int64_t)u0 * rs;
                const double* sb = S_rows + (int64_t)(two ? u1 : u0) * rs;
                const double* pa = ph + lcol_s[ea] * 6;
                const double* pb = ph + lcol_s[eb] * 6;
                double a0, a1, b0, b1;
                if (A.s_in_smem) {
                    a0 = sa[0] * pa[0];
                    b0 = sb[0] * pb[0];
                    a1 = sa[1] * pa[1];
                    b1 = sb[1] * pb[1];
                    a0 = fma(sa[2], pa[2], a0);
                    b0 = fma(sb[2], pb[2], b0);
                    a1 = fma(sa[3], pa[3], a1);
                    b1 = fma(sb[3], pb[3], b1);
                    a0 = fma(sa[4], pa[4], a0);
                    b0 = fma(sb[4], pb[4], b0);
                    a1 = fma(sa[5], pa[5], a1);
                    b1 = fma(sb[5], pb[5], b1);
                } else {
                    a0 = __ldg(sa + 0) * pa[0];
                    b0 = __ldg(sb + 0) * pb[0];
                    a1 = __ldg(sa + 1) * pa[1];
                    b1 = __ldg(sb + 1) * pb[1];
                    a0 = fma(__ldg(sa + 2), pa[2], a0);
                    b0 = fma(__ldg(sb + 2), pb[2], b0);
                    a1 = fma(__ldg(sa + 3), pa[3], a1);
                    b1 = fma(__ldg(sb + 3), pb[3], b1);
                    a0 = fma(__ldg(sa + 4), pa[4], a0);
                    b0 = fma(__ldg(sb + 4), pb[4], b0);
                    a1 = fma(__ldg(sa + 5), pa[5], a1);
                    b1 = fma(__ldg(sb + 5), pb[5], b1);
                }
                prod[u0] = a0 + a1;
                if (two) prod[u1] = b0 + b1;
            }
            lap(0);
            __syncthreads();
            lap(1);
            // step 2 (row warps): q = sums of the row's products, w = Pinv q (published), the seven partial sums
            if (warp < row_warps) {
#pragma unroll
                for (int i = 0; i < kRcmSums; ++i) v[i] = 0.0;
                LLLine* wl = A.z + (int64_t)(seq & 1u) * 6 * A.n_cams;
                for (int k0 = 0; k0 < ncam; k0 += 5 * row_warps) {
                    const int k = k0 + warp * 5 + slot;
                    const bool act = slot < 5 && k < ncam;
                    const int kk = act ? k : 0;
                    double q = 0, pk = 0, zk = 0, rk = 0;
                    if (act) {
                        double q0 = 0, q1 = 0, q2 = 0, q3 = 0;      // independent partial sums: the loads overlap
                        const int e1 = rowptr_s[k + 1];
                        int e = rowptr_s[k];
                        for (; e + 3 < e1; e += 4) {
                            q0 += prod[e * 6 + a];
                            q1 += prod[e * 6 + 6 + a];
                            q2 += prod[e * 6 + 12 + a];
                            q3 += prod[e * 6 + 18 + a];
                        }
                        for (; e < e1; ++e) q0 += prod[e * 6 + a];
                        q = (q0 + q1) + (q2 + q3);
                        vec[k * 18 + 12 + a] = q;
                        pk = ph[(own0 + k) * 6 + a];
                        zk = zh[(own0 + k) * 6 + a];
                        rk = vec[k * 18 + 6 + a];
                    }
                    const double w = precond(kk, q);
                    if (act) {
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
                publish(v, kRcmSums, seq);
            }
            lap(2);
            if (!collect(kRcmSums, seq, wh, tot)) return;
            const double pq = tot[0], rho = tot[5];
            const double inv_rho = __drcp_rn(rho), alpha = rho * __drcp_rn(pq);      // independent reciprocals (no division sequence)
            const double rho_n = fma(alpha, fma(alpha, tot[2], -2.0 * tot[1]), rho);
            double rr_n = fma(alpha, fma(alpha, tot[4], -2.0 * tot[3]), tot[6]);
            if (!(pq > 0.0) || !(rho > 0.0) || !isfinite(alpha) || !isfinite(rho_n) || !isfinite(rr_n)) {
                done = 2;      // breakdown: keep the last finite iterate
                break;
            }
            if (rr_n < 0.0) rr_n = 0.0;
            const double beta = rho_n > 0.0 ? rho_n * inv_rho : 0.0;
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
            inv_nu2 += rr_n > 0.0 ? __drcp_rn(rr_n) : INFINITY;
            if (rr_n <= A.rtol2 * b2 || rr_n <= A.atol2f || 1.0 <= inv_nu2 * A.ktol2f * (double)its) {
                done = 1;
                break;
            }
            lap(4);
            __syncthreads();
        }
#ifdef MMBA_PHASE_TIMING
        if (A.phase && blockIdx.x == 0 && tid == 0)
            for (int k = 0; k < 8; ++k) A.phase[k] += phc[k];
#endif
    }
    __syncthreads();
    for (int i = tid; i < ncam * 6; i += blockDim.x) A.x[c0 * 6 + i] = vec[(i / 6) * 18 + (i % 6)];
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
