// Small-vector kernels of the TRF / PCG driver (n-vectors and 6*Nc camera vectors).
//
// Replaces numpy vector arithmetic inside scipy's trf_no_bounds (trf.py:433-561), compute_grad /
// compute_jac_scale (common.py:590-610) and LSMR's vector updates (lsmr.py:373-377).
//
// Camera-part reductions of the PCG are *deterministic* (fixed-order sums inside one CTA): with
// cameras replicated across ranks, every rank then takes bit-identical PCG decisions without
// exchanging flags.
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"
#include "rcm.cuh"

namespace mmba {

namespace cg = cooperative_groups;

__device__ __forceinline__ double block_sum_det(double v, double* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    return t;
}

// fixed-order sum of nb per-block partials; every thread of the block gets the same value
__device__ __forceinline__ double sum_partials(const double* __restrict__ part, int nb, double* s_red) {
    double v = 0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) v += part[i];
    return block_sum_det(v, s_red);
}

// ---------------------------------------------------------------------------------------------
// scale update + gradient statistics (compute_jac_scale / compute_grad, common.py:590-610;
// Delta0 = ||x0 * scale_inv||, trf.py:443)
//   PACKED: element e of block b = e / BS is the diagonal of a packed upper-triangle block (points: V);
//   otherwise blk already holds the column sums of squares (cameras: diag(J_c^T J_c)); g_h = g / scale_inv
// ---------------------------------------------------------------------------------------------
// Grid-stride (a few CTAs per SM): every CTA reduces its partial sums once and issues ONE atomic per result — a thread
// block per 256 elements meant ~100 k atomics on four addresses at C4 (the max alone: one per warp), 94 us per launch
// for 170 MB of traffic.
template <int BS, bool PACKED>
__global__ void scale_grad_kernel(const double* __restrict__ blk, const double* __restrict__ g,
                                  const double* __restrict__ x, double* __restrict__ sinv, double* __restrict__ gh,
                                  double* __restrict__ u1, int first, int64_t n_elem, double* __restrict__ scal,
                                  int accumulate) {
    __shared__ double s_red[64];
    __shared__ unsigned long long s_max;
    constexpr int STRIDE = BS * (BS + 1) / 2;
    double acc[3] = {0, 0, 0};
    double gabs = 0;
    if (threadIdx.x == 0) s_max = 0ull;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / BS;
        const int k = (int)(e - b * BS);
        const double diag = PACKED ? blk[b * STRIDE + (k * BS - k * (k - 1) / 2)] : blk[e];
        double si = sqrt(diag);
        if (first) { if (si == 0.0) si = 1.0; }
        else si = fmax(si, sinv[e]);
        sinv[e] = si;
        const double gv = g[e], xv = x[e];
        const double h = gv / si;
        gh[e] = h;
        if (u1) u1[e] = h / si;      // u1 = d o g_h: the unscaled gradient direction (Cauchy step, subspace basis)
        gabs = fmax(gabs, fabs(gv));
        acc[0] += h * h;
        acc[1] += (xv * si) * (xv * si);
        acc[2] += xv * xv;
    }
    if (!accumulate) return;
    double* outp[3] = {scal + S_GH2, scal + S_XSI2, scal + S_X2};
    block_accumulate<3>(acc, s_red, outp);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gabs = fmax(gabs, __shfl_xor_sync(kFull, gabs, off));
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, (unsigned long long)__double_as_longlong(gabs));
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(scal + S_GINF), s_max);
}

// out = coef * a / sinv   (unscaled image of a scaled vector: d o a)
__global__ void unscale_kernel(const double* __restrict__ a, const double* __restrict__ sinv, double coef,
                               double* __restrict__ out, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = coef * a[e] / sinv[e];
}

// ---------------------------------------------------------------------------------------------
// 2-D subspace construction  S = orth[g_h, gn_h]  (trf.py:496-500; scipy uses LAPACK QR; here the orthonormal
// basis is expressed through dot products and one Gram matrix, see subspace_dots_kernel; the subspace and therefore
// the step are the same)
// ---------------------------------------------------------------------------------------------
// gn = src * (sinv if MUL_SINV) ; D0 += gh.gn ; D1 += gn.gn
template <bool MUL_SINV>
__global__ void gn_assemble_kernel(const double* __restrict__ src, const double* __restrict__ sinv,
                                   const double* __restrict__ gh, double* __restrict__ gn, int64_t n,
                                   double* __restrict__ scal, int accumulate) {
    __shared__ double s_red[64];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[2] = {0, 0};
    if (e < n) {
        const double v = MUL_SINV ? src[e] * sinv[e] : src[e];
        gn[e] = v;
        acc[0] = gh[e] * v;
        acc[1] = v * v;
    }
    if (!accumulate) return;
    double* outp[2] = {scal + S_DOT0, scal + S_DOT1};
    block_accumulate<2>(acc, s_red, outp);
}

// One pass for everything the 2-D subspace needs from the vectors (trf.py:496-500 builds S = orth[g_h, gn_h] with a QR
// and multiplies J_h S; here the basis is never materialised: with u1 = d o g_h and u2 = d o gn_h (the unscaled images,
// u2 = [pxt | dp] as the back-substitution left it) every quantity of the subspace problem is a combination of
//   D0 = g_h.gn_h, D1 = gn_h.gn_h  (scaled),  U11 = u1.u1, U12 = u1.u2, U22 = u2.u2  (unscaled: step norm)
// and of the Gram matrix of J [u1 u2] (one J pass, M_JV2).  Also writes u2 contiguously (camera part | local points).
//   cam part (e < ncam): gn_h = px, u2 = pxt;  point part: u2 = dp, gn_h = u2 * sinv
__global__ void subspace_dots_kernel(const double* __restrict__ gh, const double* __restrict__ u1, const double* __restrict__ sinv,
                                     const double* __restrict__ px, const double* __restrict__ pxt, const double* __restrict__ dp,
                                     double* __restrict__ u2, int64_t ncam, int64_t n, double* __restrict__ scal, int lead) {
    __shared__ double s_red[64];
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        double gnh, w2;
        const bool cam = e < ncam;
        if (cam) {
            gnh = px[e];
            w2 = pxt[e];
        } else {
            w2 = dp[e - ncam];
            gnh = w2 * sinv[e];
        }
        u2[e] = w2;
        if (!cam || lead) {       // the replicated camera part is counted by one rank only
            const double g = gh[e], w1 = u1[e];
            acc[0] += g * gnh;
            acc[1] += gnh * gnh;
            acc[2] += w1 * w1;
            acc[3] += w1 * w2;
            acc[4] += w2 * w2;
        }
    }
    double* outp[5] = {scal + S_DOT0, scal + S_DOT1, scal + S_DOT2, scal + S_DOT3, scal + S_DOT4};
    block_accumulate<5>(acc, s_red, outp);
}

// trial point x_new = x + p0 * v1 + p1 * v2   (x + d o step_h, trf.py:512-513)
__global__ void trial_kernel(const double* __restrict__ x, const double* __restrict__ v1, const double* __restrict__ v2,
                             double p0, double p1, double* __restrict__ xn, int64_t n) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // a zero coefficient switches its vector off entirely (the Gauss-Newton direction may be non-finite after a PCG
    // breakdown: the step then is the 1-D Cauchy step along v1)
    if (e < n) xn[e] = x[e] + p0 * v1[e] + (p1 != 0.0 ? p1 * v2[e] : 0.0);
}

// ---------------------------------------------------------------------------------------------
// PCG on the reduced camera system, one thread per camera (block-Jacobi = 6x6 Schur diagonal)
// ---------------------------------------------------------------------------------------------
// b = d o (g_c - y) ; Pinv = (d d^T o S_cc + reg I)^-1 ; x = 0, r = b, z = Pinv r, p = z, xt = d o p
__global__ void __launch_bounds__(kCamBlock) pcg_init_kernel(PcgVecs P, double reg) {
    __shared__ double s_red[8];
    const int c = blockIdx.x * kCamBlock + threadIdx.x;
    double rho = 0, b2 = 0;
    if (c < P.n_cams) {
        double d[6], b[6], z[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            d[k] = 1.0 / P.sinv[c * 6 + k];
            b[k] = d[k] * (P.gc[c * 6 + k] - P.y[c * 6 + k]);
            P.y[c * 6 + k] = 0.0;
        }
        double s[21], inv[21];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int bb = a; bb < 6; ++bb) {
                const int i = tri6(a, bb);
                s[i] = d[a] * d[bb] * P.Sd[c * 21 + i] + (a == bb ? reg : 0.0);
            }
        sym6_inverse(s, inv);
#pragma unroll
        for (int i = 0; i < 21; ++i) P.Pinv[c * 21 + i] = inv[i];
        sym6_matvec(inv, b, z);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            P.x[c * 6 + k] = 0.0;
            P.r[c * 6 + k] = b[k];
            P.z[c * 6 + k] = z[k];
            P.p[c * 6 + k] = z[k];
            P.xt[c * 6 + k] = d[k] * z[k];
            rho += b[k] * z[k];
            b2 += b[k] * b[k];
        }
    }
    rho = block_sum_det(rho, s_red);
    b2 = block_sum_det(b2, s_red);
    if (threadIdx.x == 0) {
        P.part[P_RHO0 * kMaxCamBlocks + blockIdx.x] = rho;
        P.part[P_B2 * kMaxCamBlocks + blockIdx.x] = b2;
        if (blockIdx.x == 0) { P.flags[0] = 0; P.flags[1] = 0; }
    }
}

constexpr int kPcgThreads = 256;
constexpr int kPcgCluster = 8;     // CTAs of the update kernel: one thread-block cluster (portable size)

// Deterministic sum over the whole cluster: fixed-order block sums, exchanged through distributed
// shared memory and added in cluster-rank order by every CTA (all CTAs, and all ranks, get the same bits).
template <int NV>
__device__ __forceinline__ void cluster_sum_det(cg::cluster_group& cluster, double (&v)[NV], double* s_red, double* s_xch /* [NV] */) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = block_sum_det(v[i], s_red);
    const unsigned nb = cluster.num_blocks();
    if (nb == 1) return;                       // single CTA: the block sum is the result
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) s_xch[i] = v[i];
    }
    cluster.sync();
    // lane r of warp 0 fetches CTA r's partial (one remote latency for all of them); the xor tree adds
    // them in the same order on every CTA
    if (threadIdx.x < 32) {
        const double* rem = cluster.map_shared_rank(s_xch, threadIdx.x < nb ? threadIdx.x : 0);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = threadIdx.x < nb ? rem[i] : 0.0;
            x = warp_sum(x);
            if (threadIdx.x == 0) s_red[i] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = s_red[i];
    __syncthreads();
}

// One PCG iteration's camera-vector work in ONE launch of one thread-block cluster, so that every
// reduction is a fixed-order sum (LSMR's vector updates, lsmr.py:373-377):
//   q = S p = d o y + reg p ; alpha = rho / p.q ; x += alpha p ; r -= alpha q ; z = Pinv r ;
//   stop if ||r|| <= rtol ||b|| or ||r|| <= atol ||f|| (atol2f) ; beta = r.z / rho ; p = z + beta p ; xt = d o p ; y <- 0
// y holds this iteration's sum_i Jc_i^T (Jc_i xt - Jp_i z_p) from the MATVEC pass (all-reduced over ranks).
// Launched as ONE thread-block cluster: 1 CTA for <= 256 cameras, else kPcgCluster CTAs (runtime cluster
// dimension).  Up to cluster*kPcgThreads cameras: one camera per thread, every operand loaded up front (one
// memory round trip); more cameras: a strided loop with q and z kept in global memory.
__global__ void __launch_bounds__(kPcgThreads)
pcg_update_kernel(PcgVecs P, double reg, int it, double rtol2, double atol2f, double ktol2f, int nb_init, int parity, unsigned long long seq) {
    // Programmatic dependent launch (see launch_tile): this grid may be scheduled while the MATVEC pass still
    // drains.  Everything written by the PREVIOUS update (flags, p, r, x, state) or earlier (sinv, Pinv) is
    // complete by then and may be read at once; y and every store wait for griddepcontrol.wait below.
    cg::cluster_group cluster = cg::this_cluster();
    if (P.flags[0]) return;   // uniform over the cluster: flags are only written by the previous update
    __shared__ double s_red[32];
    __shared__ double s_xa[2], s_xb[2];
    const int tid = threadIdx.x;
    const int gtid = (int)cluster.block_rank() * kPcgThreads + tid;
    const int nthr = (int)cluster.num_blocks() * kPcgThreads;
    const bool lead = gtid == 0;
    double rho, b2;
    int done = 0;
    const bool xchg = P.nranks > 1;
    const double* slots = nullptr;
    auto exchange = [&]() {
        // One-shot all-reduce of the 6*Nc Schur product over NVLink peer memory, fused into this kernel
        // (replaces a per-iteration ncclAllReduce: 30+ us at 8 GPUs for this 10-100 KB message).
        // Push: every rank stores its partial y into slot [parity][rank] of EVERY rank's receive buffer
        // (plain stores to peer-mapped addresses) and clears y; after a cluster barrier one release
        // store per peer raises this rank's arrival flag there.
        const int n = 6 * P.n_cams;
        const int64_t off = (int64_t)(parity * P.nranks + P.rank) * P.n6;
        for (int i = gtid; i < n; i += nthr) {
            const double v = P.y[i];
            P.y[i] = 0.0;
            for (int r = 0; r < P.nranks; ++r) P.peer_slots[r][off + i] = v;
        }
        cluster.sync();
        if (gtid < P.nranks) st_release_sys(P.peer_flags[gtid] + parity * P.nranks + P.rank, seq);
        // Wait: all ranks' partials have landed in this rank's slots; they are then added in rank order,
        // so every rank obtains bit-identical sums and takes identical PCG decisions.
        if (tid < P.nranks) {
            const unsigned long long* f = P.xflags + parity * P.nranks + tid;
            const long long t0 = clock64();
            while (ld_acquire_sys(f) < seq) {
                if (clock64() - t0 > (1ll << 32)) {   // ~2 s: a peer died; fail instead of hanging
                    P.flags[2] = 1;
                    break;
                }
            }
        }
        __syncthreads();
        slots = P.xslots + (int64_t)parity * P.nranks * P.n6;
    };
    if (P.n_cams <= nthr) {
        const int c = gtid;
        const bool live = c < P.n_cams;
        // register-lean: the 6x6 preconditioner block is prefetched into L1 and read when needed
        double pv[6], yv[6], si[6], rv[6], xv[6];
        if (live) {
            const char* pin = reinterpret_cast<const char*>(P.Pinv + c * 21);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pin));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pin + 64));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pin + 128));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pin + 160));
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                pv[k] = P.p[c * 6 + k];
                si[k] = P.sinv[c * 6 + k];
                rv[k] = P.r[c * 6 + k];
                xv[k] = P.x[c * 6 + k];
            }
        }
        if (it == 0) {
            rho = sum_partials(P.part + P_RHO0 * kMaxCamBlocks, nb_init, s_red);
            b2 = sum_partials(P.part + P_B2 * kMaxCamBlocks, nb_init, s_red);
        } else {
            rho = P.state[0];
            b2 = P.state[1];
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");   // the MATVEC pass (and its memory) is complete
        if (xchg) exchange();
        if (live) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                if (xchg) {
                    double acc = 0.0;
                    for (int r = 0; r < P.nranks; ++r) acc += __ldcg(slots + (int64_t)r * P.n6 + c * 6 + k);
                    yv[k] = acc;
                } else {
                    yv[k] = __ldcg(P.y + c * 6 + k);
                }
            }
        }
        double q[6], z[6];
        double s1[1] = {0};
        if (live) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                q[k] = yv[k] / si[k] + reg * pv[k];
                if (!xchg) P.y[c * 6 + k] = 0.0;
                s1[0] += pv[k] * q[k];
            }
        }
        cluster_sum_det<1>(cluster, s1, s_red, s_xa);
        const double pq = s1[0];
        const double alpha = rho / pq;
        double s2[2] = {0, 0};
        if (live) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                xv[k] += alpha * pv[k];
                rv[k] -= alpha * q[k];
            }
            sym6_matvec(P.Pinv + c * 21, rv, z);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                P.x[c * 6 + k] = xv[k];
                P.r[c * 6 + k] = rv[k];
                s2[0] += rv[k] * rv[k];
                s2[1] += rv[k] * z[k];
            }
        }
        cluster_sum_det<2>(cluster, s2, s_red, s_xb);
        const double rr = s2[0], rz = s2[1];
        // smoothed (minimal-residual) norm of the CG process, see rcm_pcg_kernel: 1 / nu^2 = sum 1 / ||r_j||^2
        const double inv_nu2 = (it == 0 ? 1.0 / b2 : P.state[3]) + (rr > 0.0 ? 1.0 / rr : INFINITY);
        if (rr <= rtol2 * b2 || rr <= atol2f || 1.0 <= inv_nu2 * ktol2f * (double)(it + 1)) done = 1;
        else if (!(pq > 0.0) || !isfinite(rr) || !(rz > 0.0)) done = 2;
        if (lead) {
            P.state[0] = rz;
            P.state[1] = b2;
            P.state[2] = rr;
            P.state[3] = inv_nu2;
            if (done) {
                P.flags[1] = it + 1;
                __threadfence();
                P.flags[0] = done;
            }
        }
        if (!done && live) {
            const double beta = rz / rho;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double pk = z[k] + beta * pv[k];
                P.p[c * 6 + k] = pk;
                P.xt[c * 6 + k] = pk / si[k];
            }
        }
    } else {
        if (it == 0) {
            rho = sum_partials(P.part + P_RHO0 * kMaxCamBlocks, nb_init, s_red);
            b2 = sum_partials(P.part + P_B2 * kMaxCamBlocks, nb_init, s_red);
        } else {
            rho = P.state[0];
            b2 = P.state[1];
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (xchg) exchange();
        double s1[1] = {0};
        for (int c = gtid; c < P.n_cams; c += nthr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double pk = P.p[c * 6 + k];
                double yk;
                if (xchg) {
                    yk = 0.0;
                    for (int r = 0; r < P.nranks; ++r) yk += __ldcg(slots + (int64_t)r * P.n6 + c * 6 + k);
                } else {
                    yk = P.y[c * 6 + k];
                    P.y[c * 6 + k] = 0.0;
                }
                const double qk = yk / P.sinv[c * 6 + k] + reg * pk;
                P.q[c * 6 + k] = qk;
                s1[0] += pk * qk;
            }
        }
        cluster_sum_det<1>(cluster, s1, s_red, s_xa);
        const double pq = s1[0];
        const double alpha = rho / pq;
        double s2[2] = {0, 0};
        for (int c = gtid; c < P.n_cams; c += nthr) {
            double r[6], z[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                P.x[c * 6 + k] += alpha * P.p[c * 6 + k];
                r[k] = P.r[c * 6 + k] - alpha * P.q[c * 6 + k];
                P.r[c * 6 + k] = r[k];
            }
            sym6_matvec(P.Pinv + c * 21, r, z);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                P.z[c * 6 + k] = z[k];
                s2[0] += r[k] * r[k];
                s2[1] += r[k] * z[k];
            }
        }
        cluster_sum_det<2>(cluster, s2, s_red, s_xb);
        const double rr = s2[0], rz = s2[1];
        // smoothed (minimal-residual) norm of the CG process, see rcm_pcg_kernel: 1 / nu^2 = sum 1 / ||r_j||^2
        const double inv_nu2 = (it == 0 ? 1.0 / b2 : P.state[3]) + (rr > 0.0 ? 1.0 / rr : INFINITY);
        if (rr <= rtol2 * b2 || rr <= atol2f || 1.0 <= inv_nu2 * ktol2f * (double)(it + 1)) done = 1;
        else if (!(pq > 0.0) || !isfinite(rr) || !(rz > 0.0)) done = 2;
        if (lead) {
            P.state[0] = rz;
            P.state[1] = b2;
            P.state[2] = rr;
            P.state[3] = inv_nu2;
            if (done) {
                P.flags[1] = it + 1;
                __threadfence();
                P.flags[0] = done;
            }
        }
        if (!done) {
            const double beta = rz / rho;
            for (int c = gtid; c < P.n_cams; c += nthr) {
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double pk = P.z[c * 6 + k] + beta * P.p[c * 6 + k];
                    P.p[c * 6 + k] = pk;
                    P.xt[c * 6 + k] = pk / P.sinv[c * 6 + k];
                }
            }
        }
    }
    cluster.sync();   // no CTA leaves while its exchange slots may still be read remotely
}

// ---------------------------------------------------------------------------------------------
// Pose-only adjustment (adjustPose, bundleAdjuster.py:214-243: points frozen, dense least_squares
// with tr_solver='exact', x_scale=1).  With the points fixed J is block diagonal (one 2k x 6 block per
// camera), so the SVD scipy takes of the whole J (trf.py:480-484) factorises per camera:
// J_c^T J_c = V_c diag(lambda) V_c^T, s = sqrt(lambda), s*uf = V_c^T g_c.
// ---------------------------------------------------------------------------------------------

// per camera: cyclic Jacobi eigen-decomposition of the 6x6 block; lam[6], suf[6] = V^T g, V[36] (columns)
__global__ void __launch_bounds__(128) pose_eig_kernel(const double* __restrict__ U, const double* __restrict__ g,
                                                       double* __restrict__ lam, double* __restrict__ suf,
                                                       double* __restrict__ Vout, int n_cams) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cams) return;
    double A[6][6], V[6][6];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            A[a][b] = U[c * 21 + (a <= b ? tri6(a, b) : tri6(b, a))];
            V[a][b] = a == b ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0, diag = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            diag += A[a][a] * A[a][a];
#pragma unroll
            for (int b = a + 1; b < 6; ++b) off += A[a][b] * A[a][b];
        }
        if (off <= 1e-34 * diag) break;
#pragma unroll
        for (int p = 0; p < 5; ++p)
#pragma unroll
            for (int q = p + 1; q < 6; ++q) {
                const double apq = A[p][q];
                if (apq != 0.0) {
                    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const double akp = A[k][p], akq = A[k][q];
                        A[k][p] = cs * akp - sn * akq;
                        A[k][q] = sn * akp + cs * akq;
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const double apk = A[p][k], aqk = A[q][k];
                        A[p][k] = cs * apk - sn * aqk;
                        A[q][k] = sn * apk + cs * aqk;
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = cs * vkp - sn * vkq;
                        V[k][q] = sn * vkp + cs * vkq;
                    }
                }
            }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        lam[c * 6 + k] = fmax(A[k][k], 0.0);
        double sv = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            sv += V[j][k] * g[c * 6 + j];
            Vout[c * 36 + j * 6 + k] = V[j][k];
        }
        suf[c * 6 + k] = sv;
    }
}

// results of pose_tr_kernel in the scalar array
enum PoseScal { PS_ALPHA = S_DOT0, PS_NITER = S_DOT1, PS_PRED = S_DOT2, PS_STEPNORM = S_DOT3, PS_GNORM = S_DOT4 };

// One CTA: scipy's solve_lsq_trust_region (common.py:106-164) on the n = 6*Nc (lambda, suf) pairs:
// Gauss-Newton step if it is inside the radius, else More's root-finding for alpha with
// ||p(alpha)|| = Delta, p = -V (suf / (lambda + alpha)), rescaled onto the boundary.  Writes the step
// coefficients w in the eigenbasis, alpha, the iteration count, the predicted reduction
// -(0.5 p^T H p + g^T p) and ||p||.
__global__ void __launch_bounds__(256) pose_tr_kernel(const double* __restrict__ lam, const double* __restrict__ suf,
                                                      double* __restrict__ w, int n, double m_rows, double Delta,
                                                      double alpha_in, double* __restrict__ scal) {
    __shared__ double s_red[32];
    const int tid = threadIdx.x;
    auto bsum = [&](double v) { return block_sum_det(v, s_red); };
    auto bmax = [&](double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, off));
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = v;
        __syncthreads();
        double t = s_red[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, s_red[i]);
        return t;
    };
    double lmax = 0, lmin_neg = -1e300, suf2 = 0, gn2 = 0, d0 = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double l = lam[i], sf = suf[i];
        lmax = fmax(lmax, l);
        lmin_neg = fmax(lmin_neg, -l);
        suf2 += sf * sf;
        if (l > 0) {
            gn2 += sf * sf / (l * l);
            d0 += sf * sf / (l * l * l);
        }
    }
    lmax = bmax(lmax);
    const double lmin = -bmax(lmin_neg);
    suf2 = bsum(suf2);
    gn2 = bsum(gn2);
    d0 = bsum(d0);
    const double EPS = 2.220446049250313e-16;
    const bool full_rank = m_rows >= n && sqrt(lmin) > EPS * m_rows * sqrt(lmax);
    double alpha = 0.0, scale = 1.0;
    int n_iter = 0;
    if (!(full_rank && sqrt(gn2) <= Delta)) {
        double alpha_upper = sqrt(suf2) / Delta;
        double alpha_lower = 0.0;
        if (full_rank) {
            const double pn = sqrt(gn2);
            alpha_lower = -(pn - Delta) / (-d0 / pn);
        }
        if (!full_rank && alpha_in == 0.0) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
        else alpha = alpha_in;
        for (int it = 0; it < 10; ++it) {
            if (alpha < alpha_lower || alpha > alpha_upper) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
            double a2 = 0, a3 = 0;
            for (int i = tid; i < n; i += blockDim.x) {
                const double dn = lam[i] + alpha, sf = suf[i];
                a2 += sf * sf / (dn * dn);
                a3 += sf * sf / (dn * dn * dn);
            }
            a2 = bsum(a2);
            a3 = bsum(a3);
            const double pn = sqrt(a2);
            const double phi = pn - Delta, phi_prime = -a3 / pn;
            if (phi < 0) alpha_upper = alpha;
            const double ratio = phi / phi_prime;
            alpha_lower = fmax(alpha_lower, alpha - ratio);
            alpha -= (phi + Delta) * ratio / Delta;
            n_iter = it + 1;
            if (fabs(phi) < 0.01 * Delta) break;
        }
        double pn2 = 0;
        for (int i = tid; i < n; i += blockDim.x) {
            const double v = suf[i] / (lam[i] + alpha);
            pn2 += v * v;
        }
        pn2 = bsum(pn2);
        scale = Delta / sqrt(pn2);
    }
    double pred = 0, wn2 = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double l = lam[i], sf = suf[i];
        const double wi = -scale * sf / (l + alpha);
        w[i] = wi;
        pred += 0.5 * l * wi * wi + sf * wi;
        wn2 += wi * wi;
    }
    pred = bsum(pred);
    wn2 = bsum(wn2);
    if (tid == 0) {
        scal[PS_ALPHA] = alpha;
        scal[PS_NITER] = (double)n_iter;
        scal[PS_PRED] = -pred;
        scal[PS_STEPNORM] = sqrt(wn2);
    }
}

// x_new (cameras) = x + V w ; the frozen points are copied
__global__ void pose_step_kernel(const double* __restrict__ x, const double* __restrict__ V, const double* __restrict__ w,
                                 double* __restrict__ xn, int n_cams, int64_t n_total) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_total) return;
    if (e >= 6 * (int64_t)n_cams) {
        xn[e] = x[e];
        return;
    }
    const int c = (int)(e / 6), j = (int)(e - 6 * c);
    double p = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) p += V[c * 36 + j * 6 + k] * w[c * 6 + k];
    xn[e] = x[e] + p;
}

// ---------------------------------------------------------------------------------------------
// Batched two-view triangulation: the per-track cv2.triangulatePoints call of
// processor.triangulatePoints (processor.py:246-261).  DLT: A = [u1 P1[2]-P1[0]; v1 P1[2]-P1[1];
// u2 P2[2]-P2[0]; v2 P2[2]-P2[1]] (4x4), X = right singular vector of the smallest singular value,
// dehomogenised.  One thread per track; one-sided (Hestenes) Jacobi SVD in registers, which keeps the
// accuracy of an SVD of A itself (no A^T A).  48 B in (2 x uv + 2 x frame index), 24 B out per track;
// the projection matrices (96 B per frame) stay in L1/L2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) triangulate_kernel(const double* __restrict__ proj, const int64_t* __restrict__ f1,
                                                          const int64_t* __restrict__ f2, const double* __restrict__ uv1,
                                                          const double* __restrict__ uv2, double* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* P1 = proj + f1[i] * 12;
    const double* P2 = proj + f2[i] * 12;
    const double u1 = uv1[2 * i], v1 = uv1[2 * i + 1], u2 = uv2[2 * i], v2 = uv2[2 * i + 1];
    double A[4][4], V[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double p10 = __ldg(P1 + k), p11 = __ldg(P1 + 4 + k), p12 = __ldg(P1 + 8 + k);
        const double p20 = __ldg(P2 + k), p21 = __ldg(P2 + 4 + k), p22 = __ldg(P2 + 8 + k);
        A[0][k] = u1 * p12 - p10;
        A[1][k] = v1 * p12 - p11;
        A[2][k] = u2 * p22 - p20;
        A[3][k] = v2 * p22 - p21;
#pragma unroll
        for (int j = 0; j < 4; ++j) V[j][k] = j == k ? 1.0 : 0.0;
    }
    // rotate column pairs until all columns of A V are mutually orthogonal
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double app = 0, aqq = 0, apq = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    app += A[r][p] * A[r][p];
                    aqq += A[r][q] * A[r][q];
                    apq += A[r][p] * A[r][q];
                }
                if (fabs(apq) > 1e-17 * sqrt(app * aqq) && apq != 0.0) {
                    off = fmax(off, fabs(apq) / sqrt(app * aqq));
                    const double zeta = (aqq - app) / (2.0 * apq);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const double ap = A[r][p], aq = A[r][q];
                        A[r][p] = cs * ap - sn * aq;
                        A[r][q] = sn * ap + cs * aq;
                        const double vp = V[r][p], vq = V[r][q];
                        V[r][p] = cs * vp - sn * vq;
                        V[r][q] = sn * vp + cs * vq;
                    }
                }
            }
        if (off < 1e-15) break;
    }
    int best = 0;
    double smin = 1e300;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double nk = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) nk += A[r][k] * A[r][k];
        if (nk < smin) {
            smin = nk;
            best = k;
        }
    }
    double x[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) x[r] = best == 0 ? V[r][0] : best == 1 ? V[r][1] : best == 2 ? V[r][2] : V[r][3];
    const double iw = 1.0 / x[3];
    out[3 * i] = x[0] * iw;
    out[3 * i + 1] = x[1] * iw;
    out[3 * i + 2] = x[2] * iw;
}

// ---------------------------------------------------------------------------------------------
// All-reduce of a handful of scalars across the ranks of a sharded solve (the sums the TRF logic reads back:
// ||J v||^2, dot products of the 2-D subspace, trial cost, ||g||_inf) as ONE tiny kernel over NVLink peer memory,
// instead of a ~30 us ncclAllReduce each.  Every value travels as a self-validating 16-byte line (ll_store /
// ll_try_load with system scope): rank r stores its values into row [parity][r] of EVERY rank's area (plain
// stores to peer-mapped addresses), then polls the rows of its own area and combines them in rank order, so all
// ranks obtain bit-identical results.  Two parities make reuse safe: a rank can write exchange s + 2 only after it
// has completed exchange s + 1, which every peer joins only after it has finished reading exchange s.
// Up to two segments (p0[n0], p1[n1]); a segment is summed as doubles or max-reduced as the bit patterns of
// non-negative doubles.
// ---------------------------------------------------------------------------------------------
constexpr int kPeerSmallMax = 16;

__device__ __forceinline__ void ll_store_sys(LLLine* line, unsigned long long bits, unsigned seq) {
    const unsigned long long tag = (unsigned long long)seq << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(line), "l"(tag | (bits & 0xffffffffull)), "l"(tag | (bits >> 32))
                 : "memory");
}
__device__ __forceinline__ bool ll_try_load_sys(const LLLine* line, unsigned seq, unsigned long long& bits) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(line) : "memory");
    bits = (w0 & 0xffffffffull) | (w1 << 32);
    return (unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq;
}

__global__ void peer_allreduce_small_kernel(double* p0, int n0, int max0, double* p1, int n1, int max1, LLLine* const* peer_ll,
                                            const LLLine* my_ll, int rank, int nranks, unsigned seq, double* err) {
    __shared__ unsigned long long s_val[8 * kPeerSmallMax];   // [rank][value], nranks <= 8 per node
    const int n = n0 + n1, tid = threadIdx.x;
    const int parity = (int)(seq & 1u);
    if (tid < n * nranks) {
        const int q = tid / n, i = tid - q * n;
        const double v = i < n0 ? p0[i] : p1[i - n0];
        ll_store_sys(peer_ll[q] + ((size_t)parity * nranks + rank) * kPeerSmallMax + i, (unsigned long long)__double_as_longlong(v), seq);
    }
    if (tid < n * nranks) {
        const int r = tid / n, i = tid - r * n;
        const LLLine* line = my_ll + ((size_t)parity * nranks + r) * kPeerSmallMax + i;
        unsigned long long bits = 0;
        const long long t0 = clock64();
        while (!ll_try_load_sys(line, seq, bits)) {
            if (clock64() - t0 > (1ll << 32)) {   // ~2 s: a peer died; report instead of hanging
                *err = 1.0;
                break;
            }
        }
        s_val[r * kPeerSmallMax + i] = bits;
    }
    __syncthreads();
    if (tid < n) {
        const bool is_max = tid < n0 ? max0 != 0 : max1 != 0;
        double* dst = tid < n0 ? p0 + tid : p1 + (tid - n0);
        if (is_max) {
            unsigned long long best = 0;
            for (int r = 0; r < nranks; ++r) best = max(best, s_val[r * kPeerSmallMax + tid]);
            *dst = __longlong_as_double((long long)best);
        } else {
            double acc = 0.0;
            for (int r = 0; r < nranks; ++r) acc += __longlong_as_double((long long)s_val[r * kPeerSmallMax + tid]);
            *dst = acc;
        }
    }
}

// Reference point for the roofline: a plain grid-stride LDG.128 read of n doubles (what a trivial
// streaming kernel achieves on the same bytes, launch overhead included).  bench/diagnostics only.
__global__ void __launch_bounds__(256) stream_read_kernel(const double2* __restrict__ src, int64_t n2, double* __restrict__ out) {
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = __ldcs(src + i);
        acc += v.x + v.y;
    }
    if (acc == 1.2345e-300) out[0] = acc;   // never true: keeps the loads alive
}

// ---------------------------------------------------------------------------------------------
// rotate / project as stand-alone row-wise kernels (bundleAdjuster.py:7-52): one thread per row, the
// row's own rotation vector (no per-camera hoisting here: the reference passes one parameter row per point)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rodrigues_row(const double* __restrict__ X, const double* __restrict__ w, double (&Y)[3]) {
    const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double t = sqrt(t2);
    if (!(t > 0.0)) {          // theta = 0: v = nan_to_num(0 / 0) = 0 -> the point itself (bundleAdjuster.py:19-28)
        Y[0] = X[0];
        Y[1] = X[1];
        Y[2] = X[2];
        return;
    }
    const double v0 = w[0] / t, v1 = w[1] / t, v2 = w[2] / t;
    double s, c;
    sincos(t, &s, &c);
    const double dot = X[0] * v0 + X[1] * v1 + X[2] * v2;
    const double c0 = v1 * X[2] - v2 * X[1], c1 = v2 * X[0] - v0 * X[2], c2 = v0 * X[1] - v1 * X[0];
    Y[0] = c * X[0] + s * c0 + dot * (1.0 - c) * v0;
    Y[1] = c * X[1] + s * c1 + dot * (1.0 - c) * v1;
    Y[2] = c * X[2] + s * c2 + dot * (1.0 - c) * v2;
}
__global__ void rotate_rows_kernel(const double* __restrict__ pts, const double* __restrict__ rvec, double* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double Y[3];
    rodrigues_row(pts + 3 * i, rvec + 3 * i, Y);
    out[3 * i] = Y[0];
    out[3 * i + 1] = Y[1];
    out[3 * i + 2] = Y[2];
}
__global__ void project_rows_kernel(const double* __restrict__ pts, const double* __restrict__ params, int stride, double k0, double k1,
                                    double k2, double k3, double k4, double k5, double k6, double k7, double k8,
                                    double* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double Y[3];
    const double* p = params + (int64_t)stride * i;
    rodrigues_row(pts + 3 * i, p, Y);
    const double C0 = Y[0] + p[3], C1 = Y[1] + p[4], C2 = Y[2] + p[5];
    const double q0 = k0 * C0 + k1 * C1 + k2 * C2, q1 = k3 * C0 + k4 * C1 + k5 * C2, q2 = k6 * C0 + k7 * C1 + k8 * C2;
    out[2 * i] = q0 / q2;
    out[2 * i + 1] = q1 / q2;
}

}  // namespace mmba
