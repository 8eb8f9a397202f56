// Shared definitions of the reduced-camera-system PCG: vector bundle, small dense helpers, system-scope
// load/store used by the peer exchange.  Included by kernels.cuh (the fused MATVEC+update kernel) and
// veckernels.cuh (the stand-alone kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmba {

// packed upper-triangle index helpers for 6x6 (21) and 3x3 (6)
__host__ __device__ constexpr int tri6(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }
__host__ __device__ constexpr int tri3(int a, int b) { return a * 3 - a * (a - 1) / 2 + (b - a); }
__host__ __device__ constexpr int tri6_row(int idx) {
    return idx < 6 ? 0 : idx < 11 ? 1 : idx < 15 ? 2 : idx < 18 ? 3 : idx < 20 ? 4 : 5;
}
__host__ __device__ constexpr int tri6_col(int idx) { return idx - tri6(tri6_row(idx), tri6_row(idx)) + tri6_row(idx); }

constexpr int kCamBlock = 128;   // threads per block of the thread-per-camera kernels
constexpr int kMaxCamBlocks = 1024;

// per-block partial sums of the PCG (deterministic reductions)
enum Part { P_RHO0 = 0, P_B2, P_COUNT };

__device__ __forceinline__ void sym6_matvec(const double* __restrict__ m, const double (&v)[6], double (&o)[6]) {
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        double s = 0;
#pragma unroll
        for (int b = 0; b < 6; ++b) s += m[a <= b ? tri6(a, b) : tri6(b, a)] * v[b];
        o[a] = s;
    }
}

// inverse of a symmetric positive definite 6x6 (packed upper triangle in/out) by Cholesky;
// falls back to the inverse diagonal when the factorisation breaks down
__device__ __forceinline__ void sym6_inverse(const double (&s)[21], double (&inv)[21]) {
    double L[6][6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = s[tri6(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        const double l = sqrt(d);
        L[j][j] = l;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = s[tri6(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v / l;
        }
    }
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 21; ++i) inv[i] = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) inv[tri6(a, a)] = s[tri6(a, a)] > 0.0 ? 1.0 / s[tri6(a, a)] : 1.0;
        return;
    }
    // W = L^-1 (lower triangular), inverse = W^T W
    double W[6][6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        W[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = 0;
#pragma unroll
            for (int k = j; k < i; ++k) v -= L[i][k] * W[k][j];
            W[i][j] = v / L[i][i];
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) {
            double v = 0;
#pragma unroll
            for (int k = b; k < 6; ++k) v += W[k][a] * W[k][b];
            inv[tri6(a, b)] = v;
        }
}

struct PcgVecs {
    const double* gc;     // [Nc][6]
    const double* sinv;   // [Nc][6] scale_inv of the camera parameters
    double* y;            // [Nc][6] scatter target of the schur kernels
    double* Sd;           // [Nc][21] diagonal blocks of the reduced system, sum (Jc^T Jc - F E^T), unscaled
    double* Pinv;         // [Nc][21]
    double *x, *r, *z, *p, *q, *xt;   // [Nc][6]
    double* part;         // [P_COUNT][kMaxCamBlocks] per-block partials of pcg_init
    double* state;        // [0] rho, [1] ||b||^2, [2] ||r||^2 carried between iterations
    int* flags;           // [0] done, [1] iterations
    int n_cams;
    // peer exchange (nranks > 1): this rank's receive slots [2][nranks][n6] and arrival flags
    // [2][nranks], plus every rank's mapped slots / flags; nranks == 1 when the exchange is off
    // (y is then complete in place, all-reduced by NCCL if there are several ranks)
    const double* xslots;
    const unsigned long long* xflags;
    double* const* peer_slots;
    unsigned long long* const* peer_flags;
    int nranks, rank, n6;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace mmba
