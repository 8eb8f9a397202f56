// Block-sparsity pattern of the reduced camera matrix  S = U - W V'^-1 W^T  (host side).
//
// Two cameras share a block of S iff some point is observed by both.  The pattern is computed once per
// problem from the FULL problem (identical on every rank of a sharded solve, so the block values can be
// all-reduced), in two forms:
//   * upper triangle (i <= j) in CSR order: the accumulation target of the S-build kernel;
//   * full (both triangles) CSR: what the PCG kernel multiplies with; every full block names its source
//     block in the upper triangle and whether it is read transposed.
// Pure C++ (no CUDA): covered by the CPU test-suite through mmba_host_rcm_pattern.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mmba {

struct RcmPattern {
    int64_t n_cams = 0;
    int64_t total_pairs = 0;            // sum over points of L (L + 1) / 2: pair-blocks one S-build evaluates
    std::vector<int32_t> up_rowptr;     // [n_cams + 1]
    std::vector<int32_t> up_cols;       // [nnz_up], ascending per row, the first entry of row i is the diagonal block
    std::vector<int32_t> rowptr;        // [n_cams + 1] full pattern
    std::vector<int32_t> cols;          // [nnz_full] ascending per row
    std::vector<int32_t> rows;          // [nnz_full] row of each full block
    std::vector<int32_t> src;           // [nnz_full] source block in the upper triangle; bit 31 set = transposed
    std::vector<int32_t> diag;          // [n_cams] full-pattern index of the diagonal block
    int64_t nnz_up() const { return (int64_t)up_cols.size(); }
    int64_t nnz_full() const { return (int64_t)cols.size(); }
};

// Partition of the cameras over the CTAs of the PCG kernel (contiguous ranges of `cpc` cameras) with, per CTA, the
// sorted union of the columns its rows touch (its "halo": the entries of the search direction it needs).
struct RcmPartition {
    int n_ctas = 0, cpc = 0;            // CTAs, cameras per CTA
    int nblk_max = 0, nh_max = 0;       // largest number of blocks / halo columns of one CTA
    std::vector<int32_t> halo_ptr;      // [n_ctas + 1]
    std::vector<int32_t> halo_cols;     // global camera ids, ascending per CTA
    std::vector<uint16_t> lcol;         // [nnz_full] column of each block as an index into its CTA's halo list
    std::vector<int32_t> own_l;         // [n_cams] index of camera c in its CTA's halo list
};
void build_rcm_partition(RcmPartition& out, const RcmPattern& pat, int max_ctas);

// max_blocks: give up (return false, pattern left empty) as soon as the upper triangle exceeds this many blocks.
// point_order (optional, n_points entries): a permutation of the points that places equal camera lists next to each
// other (the plan's internal order); only speeds the marking up.
bool build_rcm_pattern(RcmPattern& out, int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
                       const int64_t* pt_idx, int64_t max_blocks, const int32_t* point_order = nullptr);

}  // namespace mmba
