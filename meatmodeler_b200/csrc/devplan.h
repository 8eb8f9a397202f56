// Device-side tile plan and reduced-camera-matrix pattern (the GPU twin of plan.cpp / rcm.cpp).
//
// mmba_set_problem used to spend more host time ordering observations than the solve spends on the GPU
// (C2: 7.5 ms plan + 3.2 ms pattern against a 7 ms solve; C4: ~190 ms against 37 ms).  Here the same plan is
// built by hand-written kernels from the index arrays as uploaded:
//   stats      per point: observation count, first / last camera, first camera of the upper half (ring-aware key)
//   order      stable LSD radix sort of the points by key camera -> internal point order, prefix sums, shard cuts
//   group      stable LSD radix sort of the (local) observations by (internal point, camera)
//   tiles      greedy point-aligned tiles of 256 slots: the sequential "next tile starts where 256 slots are used up"
//              chain is resolved by pointer doubling (log2(tiles) rounds)
//   tile build one CTA per tile: distinct cameras, local slots, camera-sorted order and runs (bitonic sort in shared
//              memory), S-build strategy, tile-major copy of the observed pixels
//   pattern    co-visibility bitmap -> upper / full block-CSR of the reduced camera matrix, PCG partition + halos
// The result is bit-identical to build_plan / build_rcm_pattern / build_rcm_partition on the host (tests compare
// them), which stay as the CPU statement of the layout (mmba_plan_create, mmba_host_rcm_pattern).
// Replaces: pointAdjustmentSparsity (bundleAdjuster.py:55-78) and scipy's column grouping of it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "plan.h"

namespace mmba {

// grow-only device allocation
struct DevBuf {
    char* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes);
    void release();
};

// bump allocator over a DevBuf (256-byte aligned); base == nullptr: measure only
struct Carver {
    char* base = nullptr;
    size_t off = 0;
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

// scalars the kernels leave for the host (one D2H copy per synchronisation point)
struct PlanInfo {
    int n_tiles, n_obs_local, npo_local, err_track;     // err_track: 1 + caller's point index with > 256 observations, 0 = none
    int max_tile_cams, max_tile_pts, pt_begin, pt_end;
    int nnz_up, nnz_full, nblk_max, nh_max;
    long long total_pairs;
    int shard_begin[17];
    int pad[3];
};

struct DevPlan {
    int64_t n_cams = 0, n_points = 0, n_obs = 0;
    int rank = 0, nranks = 1;
    // host copies of PlanInfo
    int64_t pt_begin = 0, pt_end = 0, n_obs_local = 0, n_tiles = 0, n_slots = 0;
    int max_tile_cams = 0, max_tile_pts = 0, cam_stride = 0;
    // persistent outputs (live as long as the problem)
    int32_t* point_perm = nullptr;   // [n_points] internal -> caller
    int32_t* point_inv = nullptr;    // [n_points] caller -> internal
    TileMeta* meta = nullptr;        // [n_tiles]
    int32_t* tile_cams = nullptr;    // [n_tiles][cam_stride]
    double* uvt = nullptr;           // [n_tiles][2][256]
    int32_t* slot_obs = nullptr;     // [n_slots] caller's observation index, -1 = empty
    // reduced camera matrix pattern (valid when rcm_ok)
    bool rcm_ok = false;
    int64_t nnz_up = 0, nnz_full = 0, total_pairs = 0;
    int n_ctas = 0, cpc = 0, nblk_max = 0, nh_max = 0;
    int *up_rowptr = nullptr, *up_cols = nullptr, *rowptr = nullptr, *cols = nullptr, *rows = nullptr, *src = nullptr,
        *diag = nullptr, *halo_ptr = nullptr, *halo_cols = nullptr, *own_l = nullptr;
    uint16_t* lcol = nullptr;
};

// Everything the planner allocates; owned by the handle, reused across problems.
struct DevPlanner {
    DevBuf in, in2, work, work2, out, pat, tmp;
    PlanInfo* d_info = nullptr;      // device
    PlanInfo* h_info = nullptr;      // pinned
    // scratch of the point stages (work) and of the observation stages (work2); carved by the stage functions
    struct {
        int *count, *first, *last, *first_hi, *key, *start_all, *hist, *bsum;
        unsigned long long *k0, *k1;
        unsigned *v0, *v1;
    } a{};
    struct {
        unsigned long long *k0, *k1, *bits;
        unsigned *v0, *v1;
        int *hist, *bsum, *start, *jump0, *jump1, *mark, *tidx, *tile_p0, *pc, *cnt_up, *cnt_full, *hbits_cnt;
        unsigned long long* hbits;
        const unsigned long long* keys_sorted;
        const unsigned* vals_sorted;
        size_t words;
        int64_t tiles_max;
    } c{};
    // staged inputs of the local observation chunk (device)
    int32_t* cam = nullptr;
    int32_t* pt = nullptr;
    double* uv = nullptr;
    int32_t* gidx = nullptr;         // global observation index of every local observation (nullptr = identity)
    int64_t n_in = 0;                // observations staged
    // sharded set-up: the staged chunk grouped by destination rank (send side) and what this rank receives
    int32_t *cam_s = nullptr, *pt_s = nullptr, *gidx_s = nullptr;
    double* uv_s = nullptr;
    int32_t *cam_r = nullptr, *pt_r = nullptr, *gidx_r = nullptr;
    double* uv_r = nullptr;
    int* d_counts = nullptr;         // [nranks] observations bound for every rank, then [nranks][nranks] all-gathered
    void release();
};

// Stage A (enqueue only): per-point statistics of the staged observations.  After it, a sharded solve combines
// count / first / last / first_hi over ranks (devplan_stat_block + devplan_combine_stats).
int devplan_stats(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err);
void devplan_stat_arrays(DevPlanner& P, const DevPlan& D, int** count, int** first, int** last, int** first_hi);
// the four arrays as one contiguous block (what a sharded set-up all-gathers) and the kernel that combines the gathered
// per-rank blocks (counts add, first / first_hi: minimum, last: maximum)
void devplan_stat_block(DevPlanner& P, const DevPlan& D, int** block, size_t* n_ints);
void devplan_combine_stats(DevPlanner& P, const DevPlan& D, const int* gathered, int nranks, cudaStream_t s);
// Stage B (enqueue only): point order, prefix sums, shard cuts (info.shard_begin, pt_begin, pt_end).
int devplan_order(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err);
// Sharded set-up between stages B and C.  Every observation belongs to the rank that owns its point:
//   devplan_dispatch_pack   (enqueue) groups the staged chunk (observations o0, o0 + 1, ... of the caller's arrays) by
//                           destination rank, stable; P.d_counts[r] = observations bound for rank r
//   devplan_dispatch_recv   allocates the receive side for n_recv observations and makes it the planner's input
//                           (the caller moves the data: one grouped send / receive per peer, NCCL over NVLink)
int devplan_dispatch_pack(DevPlanner& P, DevPlan& D, int64_t o0, cudaStream_t s, std::string& err);
int devplan_dispatch_recv(DevPlanner& P, int64_t n_recv, std::string& err);
// bits |= gathered[r] for r < nranks (co-visibility bitmaps of the other ranks' points)
void devplan_or_bitmaps(unsigned long long* bits, const unsigned long long* gathered, size_t n_words, int nranks, cudaStream_t s);

// Stage C (enqueue only): observation grouping, tiles, tile tables, co-visibility bitmap and its block counts.
// The observations staged in P (cam / pt / uv / gidx) must be exactly this rank's.
int devplan_tiles(DevPlanner& P, DevPlan& D, bool want_pattern, cudaStream_t s, std::string& err);
// bitmap of the co-visibility pattern: [n_cams][words] uint64, symmetric; a sharded solve ORs it over ranks between
// devplan_tiles and devplan_pattern_sizes
void devplan_bitmap(DevPlanner& P, const DevPlan& D, unsigned long long** bits, size_t* n_words);
int devplan_pattern_sizes(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err);
// Synchronises, reads PlanInfo into D (errors: MMBA_ERR_TRACK ...).
int devplan_sync_sizes(DevPlanner& P, DevPlan& D, cudaStream_t s, std::string& err);
// Stage D (enqueue + one sync): CSR arrays of the pattern, PCG partition (n_ctas CTAs), halo lists.
int devplan_pattern_fill(DevPlanner& P, DevPlan& D, int max_ctas, cudaStream_t s, std::string& err);

// small device helpers shared with the engine
void devplan_gather_x(const double* x_caller, double* x_internal, const int32_t* point_perm, int64_t n_cams, int64_t pt_begin,
                      int64_t npl, cudaStream_t s);
void devplan_scatter_x(const double* x_internal, double* x_caller, const int32_t* point_perm, int64_t n_cams, int64_t pt_begin,
                       int64_t npl, bool with_cams, cudaStream_t s);
// rows [row0, row0 + rows) of the tile-major array src ([tile][src_rows][256]) -> caller-ordered out (n_obs, rows)
void devplan_scatter_slots(const double* src, int src_rows, int row0, int rows, const int32_t* slot_obs, int64_t n_slots,
                           double* out, cudaStream_t s);

}  // namespace mmba
