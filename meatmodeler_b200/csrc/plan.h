// Host-side tile plan: observation reordering into point-aligned tiles + point sharding.
//
// Replaces what the reference expresses as a sparsity matrix (bundleAdjuster.py:55-78): the block
// structure of J is implied by (cam_idx, pt_idx); the plan turns it into the streaming layout the
// kernels want.  Pure C++ (no CUDA) so that the CPU test-suite covers it.
#pragma once
#include <cstdint>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace mmba {

constexpr int kTileObs = 256;          // observation slots per tile == threads per CTA
constexpr uint16_t kPadKey = 0xFFFF;   // sorted-key of an empty slot

constexpr int kMaxRun = 32;            // longest camera run one thread sums (longer runs are split; 8 was measured: S-build -3 %, BUILD +5 %)
constexpr int kRcmTabCap = 8192;       // S-build (point, camera) -> slot table entries per tile (u16)

// One record per tile, bulk-copied to shared memory as a unit (2592 bytes, 16-byte multiple).
struct TileMeta {
    int32_t pt0;        // first local (shard-relative, internal-order) point of the tile
    int32_t npts;       // points in the tile
    int32_t ncams;      // distinct cameras in the tile
    int32_t nobs;       // live observation slots (the rest of the 256 are padding)
    int32_t nruns;      // camera runs in the tile's camera-sorted order (each at most kMaxRun long)
    int32_t pair_mode;  // S-build: 0 = point-pair-major; 1 = camera-pair-major (row units, flushed per tile);
                        // 2 = camera-pair-major with <= 256 camera pairs: one 6x6 block per thread, kept in registers
                        //     across consecutive tiles that touch the same cameras
    int32_t npairs;     // sum over the tile's points of L (L + 1) / 2
    int32_t pad;
    uint16_t slot_cam[kTileObs];   // local camera slot of the observation in its tile
    uint16_t slot_pt[kTileObs];    // local point index of the observation in its tile (0xFFFF = empty slot)
    uint16_t sort_src[kTileObs];   // j-th entry of the tile in camera-sorted order -> slot in tile
    uint16_t run_start[kTileObs];  // run r covers sorted positions run_start[r] .. run_start[r+1] (or nobs)
    uint16_t run_cam[kTileObs];    // local camera slot of run r
};

struct Plan {
    int64_t n_cams = 0, n_points = 0, n_obs = 0;
    int rank = 0, nranks = 1;
    // internal point order (all points, identical on every rank): sorted by first camera
    std::vector<int32_t> point_perm;     // internal -> caller
    std::vector<int64_t> shard_begin;    // nranks+1 cut positions in internal point order
    int64_t pt_begin = 0, pt_end = 0;    // this rank's internal point range
    int64_t n_obs_local = 0;
    // tiles of this rank
    int64_t n_tiles = 0, n_slots = 0;
    int max_tile_cams = 0, max_tile_pts = 0;
    int cam_stride = 0;                  // entries per tile in tile_cams (max_tile_cams rounded up to 4)
    std::vector<TileMeta> meta;
    std::vector<int32_t> tile_cams;      // [tile][cam_stride] global camera ids (ascending), -1 padded
    std::vector<int64_t> slot_obs;       // padded slot -> caller's observation index, -1 = empty

    int64_t n_points_local() const { return pt_end - pt_begin; }
};

// Static-chunked parallel loop over [0, n) on short-lived std::threads (at most 8): fn(begin, end, worker).
// No thread pool is left spinning between calls (an OpenMP runtime's idle workers were measured to slow
// the surrounding CUDA calls by several times).
template <typename F>
inline void parallel_ranges(int64_t n, int64_t min_chunk, F&& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int64_t workers = std::min<int64_t>(std::max(1u, std::min(8u, hw ? hw / 2 : 1u)), std::max<int64_t>(1, n / std::max<int64_t>(1, min_chunk)));
    if (workers <= 1) {
        fn((int64_t)0, n, 0);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t step = (n + workers - 1) / workers;
    for (int64_t w = 0; w < workers; ++w) {
        const int64_t b = w * step, e = std::min(n, b + step);
        if (b >= e) break;
        pool.emplace_back([&fn, b, e, w]() { fn(b, e, (int)w); });
    }
    for (auto& t : pool) t.join();
}

// Returns 0 or a negative MMBA_ERR_* code; `err` receives the message.
int build_plan(Plan& plan, int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
               const int64_t* pt_idx, int rank, int nranks, std::string& err);

}  // namespace mmba
