// Host-side tile plan: observation reordering into point-aligned tiles + point sharding.
//
// Replaces what the reference expresses as a sparsity matrix (bundleAdjuster.py:55-78): the block
// structure of J is implied by (cam_idx, pt_idx); the plan turns it into the streaming layout the
// kernels want.  Pure C++ (no CUDA) so that the CPU test-suite covers it.
#pragma once
#include <cstdint>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
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

// A small pool of SLEEPING workers (condition variable, no spinning: an OpenMP runtime's spinning idle workers were
// measured to slow the surrounding CUDA calls by several times; spawning std::threads per call cost ~15 us per thread,
// 2-3 ms per adjustPoints call over the staged copies).  One job at a time: a second concurrent caller (the pattern
// thread next to the main thread) falls back to short-lived threads.
class WorkerPool {
public:
    static WorkerPool& get() {
        static WorkerPool pool;
        return pool;
    }
    // fn(w) for w in [0, n): w = 0 on the calling thread, the rest on pool workers; returns false if the pool is busy
    template <typename F>
    bool run(int n, F&& fn) {
        std::unique_lock<std::mutex> owner(busy_, std::try_to_lock);
        if (!owner.owns_lock()) return false;
        {
            std::lock_guard<std::mutex> g(m_);
            while ((int)threads_.size() < n - 1) {
                const int id = (int)threads_.size() + 1;
                threads_.emplace_back([this, id]() { loop(id); });
            }
            job_ = [&fn](int w) { fn(w); };
            job_n_ = n;
            pending_ = n - 1;
            ++gen_;
        }
        cv_work_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> g(m_);
        cv_done_.wait(g, [this]() { return pending_ == 0; });
        job_ = nullptr;
        return true;
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++gen_;
        }
        cv_work_.notify_all();
        for (auto& t : threads_) t.join();
    }

private:
    void loop(int id) {
        uint64_t seen = 0;
        {
            std::lock_guard<std::mutex> g(m_);
            seen = gen_ - 1;      // created inside run(): the current generation is this worker's first job
        }
        while (true) {
            std::function<void(int)> job;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_work_.wait(g, [&]() { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                if (id >= job_n_) continue;
                job = job_;
            }
            job(id);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::mutex busy_, m_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> threads_;
    std::function<void(int)> job_;
    int job_n_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

// Static-chunked parallel loop over [0, n): fn(begin, end, worker), at most max_workers workers (half of the cores
// unless all_cores: memory-bound copies scale past that, the plan's compute phases do not).
template <typename F>
inline void parallel_ranges(int64_t n, int64_t min_chunk, F&& fn, unsigned max_workers = 8, bool all_cores = false) {
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned want = all_cores ? (hw ? hw : 1u) : (hw ? hw / 2 : 1u);
    int64_t workers = std::min<int64_t>(std::max(1u, std::min(max_workers, want)), std::max<int64_t>(1, n / std::max<int64_t>(1, min_chunk)));
    if (workers <= 1) {
        fn((int64_t)0, n, 0);
        return;
    }
    const int64_t step = (n + workers - 1) / workers;
    auto part = [&](int w) {
        const int64_t b = w * step, e = std::min(n, b + step);
        if (b < e) fn(b, e, w);
    };
    if (WorkerPool::get().run((int)workers, part)) return;
    std::vector<std::thread> pool;
    for (int64_t w = 0; w < workers; ++w) pool.emplace_back([&part, w]() { part((int)w); });
    for (auto& t : pool) t.join();
}

// Returns 0 or a negative MMBA_ERR_* code; `err` receives the message.
int build_plan(Plan& plan, int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
               const int64_t* pt_idx, int rank, int nranks, std::string& err);

}  // namespace mmba
