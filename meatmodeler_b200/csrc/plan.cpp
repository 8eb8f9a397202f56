// Host-side tile plan (see plan.h).
#include "plan.h"

#include <algorithm>
#include <numeric>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../../include/mmba.h"

namespace mmba {

int build_plan(Plan& plan, int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
               const int64_t* pt_idx, int rank, int nranks, std::string& err) {
    if (n_cams <= 0 || n_points <= 0 || n_obs <= 0 || !cam_idx || !pt_idx) {
        err = "set_problem: sizes must be positive and index arrays non-null";
        return MMBA_ERR_ARG;
    }
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        err = "set_problem: rank/nranks out of range";
        return MMBA_ERR_ARG;
    }
    if (n_cams > INT32_MAX / 32 || n_points > INT32_MAX / 16 || n_obs > (int64_t)INT32_MAX * 8) {
        err = "set_problem: problem too large for 32-bit device indices";
        return MMBA_ERR_ARG;
    }
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (getenv("MMBA_PLAN_TIMING")) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "plan %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t; } };
    plan = Plan();
    plan.n_cams = n_cams;
    plan.n_points = n_points;
    plan.n_obs = n_obs;
    plan.rank = rank;
    plan.nranks = nranks;

    // 1. per-point observation count and first camera.  "First" is the smallest camera id, except for tracks
    //    that wrap around a closed camera ring (turntable video: frames N-1 and 0 are neighbours): a track spanning
    //    more than half of the ids is keyed by its smallest camera in the upper half, so that the window
    //    {190..199, 0..9} sorts next to {190..199} instead of mixing with {0..19} (fewer cameras per tile).
    std::vector<int32_t> count(n_points, 0), first_cam(n_points, (int32_t)n_cams);
    {
        std::vector<int32_t> last_cam(n_points, -1), first_hi(n_points, (int32_t)n_cams);
        const int64_t half = n_cams / 2;
        // observation chunks in parallel: relaxed atomic count / min / max per point (the results do not depend
        // on the order)
        auto amin = [](int32_t* a, int32_t v) {
            int32_t cur = __atomic_load_n(a, __ATOMIC_RELAXED);
            while (v < cur && !__atomic_compare_exchange_n(a, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
            }
        };
        auto amax = [](int32_t* a, int32_t v) {
            int32_t cur = __atomic_load_n(a, __ATOMIC_RELAXED);
            while (v > cur && !__atomic_compare_exchange_n(a, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
            }
        };
        int64_t bad[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        parallel_ranges(n_obs, 65536, [&](int64_t i0, int64_t i1, int worker) {
            for (int64_t i = i0; i < i1; ++i) {
                const int64_t c = cam_idx[i], p = pt_idx[i];
                if (c < 0 || c >= n_cams || p < 0 || p >= n_points) {
                    if (bad[worker] < 0) bad[worker] = i;
                    continue;
                }
                __atomic_fetch_add(&count[p], 1, __ATOMIC_RELAXED);
                amin(&first_cam[p], (int32_t)c);
                amax(&last_cam[p], (int32_t)c);
                if (c >= half) amin(&first_hi[p], (int32_t)c);
            }
        });
        int64_t first_bad = -1;
        for (int w = 0; w < 8; ++w)
            if (bad[w] >= 0 && (first_bad < 0 || bad[w] < first_bad)) first_bad = bad[w];
        if (first_bad >= 0) {
            err = "set_problem: index out of range at observation " + std::to_string(first_bad);
            return MMBA_ERR_ARG;
        }
        for (int64_t p = 0; p < n_points; ++p)
            if (count[p] && last_cam[p] - first_cam[p] > half && first_hi[p] < n_cams) first_cam[p] = first_hi[p];
    }
    for (int64_t p = 0; p < n_points; ++p) {
        if (count[p] > kTileObs) {
            err = "set_problem: point " + std::to_string(p) + " has " + std::to_string(count[p]) +
                  " observations; one tile holds at most " + std::to_string(kTileObs);
            return MMBA_ERR_TRACK;
        }
    }

    lap("1 count");
    // 2. internal point order: counting sort by first camera (stable -> ties by caller index);
    //    unobserved points (first_cam == n_cams) go last
    plan.point_perm.resize(n_points);
    {
        std::vector<int64_t> bucket(n_cams + 2, 0);
        for (int64_t p = 0; p < n_points; ++p) ++bucket[first_cam[p] + 1];
        for (int64_t c = 0; c <= n_cams; ++c) bucket[c + 1] += bucket[c];
        for (int64_t p = 0; p < n_points; ++p) plan.point_perm[bucket[first_cam[p]]++] = (int32_t)p;
    }
    std::vector<int32_t> point_inv(n_points);
    for (int64_t q = 0; q < n_points; ++q) point_inv[plan.point_perm[q]] = (int32_t)q;

    lap("2 order");
    // 3. shard cuts: contiguous internal point ranges balanced by observation count
    plan.shard_begin.assign(nranks + 1, n_points);
    plan.shard_begin[0] = 0;
    {
        int64_t acc = 0;
        int next = 1;
        for (int64_t q = 0; q < n_points && next < nranks; ++q) {
            while (next < nranks && acc >= (n_obs * next + nranks - 1) / nranks) {
                plan.shard_begin[next++] = q;
            }
            acc += count[plan.point_perm[q]];
        }
        // remaining cuts (if any) stay at n_points -> empty shards
    }
    plan.pt_begin = plan.shard_begin[rank];
    plan.pt_end = plan.shard_begin[rank + 1];
    const int64_t npl = plan.pt_end - plan.pt_begin;

    // 4. observations of the local points, grouped by internal point (stable), cameras ascending
    std::vector<int64_t> start(npl + 1, 0);
    for (int64_t q = 0; q < npl; ++q) start[q + 1] = start[q] + count[plan.point_perm[plan.pt_begin + q]];
    plan.n_obs_local = start[npl];
    std::vector<int64_t> grouped(plan.n_obs_local);
    std::vector<int32_t> gcam(plan.n_obs_local);
    {
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        // scatter in parallel (atomic slot counters); the per-point sort below orders by (camera, observation),
        // so the result does not depend on the arrival order
        parallel_ranges(n_obs, 65536, [&](int64_t i0, int64_t i1, int) {
            for (int64_t i = i0; i < i1; ++i) {
                const int64_t q = (int64_t)point_inv[pt_idx[i]] - plan.pt_begin;
                if (q >= 0 && q < npl) grouped[__atomic_fetch_add(&fill[q], 1, __ATOMIC_RELAXED)] = i;
            }
        });
        parallel_ranges(npl, 4096, [&](int64_t q0, int64_t q1, int) {
            for (int64_t q = q0; q < q1; ++q) {
                // tracks are short: insertion sort (stable), usually already ascending.  The camera of every
                // grouped observation is kept (gcam) so that the tile pass below reads it sequentially.
                int64_t* g = grouped.data() + start[q];
                int32_t* gc = gcam.data() + start[q];
                const int64_t L = start[q + 1] - start[q];
                for (int64_t a = 0; a < L; ++a) gc[a] = (int32_t)cam_idx[g[a]];
                for (int64_t a = 1; a < L; ++a) {
                    const int64_t v = g[a];
                    const int32_t cv = gc[a];
                    int64_t b = a;
                    while (b > 0 && (gc[b - 1] > cv || (gc[b - 1] == cv && g[b - 1] > v))) {
                        g[b] = g[b - 1];
                        gc[b] = gc[b - 1];
                        --b;
                    }
                    g[b] = v;
                    gc[b] = cv;
                }
            }
        });
    }

    lap("4 group");
    // 5. greedy point-aligned tiles
    struct Range { int64_t p0, p1; };
    std::vector<Range> ranges;
    {
        int64_t p0 = 0, used = 0;
        for (int64_t q = 0; q < npl; ++q) {
            const int64_t L = start[q + 1] - start[q];
            if (L == 0) continue;  // unobserved points sit at the end of the order; never in a tile
            if (used + L > kTileObs) {
                ranges.push_back({p0, q});
                p0 = q;
                used = 0;
            }
            if (used == 0) p0 = q;
            used += L;
        }
        if (used > 0) {
            int64_t p1 = npl;
            while (p1 > p0 && start[p1] - start[p1 - 1] == 0) --p1;
            ranges.push_back({p0, p1});
        }
    }
    plan.n_tiles = (int64_t)ranges.size();
    plan.n_slots = plan.n_tiles * kTileObs;
    plan.meta.resize(plan.n_tiles);
    plan.slot_obs.assign(plan.n_slots, -1);

    lap("5 ranges+alloc");
    // 6. per tile: local camera table, local slots, camera-sorted order (tiles are independent)
    // camera lists: appended to one pool per worker (no per-tile allocation), copied to tile_cams afterwards
    std::vector<int32_t> pool[8];
    std::vector<int64_t> cams_at(plan.n_tiles);      // offset of tile t's list in its worker's pool
    std::vector<int8_t> cams_worker(plan.n_tiles);
    int max_cams_w[8] = {0}, max_pts_w[8] = {0};
    parallel_ranges(plan.n_tiles, 256, [&](int64_t t0, int64_t t1, int worker) {
        std::vector<int32_t> stamp(n_cams, -1), local_of(n_cams, 0);
        std::vector<int32_t>& mypool = pool[worker];
        mypool.reserve((size_t)(t1 - t0) * 32);
        std::vector<uint16_t> idx(kTileObs);
        uint16_t cnt[kTileObs + 1];
        int max_cams = 0, max_pts = 0;
        for (int64_t t = t0; t < t1; ++t) {
            const Range rg = ranges[t];
            const int64_t o0 = start[rg.p0], o1 = start[rg.p1];
            const int n = (int)(o1 - o0);
            const int64_t base = t * kTileObs;
            const size_t at = mypool.size();
            cams_at[t] = (int64_t)at;
            cams_worker[t] = (int8_t)worker;
            for (int i = 0; i < n; ++i) {
                const int32_t c = gcam[o0 + i];
                if (stamp[c] != (int32_t)t) {
                    stamp[c] = (int32_t)t;
                    mypool.push_back(c);
                }
            }
            std::sort(mypool.begin() + at, mypool.end());
            const int32_t* cams = mypool.data() + at;
            const size_t ncams_t = mypool.size() - at;
            for (size_t s = 0; s < ncams_t; ++s) local_of[cams[s]] = (int32_t)s;
            TileMeta& m = plan.meta[t];
            m.pt0 = (int32_t)rg.p0;
            m.npts = (int32_t)(rg.p1 - rg.p0);
            m.ncams = (int32_t)ncams_t;
            m.nobs = n;
            max_cams = std::max(max_cams, m.ncams);
            max_pts = std::max(max_pts, m.npts);
            m.pair_mode = m.npairs = m.pad = 0;
            for (int j = 0; j < kTileObs; ++j) {
                m.slot_cam[j] = 0;
                m.slot_pt[j] = 0xFFFF;   // empty slots carry the pad marker
                m.sort_src[j] = (uint16_t)j;
                m.run_start[j] = 0;
                m.run_cam[j] = 0;
            }
            int i = 0;
            int64_t npairs = 0;
            bool dup = false;   // a camera observing the same point twice
            for (int64_t q = rg.p0; q < rg.p1; ++q) {
                const int64_t L = start[q + 1] - start[q];
                npairs += L * (L + 1) / 2;
                for (int64_t o = start[q]; o < start[q + 1]; ++o, ++i) {
                    plan.slot_obs[base + i] = grouped[o];
                    m.slot_cam[i] = (uint16_t)local_of[gcam[o]];
                    m.slot_pt[i] = (uint16_t)(q - rg.p0);
                    if (o > start[q] && m.slot_cam[i] == m.slot_cam[i - 1]) dup = true;
                }
            }
            // S-build strategy of the tile: with few cameras every (camera pair, row) unit accumulates over the
            // tile's points in registers and issues one RED per entry; otherwise one unit per (observation pair, row)
            m.npairs = (int32_t)npairs;
            const int64_t campairs = (int64_t)m.ncams * (m.ncams + 1) / 2;
            m.pair_mode = (!dup && (int64_t)m.npts * m.ncams <= kRcmTabCap && campairs * m.npts <= 4 * npairs) ? 1 : 0;
            if (m.pair_mode == 1 && campairs <= kTileObs) m.pair_mode = 2;   // one whole 6x6 block per thread, kept across tiles
            // stable counting sort of the slots by local camera
            const int nc = m.ncams;
            for (int c = 0; c <= nc; ++c) cnt[c] = 0;
            for (int j = 0; j < n; ++j) ++cnt[m.slot_cam[j] + 1];
            for (int c = 0; c < nc; ++c) cnt[c + 1] = (uint16_t)(cnt[c + 1] + cnt[c]);
            for (int j = 0; j < n; ++j) idx[cnt[m.slot_cam[j]]++] = (uint16_t)j;
            int nruns = 0;
            for (int j = 0; j < n; ++j) {
                m.sort_src[j] = idx[j];
                const uint16_t key = m.slot_cam[idx[j]];
                if (j == 0 || key != m.slot_cam[idx[j - 1]] || j - m.run_start[nruns - 1] >= kMaxRun) {
                    m.run_start[nruns] = (uint16_t)j;
                    m.run_cam[nruns] = key;
                    ++nruns;
                }
            }
            m.nruns = nruns;
        }
        max_cams_w[worker] = max_cams;
        max_pts_w[worker] = max_pts;
    });
    lap("6 tiles");
    for (int w = 0; w < 8; ++w) {
        plan.max_tile_cams = std::max(plan.max_tile_cams, max_cams_w[w]);
        plan.max_tile_pts = std::max(plan.max_tile_pts, max_pts_w[w]);
    }
    plan.cam_stride = std::max(4, (plan.max_tile_cams + 3) / 4 * 4);
    plan.tile_cams.assign((size_t)plan.n_tiles * plan.cam_stride, -1);
    for (int64_t t = 0; t < plan.n_tiles; ++t) {
        const int32_t* src = pool[cams_worker[t]].data() + cams_at[t];
        std::copy(src, src + plan.meta[t].ncams, plan.tile_cams.begin() + t * plan.cam_stride);
    }
    lap("7 tile_cams");
    return MMBA_OK;
}

}  // namespace mmba
