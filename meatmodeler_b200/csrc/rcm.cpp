// Host-side block pattern of the reduced camera matrix (see rcm.h).
#include "rcm.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan.h"

namespace mmba {

bool build_rcm_pattern(RcmPattern& out, int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
                       const int64_t* pt_idx, int64_t max_blocks, const int32_t* point_order) {
    out = RcmPattern();
    out.n_cams = n_cams;
    if (n_cams <= 0 || n_points <= 0 || n_obs <= 0) return false;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (getenv("MMBA_PLAN_TIMING")) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "rcm %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t; } };
    // cameras of every point (caller's point order), CSR
    std::vector<int64_t> start(n_points + 1, 0);
    for (int64_t i = 0; i < n_obs; ++i) ++start[pt_idx[i] + 1];
    int64_t pairs = 0;
    for (int64_t p = 0; p < n_points; ++p) {
        const int64_t L = start[p + 1];
        pairs += L * (L + 1) / 2;
        start[p + 1] += start[p];
    }
    out.total_pairs = pairs;
    std::vector<int32_t> cams(n_obs);
    {
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n_obs; ++i) cams[fill[pt_idx[i]]++] = (int32_t)cam_idx[i];
        // canonical (ascending) camera lists, so that equal sets compare equal below
        parallel_ranges(n_points, 8192, [&](int64_t p0, int64_t p1, int) {
            for (int64_t p = p0; p < p1; ++p) {
                int32_t* g = cams.data() + start[p];
                const int64_t L = start[p + 1] - start[p];
                for (int64_t a = 1; a < L; ++a) {
                    const int32_t v = g[a];
                    int64_t b = a;
                    while (b > 0 && g[b - 1] > v) {
                        g[b] = g[b - 1];
                        --b;
                    }
                    g[b] = v;
                }
            }
        });
    }
    lap("csr");
    // co-visibility bitmap, upper triangle: bit (i, j), i <= j.  Video-like tracks repeat the same camera
    // list point after point: a point whose list equals its predecessor's is skipped, and a set bit is only
    // tested, so the marking is close to one pass over the observations.
    const int64_t W = (n_cams + 63) / 64;
    std::vector<uint64_t> bits((size_t)n_cams * W, 0);
    // point_order (the plan's internal order: points sorted by first camera) brings equal lists together
    parallel_ranges(n_points, 8192, [&](int64_t q0, int64_t q1, int) {
        int64_t prev = -1;
        for (int64_t q = q0; q < q1; ++q) {
            const int64_t p = point_order ? point_order[q] : q;
            const int64_t b = start[p], L = start[p + 1] - b;
            if (L == 0) continue;
            const bool same = prev >= 0 && start[prev + 1] - start[prev] == L &&
                              std::memcmp(&cams[b], &cams[start[prev]], (size_t)L * sizeof(int32_t)) == 0;
            prev = p;
            if (same) continue;
            for (int64_t a = 0; a < L; ++a)
                for (int64_t c = a; c < L; ++c) {
                    const int32_t i = cams[b + a], j = cams[b + c];
                    uint64_t* w = &bits[(size_t)i * W + (j >> 6)];
                    const uint64_t m = 1ull << (j & 63);
                    if (!(__atomic_load_n(w, __ATOMIC_RELAXED) & m)) __atomic_fetch_or(w, m, __ATOMIC_RELAXED);
                }
        }
    });
    lap("mark");
    // every camera owns its diagonal block (an unobserved camera keeps S_cc = reg I)
    for (int64_t i = 0; i < n_cams; ++i) bits[(size_t)i * W + (i >> 6)] |= 1ull << (i & 63);
    // upper CSR
    out.up_rowptr.assign(n_cams + 1, 0);
    for (int64_t i = 0; i < n_cams; ++i) {
        int64_t cnt = 0;
        for (int64_t w = 0; w < W; ++w) cnt += __builtin_popcountll(bits[(size_t)i * W + w]);
        if (out.up_rowptr[i] + cnt > max_blocks) {
            out = RcmPattern();
            out.n_cams = n_cams;
            out.total_pairs = pairs;
            return false;
        }
        out.up_rowptr[i + 1] = (int32_t)(out.up_rowptr[i] + cnt);
    }
    out.up_cols.resize(out.up_rowptr[n_cams]);
    parallel_ranges(n_cams, 256, [&](int64_t i0, int64_t i1, int) {
        for (int64_t i = i0; i < i1; ++i) {
            int32_t* dst = out.up_cols.data() + out.up_rowptr[i];
            for (int64_t w = 0; w < W; ++w) {
                uint64_t v = bits[(size_t)i * W + w];
                while (v) {
                    *dst++ = (int32_t)(w * 64 + __builtin_ctzll(v));
                    v &= v - 1;
                }
            }
        }
    });
    lap("upper");
    // full CSR: row i = { k < i : (k, i) set } (transposed sources, ascending k) then { j >= i : (i, j) set }
    std::vector<int32_t> cnt(n_cams, 0);
    for (int64_t i = 0; i < n_cams; ++i) {
        cnt[i] += out.up_rowptr[i + 1] - out.up_rowptr[i];
        for (int32_t e = out.up_rowptr[i]; e < out.up_rowptr[i + 1]; ++e)
            if (out.up_cols[e] != i) ++cnt[out.up_cols[e]];
    }
    out.rowptr.assign(n_cams + 1, 0);
    for (int64_t i = 0; i < n_cams; ++i) out.rowptr[i + 1] = out.rowptr[i] + cnt[i];
    const int64_t nnz = out.rowptr[n_cams];
    out.cols.resize(nnz);
    out.rows.resize(nnz);
    out.src.resize(nnz);
    out.diag.assign(n_cams, -1);
    std::vector<int32_t> pos(out.rowptr.begin(), out.rowptr.end() - 1);
    for (int64_t k = 0; k < n_cams; ++k)
        for (int32_t e = out.up_rowptr[k]; e < out.up_rowptr[k + 1]; ++e) {
            const int32_t j = out.up_cols[e];
            if (j == k) continue;
            const int32_t q = pos[j]++;
            out.cols[q] = (int32_t)k;
            out.rows[q] = j;
            out.src[q] = (int32_t)((uint32_t)e | 0x80000000u);
        }
    for (int64_t i = 0; i < n_cams; ++i)
        for (int32_t e = out.up_rowptr[i]; e < out.up_rowptr[i + 1]; ++e) {
            const int32_t q = pos[i]++;
            out.cols[q] = out.up_cols[e];
            out.rows[q] = (int32_t)i;
            out.src[q] = e;
            if (out.up_cols[e] == i) out.diag[i] = q;
        }
    lap("full");
    return true;
}

void build_rcm_partition(RcmPartition& out, const RcmPattern& pat, int max_ctas) {
    out = RcmPartition();
    const int64_t nc = pat.n_cams;
    // a few cameras per CTA even for small problems: fewer CTAs make the grid barrier cheaper
    int64_t g = std::max<int64_t>(1, std::min<int64_t>(max_ctas, (nc + 3) / 4));
    const int64_t cpc = (nc + g - 1) / g;
    g = (nc + cpc - 1) / cpc;
    out.n_ctas = (int)g;
    out.cpc = (int)cpc;
    out.halo_ptr.assign(g + 1, 0);
    out.lcol.resize(pat.nnz_full());
    out.own_l.assign(nc, 0);
    std::vector<int32_t> local_of(nc, -1), list;
    for (int64_t b = 0; b < g; ++b) {
        const int64_t c0 = b * cpc, c1 = std::min(nc, c0 + cpc);
        const int32_t e0 = pat.rowptr[c0], e1 = pat.rowptr[c1];
        list.clear();
        for (int32_t e = e0; e < e1; ++e) {
            const int32_t j = pat.cols[e];
            if (local_of[j] != (int32_t)b) {
                local_of[j] = (int32_t)b;
                list.push_back(j);
            }
        }
        std::sort(list.begin(), list.end());
        // reuse local_of as the index table of this CTA (overwritten CTA by CTA; cameras outside `list` are never read)
        std::vector<int32_t>& idx = local_of;
        for (size_t i = 0; i < list.size(); ++i) idx[list[i]] = (int32_t)i;
        for (int32_t e = e0; e < e1; ++e) out.lcol[e] = (uint16_t)idx[pat.cols[e]];
        for (int64_t c = c0; c < c1; ++c) out.own_l[c] = idx[c];
        for (int32_t j : list) idx[j] = -1 - (int32_t)b;   // never equal to a later CTA id
        out.halo_ptr[b + 1] = out.halo_ptr[b] + (int32_t)list.size();
        out.halo_cols.insert(out.halo_cols.end(), list.begin(), list.end());
        out.nblk_max = std::max(out.nblk_max, (int)(e1 - e0));
        out.nh_max = std::max(out.nh_max, (int)list.size());
    }
}

}  // namespace mmba
