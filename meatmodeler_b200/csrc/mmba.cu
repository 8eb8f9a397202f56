// libmmba.so — C-ABI (include/mmba.h) + host driver of the B200 bundle-adjustment engine.
//
// The driver re-implements the outer loop the reference reaches through
// scipy.optimize.least_squares(method='trf', tr_solver='lsmr', x_scale='jac')
// (bundleAdjuster.py:180-192 -> scipy/optimize/_lsq/trf.py:415-587): Marquardt column scaling with a
// monotone scale, Cauchy-step regulariser, a damped Gauss-Newton step (here: Schur elimination of
// the 3x3 point blocks + block-Jacobi PCG on the camera system instead of LSMR), the 2-D subspace
// trust-region problem, the radius rule and the termination tests.  Every O(No), O(Np), O(Nc)
// operation is a CUDA kernel (kernels.cuh, veckernels.cuh); the host only does O(1) scalar work.
//
// Multi-GPU (nranks > 1): observations are sharded by point (plan.cpp), cameras are replicated, and
// the camera-sized partial sums (U, g_c, Schur products) plus a handful of scalars are summed with
// ncclAllReduce.  NCCL is bound at run time with dlopen so that a single-GPU process never needs it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mmba.h"
#include "devplan.h"
#include "kernels.cuh"
#include "plan.h"
#include "rcm.cuh"
#include "rcm.h"
#include "trf_host.h"
#include "veckernels.cuh"

using namespace mmba;

// ---------------------------------------------------------------------------------------------
// run-time NCCL binding
// ---------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
thread_local std::string g_thread_error;

bool load_nccl(std::string& err) {
    if (g_nccl.lib) return true;
    // prefer a copy the process has already loaded (torch bundles its own libnccl.so.2)
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        err = std::string("cannot load libnccl: ") + dlerror();
        return false;
    }
#define MMBA_SYM(name)                                                                  \
    g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(lib, "nccl" #name));      \
    if (!g_nccl.name) {                                                                  \
        err = "libnccl lacks nccl" #name;                                                \
        return false;                                                                    \
    }
    MMBA_SYM(GetUniqueId)
    MMBA_SYM(CommInitRank)
    MMBA_SYM(CommDestroy)
    MMBA_SYM(AllReduce)
    MMBA_SYM(AllGather)
    MMBA_SYM(Send)
    MMBA_SYM(Recv)
    MMBA_SYM(GroupStart)
    MMBA_SYM(GroupEnd)
    MMBA_SYM(GetErrorString)
#undef MMBA_SYM
    g_nccl.lib = lib;
    return true;
}

// ---------------------------------------------------------------------------------------------
// device arena: one cudaMalloc per problem, carved into 256-byte aligned arrays
// ---------------------------------------------------------------------------------------------
struct Arena {
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

struct Dev {
    // plan
    TileMeta* meta;
    int32_t* tile_cams;
    double* uv;
    // linearisation
    double *J, *res;
    double* ju1;       // [slots][2] J u1 per observation (JV1 -> BACKSUB, subspace Gram sums)
    double *x, *xn, *x0, *camtab, *camtab_n;
    double *U, *Ud, *g, *V, *M, *zg, *dp;
    // n-vectors (camera part first, then the local points)
    double *sinv, *gh, *gn, *tmp;
    // PCG (camera-sized)
    double *y, *Sd, *Pinv, *px, *pr, *pz, *pp, *pq, *pxt, *part, *state;
    double *pose_lam, *pose_suf, *pose_V, *pose_w;   // pose-only adjustment (per-camera eigen data)
    int* flags;
    double* scal;
    double* x_io;      // [6 Nc + 3 Np] parameters in the caller's layout (upload / download staging on the device)
    // explicit reduced camera matrix (rcm.h / rcm.cuh); null when the implicit product is used
    double *Tup, *S, *rcm_b;
    int *up_rowptr, *up_cols, *rc_rowptr, *rc_cols, *rc_rows, *rc_src, *rc_diag;
    int *rc_halo_ptr, *rc_halo_cols, *rc_own;
    uint16_t* rc_lcol;
    RcmSlot* rcm_slots;
    LLLine* rcm_z;
    double* rcm_hist;   // (||r||^2, r.z) of every PCG iterate of the last inner solve (profile & 2)
    long long* dbg;     // [64] phase cycle counters of the instrumented kernels (profile & 4)
};

// One-shot peer all-reduce of the per-iteration Schur product (see xchg_push_kernel)
struct Xchg {
    bool on = false;
    void* base = nullptr;                 // this rank's receive buffer (exported with cudaIpc)
    std::vector<void*> peers;             // peer mappings (nullptr for self)
    double* slots = nullptr;              // [2][nranks][n6]
    unsigned long long* flags = nullptr;  // [2][nranks]
    double** d_peer_slots = nullptr;      // device arrays of per-rank pointers
    unsigned long long** d_peer_flags = nullptr;
    int n6 = 0;       // slot length in doubles (capacity: the buffers are kept across problems)
    unsigned long long seq = 0;
    // small all-reduce (peer_allreduce_small_kernel): LL lines [2 parities][nranks][kPeerSmallMax] in every rank's buffer
    LLLine* ll = nullptr;                 // this rank's area
    LLLine** d_peer_ll = nullptr;         // device array: every rank's area as mapped here
    unsigned llseq = 0;
};

struct Profile {
    int64_t launches[MMBA_K_COUNT] = {0};
    double ms[MMBA_K_COUNT] = {0};
    std::vector<cudaEvent_t> pool;
    std::vector<int> cls;   // class of pair i (events 2i, 2i+1)
    size_t used = 0;
};

}  // namespace

constexpr int kStageSlots = 3;
constexpr size_t kStageSlotBytes = (size_t)24 << 20;

struct mmba_handle {
    mmba_options opt;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ncclComm_t comm = nullptr;
    bool has_problem = false;
    DevPlanner planner;            // device-side plan builder (devplan.h) and its buffers
    DevPlan dp;                    // the current problem's plan: sizes on the host, tables on the device
    std::vector<int32_t> h_point_perm, h_rc_rows, h_rc_cols;   // host copies for the evaluation hooks (downloaded on demand)
    bool rcm_ready = false;        // pattern built and device arrays carved for the current problem
    int rcm_warps = 0, rcm_s_in_smem = 0;
    unsigned rcm_seq = 0;          // sequence numbers handed to the PCG launches of this problem (never reused)
    int hist_cap = 0;              // iterations the history buffer holds (pcg_maxit at set_problem time)
    std::vector<std::vector<double>> pcg_hist;   // per outer iteration: (||r_k||^2, r_k.z_k), k = 0 .. its
    size_t rcm_smem_bytes = 0;
    int64_t Nc = 0, npl = 0, ns = 0, nt = 0, nloc = 0;
    double K[9];
    void* arena = nullptr;
    size_t arena_bytes = 0, arena_cap = 0;
    Dev d;
    TileArgs targs;
    int sm_count = 148;
    size_t smem[M_COUNT] = {0};     // dynamic shared memory of tile_kernel<MODE>
    int grid[M_COUNT] = {0};        // persistent grid of tile_kernel<MODE>: min(tiles, SMs x resident CTAs)
    // pinned staging ring between pageable caller buffers and the device: several host threads fill a slot while
    // the copy engine drains the previous one
    char* stage_buf = nullptr;
    cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr};
    bool stage_busy[kStageSlots] = {false, false, false};
    int stage_next = 0;
    double* h_scal = nullptr;    // pinned S_COUNT
    int* h_flags = nullptr;      // pinned 4
    std::vector<mmba_iter_log> log;
    Profile prof;
    Xchg xchg;
};

struct mmba_plan {
    Plan plan;
};

namespace {

int fail(mmba_handle* h, int code, const std::string& msg) {
    g_thread_error = msg;
    if (h) h->err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, MMBA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(h, MMBA_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));  \
    } while (0)

#define TRY(call)             \
    do {                      \
        int rc_ = (call);     \
        if (rc_ != MMBA_OK) return rc_; \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- profiling -------------------------------------------------------------------------------
void prof_begin(mmba_handle* h, int cls) {
    h->prof.launches[cls]++;
    if (!(h->opt.profile & 1)) return;
    Profile& p = h->prof;
    if (p.used * 2 + 2 > p.pool.size()) {
        if (p.pool.size() >= 2 * 65536) return;
        for (int i = 0; i < 512; ++i) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            p.pool.push_back(e);
        }
    }
    if (p.cls.size() <= p.used) p.cls.resize(p.used + 1);
    p.cls[p.used] = cls;
    cudaEventRecord(p.pool[2 * p.used], h->stream);
}
void prof_end(mmba_handle* h, int) {
    if (!(h->opt.profile & 1)) return;
    Profile& p = h->prof;
    if (p.used * 2 + 2 > p.pool.size()) return;
    cudaEventRecord(p.pool[2 * p.used + 1], h->stream);
    p.used++;
}
void prof_collect(mmba_handle* h) {
    Profile& p = h->prof;
    if (!(h->opt.profile & 1)) return;
    cudaStreamSynchronize(h->stream);
    for (size_t i = 0; i < p.used; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.pool[2 * i], p.pool[2 * i + 1]) == cudaSuccess) p.ms[p.cls[i]] += ms;
    }
    p.used = 0;
}

#define LAUNCH(cls, kern, grid, block, smem, ...)                         \
    do {                                                                  \
        prof_begin(h, cls);                                               \
        kern<<<(grid), (block), (smem), h->stream>>>(__VA_ARGS__);        \
        prof_end(h, cls);                                                 \
    } while (0)

// ---- collectives ------------------------------------------------------------------------------
struct Red {
    double* p;
    size_t n;
    bool is_max;
};

// Sum (or max) the listed device arrays over ranks, in place, as one NCCL group.
int allreduce(mmba_handle* h, std::initializer_list<Red> items) {
    if (h->opt.nranks <= 1) return MMBA_OK;
    // a handful of scalars: one-shot exchange of self-validating lines over NVLink peer memory instead of NCCL
    // (≈30 us per call at 8 GPUs); every rank adds the partials in rank order -> identical results everywhere
    if (h->xchg.on && items.size() <= 2 && h->opt.nranks <= 8) {
        size_t total = 0;
        for (const Red& r : items) total += r.n;
        if (total <= (size_t)kPeerSmallMax) {
            const Red* it = items.begin();
            const Red a = it[0], b = items.size() > 1 ? it[1] : Red{nullptr, 0, false};
            const unsigned seq = ++h->xchg.llseq;
            prof_begin(h, MMBA_K_ALLREDUCE);
            peer_allreduce_small_kernel<<<1, 32 * ((kPeerSmallMax * h->opt.nranks + 31) / 32), 0, h->stream>>>(
                a.p, (int)a.n, a.is_max ? 1 : 0, b.p, (int)b.n, b.is_max ? 1 : 0, h->xchg.d_peer_ll, h->xchg.ll, h->opt.rank,
                h->opt.nranks, seq, h->d.scal + S_PEER_ERR);
            prof_end(h, MMBA_K_ALLREDUCE);
            return MMBA_OK;
        }
    }
    prof_begin(h, MMBA_K_ALLREDUCE);
    NC(g_nccl.GroupStart());
    for (const Red& r : items) {
        if (r.is_max) {
            // non-negative doubles order like their bit patterns
            NC(g_nccl.AllReduce(r.p, r.p, r.n, ncclUint64, ncclMax, h->comm, h->stream));
        } else {
            NC(g_nccl.AllReduce(r.p, r.p, r.n, ncclDouble, ncclSum, h->comm, h->stream));
        }
    }
    NC(g_nccl.GroupEnd());
    prof_end(h, MMBA_K_ALLREDUCE);
    return MMBA_OK;
}

int read_scalars(mmba_handle* h) {
    CU(cudaMemcpyAsync(h->h_scal, h->d.scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->h_scal[S_PEER_ERR] != 0.0) return fail(h, MMBA_ERR_NCCL, "peer all-reduce timed out: a rank stopped participating");
    return MMBA_OK;
}

int zero(mmba_handle* h, double* p, size_t n) {
    CU(cudaMemsetAsync(p, 0, n * sizeof(double), h->stream));
    return MMBA_OK;
}

// ---- memory layout ----------------------------------------------------------------------------
// The plan's tables (tile metadata, camera lists, tile-major pixels, block pattern) live in the planner's buffers;
// the arena holds everything the solve reads and writes.
void carve(mmba_handle* h, Arena& a) {
    const DevPlan& dp = h->dp;
    Dev& d = h->d;
    const size_t Nc = h->Nc, npl = std::max<int64_t>(h->npl, 1), ns = std::max<int64_t>(h->ns, 1), nloc = 6 * Nc + 3 * npl;
    d.meta = dp.meta;
    d.tile_cams = dp.tile_cams;
    d.uv = dp.uvt;
    d.J = a.take<double>(18 * ns);
    d.res = a.take<double>(2 * ns);
    d.ju1 = a.take<double>(2 * ns);
    d.x = a.take<double>(nloc);
    d.xn = a.take<double>(nloc);
    d.x0 = a.take<double>(nloc);
    d.camtab = a.take<double>(kCamTab * Nc);
    d.camtab_n = a.take<double>(kCamTab * Nc);
    d.U = a.take<double>(21 * Nc);
    d.Ud = a.take<double>(6 * Nc);
    d.g = a.take<double>(nloc);
    d.V = a.take<double>(6 * npl);
    d.M = a.take<double>(6 * npl);
    d.zg = a.take<double>(3 * npl);
    d.dp = a.take<double>(3 * npl);
    d.sinv = a.take<double>(nloc);
    d.gh = a.take<double>(nloc);
    d.gn = a.take<double>(nloc);
    d.tmp = a.take<double>(nloc);
    d.y = a.take<double>(6 * Nc);
    d.Sd = a.take<double>(21 * Nc);
    d.Pinv = a.take<double>(21 * Nc);
    d.px = a.take<double>(6 * Nc);
    d.pr = a.take<double>(6 * Nc);
    d.pz = a.take<double>(6 * Nc);
    d.pp = a.take<double>(6 * Nc);
    d.pq = a.take<double>(6 * Nc);
    d.pxt = a.take<double>(6 * Nc);
    d.part = a.take<double>((size_t)P_COUNT * kMaxCamBlocks);
    d.state = a.take<double>(4);
    d.pose_lam = a.take<double>(6 * Nc);
    d.pose_suf = a.take<double>(6 * Nc);
    d.pose_V = a.take<double>(36 * Nc);
    d.pose_w = a.take<double>(6 * Nc);
    d.flags = a.take<int>(4);
    d.scal = a.take<double>(S_COUNT);
    d.x_io = a.take<double>(6 * Nc + 3 * (size_t)dp.n_points);
    d.dbg = a.take<long long>(64);
    if (h->rcm_ready) {
        d.Tup = a.take<double>(36 * (size_t)dp.nnz_up);
        d.S = a.take<double>(36 * (size_t)dp.nnz_full);
        d.rcm_b = a.take<double>(6 * Nc);
        d.up_rowptr = dp.up_rowptr;
        d.up_cols = dp.up_cols;
        d.rc_rowptr = dp.rowptr;
        d.rc_cols = dp.cols;
        d.rc_rows = dp.rows;
        d.rc_src = dp.src;
        d.rc_diag = dp.diag;
        d.rc_halo_ptr = dp.halo_ptr;
        d.rc_halo_cols = dp.halo_cols;
        d.rc_own = dp.own_l;
        d.rc_lcol = dp.lcol;
        d.rcm_slots = a.take<RcmSlot>(2 * (size_t)kRcmMaxCtas);
        d.rcm_z = a.take<LLLine>(2 * 6 * Nc);
        d.rcm_hist = (h->opt.profile & 2) ? a.take<double>(2 * ((size_t)h->opt.pcg_maxit + 1)) : nullptr;
    } else {
        d.Tup = d.S = d.rcm_b = nullptr;
        d.up_rowptr = d.up_cols = d.rc_rowptr = d.rc_cols = d.rc_rows = d.rc_src = d.rc_diag = nullptr;
        d.rc_halo_ptr = d.rc_halo_cols = d.rc_own = nullptr;
        d.rc_lcol = nullptr;
        d.rcm_slots = nullptr;
        d.rcm_z = nullptr;
        d.rcm_hist = nullptr;
    }
}

// ---- host <-> device transfers through the pinned staging ring ------------------------------------
int stage_init(mmba_handle* h) {
    if (h->stage_buf) return MMBA_OK;
    CU(cudaMallocHost(&h->stage_buf, kStageSlots * kStageSlotBytes));
    for (int i = 0; i < kStageSlots; ++i) CU(cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
    return MMBA_OK;
}
// next slot of the ring, free to be written by the host
int stage_acquire(mmba_handle* h, int* idx, char** ptr) {
    TRY(stage_init(h));
    const int i = h->stage_next;
    h->stage_next = (i + 1) % kStageSlots;
    if (h->stage_busy[i]) {
        CU(cudaEventSynchronize(h->stage_ev[i]));
        h->stage_busy[i] = false;
    }
    *idx = i;
    *ptr = h->stage_buf + (size_t)i * kStageSlotBytes;
    return MMBA_OK;
}
int stage_commit(mmba_handle* h, int idx) {
    CU(cudaEventRecord(h->stage_ev[idx], h->stream));
    h->stage_busy[idx] = true;
    return MMBA_OK;
}
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    parallel_ranges((int64_t)bytes, (int64_t)1 << 20, [&](int64_t b, int64_t e, int) {
        std::memcpy(static_cast<char*>(dst) + b, static_cast<const char*>(src) + b, (size_t)(e - b));
    }, 16, true);
}
// pieces of the generic pageable copies: small enough that the DMA of one piece overlaps the host copy of the next
constexpr size_t kCopyPiece = (size_t)8 << 20;
// pageable host -> device; the caller's buffer is fully consumed on return
int h2d(mmba_handle* h, void* dst, const void* src, size_t bytes) {
    for (size_t off = 0; off < bytes; off += kCopyPiece) {
        const size_t n = std::min(kCopyPiece, bytes - off);
        int idx;
        char* slot;
        TRY(stage_acquire(h, &idx, &slot));
        parallel_memcpy(slot, static_cast<const char*>(src) + off, n);
        CU(cudaMemcpyAsync(static_cast<char*>(dst) + off, slot, n, cudaMemcpyHostToDevice, h->stream));
        TRY(stage_commit(h, idx));
    }
    return MMBA_OK;
}
// device -> pageable host; complete on return
int d2h(mmba_handle* h, void* dst, const void* src, size_t bytes) {
    struct Pending { int idx; char* slot; size_t off, n; } pend[kStageSlots] = {};
    int npend = 0;
    auto drain_one = [&]() -> int {
        const Pending q = pend[0];
        CU(cudaEventSynchronize(h->stage_ev[q.idx]));
        h->stage_busy[q.idx] = false;
        parallel_memcpy(static_cast<char*>(dst) + q.off, q.slot, q.n);
        for (int i = 1; i < npend; ++i) pend[i - 1] = pend[i];
        --npend;
        return MMBA_OK;
    };
    for (size_t off = 0; off < bytes; off += kCopyPiece) {
        if (npend == kStageSlots - 1) TRY(drain_one());
        const size_t n = std::min(kCopyPiece, bytes - off);
        int idx;
        char* slot;
        TRY(stage_acquire(h, &idx, &slot));
        CU(cudaMemcpyAsync(slot, static_cast<const char*>(src) + off, n, cudaMemcpyDeviceToHost, h->stream));
        TRY(stage_commit(h, idx));
        pend[npend++] = Pending{idx, slot, off, n};
    }
    while (npend) TRY(drain_one());
    return MMBA_OK;
}

void xchg_release(mmba_handle* h) {
    Xchg& x = h->xchg;
    for (void* p : x.peers)
        if (p) cudaIpcCloseMemHandle(p);
    x.peers.clear();
    if (x.base) cudaFree(x.base);
    if (x.d_peer_slots) cudaFree(x.d_peer_slots);
    if (x.d_peer_flags) cudaFree(x.d_peer_flags);
    if (x.d_peer_ll) cudaFree(x.d_peer_ll);
    x = Xchg();
}

// Allocate this rank's receive buffer, exchange cudaIpc handles through the NCCL communicator and map
// every peer's buffer.  On any failure the exchange stays off and the solve uses ncclAllReduce.
int xchg_setup(mmba_handle* h) {
    Xchg& x = h->xchg;
    const int nr = h->opt.nranks;
    const char* env = getenv("MMBA_PEER_XCHG");
    if (nr <= 1 || (env && env[0] == '0')) return MMBA_OK;
    if (x.on && x.n6 >= 6 * h->Nc) return MMBA_OK;   // mapped buffers of an earlier problem are large enough
    xchg_release(h);
    x.n6 = (int)std::max<int64_t>(6 * h->Nc, 8192);
    const size_t flag_bytes = 256 * ((2 * nr * sizeof(unsigned long long) + 255) / 256);
    const size_t ll_bytes = 256 * ((2 * (size_t)nr * kPeerSmallMax * sizeof(LLLine) + 255) / 256);
    const size_t bytes = flag_bytes + ll_bytes + 2 * (size_t)nr * x.n6 * sizeof(double);
    CU(cudaMalloc(&x.base, bytes));
    CU(cudaMemsetAsync(x.base, 0, bytes, h->stream));
    x.flags = static_cast<unsigned long long*>(x.base);
    x.ll = reinterpret_cast<LLLine*>(static_cast<char*>(x.base) + flag_bytes);
    x.slots = reinterpret_cast<double*>(static_cast<char*>(x.base) + flag_bytes + ll_bytes);
    cudaIpcMemHandle_t mine;
    CU(cudaIpcGetMemHandle(&mine, x.base));
    char *d_send = nullptr, *d_recv = nullptr;
    CU(cudaMalloc(&d_send, sizeof(mine)));
    CU(cudaMalloc(&d_recv, sizeof(mine) * nr));
    CU(cudaMemcpyAsync(d_send, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllGather(d_send, d_recv, sizeof(mine), ncclChar, h->comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all(nr);
    CU(cudaMemcpyAsync(all.data(), d_recv, sizeof(mine) * nr, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d_send);
    cudaFree(d_recv);
    x.peers.assign(nr, nullptr);
    std::vector<double*> ps(nr);
    std::vector<unsigned long long*> pf(nr);
    std::vector<LLLine*> pl(nr);
    bool ok = true;
    for (int r = 0; r < nr; ++r) {
        void* p = x.base;
        if (r != h->opt.rank) {
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
                p = x.base;
            } else {
                x.peers[r] = p;
            }
        }
        pf[r] = static_cast<unsigned long long*>(p);
        pl[r] = reinterpret_cast<LLLine*>(static_cast<char*>(p) + flag_bytes);
        ps[r] = reinterpret_cast<double*>(static_cast<char*>(p) + flag_bytes + ll_bytes);
    }
    // every rank must take the same path: agree on success
    double* d_ok = nullptr;
    CU(cudaMalloc(&d_ok, sizeof(double)));
    const double okv = ok ? 0.0 : 1.0;
    CU(cudaMemcpyAsync(d_ok, &okv, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclDouble, ncclSum, h->comm, h->stream));
    double bad = 0;
    CU(cudaMemcpyAsync(&bad, d_ok, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d_ok);
    if (bad != 0.0) {
        xchg_release(h);
        return MMBA_OK;
    }
    CU(cudaMalloc(&x.d_peer_slots, nr * sizeof(double*)));
    CU(cudaMalloc(&x.d_peer_flags, nr * sizeof(unsigned long long*)));
    CU(cudaMemcpyAsync(x.d_peer_slots, ps.data(), nr * sizeof(double*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(x.d_peer_flags, pf.data(), nr * sizeof(unsigned long long*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMalloc(&x.d_peer_ll, nr * sizeof(LLLine*)));
    CU(cudaMemcpyAsync(x.d_peer_ll, pl.data(), nr * sizeof(LLLine*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    x.llseq = 0;
    x.seq = 0;
    x.on = true;
    return MMBA_OK;
}

// The device arena survives a change of problem (mmba_set_problem reuses it when it is large enough):
// repeated adjustPoints calls do not pay cudaFree/cudaMalloc of hundreds of MB each time.
void release_problem(mmba_handle* h) {
    h->has_problem = false;
}

// caller's x (cameras | points in caller order) -> device x (cameras | local points, internal order)
int put_x(mmba_handle* h, const double* x, double* dst) {
    const DevPlan& dp = h->dp;
    TRY(h2d(h, h->d.x_io, x, (size_t)(6 * h->Nc + 3 * dp.n_points) * sizeof(double)));
    devplan_gather_x(h->d.x_io, dst, dp.point_perm, h->Nc, dp.pt_begin, h->npl, h->stream);
    return MMBA_OK;
}

// the same with the camera and the point parameters in separate host arrays (no packed copy on the host)
int put_x_split(mmba_handle* h, const double* cams, const double* points, double* dst) {
    const DevPlan& dp = h->dp;
    TRY(h2d(h, h->d.x_io, cams, (size_t)(6 * h->Nc) * sizeof(double)));
    TRY(h2d(h, h->d.x_io + 6 * h->Nc, points, (size_t)(3 * dp.n_points) * sizeof(double)));
    devplan_gather_x(h->d.x_io, dst, dp.point_perm, h->Nc, dp.pt_begin, h->npl, h->stream);
    return MMBA_OK;
}

// device n-vector -> caller's layout (x, or cameras and points separately when x == nullptr).  With nranks > 1 the
// point part is completed over ranks.
int get_x(mmba_handle* h, const double* src, double* x, double* cams = nullptr, double* points = nullptr) {
    const DevPlan& dp = h->dp;
    const size_t n_total = (size_t)(6 * h->Nc + 3 * dp.n_points);
    if (h->opt.nranks > 1) CU(cudaMemsetAsync(h->d.x_io + 6 * h->Nc, 0, 3 * (size_t)dp.n_points * sizeof(double), h->stream));
    devplan_scatter_x(src, h->d.x_io, dp.point_perm, h->Nc, dp.pt_begin, h->npl, true, h->stream);
    if (h->opt.nranks > 1) TRY(allreduce(h, {{h->d.x_io + 6 * h->Nc, (size_t)(3 * dp.n_points), false}}));
    if (x) return d2h(h, x, h->d.x_io, n_total * sizeof(double));
    TRY(d2h(h, cams, h->d.x_io, (size_t)(6 * h->Nc) * sizeof(double)));
    return d2h(h, points, h->d.x_io + 6 * h->Nc, (size_t)(3 * dp.n_points) * sizeof(double));
}

// tile-major rows [row0, row0 + rows) of src ([tile][src_rows][256]) -> caller-ordered (n_obs, rows) row-major;
// entries of observations held by other ranks are zero
int get_slots(mmba_handle* h, const double* src, int src_rows, int row0, int rows, double* out) {
    const DevPlan& dp = h->dp;
    const size_t n = (size_t)dp.n_obs * rows;
    double* tmp = nullptr;
    CU(cudaMalloc(&tmp, std::max<size_t>(n, 1) * sizeof(double)));
    cudaMemsetAsync(tmp, 0, n * sizeof(double), h->stream);
    devplan_scatter_slots(src, src_rows, row0, rows, dp.slot_obs, h->ns, tmp, h->stream);
    int rc = d2h(h, out, tmp, n * sizeof(double));
    cudaFree(tmp);
    return rc;
}

// Jt [tile][18][256] -> caller-ordered Jc (n_obs,2,6) and Jp (n_obs,2,3)
int get_jacobian_slots(mmba_handle* h, double* Jc, double* Jp) {
    TRY(get_slots(h, h->d.J, kJRows, 0, 12, Jc));
    return get_slots(h, h->d.J, kJRows, 12, 6, Jp);
}

// host copy of the internal point order (evaluation hooks only)
int host_point_perm(mmba_handle* h) {
    if ((int64_t)h->h_point_perm.size() == h->dp.n_points) return MMBA_OK;
    h->h_point_perm.resize(h->dp.n_points);
    CU(cudaMemcpyAsync(h->h_point_perm.data(), h->dp.point_perm, h->dp.n_points * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MMBA_OK;
}

// ---- device phases ----------------------------------------------------------------------------
PcgVecs pcg_vecs(mmba_handle* h) {
    Dev& d = h->d;
    PcgVecs P;
    P.gc = d.g;
    P.sinv = d.sinv;
    P.y = d.y;
    P.Sd = d.Sd;
    P.Pinv = d.Pinv;
    P.x = d.px;
    P.r = d.pr;
    P.z = d.pz;
    P.p = d.pp;
    P.q = d.pq;
    P.xt = d.pxt;
    P.part = d.part;
    P.flags = d.flags;
    P.state = d.state;
    P.n_cams = (int)h->Nc;
    P.xslots = h->xchg.slots;
    P.xflags = h->xchg.flags;
    P.peer_slots = h->xchg.d_peer_slots;
    P.peer_flags = h->xchg.d_peer_flags;
    P.rank = h->opt.rank;
    P.nranks = h->xchg.on ? h->opt.nranks : 1;
    P.n6 = h->xchg.n6;
    return P;
}

int cam_prep(mmba_handle* h, const double* x, double* camtab) {
    LAUNCH(MMBA_K_CAMPREP, cam_prep_kernel, cdiv(h->Nc, 128), 128, 0, x, camtab, (int)h->Nc);
    return MMBA_OK;
}

// `overlap`: programmatic dependent launch — the grid may be scheduled while the previous kernel in the
// stream drains; every kernel starts with griddepcontrol.wait, so only the launch latency overlaps.
template <int MODE>
int launch_tile(mmba_handle* h, int cls, const ModeArgs& P, bool overlap = false) {
    if (!h->nt) return MMBA_OK;
    prof_begin(h, cls);
    if (overlap) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(h->grid[MODE]);
        cfg.blockDim = dim3(Traits<MODE>::kThreads);
        cfg.dynamicSmemBytes = h->smem[MODE];
        cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, tile_kernel<MODE>, h->targs, P));
    } else {
        tile_kernel<MODE><<<h->grid[MODE], Traits<MODE>::kThreads, h->smem[MODE], h->stream>>>(h->targs, P);
    }
    prof_end(h, cls);
    return MMBA_OK;
}

// residuals, Jacobian blocks, normal-equation blocks and cost at d.x
// (full_U: additionally the complete 6x6 camera blocks, for the evaluation hook)
int linearise(mmba_handle* h, bool full_U = false) {
    Dev& d = h->d;
    TRY(cam_prep(h, d.x, d.camtab));
    TRY(zero(h, d.Ud, 6 * h->Nc));
    TRY(zero(h, d.g, 6 * h->Nc));
    TRY(zero(h, d.scal + S_COST, 1));
    ModeArgs P{};
    P.Jw = d.J;
    P.res = d.res;
    P.cam0 = d.camtab;
    P.ptA = d.x + 6 * h->Nc;
    P.Ud = d.Ud;
    P.gc = d.g;
    P.V = d.V;
    P.gp = d.g + 6 * h->Nc;
    P.scal = d.scal;
    if (full_U) {
        TRY(zero(h, d.U, 21 * h->Nc));
        P.U = d.U;
        TRY(launch_tile<M_BUILD_FULL>(h, MMBA_K_BUILD, P));
        TRY(allreduce(h, {{d.U, (size_t)(21 * h->Nc), false}}));
    } else {
        TRY(launch_tile<M_BUILD>(h, MMBA_K_BUILD, P));
    }
    TRY(allreduce(h, {{d.Ud, (size_t)(6 * h->Nc), false}, {d.g, (size_t)(6 * h->Nc), false}, {d.scal + S_COST, 1, false}}));
    CU(cudaGetLastError());
    return MMBA_OK;
}

// scale_inv (monotone), g_h, ||g_h||^2, ||x*scale_inv||^2, ||x||^2, ||g||_inf
int scale_and_grad(mmba_handle* h, bool first) {
    Dev& d = h->d;
    const int lead = h->opt.rank == 0;
    TRY(zero(h, d.scal + S_GH2, 4));
    auto sg_cam = scale_grad_kernel<6, false>;
    auto sg_pt = scale_grad_kernel<3, true>;
    const int cap = 8 * h->sm_count;      // grid-stride kernels: a few CTAs per SM, one atomic per CTA and result
    // (also leaves u1 = d o g_h in d.tmp)
    LAUNCH(MMBA_K_VEC, sg_cam, std::min(cap, cdiv(6 * h->Nc, 256)), 256, 0, d.Ud, d.g, d.x, d.sinv, d.gh, d.tmp, (int)first,
           6 * h->Nc, d.scal, lead);
    if (h->npl)
        LAUNCH(MMBA_K_VEC, sg_pt, std::min(cap, cdiv(3 * h->npl, 256)), 256, 0, d.V, d.g + 6 * h->Nc, d.x + 6 * h->Nc,
               d.sinv + 6 * h->Nc, d.gh + 6 * h->Nc, d.tmp + 6 * h->Nc, (int)first, 3 * h->npl, d.scal, 1);
    TRY(allreduce(h, {{d.scal + S_GH2, 3, false}, {d.scal + S_GINF, 1, true}}));
    return MMBA_OK;
}

// ||J v||^2 for the unscaled n-vector v (device, internal layout) -> scal[S_JV00]
int jv1(mmba_handle* h, const double* v) {
    Dev& d = h->d;
    TRY(zero(h, d.scal + S_JV00, 3));
    ModeArgs P{};
    P.J = d.J;
    P.cam0 = v;
    P.ptA = v + 6 * h->Nc;
    P.scal = d.scal;
    P.aux_w = d.ju1;
    TRY(launch_tile<M_JV1>(h, MMBA_K_JV, P));
    TRY(allreduce(h, {{d.scal + S_JV00, 3, false}}));
    return MMBA_OK;
}

ModeArgs matvec_args(mmba_handle* h) {
    Dev& d = h->d;
    ModeArgs P{};
    P.J = d.J;
    P.cam0 = d.pxt;
    P.ptA = d.M;
    P.y = d.y;
    P.done = d.flags;
    return P;
}

int schur_matvec(mmba_handle* h) {
    Dev& d = h->d;
    TRY(launch_tile<M_MATVEC>(h, MMBA_K_MATVEC, matvec_args(h), !(h->opt.profile & 1)));
    TRY(allreduce(h, {{d.y, (size_t)(6 * h->Nc), false}}));
    return MMBA_OK;
}

ModeArgs rhs_args(mmba_handle* h) {
    Dev& d = h->d;
    ModeArgs P{};
    P.J = d.J;
    P.ptA = d.M;
    P.ptB = d.zg;
    P.y = d.y;
    P.Sd = d.Sd;
    return P;
}

ModeArgs backsub_args(mmba_handle* h) {
    Dev& d = h->d;
    ModeArgs P{};
    P.J = d.J;
    P.cam0 = d.pxt;
    P.ptA = d.M;
    P.ptB = d.g + 6 * h->Nc;
    P.dp = d.dp;
    P.aux = d.ju1;       // J u1 of the preceding JV1 pass: the pass also leaves (J u1).(J u2), ||J u2||^2 in S_JV01, S_JV11
    P.scal = d.scal;
    return P;
}

// damped Gauss-Newton step in scaled variables: (D J^T J D + reg I) p = D g.
// Leaves the camera part in d.px (scaled) and the unscaled point part in d.dp.
bool use_rcm(const mmba_handle* h) { return h->rcm_ready && h->opt.schur_mode != MMBA_SCHUR_IMPLICIT; }

ModeArgs sbuild_args(mmba_handle* h) {
    Dev& d = h->d;
    ModeArgs P{};
    P.J = d.J;
    P.ptA = d.M;
    P.ptB = d.zg;
    P.y = d.y;
    P.Tup = d.Tup;
    P.up_rowptr = d.up_rowptr;
    P.up_cols = d.up_cols;
    return P;
}

RcmPcgArgs rcm_pcg_args(mmba_handle* h, double f2) {
    Dev& d = h->d;
    RcmPcgArgs A{};
    A.S = d.S;
    A.rowptr = d.rc_rowptr;
    A.lcol = d.rc_lcol;
    A.halo_ptr = d.rc_halo_ptr;
    A.halo_cols = d.rc_halo_cols;
    A.own_l = d.rc_own;
    A.Pinv = d.Pinv;
    A.b = d.rcm_b;
    A.x = d.px;
    A.z = d.rcm_z;
    A.slots = d.rcm_slots;
    A.flags = d.flags;
    A.state = d.state;
    A.n_cams = (int)h->Nc;
    A.maxit = h->opt.pcg_maxit;
    A.cpc = h->dp.cpc;
    A.nblk_max = h->dp.nblk_max;
    A.nh_max = h->dp.nh_max;
    A.s_in_smem = h->rcm_s_in_smem;
    A.rtol2 = h->opt.pcg_rtol * h->opt.pcg_rtol;
    A.atol2f = h->opt.pcg_atol * h->opt.pcg_atol * f2;
    A.ktol2f = h->opt.pcg_ktol * h->opt.pcg_ktol * f2;
    A.hist = (d.rcm_hist && h->opt.pcg_maxit <= h->hist_cap) ? d.rcm_hist : nullptr;
    A.phase = (h->opt.profile & 4) ? d.dbg : nullptr;
    {
        const char* f = getenv("MMBA_FAULT_PCG_CTA");
        A.fault_cta = f ? atoi(f) : -1;
    }
    A.seq0 = h->rcm_seq;
    h->rcm_seq += (unsigned)h->opt.pcg_maxit + 2u;
    return A;
}

// S-build pass: unscaled upper blocks of the reduced camera matrix + Schur right-hand side (summed over ranks)
int rcm_build(mmba_handle* h) {
    Dev& d = h->d;
    TRY(zero(h, d.y, 6 * h->Nc));
    TRY(zero(h, d.Tup, 36 * (size_t)h->dp.nnz_up));
    TRY(launch_tile<M_SBUILD>(h, MMBA_K_SBUILD, sbuild_args(h)));
    TRY(allreduce(h, {{d.y, (size_t)(6 * h->Nc), false}, {d.Tup, 36 * (size_t)h->dp.nnz_up, false}}));
    return MMBA_OK;
}

// scaled full-pattern blocks, block-Jacobi preconditioner and right-hand side
int rcm_finalize(mmba_handle* h, double reg) {
    Dev& d = h->d;
    const int64_t n_entries = 36 * h->dp.nnz_full;
    LAUNCH(MMBA_K_VEC, rcm_finalize_kernel, cdiv(n_entries, 256), 256, 0, d.Tup, d.rc_rows, d.rc_cols, d.rc_src, d.sinv, reg, d.S,
           n_entries);
    LAUNCH(MMBA_K_VEC, rcm_prepare_kernel, cdiv(h->Nc, kCamBlock), kCamBlock, 0, d.S, d.rc_diag, d.g, d.y, d.sinv, d.Pinv,
           d.rcm_b, (int)h->Nc);
    return MMBA_OK;
}

// the whole PCG solve: one cooperative launch
int rcm_pcg(mmba_handle* h, double f2) {
    Dev& d = h->d;
    // neither the exchange lines nor the slots are cleared: every launch uses fresh sequence numbers (seq0)
    CU(cudaMemsetAsync(d.flags, 0, 3 * sizeof(int), h->stream));
    RcmPcgArgs A = rcm_pcg_args(h, f2);
    void* args[] = {&A};
    prof_begin(h, MMBA_K_PCG);
    CU(cudaLaunchCooperativeKernel((const void*)rcm_pcg_kernel, dim3(h->dp.n_ctas), dim3(32 * h->rcm_warps), args,
                                   h->rcm_smem_bytes, h->stream));
    prof_end(h, MMBA_K_PCG);
    return MMBA_OK;
}

// f2 = ||f||^2 at the linearisation point (0 switches the LSMR-like absolute stopping rule off)
int gn_step(mmba_handle* h, double reg, double f2, int64_t* its_out, double* relres_out) {
    Dev& d = h->d;
    const int camblocks = cdiv(h->Nc, kCamBlock);
    PcgVecs P = pcg_vecs(h);
    if (h->npl)
        LAUNCH(MMBA_K_PTINV, point_invert_kernel, cdiv(h->npl, 256), 256, 0, d.V, d.g + 6 * h->Nc, d.sinv + 6 * h->Nc, reg,
               d.M, d.zg, h->npl);
    if (use_rcm(h)) {
        // explicit reduced camera matrix: one streaming pass builds S, the PCG runs on-chip in one kernel
        TRY(rcm_build(h));
        TRY(rcm_finalize(h, reg));
        TRY(rcm_pcg(h, f2));
        CU(cudaMemcpyAsync(h->h_flags, d.flags, 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        double st[4] = {0, 0, 0, 0};
        if (relres_out) CU(cudaMemcpyAsync(st, d.state, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (h->h_flags[2]) return fail(h, MMBA_ERR_CUDA, "reduced-system PCG: grid barrier timed out");
        if (its_out) *its_out = h->h_flags[1];
        if (relres_out) *relres_out = (h->h_flags[1] > 0 && st[1] > 0) ? std::sqrt(st[2] / st[1]) : 0.0;
        if (d.rcm_hist && h->opt.pcg_maxit <= h->hist_cap) {
            std::vector<double> hist(2 * ((size_t)h->h_flags[1] + 1));
            CU(cudaMemcpyAsync(hist.data(), d.rcm_hist, hist.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            h->pcg_hist.push_back(std::move(hist));
        }
        LAUNCH(MMBA_K_VEC, unscale_kernel, cdiv(6 * h->Nc, 256), 256, 0, d.px, d.sinv, 1.0, d.pxt, 6 * h->Nc);
        TRY(launch_tile<M_BACKSUB>(h, MMBA_K_BACKSUB, backsub_args(h)));
        CU(cudaGetLastError());
        return MMBA_OK;
    }
    TRY(zero(h, d.y, 6 * h->Nc));
    TRY(zero(h, d.Sd, 21 * h->Nc));
    TRY(launch_tile<M_RHS>(h, MMBA_K_RHS, rhs_args(h)));
    TRY(allreduce(h, {{d.y, (size_t)(6 * h->Nc), false}, {d.Sd, (size_t)(21 * h->Nc), false}}));
    LAUNCH(MMBA_K_VEC, pcg_init_kernel, camblocks, kCamBlock, 0, P, reg);

    const double rtol2 = h->opt.pcg_rtol * h->opt.pcg_rtol;
    const double atol2f = h->opt.pcg_atol * h->opt.pcg_atol * f2;
    const double ktol2f = h->opt.pcg_ktol * h->opt.pcg_ktol * f2;
    const int maxit = h->opt.pcg_maxit;
    int it = 0, done = 0;
    int chunk = 8;
    while (it < maxit && !done) {
        const int stop = std::min(maxit, it + chunk);
        for (; it < stop; ++it) {
            int parity = 0;
            unsigned long long seq = 0;
            if (h->xchg.on) {
                // MATVEC + peer push; the all-reduce completes inside pcg_update
                TRY(launch_tile<M_MATVEC>(h, MMBA_K_MATVEC, matvec_args(h), !(h->opt.profile & 1)));
                seq = ++h->xchg.seq;
                parity = (int)(seq & 1);
            } else {
                TRY(schur_matvec(h));
            }
            {
                const unsigned csize = h->Nc <= kPcgThreads ? 1u : (unsigned)kPcgCluster;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(csize);
                cfg.blockDim = dim3(kPcgThreads);
                cfg.dynamicSmemBytes = 0;
                cfg.stream = h->stream;
                cudaLaunchAttribute attr[2];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = csize;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[1].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr;
                cfg.numAttrs = (h->opt.profile & 1) ? 1 : 2;   // profile mode brackets launches with events: no overlap
                prof_begin(h, MMBA_K_VEC);
                CU(cudaLaunchKernelEx(&cfg, pcg_update_kernel, P, reg, it, rtol2, atol2f, ktol2f, camblocks, parity, seq));
                prof_end(h, MMBA_K_VEC);
            }
        }
        CU(cudaMemcpyAsync(h->h_flags, d.flags, 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (h->h_flags[2]) return fail(h, MMBA_ERR_NCCL, "peer exchange timed out: a rank stopped participating");
        done = h->h_flags[0];
        chunk = 16;
    }
    const int64_t its = done ? h->h_flags[1] : it;
    if (its_out) *its_out = its;
    if (relres_out) {
        double st[4];
        CU(cudaMemcpyAsync(st, d.state, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        *relres_out = (its > 0 && st[1] > 0) ? std::sqrt(st[2] / st[1]) : 0.0;
    }
    // back-substitution with the unscaled camera step
    LAUNCH(MMBA_K_VEC, unscale_kernel, cdiv(6 * h->Nc, 256), 256, 0, d.px, d.sinv, 1.0, d.pxt, 6 * h->Nc);
    TRY(launch_tile<M_BACKSUB>(h, MMBA_K_BACKSUB, backsub_args(h)));
    CU(cudaGetLastError());
    return MMBA_OK;
}

int trial_cost(mmba_handle* h, const double* x, double* camtab) {
    Dev& d = h->d;
    TRY(cam_prep(h, x, camtab));
    TRY(zero(h, d.scal + S_COST_NEW, 1));
    ModeArgs P{};
    P.cam0 = camtab;
    P.ptA = x + 6 * h->Nc;
    P.cost = d.scal + S_COST_NEW;
    TRY(launch_tile<M_RESID>(h, MMBA_K_RESID, P));
    TRY(allreduce(h, {{d.scal + S_COST_NEW, 1, false}}));
    return MMBA_OK;
}

double ginf_of(const double* scal) {
    double v;
    std::memcpy(&v, scal + S_GINF, sizeof(double));
    return v;
}

// ---- the TRF outer loop (trf.py:415-587) --------------------------------------------------------
int run_trf(mmba_handle* h, mmba_result* out) {
    Dev& d = h->d;
    const mmba_options& o = h->opt;
    const int lead = o.rank == 0;
    const int64_t n_total = 6 * h->Nc + 3 * h->dp.n_points;
    const int64_t max_nfev = o.max_nfev > 0 ? o.max_nfev : 100 * n_total;
    const int64_t nloc = h->nloc, ncam = 6 * h->Nc, npt = 3 * h->npl;
    const int gv = cdiv(nloc, 256), gc_ = cdiv(ncam, 256), gp_ = std::max(1, cdiv(npt, 256));
    h->log.clear();
    h->pcg_hist.clear();

    TRY(linearise(h));
    TRY(scale_and_grad(h, true));
    TRY(read_scalars(h));
    double cost = 0.5 * h->h_scal[S_COST];
    if (!std::isfinite(cost)) return fail(h, MMBA_ERR_NONFINITE, "Residuals are not finite in the initial point.");
    double gh2 = h->h_scal[S_GH2], x_norm = std::sqrt(h->h_scal[S_X2]), g_norm = ginf_of(h->h_scal);
    double Delta = std::sqrt(h->h_scal[S_XSI2]);
    if (Delta == 0) Delta = 1.0;
    out->initial_cost = cost;
    int64_t nfev = 1, njev = 1, nit = 0, pcg_total = 0;
    int status = -1;   // -1 = running
    double step_norm = 0, actual = 0;
    bool have_step = false;

    while (true) {
        if (g_norm < o.gtol) status = 1;
        {
            mmba_iter_log row;
            row.iteration = nit;
            row.nfev = nfev;
            row.cost = cost;
            row.cost_reduction = have_step ? actual : NAN;
            row.step_norm = have_step ? step_norm : NAN;
            row.optimality = g_norm;
            row.reg = NAN;
            row.delta = Delta;
            row.pcg_iterations = 0;
            h->log.push_back(row);
        }
        if (status != -1 || nfev >= max_nfev) break;

        // Cauchy-step regulariser (trf.py:488-492): a = 0.5 ||J_h g_h||^2, b = -||g_h||^2
        // (u1 = d o g_h is in d.tmp: scale_and_grad)
        TRY(jv1(h, d.tmp));
        TRY(read_scalars(h));
        const double qa = 0.5 * h->h_scal[S_JV00], qb = -gh2;
        const double gh_norm = std::sqrt(gh2);
        double t_best, ag;
        min_quadratic_1d(qa, qb, 0.0, Delta / gh_norm, &t_best, &ag);
        const double reg = -ag / (Delta * Delta);

        // damped Gauss-Newton direction (replaces lsmr(J_h, f, damp=sqrt(reg)), trf.py:494-495)
        int64_t its = 0;
        TRY(gn_step(h, reg, 2.0 * cost, &its, nullptr));
        pcg_total += its;
        h->log.back().reg = reg;
        h->log.back().pcg_iterations = its;

        // S = orth[g_h, gn_h], B_S = (J_h S)^T (J_h S), g_S = S^T g_h   (trf.py:496-500).  The orthonormal basis is
        // s1 = g_h / ||g_h||, s2 = (gn_h - c1 s1) / n2 with c1 = s1.gn_h, n2^2 = ||gn_h||^2 - c1^2; it is never formed:
        // with u1 = d o g_h (d.tmp) and u2 = d o gn_h (d.gn) the products J_h s1 = J u1 / ||g_h|| and
        // J_h s2 = (J u2 - kappa J u1) / n2, kappa = c1 / ||g_h||, follow from the Gram matrix of J [u1 u2]
        // and five dot products (one vector pass): one host round trip instead of three.  The Gram matrix needs no
        // pass of its own: ||J u1||^2 is the Cauchy step's (JV1, which also stores J u1 per observation) and the
        // back-substitution pass, which completes u2, adds (J u1).(J u2) and ||J u2||^2 on its way.
        TRY(zero(h, d.scal + S_DOT0, 10));
        LAUNCH(MMBA_K_VEC, subspace_dots_kernel, std::min(gv, 8 * h->sm_count), 256, 0, d.gh, d.tmp, d.sinv, d.px, d.pxt, d.dp, d.gn,
               ncam, nloc, d.scal, lead);
        TRY(allreduce(h, {{d.scal + S_DOT0, 5, false}, {d.scal + S_JV01, 2, false}}));
        TRY(read_scalars(h));
        const double D0 = h->h_scal[S_DOT0], D1 = h->h_scal[S_DOT1];
        const double U11 = h->h_scal[S_DOT2], U12 = h->h_scal[S_DOT3], U22 = h->h_scal[S_DOT4];
        const double G11 = h->h_scal[S_JV00], G12 = h->h_scal[S_JV01], G22 = h->h_scal[S_JV11];
        const double inv_gh = gh_norm > 0 ? 1.0 / gh_norm : 0.0;
        const double n2sq = D1 - D0 * D0 * inv_gh * inv_gh;         // ||s2||^2 before normalisation
        // second basis vector is s2 / n2; degenerate (gn_h parallel to g_h up to rounding, or not finite after a PCG
        // breakdown) -> 1-D problem along s1: every term that carries gn_h is switched off, not multiplied by zero
        const bool two_d = std::isfinite(n2sq) && std::isfinite(D0) && std::isfinite(D1) && n2sq > 1e-24 * D1 && D1 > 0.0 &&
                           std::isfinite(G12) && std::isfinite(G22) && std::isfinite(U12) && std::isfinite(U22);
        const double kappa = two_d ? D0 * inv_gh * inv_gh : 0.0;    // c1 / ||g_h||
        const double i2 = two_d ? 1.0 / std::sqrt(n2sq) : 0.0;
        double B[3] = {G11 * inv_gh * inv_gh, two_d ? (G12 - kappa * G11) * inv_gh * i2 : 0.0,
                       two_d ? (G22 - 2.0 * kappa * G12 + kappa * kappa * G11) * i2 * i2 : 1.0};
        double gS[2] = {gh_norm, 0.0};                              // S^T g_h = (||g_h||, 0)
        const double vv00 = U11 * inv_gh * inv_gh, vv01 = two_d ? (U12 - kappa * U11) * inv_gh * i2 : 0.0,
                     vv11 = two_d ? (U22 - 2.0 * kappa * U12 + kappa * kappa * U11) * i2 * i2 : 0.0;
        if (i2 == 0.0) {
            B[1] = 0.0;
            B[2] = 1.0;
        }

        actual = -1;
        double cost_new = cost;
        while (actual <= 0 && nfev < max_nfev) {
            double p[2];
            bool newton;
            tr2d(B, gS, Delta, p, &newton);
            if (i2 == 0.0) p[1] = 0.0;
            const double predicted = -(0.5 * (B[0] * p[0] * p[0] + 2 * B[1] * p[0] * p[1] + B[2] * p[1] * p[1]) +
                                       gS[0] * p[0] + gS[1] * p[1]);
            // x + d o (p0 s1 + p1 s2) = x + (p0 / ||g_h|| - p1 kappa / n2) u1 + (p1 / n2) u2
            LAUNCH(MMBA_K_VEC, trial_kernel, gv, 256, 0, d.x, d.tmp, d.gn, p[0] * inv_gh - p[1] * kappa * i2, p[1] * i2, d.xn, nloc);
            TRY(trial_cost(h, d.xn, d.camtab_n));
            TRY(read_scalars(h));
            nfev++;
            const double step_h_norm = std::hypot(p[0], p[1]);
            const double f2 = h->h_scal[S_COST_NEW];
            if (!std::isfinite(f2)) {
                Delta = 0.25 * step_h_norm;
                continue;
            }
            cost_new = 0.5 * f2;
            actual = cost - cost_new;
            double Delta_new, ratio;
            update_tr_radius(Delta, actual, predicted, step_h_norm, step_h_norm > 0.95 * Delta, &Delta_new, &ratio);
            step_norm = std::sqrt(std::max(0.0, vv00 * p[0] * p[0] + 2 * vv01 * p[0] * p[1] + vv11 * p[1] * p[1]));
            have_step = true;
            const int term = check_termination(actual, cost, step_norm, x_norm, ratio, o.ftol, o.xtol);
            if (term) {
                status = term;
                break;
            }
            Delta = Delta_new;
        }
        if (actual > 0) {
            std::swap(d.x, d.xn);
            cost = cost_new;
            TRY(linearise(h));
            njev++;
            TRY(scale_and_grad(h, false));
            TRY(read_scalars(h));
            gh2 = h->h_scal[S_GH2];
            x_norm = std::sqrt(h->h_scal[S_X2]);
            g_norm = ginf_of(h->h_scal);
        } else {
            step_norm = 0;
            actual = 0;
            have_step = true;
        }
        nit++;
    }
    if (status == -1) status = 0;
    out->cost = cost;
    out->optimality = g_norm;
    out->nfev = nfev;
    out->njev = njev;
    out->nit = nit;
    out->status = status;
    out->reserved = 0;
    out->pcg_iterations = pcg_total;
    return MMBA_OK;
}

template <int MODE>
int configure_mode(mmba_handle* h) {
    SmemLayout L = smem_layout<MODE>(h->targs.max_cams, h->targs.max_pts, h->targs.ytab_cams);
    if (MODE == M_SBUILD) {
        // three stages unless the tiles' point payloads are large (tiles of single-observation points)
        h->targs.sb_stages = L.total <= 227 * 1024 ? Traits<MODE>::kStages : 2;
        L = smem_layout<MODE>(h->targs.max_cams, h->targs.max_pts, h->targs.ytab_cams, h->targs.sb_stages);
    }
    h->smem[MODE] = (size_t)L.total;
    if (L.total > 227 * 1024) return fail(h, MMBA_ERR_NOMEM, "tile_kernel needs " + std::to_string(L.total) + " bytes of shared memory");
    CU(cudaFuncSetAttribute(tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_kernel<MODE>, Traits<MODE>::kThreads, (size_t)L.total));
    if (occ < 1) return fail(h, MMBA_ERR_CUDA, "tile_kernel does not fit on an SM");
    h->grid[MODE] = (int)std::max<int64_t>(1, std::min<int64_t>(h->nt, (int64_t)h->sm_count * occ));
    return MMBA_OK;
}

int configure_kernels(mmba_handle* h) {
    TRY(configure_mode<M_BUILD>(h));
    TRY(configure_mode<M_RESID>(h));
    TRY(configure_mode<M_RESID_STORE>(h));
    TRY(configure_mode<M_MATVEC>(h));
    TRY(configure_mode<M_RHS>(h));
    TRY(configure_mode<M_BACKSUB>(h));
    TRY(configure_mode<M_JV1>(h));
    TRY(configure_mode<M_JV2>(h));
    TRY(configure_mode<M_BUILD_FULL>(h));
    if (h->rcm_ready) {
        TRY(configure_mode<M_SBUILD>(h));
        // PCG grid (rcm_part): contiguous camera ranges, at most one CTA per SM (all co-resident), one warp per
        // camera of the range.  The CTA's rows of S stay in shared memory when they fit, else they are re-read
        // from L2 every iteration.
        const DevPlan& pt = h->dp;
        h->rcm_warps = kRcmPcgThreads / 32;
        h->rcm_s_in_smem = rcm_smem(pt.cpc, pt.nblk_max, pt.nh_max, 1, pt.n_ctas).total <= 200 * 1024 ? 1 : 0;
        const RcmSmem rl = rcm_smem(pt.cpc, pt.nblk_max, pt.nh_max, h->rcm_s_in_smem, pt.n_ctas);
        if (rl.total > 200 * 1024) return fail(h, MMBA_ERR_NOMEM, "reduced-system PCG: too many cameras per CTA");
        h->rcm_smem_bytes = (size_t)rl.total;
        CU(cudaFuncSetAttribute(rcm_pcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rl.total));
    }
    return MMBA_OK;
}

// ---- pose-only TRF (adjustPose, bundleAdjuster.py:232-241): dense least_squares defaults ----------
// method='trf', tr_solver='exact', x_scale=1.0, ftol from the call site; the points are constants.
// Same outer loop as run_trf (trf.py:415-587) with scale = 1 and the exact trust-region step of
// solve_lsq_trust_region (common.py:60-164) evaluated per camera block on the device.
int run_trf_pose(mmba_handle* h, mmba_result* out) {
    Dev& d = h->d;
    const mmba_options& o = h->opt;
    const int64_t ncam = 6 * h->Nc, nloc = h->nloc;
    const int64_t max_nfev = o.max_nfev > 0 ? o.max_nfev : 100 * ncam;
    const double m_rows = 2.0 * (double)h->dp.n_obs;
    h->log.clear();
    auto sg_cam = scale_grad_kernel<6, false>;

    auto grad_stats = [&]() -> int {
        // ||g||_inf and ||x||^2 over the camera parameters (the variables of this problem)
        TRY(zero(h, d.scal + S_GH2, 4));
        LAUNCH(MMBA_K_VEC, sg_cam, std::min(8 * h->sm_count, cdiv(ncam, 256)), 256, 0, d.Ud, d.g, d.x, d.sinv, d.gh, (double*)nullptr, 1, ncam,
               d.scal, 1);
        return MMBA_OK;
    };
    TRY(linearise(h, true));
    TRY(grad_stats());
    TRY(read_scalars(h));
    double cost = 0.5 * h->h_scal[S_COST];
    if (!std::isfinite(cost)) return fail(h, MMBA_ERR_NONFINITE, "Residuals are not finite in the initial point.");
    double x_norm = std::sqrt(h->h_scal[S_X2]), g_norm = ginf_of(h->h_scal);
    double Delta = x_norm;   // ||x0 * scale_inv|| with scale_inv = 1 (trf.py:443-445)
    if (Delta == 0) Delta = 1.0;
    out->initial_cost = cost;
    int64_t nfev = 1, njev = 1, nit = 0;
    int status = -1;
    double step_norm = 0, actual = 0, alpha = 0.0;
    bool have_step = false;
    while (true) {
        if (g_norm < o.gtol) status = 1;
        {
            mmba_iter_log row;
            row.iteration = nit;
            row.nfev = nfev;
            row.cost = cost;
            row.cost_reduction = have_step ? actual : NAN;
            row.step_norm = have_step ? step_norm : NAN;
            row.optimality = g_norm;
            row.reg = alpha;
            row.delta = Delta;
            row.pcg_iterations = 0;
            h->log.push_back(row);
        }
        if (status != -1 || nfev >= max_nfev) break;
        LAUNCH(MMBA_K_VEC, pose_eig_kernel, cdiv(h->Nc, 128), 128, 0, d.U, d.g, d.pose_lam, d.pose_suf, d.pose_V, (int)h->Nc);
        actual = -1;
        double cost_new = cost;
        while (actual <= 0 && nfev < max_nfev) {
            LAUNCH(MMBA_K_VEC, pose_tr_kernel, 1, 256, 0, d.pose_lam, d.pose_suf, d.pose_w, (int)ncam, m_rows, Delta, alpha, d.scal);
            LAUNCH(MMBA_K_VEC, pose_step_kernel, cdiv(nloc, 256), 256, 0, d.x, d.pose_V, d.pose_w, d.xn, (int)h->Nc, nloc);
            TRY(trial_cost(h, d.xn, d.camtab_n));
            TRY(read_scalars(h));
            nfev++;
            alpha = h->h_scal[PS_ALPHA];
            const double predicted = h->h_scal[PS_PRED], step_h_norm = h->h_scal[PS_STEPNORM];
            const double f2 = h->h_scal[S_COST_NEW];
            if (!std::isfinite(f2)) {
                Delta = 0.25 * step_h_norm;
                continue;
            }
            cost_new = 0.5 * f2;
            actual = cost - cost_new;
            double Delta_new, ratio;
            update_tr_radius(Delta, actual, predicted, step_h_norm, step_h_norm > 0.95 * Delta, &Delta_new, &ratio);
            step_norm = step_h_norm;   // scale = 1: the step and the scaled step coincide
            have_step = true;
            const int term = check_termination(actual, cost, step_norm, x_norm, ratio, o.ftol, o.xtol);
            if (term) {
                status = term;
                break;
            }
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual > 0) {
            std::swap(d.x, d.xn);
            cost = cost_new;
            TRY(linearise(h, true));
            njev++;
            TRY(grad_stats());
            TRY(read_scalars(h));
            x_norm = std::sqrt(h->h_scal[S_X2]);
            g_norm = ginf_of(h->h_scal);
        } else {
            step_norm = 0;
            actual = 0;
            have_step = true;
        }
        nit++;
    }
    if (status == -1) status = 0;
    out->cost = cost;
    out->optimality = g_norm;
    out->nfev = nfev;
    out->njev = njev;
    out->nit = nit;
    out->status = status;
    out->reserved = 0;
    out->pcg_iterations = 0;
    return MMBA_OK;
}

int need_problem(mmba_handle* h) {
    if (!h) return fail(nullptr, MMBA_ERR_ARG, "null handle");
    if (!h->has_problem) return fail(h, MMBA_ERR_STATE, "mmba_set_problem has not been called");
    cudaError_t e = cudaSetDevice(h->opt.device);
    if (e != cudaSuccess) return fail(h, MMBA_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return MMBA_OK;
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

int mmba_version(void) { return MMBA_VERSION; }

const char* mmba_last_error(const mmba_handle* h) { return h ? h->err.c_str() : g_thread_error.c_str(); }

void mmba_default_options(mmba_options* opt) {
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->device = 0;
    opt->rank = 0;
    opt->nranks = 1;
    opt->verbose = 0;
    opt->ftol = 1e-4;
    opt->xtol = 1e-8;
    opt->gtol = 1e-8;
    opt->max_nfev = 0;
    opt->pcg_rtol = 1e-8;
    opt->pcg_maxit = 1000;
    opt->profile = 0;
    opt->schur_mode = MMBA_SCHUR_AUTO;
    opt->reserved = 0;
    opt->pcg_atol = 1e-7;
    opt->pcg_ktol = 1.23e-6;   // LSMR's atol (1e-6) x the growth of its ||A|| estimate (1.23 sqrt(k)): no calibration
}

int mmba_nccl_unique_id(uint8_t out[128]) {
    std::string err;
    if (!out) return fail(nullptr, MMBA_ERR_ARG, "null output");
    if (!load_nccl(err)) return fail(nullptr, MMBA_ERR_NCCL, err);
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, MMBA_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    std::memcpy(out, &id, 128);
    return MMBA_OK;
}

int mmba_create(mmba_handle** out, const mmba_options* opt) {
    if (!out) return fail(nullptr, MMBA_ERR_ARG, "null output handle");
    *out = nullptr;
    mmba_options o;
    if (opt) o = *opt;
    else mmba_default_options(&o);
    if (o.nranks < 1 || o.rank < 0 || o.rank >= o.nranks) return fail(nullptr, MMBA_ERR_ARG, "rank/nranks out of range");
    if (o.pcg_maxit <= 0 || !(o.pcg_rtol > 0)) return fail(nullptr, MMBA_ERR_ARG, "pcg_maxit and pcg_rtol must be positive");
    mmba_handle* h = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MMBA_ERR_CUDA,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (libmmba has no CPU fallback)");
    if (o.device < 0 || o.device >= ndev) return fail(nullptr, MMBA_ERR_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, o.device);
    if (e != cudaSuccess) return fail(nullptr, MMBA_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MMBA_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                                std::to_string(prop.minor) + "; libmmba is built for sm_100a only");
    h = new mmba_handle();
    h->opt = o;
    h->sm_count = prop.multiProcessorCount;
    CU(cudaSetDevice(o.device));
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaMallocHost(&h->h_scal, S_COUNT * sizeof(double)));
    CU(cudaMallocHost(&h->h_flags, 4 * sizeof(int)));
    if (o.nranks > 1) {
        std::string err;
        if (!load_nccl(err)) {
            delete h;
            return fail(nullptr, MMBA_ERR_NCCL, err);
        }
        ncclUniqueId id;
        std::memcpy(&id, o.nccl_id, 128);
        ncclResult_t r = g_nccl.CommInitRank(&h->comm, o.nranks, id, o.rank);
        if (r != ncclSuccess) {
            std::string msg = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
            delete h;
            return fail(nullptr, MMBA_ERR_NCCL, msg);
        }
    }
    *out = h;
    return MMBA_OK;
}

int mmba_set_options(mmba_handle* h, const mmba_options* opt) {
    if (!h || !opt) return fail(h, MMBA_ERR_ARG, "set_options: null argument");
    if (opt->pcg_maxit <= 0 || !(opt->pcg_rtol > 0)) return fail(h, MMBA_ERR_ARG, "pcg_maxit and pcg_rtol must be positive");
    h->opt.verbose = opt->verbose;
    h->opt.ftol = opt->ftol;
    h->opt.xtol = opt->xtol;
    h->opt.gtol = opt->gtol;
    h->opt.max_nfev = opt->max_nfev;
    h->opt.pcg_rtol = opt->pcg_rtol;
    h->opt.pcg_maxit = opt->pcg_maxit;
    h->opt.profile = opt->profile;
    if (opt->schur_mode < MMBA_SCHUR_AUTO || opt->schur_mode > MMBA_SCHUR_EXPLICIT) return fail(h, MMBA_ERR_ARG, "schur_mode out of range");
    if (!(opt->pcg_atol >= 0)) return fail(h, MMBA_ERR_ARG, "pcg_atol must be >= 0");
    h->opt.pcg_atol = opt->pcg_atol;
    if (!(opt->pcg_ktol >= 0)) return fail(h, MMBA_ERR_ARG, "pcg_ktol must be >= 0");
    h->opt.pcg_ktol = opt->pcg_ktol;
    h->opt.schur_mode = opt->schur_mode;   // explicit / auto take effect at the next mmba_set_problem
    return MMBA_OK;
}

void mmba_destroy(mmba_handle* h) {
    if (!h) return;
    cudaSetDevice(h->opt.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    release_problem(h);
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr;
    xchg_release(h);
    h->planner.release();
    if (h->stage_buf) cudaFreeHost(h->stage_buf);
    for (int i = 0; i < kStageSlots; ++i)
        if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    if (h->comm) g_nccl.CommDestroy(h->comm);
    for (cudaEvent_t e : h->prof.pool) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// Sharded set-up: the observations a rank staged (a contiguous chunk of the caller's arrays) travel to the ranks that
// own their points.  One all-gather of the per-destination counts, then one grouped send / receive per peer over NVLink.
// Chunks arrive in rank order and keep their order, so every rank ends up with its observations in ascending caller
// order: the plan below is the one a single GPU would build for these points.
static int exchange_observations(mmba_handle* h, int64_t o0) {
    DevPlanner& P = h->planner;
    const int nr = h->opt.nranks, me = h->opt.rank;
    std::string err;
    int rc = devplan_dispatch_pack(P, h->dp, o0, h->stream, err);
    if (rc != MMBA_OK) return fail(h, rc, err);
    int* d_all = P.d_counts + 16;
    NC(g_nccl.AllGather(P.d_counts, d_all, (size_t)nr, ncclInt32, h->comm, h->stream));
    std::vector<int> all((size_t)nr * nr);
    CU(cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int64_t n_recv = 0;
    for (int r = 0; r < nr; ++r) n_recv += all[(size_t)r * nr + me];
    const int32_t *cam_s = P.cam_s, *pt_s = P.pt_s, *gidx_s = P.gidx_s;
    const double* uv_s = P.uv_s;
    rc = devplan_dispatch_recv(P, n_recv, err);
    if (rc != MMBA_OK) return fail(h, rc, err);
    prof_begin(h, MMBA_K_ALLREDUCE);
    NC(g_nccl.GroupStart());
    int64_t soff = 0, roff = 0;
    for (int r = 0; r < nr; ++r) {
        const size_t sc = (size_t)all[(size_t)me * nr + r], rcn = (size_t)all[(size_t)r * nr + me];
        if (sc) {
            NC(g_nccl.Send(cam_s + soff, sc, ncclInt32, r, h->comm, h->stream));
            NC(g_nccl.Send(pt_s + soff, sc, ncclInt32, r, h->comm, h->stream));
            NC(g_nccl.Send(gidx_s + soff, sc, ncclInt32, r, h->comm, h->stream));
            NC(g_nccl.Send(uv_s + 2 * soff, 2 * sc, ncclDouble, r, h->comm, h->stream));
        }
        if (rcn) {
            NC(g_nccl.Recv(P.cam_r + roff, rcn, ncclInt32, r, h->comm, h->stream));
            NC(g_nccl.Recv(P.pt_r + roff, rcn, ncclInt32, r, h->comm, h->stream));
            NC(g_nccl.Recv(P.gidx_r + roff, rcn, ncclInt32, r, h->comm, h->stream));
            NC(g_nccl.Recv(P.uv_r + 2 * roff, 2 * rcn, ncclDouble, r, h->comm, h->stream));
        }
        soff += (int64_t)sc;
        roff += (int64_t)rcn;
    }
    NC(g_nccl.GroupEnd());
    prof_end(h, MMBA_K_ALLREDUCE);
    return MMBA_OK;
}

// Sharded set-up: every rank marked the camera pairs of its own points; the block pattern is their union.
static int or_bitmaps(mmba_handle* h) {
    DevPlanner& P = h->planner;
    unsigned long long* bits;
    size_t n_words;
    devplan_bitmap(P, h->dp, &bits, &n_words);
    const int nr = h->opt.nranks;
    cudaError_t e = P.tmp.ensure((size_t)nr * n_words * sizeof(unsigned long long));
    if (e != cudaSuccess) return fail(h, MMBA_ERR_NOMEM, std::string("set_problem: bitmap exchange buffer: ") + cudaGetErrorString(e));
    NC(g_nccl.AllGather(bits, P.tmp.p, n_words, ncclUint64, h->comm, h->stream));
    devplan_or_bitmaps(bits, reinterpret_cast<unsigned long long*>(P.tmp.p), n_words, nr, h->stream);
    return MMBA_OK;
}

// Stage observations [o0, o1) of the caller's arrays through the pinned ring into the planner's input buffers:
// indices narrowed to int32 and range-checked on the way (several host threads per slot), pixels copied as they are.
static int upload_observations(mmba_handle* h, int64_t n_cams, int64_t n_points, const int64_t* cam_idx, const int64_t* pt_idx,
                               const double* uv, int64_t o0, int64_t o1) {
    DevPlanner& P = h->planner;
    const int64_t n = o1 - o0;
    const bool sharded = h->opt.nranks > 1;
    for (int pass = 0; pass < 2; ++pass) {
        Carver c;
        c.base = pass ? P.in.p : nullptr;
        P.cam = c.take<int32_t>((size_t)std::max<int64_t>(n, 1));
        P.pt = c.take<int32_t>((size_t)std::max<int64_t>(n, 1));
        P.uv = c.take<double>(2 * (size_t)std::max<int64_t>(n, 1));
        if (sharded) {   // the same observations grouped by destination rank (devplan_dispatch_pack)
            P.cam_s = c.take<int32_t>((size_t)std::max<int64_t>(n, 1));
            P.pt_s = c.take<int32_t>((size_t)std::max<int64_t>(n, 1));
            P.gidx_s = c.take<int32_t>((size_t)std::max<int64_t>(n, 1));
            P.uv_s = c.take<double>(2 * (size_t)std::max<int64_t>(n, 1));
        }
        if (!pass) {
            cudaError_t e = P.in.ensure(c.off + 256);
            if (e != cudaSuccess) return fail(h, MMBA_ERR_NOMEM, std::string("set_problem: input buffers: ") + cudaGetErrorString(e));
        }
    }
    P.gidx = nullptr;
    P.n_in = n;
    const int64_t chunk = (int64_t)(kStageSlotBytes / 24);   // 4 + 4 + 16 bytes per observation
    int64_t first_bad = -1;
    for (int64_t b = o0; b < o1; b += chunk) {
        const int64_t m = std::min(chunk, o1 - b);
        int idx;
        char* slot;
        TRY(stage_acquire(h, &idx, &slot));
        int32_t* s_cam = reinterpret_cast<int32_t*>(slot);
        int32_t* s_pt = s_cam + m;
        double* s_uv = reinterpret_cast<double*>(slot + 8 * (size_t)((m + 1) / 2 * 2));
        int64_t bad[16];
        for (int w = 0; w < 16; ++w) bad[w] = -1;
        parallel_ranges(m, 32768, [&](int64_t i0, int64_t i1, int worker) {
            for (int64_t i = i0; i < i1; ++i) {
                const int64_t c = cam_idx[b + i], p = pt_idx[b + i];
                if ((uint64_t)c >= (uint64_t)n_cams || (uint64_t)p >= (uint64_t)n_points) {
                    if (bad[worker] < 0) bad[worker] = b + i;
                    s_cam[i] = 0;
                    s_pt[i] = 0;
                } else {
                    s_cam[i] = (int32_t)c;
                    s_pt[i] = (int32_t)p;
                }
            }
            std::memcpy(s_uv + 2 * i0, uv + 2 * (b + i0), (size_t)(i1 - i0) * 16);
        }, 16, true);
        for (int w = 0; w < 16; ++w)
            if (bad[w] >= 0 && (first_bad < 0 || bad[w] < first_bad)) first_bad = bad[w];
        CU(cudaMemcpyAsync(P.cam + (b - o0), s_cam, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(P.pt + (b - o0), s_pt, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(P.uv + 2 * (b - o0), s_uv, (size_t)m * 16, cudaMemcpyHostToDevice, h->stream));
        TRY(stage_commit(h, idx));
        if (first_bad >= 0) break;
    }
    if (sharded) {   // every rank checked its own chunk: agree on the first bad observation, fail together
        if (!P.d_counts) CU(cudaMalloc(&P.d_counts, (16 + 16 * 16) * sizeof(int)));
        long long v = first_bad < 0 ? INT64_MAX : first_bad;
        long long* d_v = reinterpret_cast<long long*>(P.d_counts);
        long long* d_all = reinterpret_cast<long long*>(P.d_counts + 16);
        std::vector<long long> all(h->opt.nranks);
        CU(cudaMemcpyAsync(d_v, &v, sizeof(v), cudaMemcpyHostToDevice, h->stream));
        NC(g_nccl.AllGather(d_v, d_all, 1, ncclInt64, h->comm, h->stream));
        CU(cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (long long w : all) v = std::min(v, w);
        first_bad = v == INT64_MAX ? -1 : (int64_t)v;
    }
    if (first_bad >= 0) {
        CU(cudaStreamSynchronize(h->stream));
        return fail(h, MMBA_ERR_ARG, "set_problem: index out of range at observation " + std::to_string(first_bad));
    }
    return MMBA_OK;
}

int mmba_set_problem(mmba_handle* h, int64_t n_cams, int64_t n_points, int64_t n_obs, const double K[9],
                     const int64_t* cam_idx, const int64_t* pt_idx, const double* uv) {
    if (!h) return fail(nullptr, MMBA_ERR_ARG, "null handle");
    if (!K || !uv) return fail(h, MMBA_ERR_ARG, "set_problem: null K or uv");
    if (n_cams <= 0 || n_points <= 0 || n_obs <= 0 || !cam_idx || !pt_idx)
        return fail(h, MMBA_ERR_ARG, "set_problem: sizes must be positive and index arrays non-null");
    if (n_cams > (int64_t)kMaxCamBlocks * kCamBlock) return fail(h, MMBA_ERR_ARG, "set_problem: too many cameras");
    CU(cudaSetDevice(h->opt.device));
    release_problem(h);
    // MMBA_PLAN_TIMING=1: host-side phase times of this call on stderr (diagnostics; synchronises after every phase)
    auto T0 = std::chrono::steady_clock::now();
    const bool timing = getenv("MMBA_PLAN_TIMING") != nullptr;
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(h->stream);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "set_problem %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count());
        T0 = t;
    };
    std::string err;
    DevPlanner& P = h->planner;
    DevPlan& D = h->dp;
    D = DevPlan();
    D.n_cams = n_cams;
    D.n_points = n_points;
    D.n_obs = n_obs;
    D.rank = h->opt.rank;
    D.nranks = h->opt.nranks;
    h->h_point_perm.clear();
    const int nr = h->opt.nranks;
    // 1. this rank's share of the caller's arrays goes to the device (rank r stages observations
    //    [n_obs r / nranks, n_obs (r + 1) / nranks): no rank reads the whole problem)
    const int64_t o0 = n_obs * h->opt.rank / nr, o1 = n_obs * (h->opt.rank + 1) / nr;
    TRY(upload_observations(h, n_cams, n_points, cam_idx, pt_idx, uv, o0, o1));
    lap("stage + upload");
    // 2. per-point statistics (summed over ranks), internal point order, shard cuts
    int rc = devplan_stats(P, D, h->stream, err);
    if (rc != MMBA_OK) return fail(h, rc, err);
    if (nr > 1) {
        // Per-point statistics over all ranks: ONE all-gather of the per-rank blocks, combined by a kernel of ours.
        // (Four grouped ncclAllReduce calls — sum / min / max / min over int32 — left a block of ~70 k entries of one of
        // the 4 MB arrays unreduced on 8 GPUs, NCCL 2.28.9: measured twice, a different array each time; an all-gather
        // has no reduction to get wrong, and the combine is deterministic.)
        int* block;
        size_t n_ints;
        devplan_stat_block(P, D, &block, &n_ints);
        cudaError_t e = P.tmp.ensure((size_t)nr * n_ints * sizeof(int));
        if (e != cudaSuccess) return fail(h, MMBA_ERR_NOMEM, std::string("set_problem: statistics exchange buffer: ") + cudaGetErrorString(e));
        prof_begin(h, MMBA_K_ALLREDUCE);
        NC(g_nccl.AllGather(block, P.tmp.p, n_ints, ncclInt32, h->comm, h->stream));
        prof_end(h, MMBA_K_ALLREDUCE);
        devplan_combine_stats(P, D, reinterpret_cast<const int*>(P.tmp.p), nr, h->stream);
    }
    rc = devplan_order(P, D, h->stream, err);
    if (rc != MMBA_OK) return fail(h, rc, err);
    lap("point order");
    // 3. (sharded) every observation travels to the rank that owns its point: one all-to-all over NVLink
    if (nr > 1) TRY(exchange_observations(h, o0));
    // 4. observation grouping, tiles, per-tile tables, co-visibility bitmap
    const bool want_pattern = h->opt.schur_mode != MMBA_SCHUR_IMPLICIT && n_cams <= 20000;
    rc = devplan_tiles(P, D, want_pattern, h->stream, err);
    if (rc != MMBA_OK) return fail(h, rc, err);
    if (want_pattern) {
        if (nr > 1) TRY(or_bitmaps(h));
        rc = devplan_pattern_sizes(P, D, h->stream, err);
        if (rc != MMBA_OK) return fail(h, rc, err);
    }
    // (sharded: a too-long track is seen by the rank that owns the point only; all ranks must fail together)
    std::vector<int> err_all;
    if (nr > 1) {
        NC(g_nccl.AllGather(&P.d_info->err_track, P.d_counts + 16, 1, ncclInt32, h->comm, h->stream));
        err_all.resize(nr);
        CU(cudaMemcpyAsync(err_all.data(), P.d_counts + 16, nr * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    rc = devplan_sync_sizes(P, D, h->stream, err);
    for (int v : err_all)
        if (rc == MMBA_OK && v > 0) {
            err = "set_problem: point " + std::to_string(v - 1) + " has more observations than one tile holds (" + std::to_string(kTileObs) + ")";
            rc = MMBA_ERR_TRACK;
        }
    if (rc != MMBA_OK) return fail(h, rc, err);
    lap("tiles + pattern bitmap");
    h->Nc = n_cams;
    h->npl = D.pt_end - D.pt_begin;
    h->ns = D.n_slots;
    h->nt = D.n_tiles;
    h->nloc = 6 * h->Nc + 3 * h->npl;
    std::memcpy(h->K, K, sizeof(h->K));
    // Explicit reduced camera matrix (rcm.h) when it is small next to the observation stream: at most
    // 20 000 cameras (co-visibility bitmap), blocks within ~1/2 of the J bytes a PCG iteration would stream
    // and L2-sized, and a bounded S-build cost (pair-blocks per observation).
    h->rcm_ready = false;
    if (want_pattern) {
        const bool forced = h->opt.schur_mode == MMBA_SCHUR_EXPLICIT;
        const int64_t cap_full = forced ? (int64_t)8 << 20 : std::min<int64_t>(((int64_t)96 << 20) / 288, (152 * n_obs / 2) / 288);
        const int64_t cap = std::max<int64_t>(cap_full, n_cams);
        // (the host builder bounds the upper triangle by the same cap while counting)
        h->rcm_ready = D.nnz_up <= cap && (forced || (D.nnz_full <= cap && D.total_pairs <= 64 * n_obs));
        if (h->rcm_ready) {
            rc = devplan_pattern_fill(P, D, std::min(h->sm_count, kRcmMaxCtas), h->stream, err);
            if (rc != MMBA_OK) return fail(h, rc, err);
            // the PCG kernel keeps the search direction on every CTA's halo in shared memory
            if (rcm_smem(D.cpc, D.nblk_max, D.nh_max, 0, D.n_ctas).total > 200 * 1024) h->rcm_ready = false;
        }
    }
    D.rcm_ok = h->rcm_ready;
    if (h->opt.schur_mode == MMBA_SCHUR_EXPLICIT && !h->rcm_ready)
        return fail(h, MMBA_ERR_ARG, "set_problem: the reduced camera matrix is too large to be formed explicitly");
    h->hist_cap = (h->rcm_ready && (h->opt.profile & 2)) ? h->opt.pcg_maxit : 0;
    lap("pattern CSR + partition");
    Arena measure;
    carve(h, measure);
    const size_t need = measure.off + 256;
    if (!h->arena || h->arena_cap < need) {
        if (h->arena) cudaFree(h->arena);
        h->arena = nullptr;
        h->arena_cap = 0;
        cudaError_t e = cudaMalloc(&h->arena, need);
        if (e != cudaSuccess) {
            h->arena = nullptr;
            return fail(h, MMBA_ERR_NOMEM, "set_problem: cudaMalloc of " + std::to_string(need) + " bytes failed: " +
                                               cudaGetErrorString(e));
        }
        h->arena_cap = need;
    }
    h->arena_bytes = need;
    Arena a;
    a.base = static_cast<char*>(h->arena);
    carve(h, a);
    Dev& d = h->d;
    CU(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
    lap("arena");
    TileArgs& A = h->targs;
    A.meta = d.meta;
    A.tile_cams = d.tile_cams;
    A.uv = d.uv;
    A.n_tiles = (int)h->nt;
    A.cam_stride = D.cam_stride;
    A.max_cams = std::max(D.max_tile_cams, 1);
    A.max_pts = std::max(D.max_tile_pts, 1);
    A.n_cams = (int)h->Nc;
    A.ytab_cams = h->Nc <= 340 ? (int)h->Nc : 0;   // <= 16 KB of shared memory
    A.sb_stages = Traits<M_SBUILD>::kStages;
    A.dbg = (h->opt.profile & 4) ? d.dbg : nullptr;
    std::memcpy(A.K, K, sizeof(A.K));
    TRY(configure_kernels(h));
    lap("configure kernels");
    TRY(xchg_setup(h));
    lap("peer exchange setup");
    h->has_problem = true;
    return MMBA_OK;
}

// TRF solve from the parameters in d.x; device time of the solve goes to result->solve_ms
static int solve_on_device(mmba_handle* h, mmba_result* result) {
    std::memset(result, 0, sizeof(*result));
    std::memset(h->prof.launches, 0, sizeof(h->prof.launches));
    std::memset(h->prof.ms, 0, sizeof(h->prof.ms));
    CU(cudaEventRecord(h->ev0, h->stream));
    int rc = run_trf(h, result);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    result->solve_ms = ms;
    prof_collect(h);
    return rc;
}

int mmba_solve(mmba_handle* h, double* x, mmba_result* result, double* fun_out) {
    TRY(need_problem(h));
    if (!x || !result) return fail(h, MMBA_ERR_ARG, "solve: null x or result");
    TRY(put_x(h, x, h->d.x));
    TRY(solve_on_device(h, result));
    TRY(get_x(h, h->d.x, x));
    if (fun_out) TRY(get_slots(h, h->d.res, 2, 0, 2, fun_out));
    return MMBA_OK;
}

int mmba_solve_split(mmba_handle* h, const double* cams_in, const double* points_in, double* cams_out, double* points_out,
                     mmba_result* result, double* fun_out) {
    TRY(need_problem(h));
    if (!cams_in || !points_in || !cams_out || !points_out || !result) return fail(h, MMBA_ERR_ARG, "solve_split: null argument");
    TRY(put_x_split(h, cams_in, points_in, h->d.x));
    TRY(solve_on_device(h, result));
    TRY(get_x(h, h->d.x, nullptr, cams_out, points_out));
    if (fun_out) TRY(get_slots(h, h->d.res, 2, 0, 2, fun_out));
    return MMBA_OK;
}

int mmba_solve_pose(mmba_handle* h, double* x, mmba_result* result, double* fun_out) {
    TRY(need_problem(h));
    if (!x || !result) return fail(h, MMBA_ERR_ARG, "solve_pose: null x or result");
    if (h->opt.nranks > 1) return fail(h, MMBA_ERR_STATE, "solve_pose runs on a single-GPU handle (nranks == 1)");
    TRY(put_x(h, x, h->d.x));
    std::memset(result, 0, sizeof(*result));
    std::memset(h->prof.launches, 0, sizeof(h->prof.launches));
    std::memset(h->prof.ms, 0, sizeof(h->prof.ms));
    CU(cudaEventRecord(h->ev0, h->stream));
    int rc = run_trf_pose(h, result);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    result->solve_ms = ms;
    prof_collect(h);
    if (rc != MMBA_OK) return rc;
    TRY(get_x(h, h->d.x, x));
    if (fun_out) TRY(get_slots(h, h->d.res, 2, 0, 2, fun_out));
    return MMBA_OK;
}

int mmba_set_x(mmba_handle* h, const double* x) {
    TRY(need_problem(h));
    if (!x) return fail(h, MMBA_ERR_ARG, "set_x: null x");
    return put_x(h, x, h->d.x0);
}

int mmba_solve_resident(mmba_handle* h, mmba_result* result) {
    TRY(need_problem(h));
    if (!result) return fail(h, MMBA_ERR_ARG, "solve_resident: null result");
    CU(cudaMemcpyAsync(h->d.x, h->d.x0, h->nloc * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return solve_on_device(h, result);
}

int mmba_get_x(mmba_handle* h, double* x) {
    TRY(need_problem(h));
    if (!x) return fail(h, MMBA_ERR_ARG, "get_x: null x");
    return get_x(h, h->d.x, x);
}

int mmba_get_log(const mmba_handle* h, mmba_iter_log* out, int capacity) {
    if (!h) return MMBA_ERR_ARG;
    const int n = (int)h->log.size();
    if (out)
        for (int i = 0; i < n && i < capacity; ++i) out[i] = h->log[i];
    return n;
}

int mmba_get_pcg_history(const mmba_handle* h, int outer_iteration, double* out, int capacity) {
    if (!h) return MMBA_ERR_ARG;
    if (outer_iteration < 0) return (int)h->pcg_hist.size();
    if ((size_t)outer_iteration >= h->pcg_hist.size()) return MMBA_ERR_ARG;
    const std::vector<double>& v = h->pcg_hist[outer_iteration];
    if (out)
        for (int i = 0; i < (int)v.size() && i < capacity; ++i) out[i] = v[i];
    return (int)v.size();
}

int mmba_get_phase_cycles(mmba_handle* h, int64_t out[64], int reset) {
    TRY(need_problem(h));
    if (!out) return fail(h, MMBA_ERR_ARG, "get_phase_cycles: null argument");
    static_assert(sizeof(long long) == sizeof(int64_t), "counter width");
    CU(cudaMemcpyAsync(out, h->d.dbg, 64 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (reset) CU(cudaMemsetAsync(h->d.dbg, 0, 64 * sizeof(int64_t), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MMBA_OK;
}

int mmba_get_profile(const mmba_handle* h, int64_t launches[MMBA_K_COUNT], double ms[MMBA_K_COUNT]) {
    if (!h) return MMBA_ERR_ARG;
    for (int i = 0; i < MMBA_K_COUNT; ++i) {
        if (launches) launches[i] = h->prof.launches[i];
        if (ms) ms[i] = h->prof.ms[i];
    }
    return MMBA_OK;
}

int mmba_get_shard(const mmba_handle* h, int64_t* n_obs_local, int64_t* n_points_local, int64_t* n_tiles) {
    if (!h || !h->has_problem) return MMBA_ERR_STATE;
    if (n_obs_local) *n_obs_local = h->dp.n_obs_local;
    if (n_points_local) *n_points_local = h->npl;
    if (n_tiles) *n_tiles = h->nt;
    return MMBA_OK;
}

// ---- evaluation hooks ---------------------------------------------------------------------------
int mmba_eval_residual(mmba_handle* h, const double* x, double* f) {
    TRY(need_problem(h));
    if (!x || !f) return fail(h, MMBA_ERR_ARG, "eval_residual: null argument");
    Dev& d = h->d;
    TRY(put_x(h, x, d.x));
    TRY(cam_prep(h, d.x, d.camtab));
    TRY(zero(h, d.scal + S_COST_NEW, 1));
    {
        ModeArgs P{};
        P.res = d.res;
        P.cam0 = d.camtab;
        P.ptA = d.x + 6 * h->Nc;
        P.cost = d.scal + S_COST_NEW;
        TRY(launch_tile<M_RESID_STORE>(h, MMBA_K_RESID, P));
    }
    CU(cudaGetLastError());
    return get_slots(h, d.res, 2, 0, 2, f);
}

int mmba_eval_jacobian(mmba_handle* h, const double* x, double* Jc, double* Jp) {
    TRY(need_problem(h));
    if (!x || !Jc || !Jp) return fail(h, MMBA_ERR_ARG, "eval_jacobian: null argument");
    Dev& d = h->d;
    TRY(put_x(h, x, d.x));
    TRY(linearise(h));
    return get_jacobian_slots(h, Jc, Jp);
}

int mmba_eval_blocks(mmba_handle* h, const double* x, double* U, double* V, double* gc, double* gp, double* cost) {
    TRY(need_problem(h));
    if (!x) return fail(h, MMBA_ERR_ARG, "eval_blocks: null x");
    Dev& d = h->d;
    const DevPlan& pl = h->dp;
    TRY(put_x(h, x, d.x));
    TRY(linearise(h, true));
    TRY(host_point_perm(h));
    std::vector<double> hU(21 * h->Nc), hg(h->nloc), hV(6 * std::max<int64_t>(h->npl, 1));
    CU(cudaMemcpyAsync(hU.data(), d.U, hU.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(hg.data(), d.g, hg.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->npl) CU(cudaMemcpyAsync(hV.data(), d.V, 6 * h->npl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TRY(read_scalars(h));
    if (cost) *cost = 0.5 * h->h_scal[S_COST];
    if (U)
        for (int64_t c = 0; c < h->Nc; ++c)
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) U[c * 36 + a * 6 + b] = hU[c * 21 + (a <= b ? tri6(a, b) : tri6(b, a))];
    if (gc) std::memcpy(gc, hg.data(), 6 * h->Nc * sizeof(double));
    if (V && h->opt.nranks > 1) std::memset(V, 0, 9 * pl.n_points * sizeof(double));
    if (gp && h->opt.nranks > 1) std::memset(gp, 0, 3 * pl.n_points * sizeof(double));
    for (int64_t q = 0; q < h->npl; ++q) {
        const int64_t p = h->h_point_perm[pl.pt_begin + q];
        if (V)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) V[p * 9 + a * 3 + b] = hV[q * 6 + (a <= b ? tri3(a, b) : tri3(b, a))];
        if (gp)
            for (int k = 0; k < 3; ++k) gp[p * 3 + k] = hg[6 * h->Nc + q * 3 + k];
    }
    return MMBA_OK;
}

int mmba_eval_gn_step(mmba_handle* h, const double* x, const double* scale, double reg, double* p, int64_t* pcg_iterations,
                      double* pcg_relres) {
    TRY(need_problem(h));
    if (!x || !scale || !p) return fail(h, MMBA_ERR_ARG, "eval_gn_step: null argument");
    Dev& d = h->d;
    TRY(put_x(h, x, d.x));
    TRY(linearise(h));
    // scale_inv = 1 / scale in the internal layout
    {
        std::vector<double> si(6 * h->Nc + 3 * h->dp.n_points);
        for (size_t i = 0; i < si.size(); ++i) si[i] = 1.0 / scale[i];
        TRY(put_x(h, si.data(), d.sinv));
    }
    TRY(gn_step(h, reg, 0.0, pcg_iterations, pcg_relres));
    // p = [px | dp * sinv_p]
    const int64_t ncam = 6 * h->Nc, npt = 3 * h->npl;
    TRY(zero(h, d.scal + S_DOT0, 2));
    LAUNCH(MMBA_K_VEC, gn_assemble_kernel<false>, cdiv(ncam, 256), 256, 0, d.px, d.sinv, d.gh, d.gn, ncam, d.scal, 0);
    if (npt)
        LAUNCH(MMBA_K_VEC, gn_assemble_kernel<true>, cdiv(npt, 256), 256, 0, d.dp, d.sinv + ncam, d.gh + ncam, d.gn + ncam, npt,
               d.scal, 0);
    CU(cudaGetLastError());
    return get_x(h, d.gn, p);
}

int mmba_eval_reduced_system(mmba_handle* h, const double* x, const double* scale, double reg, double* S_dense, double* rhs) {
    TRY(need_problem(h));
    if (!x || !scale || !S_dense || !rhs) return fail(h, MMBA_ERR_ARG, "eval_reduced_system: null argument");
    if (!use_rcm(h)) return fail(h, MMBA_ERR_STATE, "eval_reduced_system: the reduced camera matrix is not formed explicitly");
    Dev& d = h->d;
    TRY(put_x(h, x, d.x));
    TRY(linearise(h));
    {
        std::vector<double> si(6 * h->Nc + 3 * h->dp.n_points);
        for (size_t i = 0; i < si.size(); ++i) si[i] = 1.0 / scale[i];
        TRY(put_x(h, si.data(), d.sinv));
    }
    if (h->npl)
        LAUNCH(MMBA_K_PTINV, point_invert_kernel, cdiv(h->npl, 256), 256, 0, d.V, d.g + 6 * h->Nc, d.sinv + 6 * h->Nc, reg,
               d.M, d.zg, h->npl);
    TRY(rcm_build(h));
    TRY(rcm_finalize(h, reg));
    const int64_t nnz = h->dp.nnz_full;
    h->h_rc_rows.resize(nnz);
    h->h_rc_cols.resize(nnz);
    CU(cudaMemcpyAsync(h->h_rc_rows.data(), h->dp.rows, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->h_rc_cols.data(), h->dp.cols, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    std::vector<double> hS(36 * (size_t)nnz);
    CU(cudaMemcpyAsync(hS.data(), d.S, hS.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(rhs, d.rcm_b, 6 * h->Nc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    const int64_t n6 = 6 * h->Nc;
    std::memset(S_dense, 0, (size_t)n6 * n6 * sizeof(double));
    for (int64_t k = 0; k < nnz; ++k)
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b)
                S_dense[(6 * (int64_t)h->h_rc_rows[k] + a) * n6 + 6 * (int64_t)h->h_rc_cols[k] + b] = hS[k * 36 + a * 6 + b];
    return MMBA_OK;
}

int mmba_eval_jnorm2(mmba_handle* h, const double* x, const double* s, double* jnorm2) {
    TRY(need_problem(h));
    if (!x || !s || !jnorm2) return fail(h, MMBA_ERR_ARG, "eval_jnorm2: null argument");
    Dev& d = h->d;
    TRY(put_x(h, x, d.x));
    TRY(linearise(h));
    TRY(put_x(h, s, d.tmp));
    TRY(jv1(h, d.tmp));
    TRY(read_scalars(h));
    *jnorm2 = h->h_scal[S_JV00];
    return MMBA_OK;
}

int mmba_bench_kernel(mmba_handle* h, const double* x, int kernel_class, int iters, double* avg_ms) {
    TRY(need_problem(h));
    if (!x || !avg_ms || iters <= 0) return fail(h, MMBA_ERR_ARG, "bench_kernel: bad argument");
    Dev& d = h->d;
    TRY(put_x(h, x, d.x));
    TRY(linearise(h));
    TRY(scale_and_grad(h, true));
    const double reg = 1e-4;
    int64_t its;
    const int saved_maxit = h->opt.pcg_maxit;
    h->opt.pcg_maxit = 2;
    int rc = gn_step(h, reg, 0.0, &its, nullptr);
    h->opt.pcg_maxit = saved_maxit;
    TRY(rc);
    CU(cudaMemsetAsync(d.flags, 0, 2 * sizeof(int), h->stream));
    LAUNCH(MMBA_K_VEC, unscale_kernel, cdiv(h->nloc, 256), 256, 0, d.gh, d.sinv, 1.0, d.tmp, h->nloc);
    CU(cudaStreamSynchronize(h->stream));
    if (!h->nt) return fail(h, MMBA_ERR_STATE, "bench_kernel: this shard has no observations");
    ModeArgs Pb{};
    Pb.Jw = d.J;
    Pb.res = d.res;
    Pb.cam0 = d.camtab;
    Pb.ptA = d.x + 6 * h->Nc;
    Pb.Ud = d.Ud;
    Pb.gc = d.g;
    Pb.V = d.V;
    Pb.gp = d.g + 6 * h->Nc;
    Pb.scal = d.scal;
    ModeArgs Pr = Pb;
    Pr.cost = d.scal + S_COST_NEW;
    ModeArgs Pj{};
    Pj.J = d.J;
    Pj.cam0 = d.tmp;
    Pj.ptA = d.tmp + 6 * h->Nc;
    Pj.cam1 = d.gh;
    Pj.ptB = d.gh + 6 * h->Nc;
    Pj.scal = d.scal;
    if (kernel_class == 102) {
        // LL-line round trip between the first and the last CTA of a device-filling grid (all CTAs co-resident)
        if (!use_rcm(h)) return fail(h, MMBA_ERR_STATE, "bench_kernel: the reduced camera matrix is not formed explicitly");
        LLLine* a = d.rcm_z;
        LLLine* b = d.rcm_z + 64;
        for (int pass = 0; pass < 2; ++pass) {
            const int n = pass == 0 ? 16 : iters;
            if (pass == 1) CU(cudaEventRecord(h->ev0, h->stream));
            ll_pingpong_kernel<<<h->sm_count, 32, 0, h->stream>>>(a, b, n, h->rcm_seq);
            h->rcm_seq += (unsigned)n + 2u;
            if (pass == 1) CU(cudaEventRecord(h->ev1, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaGetLastError());
        }
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *avg_ms = ms / iters;
        return MMBA_OK;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const int n = pass == 0 ? 2 : iters;
        if (pass == 1) CU(cudaEventRecord(h->ev0, h->stream));
        for (int i = 0; i < n; ++i) {
            switch (kernel_class) {
                case MMBA_K_BUILD: TRY(launch_tile<M_BUILD>(h, MMBA_K_BUILD, Pb)); break;
                case MMBA_K_RESID: TRY(launch_tile<M_RESID>(h, MMBA_K_RESID, Pr)); break;
                case MMBA_K_RHS: TRY(launch_tile<M_RHS>(h, MMBA_K_RHS, rhs_args(h))); break;
                case MMBA_K_MATVEC: TRY(launch_tile<M_MATVEC>(h, MMBA_K_MATVEC, matvec_args(h))); break;
                case MMBA_K_BACKSUB: TRY(launch_tile<M_BACKSUB>(h, MMBA_K_BACKSUB, backsub_args(h))); break;
                case MMBA_K_JV: TRY(launch_tile<M_JV2>(h, MMBA_K_JV, Pj)); break;
                case 100: TRY(launch_tile<M_JV1>(h, MMBA_K_JV, Pj)); break;
                case 101:   // plain streaming read of Jt: the ceiling a trivial kernel reaches on these bytes
                    stream_read_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(reinterpret_cast<const double2*>(d.J),
                                                                                (int64_t)kJRows * h->ns / 2, d.scal + S_DOT9);
                    break;
                case MMBA_K_SBUILD:
                    if (!use_rcm(h)) return fail(h, MMBA_ERR_STATE, "bench_kernel: the reduced camera matrix is not formed explicitly");
                    TRY(launch_tile<M_SBUILD>(h, MMBA_K_SBUILD, sbuild_args(h)));
                    break;
                case MMBA_K_PCG:
                    if (!use_rcm(h)) return fail(h, MMBA_ERR_STATE, "bench_kernel: the reduced camera matrix is not formed explicitly");
                    TRY(rcm_pcg(h, 0.0));
                    break;
                case MMBA_K_PTINV:
                    point_invert_kernel<<<cdiv(h->npl, 256), 256, 0, h->stream>>>(d.V, d.g + 6 * h->Nc, d.sinv + 6 * h->Nc, reg,
                                                                                   d.M, d.zg, h->npl);
                    break;
                default:
                    return fail(h, MMBA_ERR_ARG, "bench_kernel: unsupported kernel class");
            }
        }
        if (pass == 1) CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *avg_ms = ms / iters;
    return MMBA_OK;
}

// ---- batched two-view triangulation (no handle: a free function on one device) --------------------
int mmba_triangulate(int device, int64_t n_frames, const double* projections, int64_t n, const int64_t* f1,
                     const int64_t* f2, const double* uv1, const double* uv2, double* points, double* kernel_ms) {
    mmba_handle* h = nullptr;
    if (n_frames <= 0 || n < 0 || !projections || (n > 0 && (!f1 || !f2 || !uv1 || !uv2 || !points)))
        return fail(h, MMBA_ERR_ARG, "triangulate: bad argument");
    for (int64_t i = 0; i < n; ++i)
        if (f1[i] < 0 || f1[i] >= n_frames || f2[i] < 0 || f2[i] >= n_frames)
            return fail(h, MMBA_ERR_ARG, "triangulate: frame index out of range at track " + std::to_string(i));
    if (n == 0) return MMBA_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(h, MMBA_ERR_CUDA, "no CUDA device (libmmba has no CPU fallback)");
    CU(cudaSetDevice(device));
    double *d_proj = nullptr, *d_uv1 = nullptr, *d_uv2 = nullptr, *d_out = nullptr;
    int64_t *d_f1 = nullptr, *d_f2 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_proj);
        cudaFree(d_uv1);
        cudaFree(d_uv2);
        cudaFree(d_out);
        cudaFree(d_f1);
        cudaFree(d_f2);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define TRI(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cleanup();                                                                             \
            return fail(h, MMBA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
        }                                                                                          \
    } while (0)
    TRI(cudaMalloc(&d_proj, n_frames * 12 * sizeof(double)));
    TRI(cudaMalloc(&d_uv1, n * 2 * sizeof(double)));
    TRI(cudaMalloc(&d_uv2, n * 2 * sizeof(double)));
    TRI(cudaMalloc(&d_out, n * 3 * sizeof(double)));
    TRI(cudaMalloc(&d_f1, n * sizeof(int64_t)));
    TRI(cudaMalloc(&d_f2, n * sizeof(int64_t)));
    TRI(cudaEventCreate(&e0));
    TRI(cudaEventCreate(&e1));
    TRI(cudaMemcpy(d_proj, projections, n_frames * 12 * sizeof(double), cudaMemcpyHostToDevice));
    TRI(cudaMemcpy(d_uv1, uv1, n * 2 * sizeof(double), cudaMemcpyHostToDevice));
    TRI(cudaMemcpy(d_uv2, uv2, n * 2 * sizeof(double), cudaMemcpyHostToDevice));
    TRI(cudaMemcpy(d_f1, f1, n * sizeof(int64_t), cudaMemcpyHostToDevice));
    TRI(cudaMemcpy(d_f2, f2, n * sizeof(int64_t), cudaMemcpyHostToDevice));
    TRI(cudaEventRecord(e0));
    triangulate_kernel<<<cdiv(n, 128), 128>>>(d_proj, d_f1, d_f2, d_uv1, d_uv2, d_out, n);
    TRI(cudaEventRecord(e1));
    TRI(cudaGetLastError());
    TRI(cudaMemcpy(points, d_out, n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (kernel_ms) {
        float ms = 0;
        TRI(cudaEventElapsedTime(&ms, e0, e1));
        *kernel_ms = ms;
    }
#undef TRI
    cleanup();
    return MMBA_OK;
}

// ---- rotate / project (no handle: free functions on one device) ------------------------------------
static int rows_op(int device, int64_t n, const double* pts, const double* params, int stride, const double* K, double* out, int out_cols) {
    mmba_handle* h = nullptr;
    if (n < 0 || (n > 0 && (!pts || !params || !out))) return fail(h, MMBA_ERR_ARG, "rotate/project: bad argument");
    if (n == 0) return MMBA_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(h, MMBA_ERR_CUDA, "no CUDA device (libmmba has no CPU fallback)");
    CU(cudaSetDevice(device));
    double *d_p = nullptr, *d_q = nullptr, *d_o = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_p);
        cudaFree(d_q);
        cudaFree(d_o);
    };
#define ROW(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cleanup();                                                                             \
            return fail(h, MMBA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
        }                                                                                          \
    } while (0)
    ROW(cudaMalloc(&d_p, n * 3 * sizeof(double)));
    ROW(cudaMalloc(&d_q, n * stride * sizeof(double)));
    ROW(cudaMalloc(&d_o, n * out_cols * sizeof(double)));
    ROW(cudaMemcpy(d_p, pts, n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    ROW(cudaMemcpy(d_q, params, n * stride * sizeof(double), cudaMemcpyHostToDevice));
    if (K)
        project_rows_kernel<<<cdiv(n, 256), 256>>>(d_p, d_q, stride, K[0], K[1], K[2], K[3], K[4], K[5], K[6], K[7], K[8], d_o, n);
    else
        rotate_rows_kernel<<<cdiv(n, 256), 256>>>(d_p, d_q, d_o, n);
    ROW(cudaGetLastError());
    ROW(cudaMemcpy(out, d_o, n * out_cols * sizeof(double), cudaMemcpyDeviceToHost));
#undef ROW
    cleanup();
    return MMBA_OK;
}

int mmba_rotate(int device, int64_t n, const double* points, const double* rot_vecs, double* out) {
    return rows_op(device, n, points, rot_vecs, 3, nullptr, out, 3);
}

int mmba_project(int device, int64_t n, const double* points, const double* frame_params, int64_t param_stride, const double K[9],
                 double* out) {
    if (!K || param_stride < 6) return fail(nullptr, MMBA_ERR_ARG, "project: null K or fewer than 6 parameters per row");
    return rows_op(device, n, points, frame_params, (int)param_stride, K, out, 2);
}

// ---- host-only functions ------------------------------------------------------------------------
int mmba_host_rcm_pattern(int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx, const int64_t* pt_idx,
                          int64_t sizes[3], int32_t* up_rowptr, int32_t* up_cols, int64_t capacity) {
    if (n_cams <= 0 || n_points <= 0 || n_obs <= 0 || !cam_idx || !pt_idx || !sizes)
        return fail(nullptr, MMBA_ERR_ARG, "rcm_pattern: bad argument");
    for (int64_t i = 0; i < n_obs; ++i)
        if (cam_idx[i] < 0 || cam_idx[i] >= n_cams || pt_idx[i] < 0 || pt_idx[i] >= n_points)
            return fail(nullptr, MMBA_ERR_ARG, "rcm_pattern: index out of range at observation " + std::to_string(i));
    RcmPattern r;
    // the same call mmba_set_problem makes: marking in the plan's internal point order
    Plan plan;
    std::string perr;
    const bool have_plan = build_plan(plan, n_cams, n_points, n_obs, cam_idx, pt_idx, 0, 1, perr) == MMBA_OK;
    if (!build_rcm_pattern(r, n_cams, n_points, n_obs, cam_idx, pt_idx, INT32_MAX / 36, have_plan ? plan.point_perm.data() : nullptr))
        return fail(nullptr, MMBA_ERR_NOMEM, "rcm_pattern: too many blocks");
    sizes[0] = r.nnz_up();
    sizes[1] = r.nnz_full();
    sizes[2] = r.total_pairs;
    if (up_rowptr) std::memcpy(up_rowptr, r.up_rowptr.data(), (n_cams + 1) * sizeof(int32_t));
    if (up_cols) {
        if (capacity < r.nnz_up()) return fail(nullptr, MMBA_ERR_NOMEM, "rcm_pattern: capacity too small");
        std::memcpy(up_cols, r.up_cols.data(), r.nnz_up() * sizeof(int32_t));
    }
    return MMBA_OK;
}

int mmba_host_tr2d(const double B[3], const double g[2], double delta, double p[2], int* newton) {
    if (!B || !g || !p) return MMBA_ERR_ARG;
    bool nw = false;
    tr2d(B, g, delta, p, &nw);
    if (newton) *newton = nw ? 1 : 0;
    return MMBA_OK;
}

int mmba_host_min_quadratic_1d(double a, double b, double lb, double ub, double* t, double* y) {
    if (!t || !y) return MMBA_ERR_ARG;
    min_quadratic_1d(a, b, lb, ub, t, y);
    return MMBA_OK;
}

int mmba_host_update_tr_radius(double delta, double actual, double predicted, double step_norm, int bound_hit,
                               double* delta_new, double* ratio) {
    if (!delta_new || !ratio) return MMBA_ERR_ARG;
    update_tr_radius(delta, actual, predicted, step_norm, bound_hit != 0, delta_new, ratio);
    return MMBA_OK;
}

int mmba_host_check_termination(double dF, double F, double dx_norm, double x_norm, double ratio, double ftol, double xtol) {
    return check_termination(dF, F, dx_norm, x_norm, ratio, ftol, xtol);
}

int mmba_plan_create(mmba_plan** out, int64_t n_cams, int64_t n_points, int64_t n_obs, const int64_t* cam_idx,
                     const int64_t* pt_idx, int rank, int nranks) {
    if (!out) return fail(nullptr, MMBA_ERR_ARG, "null output plan");
    *out = nullptr;
    mmba_plan* p = new mmba_plan();
    std::string err;
    int rc = build_plan(p->plan, n_cams, n_points, n_obs, cam_idx, pt_idx, rank, nranks, err);
    if (rc != MMBA_OK) {
        delete p;
        return fail(nullptr, rc, err);
    }
    *out = p;
    return MMBA_OK;
}

void mmba_plan_destroy(mmba_plan* p) { delete p; }

int mmba_plan_sizes(const mmba_plan* p, int64_t sizes[8]) {
    if (!p || !sizes) return MMBA_ERR_ARG;
    const Plan& pl = p->plan;
    sizes[0] = pl.n_tiles;
    sizes[1] = pl.n_obs_local;
    sizes[2] = pl.n_points_local();
    sizes[3] = pl.pt_begin;
    sizes[4] = pl.pt_end;
    sizes[5] = kTileObs;
    sizes[6] = pl.max_tile_cams;
    sizes[7] = pl.n_slots;
    return MMBA_OK;
}

int mmba_plan_export(const mmba_plan* p, int64_t* obs_perm, int64_t* point_perm, int32_t* slot_cam_global,
                     int32_t* slot_point_local) {
    if (!p) return MMBA_ERR_ARG;
    const Plan& pl = p->plan;
    if (obs_perm) std::copy(pl.slot_obs.begin(), pl.slot_obs.end(), obs_perm);
    if (point_perm)
        for (int64_t q = 0; q < pl.n_points; ++q) point_perm[q] = pl.point_perm[q];
    for (int64_t t = 0; t < pl.n_tiles; ++t) {
        for (int j = 0; j < kTileObs; ++j) {
            const int64_t s = t * kTileObs + j;
            const bool live = pl.slot_obs[s] >= 0;
            if (slot_cam_global) slot_cam_global[s] = live ? pl.tile_cams[t * pl.cam_stride + pl.meta[t].slot_cam[j]] : -1;
            if (slot_point_local) slot_point_local[s] = live ? pl.meta[t].pt0 + pl.meta[t].slot_pt[j] : -1;
        }
    }
    return MMBA_OK;
}

int mmba_plan_raw(const mmba_plan* p, void* meta, int32_t* tile_cams) {
    if (!p) return MMBA_ERR_ARG;
    const Plan& pl = p->plan;
    if (meta && pl.n_tiles) std::memcpy(meta, pl.meta.data(), (size_t)pl.n_tiles * sizeof(TileMeta));
    if (tile_cams)
        for (int64_t t = 0; t < pl.n_tiles; ++t)
            for (int c = 0; c < kTileObs; ++c)
                tile_cams[t * kTileObs + c] = c < pl.meta[t].ncams ? pl.tile_cams[t * pl.cam_stride + c] : -1;
    return MMBA_OK;
}

int mmba_get_plan_sizes(mmba_handle* h, int64_t sizes[8]) {
    TRY(need_problem(h));
    if (!sizes) return fail(h, MMBA_ERR_ARG, "get_plan_sizes: null argument");
    const DevPlan& D = h->dp;
    sizes[0] = D.n_tiles;
    sizes[1] = D.n_obs_local;
    sizes[2] = D.pt_end - D.pt_begin;
    sizes[3] = D.pt_begin;
    sizes[4] = D.pt_end;
    sizes[5] = kTileObs;
    sizes[6] = D.max_tile_cams;
    sizes[7] = D.n_slots;
    return MMBA_OK;
}

int mmba_get_plan_raw(mmba_handle* h, void* meta, int32_t* tile_cams, int64_t* obs_perm, int64_t* point_perm) {
    TRY(need_problem(h));
    const DevPlan& D = h->dp;
    if (meta && D.n_tiles) CU(cudaMemcpyAsync(meta, D.meta, (size_t)D.n_tiles * sizeof(TileMeta), cudaMemcpyDeviceToHost, h->stream));
    std::vector<int32_t> tc, so, pp;
    if (tile_cams) {
        tc.resize((size_t)D.n_tiles * D.cam_stride + 1);
        CU(cudaMemcpyAsync(tc.data(), D.tile_cams, (size_t)D.n_tiles * D.cam_stride * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    if (obs_perm) {
        so.resize((size_t)D.n_slots + 1);
        CU(cudaMemcpyAsync(so.data(), D.slot_obs, (size_t)D.n_slots * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    if (point_perm) {
        pp.resize((size_t)D.n_points);
        CU(cudaMemcpyAsync(pp.data(), D.point_perm, (size_t)D.n_points * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    if (tile_cams)
        for (int64_t t = 0; t < D.n_tiles; ++t)
            for (int c = 0; c < kTileObs; ++c) tile_cams[t * kTileObs + c] = c < D.cam_stride ? tc[t * D.cam_stride + c] : -1;
    if (obs_perm)
        for (int64_t i = 0; i < D.n_slots; ++i) obs_perm[i] = so[i];
    if (point_perm)
        for (int64_t i = 0; i < D.n_points; ++i) point_perm[i] = pp[i];
    return MMBA_OK;
}

int mmba_get_plan_stats(mmba_handle* h, int32_t* count, int32_t* first_cam, int32_t* last_cam, int32_t* first_hi) {
    TRY(need_problem(h));
    int *c, *f, *l, *fh;
    devplan_stat_arrays(h->planner, h->dp, &c, &f, &l, &fh);
    const size_t bytes = (size_t)h->dp.n_points * sizeof(int32_t);
    if (count) CU(cudaMemcpyAsync(count, c, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (first_cam) CU(cudaMemcpyAsync(first_cam, f, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (last_cam) CU(cudaMemcpyAsync(last_cam, l, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (first_hi) CU(cudaMemcpyAsync(first_hi, fh, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return MMBA_OK;
}

int mmba_get_rcm_pattern(mmba_handle* h, int64_t sizes[8], int32_t* up_rowptr, int32_t* up_cols, int64_t capacity) {
    TRY(need_problem(h));
    if (!sizes) return fail(h, MMBA_ERR_ARG, "get_rcm_pattern: null argument");
    if (!h->rcm_ready) return fail(h, MMBA_ERR_STATE, "get_rcm_pattern: the reduced camera matrix is not formed explicitly");
    const DevPlan& D = h->dp;
    sizes[0] = D.nnz_up;
    sizes[1] = D.nnz_full;
    sizes[2] = D.total_pairs;
    sizes[3] = D.n_ctas;
    sizes[4] = D.cpc;
    sizes[5] = D.nblk_max;
    sizes[6] = D.nh_max;
    sizes[7] = 0;
    if (up_rowptr) CU(cudaMemcpyAsync(up_rowptr, D.up_rowptr, (size_t)(h->Nc + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (up_cols) {
        if (capacity < D.nnz_up) return fail(h, MMBA_ERR_NOMEM, "get_rcm_pattern: capacity too small");
        CU(cudaMemcpyAsync(up_cols, D.up_cols, (size_t)D.nnz_up * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return MMBA_OK;
}

int mmba_plan_tile_stats(const mmba_plan* p, int32_t* n_cams, int32_t* n_points, int32_t* pair_mode, int32_t* n_pairs) {
    if (!p) return MMBA_ERR_ARG;
    const Plan& pl = p->plan;
    for (int64_t t = 0; t < pl.n_tiles; ++t) {
        if (n_cams) n_cams[t] = pl.meta[t].ncams;
        if (n_points) n_points[t] = pl.meta[t].npts;
        if (pair_mode) pair_mode[t] = pl.meta[t].pair_mode;
        if (n_pairs) n_pairs[t] = pl.meta[t].npairs;
    }
    return MMBA_OK;
}

}  // extern "C"
