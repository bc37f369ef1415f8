// libjabble_b200: host side of the C ABI declared in include/jabble_b200.h.
//
// A plan owns: the data block in HBM (canonicalised: every row ascending in x, padded to a
// multiple of the 128-pixel warp tile), the flat parameter vector, the fit map, the scratch for
// per-task partials, one CUDA stream.  An evaluation is six launches on that stream:
//   k_scatter_params -> k_prepare_rows | k_prepare_strip -> k_strip | k_fused2 | k_fused32 | k_fused ->
//   k_reduce_stage1 -> k_reduce_stage2 -> k_finish
// (plan_variant picks the fused kernel from the plan's layout, components, fit state and flags:
// k_strip, jb_strip.cuh, when the rows are many and regular enough for the column walk; else k_fused2
// for smooth tiles; else the general k_fused).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/jabble_b200.h"
#include "jb_kernels.cuh"
#include "jb_fused2.cuh"
#include "jb_fused32.cuh"
#include "jb_strip.cuh"

namespace {

thread_local std::string g_err;

// JB_TIMING=1: phases of plan creation on stderr (milliseconds)
struct PhaseTimer {
    bool on; double t0; const char* what;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit PhaseTimer(const char* w) : on(getenv("JB_TIMING") != nullptr), t0(on ? now() : 0.0), what(w) {}
    void lap(const char* next) { if (on) { const double t = now(); fprintf(stderr, "[jb timing] %-28s %8.2f ms\n", what, t - t0); t0 = t; what = next; } }
    ~PhaseTimer() { lap(""); }
};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define JB_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(JB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct jb_plan {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t E = 0, P = 0, ppad = 0;
    int n_tiles = 0;
    bool has_data = false;   // ys / yivar present
    double coef = 1.0;

    int n_comp = 0;            // components, stored globals (SPLINE) first, per-key (NORM_SPLINE) last
    int n_glob = 0, has_norm = 0;
    int ext_of[jb::kMaxComp] = {0, 1, 2};   // caller's index of the component stored at position i
    bool norm_smooth = true;   // every lane run (4 pixels) spans less than one knot of the normalisation grid
    jb_component_desc desc[jb::kMaxComp];
    jb::CompDev comp[jb::kMaxComp];

    std::vector<DevBuf> bufs;   // everything cudaMalloc'ed, freed in destroy
    int64_t device_bytes = 0, scratch_bytes = 0;

    double* d_data = nullptr;   // tile-interleaved x | y | ivar (jb_kernels.cuh)
    double* d_x = nullptr;      // [E][ppad] original x, canonical orientation (forward kernel)
    uint8_t* d_row_rev = nullptr;
    int* d_row_perm = nullptr;
    void* d_grouprec = nullptr;            // [E] RowRec, sorted-row order
    jb::TileInfo* d_tinfo = nullptr;       // [n_tiles][E], sorted-row order
    std::vector<double> h_txmin, h_txmax;  // [E][n_tiles]
    std::vector<uint8_t> h_flags;          // [E][n_tiles]
    std::vector<uint32_t> h_live;          // [E][n_tiles] lanes of the tile's warp that own a weighted pixel
    std::vector<double> h_xfirst;   // per row: smallest unmasked x (sort key), +inf if none
    std::vector<double> h_xrel;     // per row: offset of its pixel grid from a reference row's (+inf: nothing weighted); empty: undefined
    std::vector<uint8_t> h_row_rev;

    int64_t n_params = 0, n_fit = 0;
    bool fit_bufs = false;      // d_fit_map / d_pfit / d_out / h_pfit / h_out sized for n_fit
    std::vector<int64_t> h_fit_map;   // host copy of the fit map (jb_group: which outputs are sums over all rows)
    double *d_params = nullptr, *d_grad_all = nullptr, *d_fisher_all = nullptr;
    std::vector<double> h_params;
    std::vector<double> sorted_params;   // the parameters the rows were last sorted for
    bool params_set = false, sort_dirty = true;
    bool fit_shift[jb::kMaxComp] = {false, false, false};   // some shift of the component is in fit mode
    bool force_der = false;    // jb_plan_fisher: derivative sums whatever the fit flags say
    bool f2_ok = true;         // every tile suits k_fused2: live lanes smooth, lanes sharing a window copy never on one knot
    bool general_only = false; // JB_PLAN_GENERAL_KERNEL
    bool f32 = false;          // JB_PLAN_FLOAT32: float records, k_fused32
    float* d_data32 = nullptr; // tile-interleaved x_rel | y | ivar as floats (f32 plans only)
    bool force_v1 = false;     // a batch whose plans disagree on the kernel variant falls back to k_fused
    bool force_nostrip = false; // a batch whose plans disagree on k_strip (variant or rows per block) leaves it
    int variant = -1;          // fused kernel of this plan: -1 k_fused (general), else the FL flags of k_fused2
    int64_t* d_fit_map = nullptr;
    double *d_pfit = nullptr, *d_out = nullptr;
    double *h_pfit = nullptr, *h_out = nullptr;   // pinned
    int *d_status = nullptr, *h_status = nullptr; // h pinned

    int G = 0, n_groups = 0, wincap = 0;
    int64_t n_tasks = 0;
    double *d_part = nullptr, *d_rowpart = nullptr, *d_tmp = nullptr, *d_rowsum = nullptr, *d_losspart = nullptr;
    int* d_tile_rng = nullptr;
    double *d_normpart = nullptr, *d_normrow = nullptr;
    int norm_stride = 0;
    int* d_part_base = nullptr;
    int64_t max_knots = 0;
    int* d_key_ptr[jb::kMaxComp][3] = {};   // [comp][0 shift, 1 stretch, 2 per-key control points] CSR rowptr
    int* d_key_rows[jb::kMaxComp][3] = {};
    int* d_keys[jb::kMaxComp][3] = {};      // device copies of shift/stretch/ctrl key arrays
    std::vector<int> h_shift_key[jb::kMaxComp];   // host copy of the shift keys (jb_plan_grid_search sums rows into keys)

    // ---- strip path (k_strip, jb_strip.cuh): rows in sorted order, 32-pixel strips
    bool strip_allowed = true;      // not JB_PLAN_NO_STRIP / JB_NO_STRIP, model shape supported
    bool strip_ok = false;          // layout built for the current sort and its preconditions hold
    bool strip_disabled = false;    // a precondition failed at run time even after a re-sort: stay on k_fused2 / k_fused
    int strip_fail = 0;             // consecutive evaluations that raised a strip precondition
    int strip_reason = 0;           // jb_plan_info_t.strip_reason
    double s_tol = 0.05;            // StripParams::tol of this plan
    int sK_force = 0;               // rows per block imposed by a batch (its plans share one launch, hence one K)
    int sK = 0, sBPT = 0, s_nblocks = 0, s_nrgroups = 0, s_ncls = 1, s_nstrips = 0, s_hint = 0;
    int64_t s_ntasks = 0;
    double *d_sdata = nullptr, *d_sxoff = nullptr, *d_sxoff_rng = nullptr, *d_sxref = nullptr;
    int* d_srowid = nullptr;
    void* d_srowc = nullptr;
    jb::StripBlk* d_sblk = nullptr;
    unsigned long long* d_sstats = nullptr;
    double *d_spart = nullptr, *d_srowpart = nullptr, *d_stmp = nullptr, *d_slosspart = nullptr;
    int *d_spart_base = nullptr, *d_stile_rng = nullptr;
    int64_t s_scratch_tasks = 0;    // tasks the strip scratch was allocated for

    int64_t launches = 0, evaluations = 0, serial_rows = 0;
    int64_t tiles_kind[4] = {0, 0, 0, 0};   // smooth, general, serial, empty

    // device-resident parameter blocks of the evaluation kernels (one launch can serve many plans)
    struct Pack { jb::ScatterParams sc; jb::PrepareParams pr; jb::FusedParams fu; jb::ReduceParams re; jb::StripParams st; jb::StripPrepParams sp; };
    Pack* h_pack = nullptr;   // pinned
    Pack* d_pack = nullptr;
    int* d_off = nullptr;     // [6][2] block ranges {0, nb} of the six kernels for this plan alone
    int nb[6] = {0, 0, 0, 0, 0, 0};
    bool pack_dirty = true;

    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;   // pool, reused
    size_t prof_used = 0;
    cudaStream_t prof_stream = nullptr;
};

struct jb_batch {
    std::vector<jb_plan*> plans;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<const double*> p_fit;
    std::vector<double*> out;
    std::vector<void*> dev;
    jb::ScatterParams* d_sc = nullptr; jb::PrepareParams* d_pr = nullptr;
    jb::FusedParams* d_fu = nullptr; jb::ReduceParams* d_re = nullptr;
    jb::StripParams* d_st = nullptr; jb::StripPrepParams* d_sp = nullptr;
    int* d_off = nullptr;
    int total[6] = {0, 0, 0, 0, 0, 0};
    bool scatter = true, dirty = true;
    int64_t launches = 0;
    // jb_batch_eval_host: flat device buffers + pinned staging for all plans
    double *d_flat_p = nullptr, *d_flat_out = nullptr, *h_flat_p = nullptr, *h_flat_out = nullptr;
    std::vector<int64_t> off_p, off_out;
    int** d_stat_ptrs = nullptr; int* d_stat_flat = nullptr; int* h_stat_flat = nullptr;
    bool host_bound = false, host_pending = false;
    bool soft_range = false;   // jb_batch_soft_range: an out-of-range plan yields a NaN loss instead of an error
    // The launch sequence of an evaluation recorded as a CUDA graph once nothing changes between two calls
    // (same packs, same stream) and replayed from then on: one launch per evaluation instead of six.
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t gstream = nullptr;
    int steady = 0;
    bool graphs_ok = true;
};

namespace {

template <typename T>
int dev_alloc(jb_plan* pl, T** out, size_t count, bool scratch = false) {
    DevBuf b;
    b.bytes = std::max<size_t>(count * sizeof(T), 16);
    JB_CUDA(cudaMalloc(&b.p, b.bytes));
    pl->bufs.push_back(b);
    pl->device_bytes += (int64_t)b.bytes;
    if (scratch) pl->scratch_bytes += (int64_t)b.bytes;
    *out = static_cast<T*>(b.p);
    return JB_OK;
}

// uninitialised host memory for the canonical copy of a block (hundreds of MB: the loop that fills it
// writes every element, in parallel -- no single-threaded zero fill, pages first touched by their writer)
template <typename T>
struct HostBuf {
    T* p = nullptr;
    size_t n = 0;
    explicit HostBuf(size_t n_) : p(static_cast<T*>(malloc(std::max<size_t>(n_ * sizeof(T), 16)))), n(n_) {}
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
    ~HostBuf() { free(p); }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

template <typename T>
int dev_upload(jb_plan* pl, T** out, const HostBuf<T>& h) {
    int rc = dev_alloc(pl, out, h.size());
    if (rc) return rc;
    if (!h.empty()) JB_CUDA(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return JB_OK;
}

template <typename T>
int dev_upload(jb_plan* pl, T** out, const std::vector<T>& h) {
    int rc = dev_alloc(pl, out, h.size());
    if (rc) return rc;
    if (!h.empty()) JB_CUDA(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return JB_OK;
}

// free one of the plan's device buffers (nullptr is fine) and forget it
template <typename T>
void dev_free(jb_plan* pl, T*& ptr) {
    if (!ptr) return;
    for (size_t i = 0; i < pl->bufs.size(); ++i)
        if (pl->bufs[i].p == (void*)ptr) {
            pl->device_bytes -= (int64_t)pl->bufs[i].bytes;
            cudaFree(pl->bufs[i].p);
            pl->bufs.erase(pl->bufs.begin() + i);
            break;
        }
    ptr = nullptr;
}

int choose_rows_per_group(int64_t E, int n_tiles, int requested) {
    if (requested > 0) return (int)std::min<int64_t>(std::min(requested, jb::kMaxGroup), std::max<int64_t>(E, 1));
    // A task pays its set-up (row records, control-point window, final window write) once however
    // many rows it streams, so groups are as long as possible (kMaxGroup) -- unless the whole plan
    // then leaves warp slots empty: a lone plan is one wave of equal tasks, and its time is the time
    // of one task.  16 warps are resident per SM (128 registers, < 14 KB of shared memory each).
    const int64_t slots = 148LL * 16;
    const int64_t groups_that_fit = std::max<int64_t>(1, slots / std::max(n_tiles, 1));
    int64_t g = (E + groups_that_fit - 1) / groups_that_fit;
    g = std::max<int64_t>(std::min<int64_t>(g, jb::kMaxGroup), std::min<int64_t>(8, jb::kMaxGroup));
    return (int)std::min<int64_t>(g, std::max<int64_t>(E, 1));
}

int alloc_scratch(jb_plan* pl) {
    // (re)allocate the buffers whose size depends on rows_per_group
    pl->n_groups = (int)((pl->E + pl->G - 1) / pl->G);
    pl->n_tasks = (int64_t)pl->n_groups * pl->n_tiles;
    int rc;
    JB_CUDA(cudaStreamSynchronize(pl->stream));
    dev_free(pl, pl->d_part); dev_free(pl, pl->d_part_base); dev_free(pl, pl->d_losspart);
    if ((rc = dev_alloc(pl, &pl->d_part, (size_t)pl->n_glob * pl->n_tasks * pl->wincap, true))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_part_base, (size_t)pl->n_glob * pl->n_tasks, true))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_losspart, (size_t)pl->n_tasks, true))) return rc;
    pl->pack_dirty = true;
    return JB_OK;
}

constexpr int kStripWin = 64;      // doubles per published gradient window of a k_strip task (>= kStripWs + p_val + 1)
constexpr int kStripRedThreads = 128;   // threads per block of k_reduce_stage1 for k_strip plans (a strip's windows span ~70 knots)

// Does the model shape suit k_strip at all?
bool strip_shape_ok(const jb_plan* pl) {
    static const bool env_off = getenv("JB_NO_STRIP") != nullptr;
    return pl->strip_allowed && !env_off && !pl->general_only && !pl->f32 && pl->has_data && !pl->has_norm && pl->n_glob >= 1 &&
           pl->n_glob <= 2 && pl->ppad / jb::kStripPx <= jb::kRedBases && !pl->h_xrel.empty();
}

// (Re)build the strip layout for the row order `perm` (sorted by `key`): pick K from the key ranges
// of the blocks, lay the rows out in strips on the device, derive the column offsets and check the
// static preconditions of k_strip (jb_strip.cuh).  Sets pl->strip_ok.
int strip_build(jb_plan* pl, const std::vector<int>& perm) {
    pl->strip_ok = false;
    pl->pack_dirty = true;
    pl->strip_reason = pl->strip_disabled ? 6 : 1;
    if (!strip_shape_ok(pl) || pl->strip_disabled) return JB_OK;
    const int64_t E = pl->E;
    const int NC = pl->n_glob;
    // per-component keys of the sorted rows
    std::vector<double> xref(E);
    std::vector<std::vector<double>> key(NC, std::vector<double>(E));
    for (int64_t pos = 0; pos < E; ++pos) {
        const int r = perm[pos];
        xref[pos] = pl->h_xrel[r];
        for (int c = 0; c < NC; ++c) {
            const jb_component_desc& cd = pl->desc[c];
            const double sh = cd.shift_offset >= 0 ? pl->h_params[cd.shift_offset + pl->h_shift_key[c][r]] : 0.0;
            key[c][pos] = xref[pos] - sh;
        }
    }
    const int n_strips = (int)(pl->ppad / jb::kStripPx);
    int rc;
    cudaStream_t st = pl->stream;
    if (!pl->d_sdata) {
        const int max_blocks = (int)((E + 7) / 8);
        if ((rc = dev_alloc(pl, &pl->d_sdata, (size_t)n_strips * E * jb::kStripRec))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_sxoff, (size_t)n_strips * max_blocks * 32))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_sxoff_rng, (size_t)n_strips * max_blocks * 2))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_sxref, (size_t)E))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_srowid, (size_t)E))) return rc;
        {
            char* p = nullptr;
            if ((rc = dev_alloc(pl, &p, (size_t)E * sizeof(jb::StripRow<2>)))) return rc;
            pl->d_srowc = p;
        }
        if ((rc = dev_alloc(pl, &pl->d_sblk, (size_t)max_blocks))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_sstats, (size_t)16))) return rc;
    }
    double dx_max = 0.0, dx_min = INFINITY;
    for (int c = 0; c < NC; ++c) { dx_max = std::max(dx_max, pl->desc[c].dx); dx_min = std::min(dx_min, pl->desc[c].dx); }
    JB_CUDA(cudaMemcpyAsync(pl->d_srowid, perm.data(), E * sizeof(int), cudaMemcpyHostToDevice, st));
    JB_CUDA(cudaMemcpyAsync(pl->d_sxref, xref.data(), E * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        jb::StripLayoutParams lp;
        lp.tiled = pl->d_data; lp.rowid = pl->d_srowid; lp.sdata = pl->d_sdata; lp.n_rows = E; lp.n_tiles = pl->n_tiles;
        const int64_t warps = E * pl->n_tiles;
        jb::k_strip_layout<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(lp);
        JB_CUDA(cudaGetLastError());
        pl->launches++;
    }
    unsigned long long stats[16];
    // column offsets of the blocks of K rows + the checks of preconditions (1) and (3)
    auto run_xoff = [&](int K) -> int {
        unsigned long long init[16];
        for (int i = 0; i < 16; ++i) init[i] = i >= 3 ? ~0ULL : 0ULL;     // [3..10]: minima
        JB_CUDA(cudaMemcpyAsync(pl->d_sstats, init, sizeof init, cudaMemcpyHostToDevice, st));
        jb::StripXoffParams xp;
        xp.sdata = pl->d_sdata; xp.xref = pl->d_sxref; xp.xoff = pl->d_sxoff; xp.n_rows = E; xp.n_strips = n_strips; xp.K = K;
        xp.n_blocks = (int)((E + K - 1) / K); xp.tol_x = 0.999 * jb::kStripTol * dx_min; xp.stats = pl->d_sstats;
        const int64_t w2 = (int64_t)n_strips * xp.n_blocks;
        jb::k_strip_xoff<<<(unsigned)((w2 + 7) / 8), 256, 0, st>>>(xp);
        JB_CUDA(cudaGetLastError());
        pl->launches++;
        JB_CUDA(cudaMemcpyAsync(stats, pl->d_sstats, sizeof stats, cudaMemcpyDeviceToHost, st));
        JB_CUDA(cudaStreamSynchronize(st));
        return JB_OK;
    };
    // first with the longest blocks: how far do the rows' pixel grids deviate from one another?  That
    // sets the plan's tolerance, and with it how far apart the keys of a block may be.
    if ((rc = run_xoff(32))) return rc;
    pl->strip_reason = 3;
    if (stats[0] != 0) return JB_OK;           // precondition (1): the pixel grid differs from row to row
    double dev;
    memcpy(&dev, &stats[2], sizeof dev);
    const double tol = std::min(jb::kStripTol, 1.5 * dev / dx_min + 2.0e-3);
    pl->s_tol = tol;
    // rows per block: the largest K whose blocks keep every component's keys 0.04 knots inside the
    // 1 - 2 tol that k_prepare_strip refuses at run time (the margin is for the optimiser's steps)
    // (K <= 20: a block's row records live in shared memory, and 20 rows is what leaves room for 16 warps per SM)
    int K = 0;
    for (int cand : {20, 16, 12, 8}) {
        if (pl->sK_force && cand > pl->sK_force) continue;
        bool ok = true;
        for (int64_t r0 = 0; r0 < E && ok; r0 += cand)
            for (int c = 0; c < NC && ok; ++c) {
                double lo = INFINITY, hi = -INFINITY;
                for (int64_t pos = r0; pos < std::min<int64_t>(E, r0 + cand); ++pos)
                    if (std::isfinite(xref[pos])) { lo = std::min(lo, key[c][pos]); hi = std::max(hi, key[c][pos]); }
                if (lo <= hi && !((hi - lo) / pl->desc[c].dx < 1.0 - 2.0 * tol - 0.04)) ok = false;
            }
        if (ok) { K = cand; break; }
    }
    pl->strip_reason = 2;
    if (K == 0) return JB_OK;      // rows too far apart for the column walk: k_fused2 / k_fused
    if (pl->s_hint <= 0 && !pl->sK_force) {
        // a lone plan runs as ONE wave of equal tasks of whole blocks: its time is the longest task's, rows
        // plus ~4 rows' worth of checkpoint per block.  A shorter block can cut the rows per task finer
        // (2000 rows on 18 row groups: 6 x 20 = 120 rows, but 7 x 16 = 112).
        const int64_t groups_fit = std::max<int64_t>(1, 148LL * 16 / n_strips);
        double best = INFINITY;
        int best_k = K;
        for (int cand : {20, 16, 12, 8}) {
            if (cand > K) continue;
            const int64_t nb = (E + cand - 1) / cand;
            const double cost = (double)((nb + groups_fit - 1) / groups_fit) * (cand + 4.2);
            if (cost < best - 1.0e-9) { best = cost; best_k = cand; }
        }
        K = best_k;
    }
    const int n_blocks = (int)((E + K - 1) / K);
    if (K != 32 && (rc = run_xoff(K))) return rc;
    pl->strip_reason = 3;
    if (stats[0] != 0) return JB_OK;
    int ncls = 0;
    for (int nn = 1; nn <= 8 && !ncls; ++nn) {
        double gap;
        if (stats[2 + nn] == ~0ULL) gap = INFINITY;        // no two weighted columns that far apart
        else memcpy(&gap, &stats[2 + nn], sizeof gap);
        if (gap >= dx_max * (1.0 + 1.0e-9)) ncls = nn;
    }
    pl->strip_reason = 4;
    if (!ncls) return JB_OK;                   // precondition (3): more than 8 pixels per knot
    double span;
    memcpy(&span, &stats[1], sizeof span);
    // rows per task.  The window of a task holds kStripWs base intervals: the strip itself plus the drift
    // of the keys over the task's blocks.  Within that, a lone plan wants at least one task per warp
    // slot (148 SMs x 16 warps); a batch (hint > 0) wants long tasks.
    const double room = (double)jb::kStripWs - 3.0 - span / dx_min;      // knots left for the drift
    pl->strip_reason = 5;
    if (!(room > 1.0)) return JB_OK;
    int bpt_fit = 0;
    {
        for (int cand = 32; cand >= 1 && !bpt_fit; --cand) {
            bool ok = true;
            for (int b0 = 0; b0 < n_blocks && ok; b0 += cand)
                for (int c = 0; c < NC && ok; ++c) {
                    double lo = INFINITY, hi = -INFINITY;
                    const int64_t p0 = (int64_t)b0 * K, p1 = std::min<int64_t>(E, (int64_t)(b0 + cand) * K);
                    for (int64_t pos = p0; pos < p1; ++pos)
                        if (std::isfinite(xref[pos])) { lo = std::min(lo, key[c][pos]); hi = std::max(hi, key[c][pos]); }
                    if (lo <= hi && !((hi - lo) / pl->desc[c].dx < room)) ok = false;
                }
            if (ok) bpt_fit = cand;
        }
    }
    if (!bpt_fit) return JB_OK;
    int bpt;
    if (pl->s_hint > 0) bpt = 12;       // a batch keeps every warp slot busy anyway: long tasks (their set-up is ~15 % of a 120-row task;
                                        // 64 chunks: 2.838 ms with 12 blocks per task, 2.819 with 16 where the knot window has room for them)
    else {
        // as many row groups as fill the warp slots ONCE: a lone plan is one wave of equal tasks, and a few
        // tasks spilling into a second wave would double its time
        const int64_t slots = 148LL * 16;
        const int64_t groups_fit = std::max<int64_t>(1, slots / n_strips);
        bpt = (int)std::max<int64_t>(1, (n_blocks + groups_fit - 1) / groups_fit);
    }
    // (a batched plan keeps one block of the window's room for the optimiser's steps: an overflow costs a re-sort)
    bpt = std::min(bpt, pl->s_hint > 0 ? std::max(1, bpt_fit - 1) : bpt_fit);
    pl->sK = K; pl->sBPT = bpt; pl->s_nblocks = n_blocks; pl->s_nstrips = n_strips; pl->s_ncls = ncls;
    pl->s_nrgroups = (n_blocks + bpt - 1) / bpt;
    pl->s_ntasks = (int64_t)pl->s_nrgroups * n_strips;
    {
        jb::StripRngParams rp;
        rp.xoff = pl->d_sxoff; rp.rng = pl->d_sxoff_rng; rp.n_strips = n_strips; rp.n_blocks = n_blocks;
        rp.blocks_per_task = bpt; rp.n_rgroups = pl->s_nrgroups;
        jb::k_strip_rng<<<(unsigned)((pl->s_ntasks + 7) / 8), 256, 0, st>>>(rp);
        JB_CUDA(cudaGetLastError());
        pl->launches++;
    }
    if (pl->s_scratch_tasks < pl->s_ntasks || !pl->d_spart) {
        JB_CUDA(cudaStreamSynchronize(st));
        dev_free(pl, pl->d_spart); dev_free(pl, pl->d_spart_base); dev_free(pl, pl->d_slosspart);
        if ((rc = dev_alloc(pl, &pl->d_spart, (size_t)NC * pl->s_ntasks * kStripWin, true))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_spart_base, (size_t)NC * pl->s_ntasks, true))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_slosspart, (size_t)pl->s_ntasks, true))) return rc;
        pl->s_scratch_tasks = pl->s_ntasks;
    }
    if (!pl->d_srowpart) {
        const size_t nq = 8;
        if ((rc = dev_alloc(pl, &pl->d_srowpart, (size_t)E * n_strips * nq, true))) return rc;
        JB_CUDA(cudaMemsetAsync(pl->d_srowpart, 0, (size_t)E * n_strips * nq * sizeof(double), st));
        if ((rc = dev_alloc(pl, &pl->d_stmp, (size_t)NC * n_strips * pl->max_knots, true))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_stile_rng, (size_t)2 * NC * n_strips, true))) return rc;
    }
    pl->strip_ok = true;
    pl->strip_reason = 0;
    return JB_OK;
}

// Sort rows by (smallest unmasked x) - (shift of the first component that has one), so that the
// rows of a group land on nearly the same knots and the on-chip window stays small.
int resort_rows(jb_plan* pl) {
    PhaseTimer phase_all("resort: sort, tile info");
    std::vector<double> key(pl->E);
    const jb_component_desc* sc = nullptr;
    for (int c = 0; c < pl->n_comp; ++c)
        if (pl->desc[c].shift_offset >= 0) { sc = &pl->desc[c]; break; }
    const std::vector<int>* skey = (sc && pl->params_set) ? &pl->h_shift_key[sc - pl->desc] : nullptr;
    const std::vector<double>& xpos = pl->h_xrel.empty() ? pl->h_xfirst : pl->h_xrel;     // where the row's pixel grid sits
    for (int64_t r = 0; r < pl->E; ++r) {
        double k = xpos[r];
        if (skey && !skey->empty()) k -= pl->h_params[sc->shift_offset + (*skey)[r]];
        key[r] = k;
    }
    std::vector<int> perm(pl->E);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<jb::TileInfo> tinfo((size_t)pl->n_tiles * pl->E);
    for (int t = 0; t < pl->n_tiles; ++t)
        for (int64_t pos = 0; pos < pl->E; ++pos) {
            const size_t src = (size_t)perm[pos] * pl->n_tiles + t;
            jb::TileInfo& ti = tinfo[(size_t)t * pl->E + pos];
            ti.xmin = pl->h_txmin[src]; ti.xmax = pl->h_txmax[src]; ti.flags = pl->h_flags[src];
            ti.live = pl->h_live[src]; ti.pad_[0] = ti.pad_[1] = 0;
        }
    JB_CUDA(cudaMemcpyAsync(pl->d_row_perm, perm.data(), pl->E * sizeof(int), cudaMemcpyHostToDevice, pl->stream));
    JB_CUDA(cudaMemcpyAsync(pl->d_tinfo, tinfo.data(), tinfo.size() * sizeof(jb::TileInfo), cudaMemcpyHostToDevice, pl->stream));
    JB_CUDA(cudaStreamSynchronize(pl->stream));
    pl->sort_dirty = false;
    pl->sorted_params = pl->h_params;
    if (pl->params_set) {
        PhaseTimer phase("resort: strip layout");
        int rc = strip_build(pl, perm);
        if (rc) return rc;
    }
    return JB_OK;
}

// Kernel ids in launch order.
enum { K_SCATTER = 0, K_PREPARE, K_FUSED, K_STAGE1, K_STAGE2, K_FINISH, K_COUNT };

int check_supported_for_eval(const jb_plan* pl) {
    if (!pl->has_data) return fail(JB_ERR_BAD_ARG, "plan was created without ys/yivar (forward only)");
    if (pl->f32 && (pl->has_norm || pl->n_glob > 2 || !pl->f2_ok))
        return fail(JB_ERR_UNSUPPORTED, "a float32 plan needs smooth tiles (consecutive pixels less than a knot apart) and no NormalizationModel component");
    if (pl->n_glob < 1 || pl->n_glob > 2)
        return fail(JB_ERR_UNSUPPORTED, "%d CardinalSplineMixture components (the fused kernel takes 1 or 2)", pl->n_glob);
    if (pl->n_comp - pl->n_glob > 1) return fail(JB_ERR_UNSUPPORTED, "more than one NormalizationModel component");
    for (int c = 0; c < pl->n_glob; ++c)
        if (pl->desc[c].p_val != pl->desc[0].p_val)
            return fail(JB_ERR_UNSUPPORTED, "components with different spline degrees (%d vs %d)", pl->desc[c].p_val, pl->desc[0].p_val);
    if (pl->has_norm) {
        if (pl->desc[pl->n_glob].n_knots > jb::kNormMaxKnots)
            return fail(JB_ERR_UNSUPPORTED, "NormalizationModel with %lld knots per key (<= %d supported)", (long long)pl->desc[pl->n_glob].n_knots, jb::kNormMaxKnots);
        if (!pl->norm_smooth)
            return fail(JB_ERR_UNSUPPORTED, "normalisation knot grid finer than 4 pixels somewhere");
    }
    return JB_OK;
}

// Which fused kernel serves the plan.  k_fused2 (jb_fused2.cuh) takes plans without a per-key
// component whose every tile is smooth; its flags say what the fit state asks for.  JB_FUSED_V1 in
// the environment keeps the general kernel (A/B measurements).
int plan_variant(const jb_plan* pl) {
    static const bool env_v1 = getenv("JB_FUSED_V1") != nullptr;
    int fl = 0;
    for (int c = 0; c < pl->n_glob && c < 2; ++c) {
        if (pl->desc[c].shift_offset >= 0 && (pl->fit_shift[c] || pl->force_der)) fl |= 1 << c;
        if (pl->desc[c].stretch_offset >= 0) fl |= 4 << c;
    }
    // k_strip (bit 7): cubic splines get the kernel of their fit state, the other degrees one kernel
    // per shape that forms every sum (a component without a StretchingModel has stretch 1)
    if (pl->strip_ok && !pl->strip_disabled && !pl->force_nostrip && !pl->force_v1 && !env_v1 && strip_shape_ok(pl))
        return 128 | (pl->desc[0].p_val == 3 ? fl : (pl->n_glob == 1 ? 5 : 15));
    if ((env_v1 && !pl->f32) || pl->general_only || pl->force_v1 || !pl->f2_ok || pl->has_norm || pl->n_glob > 2) return -1;
    // float32 tier (bit 6): cubic splines get the kernel of their fit state, the other degrees one
    // kernel per shape that forms every sum
    if (pl->f32) return 64 | (pl->desc[0].p_val == 3 ? fl : (pl->n_glob == 1 ? 5 : 15));
    return fl;
}

// (Re)build the plan's parameter blocks and block counts; uploaded on the plan's stream.
int refresh_pack(jb_plan* pl) {
    if (!pl->pack_dirty) return JB_OK;
    pl->variant = plan_variant(pl);
    JB_CUDA(cudaStreamSynchronize(pl->stream));   // the pinned copy may still be in flight
    jb_plan::Pack& k = *pl->h_pack;
    memset(&k, 0, sizeof k);
    k.sc.p_fit = pl->d_pfit; k.sc.fit_map = pl->d_fit_map; k.sc.params = pl->d_params; k.sc.n_fit = pl->n_fit;
    k.pr.row_perm = pl->d_row_perm; k.pr.grouprec = pl->d_grouprec; k.pr.n_rows = pl->E; k.pr.n_comp = pl->n_comp;
    k.pr.status = pl->d_status;
    jb::FusedParams& fp = k.fu;
    fp.data = pl->f32 ? reinterpret_cast<const double*>(pl->d_data32) : pl->d_data;
    fp.grouprec = pl->d_grouprec; fp.tinfo = pl->d_tinfo;
    fp.n_rows = pl->E; fp.n_tiles = pl->n_tiles;
    fp.rows_per_group = pl->G; fp.n_groups = pl->n_groups; fp.n_tasks = pl->n_tasks; fp.wincap = pl->wincap;
    fp.n_comp = pl->n_glob;
    fp.normpart = pl->d_normpart; fp.norm_stride = pl->norm_stride;
    fp.part = pl->d_part; fp.part_base = pl->d_part_base; fp.rowpart = pl->d_rowpart;
    fp.losspart = pl->d_losspart; fp.status = pl->d_status;
    jb::ReduceParams& rp = k.re;
    rp.n_comp = pl->n_glob;
    rp.has_norm = pl->has_norm;
    if (pl->has_norm) {
        const int c = pl->n_glob;
        rp.norm_stride = pl->norm_stride; rp.norm_knots = (int)pl->desc[c].n_knots; rp.n_norm_keys = (int)pl->desc[c].n_ctrl_keys;
        rp.norm_ctrl_off = pl->desc[c].ctrl_offset;
        rp.normpart = pl->d_normpart; rp.normrow = pl->d_normrow;
        rp.norm_rowptr = pl->d_key_ptr[c][2]; rp.norm_rows = pl->d_key_rows[c][2];
    }
    for (int c = 0; c < pl->n_comp; ++c) {
        k.pr.comp[c] = fp.comp[c] = rp.comp[c] = pl->comp[c];
        rp.ctrl_off[c] = pl->desc[c].ctrl_offset;
        rp.shift_off[c] = pl->desc[c].shift_offset;
        rp.stretch_off[c] = pl->desc[c].stretch_offset;
        rp.n_shift[c] = pl->desc[c].shift_offset >= 0 ? (int)pl->desc[c].n_shift : 0;
        rp.n_stretch[c] = pl->desc[c].stretch_offset >= 0 ? (int)pl->desc[c].n_stretch : 0;
        rp.shift_rowptr[c] = pl->d_key_ptr[c][0];   rp.shift_rows[c] = pl->d_key_rows[c][0];
        rp.stretch_rowptr[c] = pl->d_key_ptr[c][1]; rp.stretch_rows[c] = pl->d_key_rows[c][1];
    }
    const bool strip = pl->variant >= 0 && (pl->variant & 128);
    rp.n_rows = pl->E; rp.n_tasks = pl->n_tasks; rp.n_tiles = pl->n_tiles; rp.n_groups = pl->n_groups; rp.wincap = pl->wincap;
    rp.part = pl->d_part; rp.part_base = pl->d_part_base; rp.rowpart = pl->d_rowpart; rp.losspart = pl->d_losspart;
    rp.tilepart = pl->d_tmp; rp.tile_rng = pl->d_tile_rng; rp.max_knots = pl->max_knots;
    if (strip) {
        // the reducers take the partials of k_strip as they take those of k_fused2: a "tile" is a strip,
        // a "group" the row range of a task
        rp.n_tasks = pl->s_ntasks; rp.n_tiles = pl->s_nstrips; rp.n_groups = pl->s_nrgroups; rp.wincap = kStripWin;
        rp.part = pl->d_spart; rp.part_base = pl->d_spart_base; rp.rowpart = pl->d_srowpart; rp.losspart = pl->d_slosspart;
        rp.tilepart = pl->d_stmp; rp.tile_rng = pl->d_stile_rng; rp.row_of_pos = pl->d_srowid;
        jb::StripParams& sp = k.st;
        sp.data = pl->d_sdata; sp.xoff = pl->d_sxoff; sp.xoff_rng = pl->d_sxoff_rng; sp.rowc = pl->d_srowc; sp.blk = pl->d_sblk;
        sp.rowid = pl->d_srowid; sp.n_rows = pl->E; sp.n_strips = pl->s_nstrips; sp.K = pl->sK; sp.blocks_per_task = pl->sBPT;
        sp.n_blocks = pl->s_nblocks; sp.n_rgroups = pl->s_nrgroups; sp.ncls = pl->s_ncls; sp.n_tasks = pl->s_ntasks;
        sp.wincap = kStripWin; sp.n_comp = pl->n_glob; sp.tol = pl->s_tol;
        sp.part = pl->d_spart; sp.part_base = pl->d_spart_base; sp.rowpart = pl->d_srowpart; sp.losspart = pl->d_slosspart;
        sp.status = pl->d_status;
        jb::StripPrepParams& pp = k.sp;
        pp.rowid = pl->d_srowid; pp.xref = pl->d_sxref; pp.rowc = pl->d_srowc; pp.blk = pl->d_sblk; pp.n_rows = pl->E;
        pp.K = pl->sK; pp.n_blocks = pl->s_nblocks; pp.n_comp = pl->n_glob; pp.status = pl->d_status; pp.tol = pl->s_tol;
        for (int c = 0; c < pl->n_glob; ++c) sp.comp[c] = pp.comp[c] = pl->comp[c];
    }
    rp.rowsum = pl->d_rowsum;
    rp.coef = pl->coef;
    rp.grad_all = pl->d_grad_all; rp.fisher_all = pl->d_fisher_all;
    rp.fit_map = pl->d_fit_map; rp.n_fit = pl->n_fit;
    rp.out = pl->d_out; rp.status = pl->d_status;
    int NQ = jb::row_sum_stride(pl->n_glob, pl->has_norm);
    for (int c = 0; c < jb::kMaxComp; ++c)
        for (int q = 0; q < 3; ++q) rp.rq[c][q] = -1;
    if (pl->variant < 0) {
        for (int c = 0; c < pl->n_glob; ++c)
            for (int q = 0; q < 3; ++q) rp.rq[c][q] = c * jb::kRowQ + q;
        if (pl->has_norm) { rp.rq[pl->n_glob][0] = pl->n_glob * jb::kRowQ; rp.rq[pl->n_glob][1] = pl->n_glob * jb::kRowQ + 1; }
    } else {
        const int fl = pl->variant & 15;       // bit 6 marks the float32 tier, bit 7 k_strip
        NQ = strip ? std::max(1, jb::f2_nsums(pl->n_glob, fl)) : jb::f2_pad(jb::f2_nsums(pl->n_glob, fl));     // k_strip does not pad its records
        for (int c = 0; c < pl->n_glob; ++c)
            for (int q = 0; q < 3; ++q) rp.rq[c][q] = jb::f2_slot(pl->n_glob, fl, c, q);
    }
    rp.nq = NQ;
    rp.nb_ctrl = pl->n_glob * rp.n_tiles;
    const int red_threads = strip ? kStripRedThreads : jb::kRedSlab;      // k_reduce_stage1's block: one thread per knot of a tile's window union
    rp.nb_stage1_rows = (int)((pl->E + red_threads / 32 - 1) / (red_threads / 32));     // a warp per row
    int64_t n2 = 0;       // k_reduce_stage2: one thread per differentiated parameter, every segment padded to whole warps
    auto pad32 = [](int64_t v) { return (v + 31) & ~(int64_t)31; };
    for (int c = 0; c < pl->n_comp; ++c)
        n2 += pad32(c < pl->n_glob ? pl->comp[c].n_knots : pl->desc[c].n_knots * pl->desc[c].n_ctrl_keys) + pad32(rp.n_shift[c]) + pad32(rp.n_stretch[c]);
    pl->nb[K_SCATTER] = (int)std::max<int64_t>(1, (pl->n_fit + 255) / 256);
    pl->nb[K_PREPARE] = strip ? (pl->s_nblocks + 3) / 4 : (int)((pl->E + 127) / 128);
    pl->nb[K_FUSED] = strip ? (int)((pl->s_ntasks + jb::kStripWarps - 1) / jb::kStripWarps)
                            : (int)((pl->n_tasks + jb::kWarpsPerBlock - 1) / jb::kWarpsPerBlock);
    pl->nb[K_STAGE1] = rp.nb_ctrl + rp.nb_stage1_rows + (pl->has_norm ? (int)((pl->E * pl->norm_stride + red_threads - 1) / red_threads) : 0);
    pl->nb[K_STAGE2] = (int)((n2 + 255) / 256);
    pl->nb[K_FINISH] = (int)(1 + (pl->n_fit + 1023) / 1024);
    int off[K_COUNT][2];
    for (int i = 0; i < K_COUNT; ++i) { off[i][0] = 0; off[i][1] = pl->nb[i]; }
    JB_CUDA(cudaMemcpyAsync(pl->d_pack, pl->h_pack, sizeof k, cudaMemcpyHostToDevice, pl->stream));
    JB_CUDA(cudaMemcpyAsync(pl->d_off, off, sizeof off, cudaMemcpyHostToDevice, pl->stream));
    JB_CUDA(cudaStreamSynchronize(pl->stream));
    pl->pack_dirty = false;
    return JB_OK;
}

// What one launch sequence needs: where the parameter blocks of the n plans live.
struct LaunchSet {
    const jb::ScatterParams* sc; const jb::PrepareParams* pr; const jb::FusedParams* fu; const jb::ReduceParams* re;
    const jb::StripParams* st; const jb::StripPrepParams* sp;
    int strip_K;             // longest block (rows) among the k_strip plans of the launch: its shared-memory layout
    const int* off[K_COUNT]; // n+1 block offsets per kernel
    int total[K_COUNT];
    int n;
    int n_comp, p_val, pvn, wincap;   // global components, their degree, degree of the normalisation spline (0: none)
    int variant;                      // -1: k_fused, else the FL flags of k_fused2 (same for every plan of the launch)
    int64_t max_knots;                // longest knot axis among the plans of the launch
    bool scatter;
};

template <int NC, int PV, int PVN>
int launch_fused_t(const LaunchSet& L, cudaStream_t st) {
    const size_t smem = (size_t)jb::kWarpsPerBlock * jb::fused_warp_smem<NC, (PVN > 0)>(L.wincap);
    JB_CUDA(cudaFuncSetAttribute(jb::k_fused<NC, PV, PVN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jb::k_fused<NC, PV, PVN><<<(unsigned)L.total[K_FUSED], jb::kWarpsPerBlock * 32, smem, st>>>(L.fu, L.off[K_FUSED], L.n);
    JB_CUDA(cudaGetLastError());
    return JB_OK;
}

template <int NC, int PV>
int launch_fused_n(const LaunchSet& L, cudaStream_t st) {
    switch (L.pvn) {
        case 0: return launch_fused_t<NC, PV, 0>(L, st);
        case 1: return launch_fused_t<NC, PV, 1>(L, st);
        case 2: return launch_fused_t<NC, PV, 2>(L, st);
        case 3: return launch_fused_t<NC, PV, 3>(L, st);
    }
    return fail(JB_ERR_UNSUPPORTED, "normalisation spline of degree %d", L.pvn);
}

template <int NC, int PV, int FL>
int launch_fused2_t(const LaunchSet& L, cudaStream_t st) {
    const size_t smem = (size_t)jb::kWarpsPerBlock * jb::fused2_warp_smem<NC>(L.wincap);
    JB_CUDA(cudaFuncSetAttribute(jb::k_fused2<NC, PV, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jb::k_fused2<NC, PV, FL><<<(unsigned)L.total[K_FUSED], jb::kWarpsPerBlock * 32, smem, st>>>(L.fu, L.off[K_FUSED], L.n);
    JB_CUDA(cudaGetLastError());
    return JB_OK;
}

template <int NC, int PV, int FL>
int launch_fused32_t(const LaunchSet& L, cudaStream_t st) {
    const size_t smem = (size_t)jb::kWarpsPerBlock * jb::fused32_warp_smem<NC>(L.wincap);
    JB_CUDA(cudaFuncSetAttribute(jb::k_fused32<NC, PV, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jb::k_fused32<NC, PV, FL><<<(unsigned)L.total[K_FUSED], jb::kWarpsPerBlock * 32, smem, st>>>(L.fu, L.off[K_FUSED], L.n);
    JB_CUDA(cudaGetLastError());
    return JB_OK;
}

template <int NC, int PV>
int launch_fused32_f(const LaunchSet& L, cudaStream_t st) {
    if constexpr (PV != 3) return launch_fused32_t<NC, PV, (NC == 1 ? 5 : 15)>(L, st);
    else {
#define JB_F32(f) case f: return launch_fused32_t<NC, PV, f>(L, st);
        if constexpr (NC == 1) {
            switch (L.variant & 15) { JB_F32(0) JB_F32(1) JB_F32(4) JB_F32(5) }
        } else {
            switch (L.variant & 15) {
                JB_F32(0) JB_F32(1) JB_F32(2) JB_F32(3) JB_F32(4) JB_F32(5) JB_F32(6) JB_F32(7)
                JB_F32(8) JB_F32(9) JB_F32(10) JB_F32(11) JB_F32(12) JB_F32(13) JB_F32(14) JB_F32(15)
            }
        }
#undef JB_F32
        return fail(JB_ERR_UNSUPPORTED, "no k_fused32 for flags %d", L.variant & 15);
    }
}

template <int NC, int PV>
int launch_fused2_f(const LaunchSet& L, cudaStream_t st) {
    if (L.variant & 64) return launch_fused32_f<NC, PV>(L, st);
#define JB_F2(f) case f: return launch_fused2_t<NC, PV, f>(L, st);
#ifdef JB_F2_FAST_BUILD   // kernel-tuning builds: only the benchmark's variant
    if constexpr (NC == 2 && PV == 3) { switch (L.variant) { JB_F2(9) } }
    else
#endif
    if constexpr (NC == 1) {
        switch (L.variant) { JB_F2(0) JB_F2(1) JB_F2(4) JB_F2(5) }
    } else {
        switch (L.variant) {
            JB_F2(0) JB_F2(1) JB_F2(2) JB_F2(3) JB_F2(4) JB_F2(5) JB_F2(6) JB_F2(7)
            JB_F2(8) JB_F2(9) JB_F2(10) JB_F2(11) JB_F2(12) JB_F2(13) JB_F2(14) JB_F2(15)
        }
    }
#undef JB_F2
    return fail(JB_ERR_UNSUPPORTED, "no k_fused2 for flags %d", L.variant);
}

template <int NC, int PV, int FL>
int launch_strip_t(const LaunchSet& L, cudaStream_t st) {
    using Lay = jb::StripLay<NC, PV, FL>;
    const size_t smem = Lay::launch_bytes(L.strip_K);
    JB_CUDA(cudaFuncSetAttribute(jb::k_strip<NC, PV, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jb::k_strip<NC, PV, FL><<<(unsigned)L.total[K_FUSED], jb::kStripWarps * 32, smem, st>>>(L.st, L.off[K_FUSED], L.n, L.strip_K);
    JB_CUDA(cudaGetLastError());
    return JB_OK;
}

template <int NC, int PV>
int launch_strip_f(const LaunchSet& L, cudaStream_t st) {
    if constexpr (PV != 3) return launch_strip_t<NC, PV, (NC == 1 ? 5 : 15)>(L, st);
    else {
#define JB_ST(f) case f: return launch_strip_t<NC, PV, f>(L, st);
#ifdef JB_F2_FAST_BUILD   // kernel-tuning builds: only the benchmark's variant
        if constexpr (NC == 2) { switch (L.variant & 15) { JB_ST(9) } }
        else
#endif
        if constexpr (NC == 1) {
            switch (L.variant & 15) { JB_ST(0) JB_ST(1) JB_ST(4) JB_ST(5) }
        } else {
            switch (L.variant & 15) {
                JB_ST(0) JB_ST(1) JB_ST(2) JB_ST(3) JB_ST(4) JB_ST(5) JB_ST(6) JB_ST(7)
                JB_ST(8) JB_ST(9) JB_ST(10) JB_ST(11) JB_ST(12) JB_ST(13) JB_ST(14) JB_ST(15)
            }
        }
#undef JB_ST
        return fail(JB_ERR_UNSUPPORTED, "no k_strip for flags %d", L.variant & 15);
    }
}

// Six launches for the n plans of L (n = 1: the plan's own blocks; n > 1: the dense arrays of a jb_batch).
int launch_all(const LaunchSet& L, cudaStream_t st, const double* p_fit_override, double* out_override,
               jb_plan* prof /* may be null */) {
    if (L.scatter)
        jb::k_scatter_params<<<(unsigned)L.total[K_SCATTER], 256, 0, st>>>(L.sc, L.off[K_SCATTER], L.n, p_fit_override);
    const bool strip = L.variant >= 0 && (L.variant & 128);
    if (strip) {
        const unsigned nb = (unsigned)L.total[K_PREPARE];
        if (L.n_comp == 1) jb::k_prepare_strip<1><<<nb, 128, 0, st>>>(L.sp, L.off[K_PREPARE], L.n);
        else jb::k_prepare_strip<2><<<nb, 128, 0, st>>>(L.sp, L.off[K_PREPARE], L.n);
    } else {
        const unsigned nb = (unsigned)L.total[K_PREPARE];
        if (L.n_comp == 1 && !L.pvn) jb::k_prepare_rows<1, 0><<<nb, 128, 0, st>>>(L.pr, L.off[K_PREPARE], L.n);
        else if (L.n_comp == 1) jb::k_prepare_rows<1, 1><<<nb, 128, 0, st>>>(L.pr, L.off[K_PREPARE], L.n);
        else if (!L.pvn) jb::k_prepare_rows<2, 0><<<nb, 128, 0, st>>>(L.pr, L.off[K_PREPARE], L.n);
        else jb::k_prepare_rows<2, 1><<<nb, 128, 0, st>>>(L.pr, L.off[K_PREPARE], L.n);
    }
    if (prof && prof->profiling) {
        if (prof->prof_used == prof->prof_events.size()) {
            cudaEvent_t a, b;
            JB_CUDA(cudaEventCreate(&a));
            JB_CUDA(cudaEventCreate(&b));
            prof->prof_events.emplace_back(a, b);
        }
        JB_CUDA(cudaEventRecord(prof->prof_events[prof->prof_used].first, st));
    }
    int rc;
    if (strip) {
        switch (L.n_comp * 10 + L.p_val) {
            case 11: rc = launch_strip_f<1, 1>(L, st); break;
            case 12: rc = launch_strip_f<1, 2>(L, st); break;
            case 13: rc = launch_strip_f<1, 3>(L, st); break;
            case 21: rc = launch_strip_f<2, 1>(L, st); break;
            case 22: rc = launch_strip_f<2, 2>(L, st); break;
            case 23: rc = launch_strip_f<2, 3>(L, st); break;
            default: rc = fail(JB_ERR_UNSUPPORTED, "no strip kernel for %d components of degree %d", L.n_comp, L.p_val);
        }
    } else if (L.variant >= 0) {
        switch (L.n_comp * 10 + L.p_val) {
            case 11: rc = launch_fused2_f<1, 1>(L, st); break;
            case 12: rc = launch_fused2_f<1, 2>(L, st); break;
            case 13: rc = launch_fused2_f<1, 3>(L, st); break;
            case 21: rc = launch_fused2_f<2, 1>(L, st); break;
            case 22: rc = launch_fused2_f<2, 2>(L, st); break;
            case 23: rc = launch_fused2_f<2, 3>(L, st); break;
            default: rc = fail(JB_ERR_UNSUPPORTED, "no fused kernel for %d components of degree %d", L.n_comp, L.p_val);
        }
    } else
    switch (L.n_comp * 10 + L.p_val) {
        case 11: rc = launch_fused_n<1, 1>(L, st); break;
        case 12: rc = launch_fused_n<1, 2>(L, st); break;
        case 13: rc = launch_fused_n<1, 3>(L, st); break;
        case 21: rc = launch_fused_n<2, 1>(L, st); break;
        case 22: rc = launch_fused_n<2, 2>(L, st); break;
        case 23: rc = launch_fused_n<2, 3>(L, st); break;
        default: rc = fail(JB_ERR_UNSUPPORTED, "no fused kernel for %d components of degree %d", L.n_comp, L.p_val);
    }
    if (rc) return rc;
    if (prof && prof->profiling) {
        JB_CUDA(cudaEventRecord(prof->prof_events[prof->prof_used].second, st));
        prof->prof_used++;
        prof->prof_stream = st;
    }
    jb::k_reduce_stage1<<<(unsigned)L.total[K_STAGE1], strip ? kStripRedThreads : jb::kRedSlab, 0, st>>>(L.re, L.off[K_STAGE1], L.n);
    jb::k_reduce_stage2<<<(unsigned)L.total[K_STAGE2], 256, 0, st>>>(L.re, L.off[K_STAGE2], L.n);
    jb::k_finish<<<(unsigned)L.total[K_FINISH], 1024, 0, st>>>(L.re, L.off[K_FINISH], L.n, out_override);
    JB_CUDA(cudaGetLastError());
    return JB_OK;
}

// Enqueue one evaluation on `st`.  d_p_fit may be null (keep current parameters).
int enqueue_eval(jb_plan* pl, const double* d_p_fit, double* d_out, cudaStream_t st) {
    int rc = check_supported_for_eval(pl);
    if (rc) return rc;
    if (!pl->params_set) return fail(JB_ERR_BAD_ARG, "jb_plan_set_params has not been called");
    if (pl->sort_dirty && (rc = resort_rows(pl))) return rc;
    if ((rc = refresh_pack(pl))) return rc;
    LaunchSet L;
    L.sc = &pl->d_pack->sc; L.pr = &pl->d_pack->pr; L.fu = &pl->d_pack->fu; L.re = &pl->d_pack->re;
    L.st = &pl->d_pack->st; L.sp = &pl->d_pack->sp; L.strip_K = pl->sK;
    for (int i = 0; i < K_COUNT; ++i) { L.off[i] = pl->d_off + 2 * i; L.total[i] = pl->nb[i]; }
    L.n = 1; L.n_comp = pl->n_glob; L.p_val = pl->desc[0].p_val; L.wincap = pl->wincap;
    L.pvn = pl->has_norm ? pl->desc[pl->n_glob].p_val : 0;
    L.max_knots = pl->max_knots;
    L.variant = pl->variant;
    L.scatter = d_p_fit != nullptr && pl->n_fit > 0;
    if ((rc = launch_all(L, st, d_p_fit, d_out, pl))) return rc;
    pl->launches += L.scatter ? 6 : 5;
    pl->evaluations++;
    return JB_OK;
}

int read_status(jb_plan* pl, cudaStream_t st, bool* window) {
    JB_CUDA(cudaMemcpyAsync(pl->h_status, pl->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    const int bits = pl->h_status[0];
    pl->serial_rows = pl->h_status[1];
    if (window) *window = false;
    if (bits || pl->h_status[1]) JB_CUDA(cudaMemsetAsync(pl->d_status, 0, 2 * sizeof(int), st));
    if (bits & jb::kStatBounds)
        return fail(JB_ERR_CUDA, "JB_DEBUG_BOUNDS: a window address left its shared-memory copy (kernel bug)");
    if (bits & jb::kStatOutOfRange)
        return fail(JB_ERR_OUT_OF_RANGE, "an unmasked warped pixel lies outside the knot grid of a spline component");
    if (bits & (jb::kStatWindow | jb::kStatStrip)) {
        if (window) *window = true;
        if (pl->variant >= 0 && (pl->variant & 128))
            return fail(JB_ERR_WINDOW, "rows of a block moved more than a knot apart (the shifts drifted since the rows were sorted)");
        return fail(JB_ERR_WINDOW, "knot window of a pixel tile exceeds %d knots (grid much finer than the pixels, or rows of one group far apart)", pl->wincap);
    }
    return JB_OK;
}

// A window overflow was reported for `pl`: what to change before evaluating again.  k_strip plans
// first re-sort (and re-lay out) their rows with the current parameters, then leave k_strip; the
// other kernels re-sort, then widen the window, then shorten the row groups.  false: give up.
int window_retry(jb_plan* pl, int attempt, bool* again) {
    *again = false;
    JB_CUDA(cudaMemcpy(pl->h_params.data(), pl->d_params, pl->n_params * sizeof(double), cudaMemcpyDeviceToHost));
    if (pl->variant >= 0 && (pl->variant & 128)) {
        if (pl->strip_fail++ > 0) { pl->strip_disabled = true; pl->strip_ok = false; pl->pack_dirty = true; pl->strip_reason = 6; }
    } else if (attempt > 0) {
        if (pl->wincap < jb::kWinCapMax) pl->wincap = std::min(jb::kWinCapMax, pl->wincap + 32);
        else if (pl->G > 1) pl->G = std::max(1, pl->G / 4);
        else return JB_OK;
        int rc2 = alloc_scratch(pl);
        if (rc2) return rc2;
    }
    int rc2 = resort_rows(pl);
    if (rc2) return rc2;
    *again = true;
    return JB_OK;
}

// Evaluate on the plan's stream and wait; d_p_fit null = at the current parameters.  Leaves
// [loss, grad_fit] in pl->h_out.  Retries after a window overflow: first re-sort the rows with
// the current parameters (the shifts may have drifted), then shrink the row groups.
int eval_sync(jb_plan* pl, const double* d_p_fit) {
    cudaStream_t st = pl->stream;
    for (int attempt = 0;; ++attempt) {
        int rc = enqueue_eval(pl, d_p_fit, pl->d_out, st);
        if (rc) return rc;
        JB_CUDA(cudaMemcpyAsync(pl->h_out, pl->d_out, (pl->n_fit + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
        bool window = false;
        rc = read_status(pl, st, &window);
        if (rc == JB_OK) { pl->strip_fail = 0; return JB_OK; }
        if (!window || attempt >= 14) return rc;
        const std::string msg = g_err;
        bool again = false;
        int rc2 = window_retry(pl, attempt, &again);
        if (rc2) return rc2;
        if (!again) { g_err = msg; return rc; }
    }
}

}  // namespace

// ================================================================================ C ABI
namespace {
// JB_GRAPHS=0 turns the recorded launch sequences off (plain stream launches everywhere)
bool graphs_enabled() {
    static const bool on = [] { const char* e = getenv("JB_GRAPHS"); return !(e && e[0] == '0'); }();
    return on;
}
bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs != cudaStreamCaptureStatusNone;
}
void drop_graph(cudaGraphExec_t& g) {
    if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}
// Record what `enqueue` puts on `st` and instantiate it.  false: recording is not possible here (the work
// was NOT done: the caller enqueues it the plain way); rc carries an error of `enqueue` itself.
template <typename F>
bool record_graph(cudaStream_t st, cudaGraphExec_t* exec, int* rc, F enqueue) {
    *rc = JB_OK;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    *rc = enqueue();
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    bool ok = *rc == JB_OK && e == cudaSuccess && g != nullptr;
    if (ok && cudaGraphInstantiate(exec, g, 0) != cudaSuccess) { *exec = nullptr; ok = false; }
    if (g) cudaGraphDestroy(g);
    if (!ok) cudaGetLastError();
    return ok;
}
}  // namespace

extern "C" {

const char* jb_last_error(void) { return g_err.c_str(); }
int jb_abi_version(void) { return JB_ABI_VERSION; }

int jb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// 64-bit position-sensitive checksums of n host buffers (OpenMP over the buffers): what
// Data.blockify compares to notice an in-place edit of a frame before it reuses the resident block.
int jb_host_checksum(const void* const* ptrs, const int64_t* nbytes, int64_t n, uint64_t* out) {
    if (n < 0 || (n && (!ptrs || !nbytes || !out))) return fail(JB_ERR_BAD_ARG, "null argument");
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        const unsigned char* p = static_cast<const unsigned char*>(ptrs[i]);
        const int64_t nb = nbytes[i] > 0 ? nbytes[i] : 0, nw = nb / 8;
        uint64_t s1 = 0x9E3779B97F4A7C15ull, s2 = 0;
        for (int64_t k = 0; k < nw; ++k) {
            uint64_t w;
            memcpy(&w, p + 8 * k, 8);
            s1 += w;
            s2 += w * (uint64_t)(2 * k + 1);
        }
        uint64_t tail = 0;
        if (nb > 8 * nw) memcpy(&tail, p + 8 * nw, (size_t)(nb - 8 * nw));
        s1 += tail; s2 += tail * (uint64_t)(2 * nw + 1);
        out[i] = s1 ^ (s2 * 0xff51afd7ed558ccdull) ^ (uint64_t)nb;
    }
    return JB_OK;
}

void jb_plan_destroy(jb_plan* pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    if (pl->stream) cudaStreamSynchronize(pl->stream);
    for (auto& b : pl->bufs) cudaFree(b.p);
    for (auto& e : pl->prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (pl->h_pfit) cudaFreeHost(pl->h_pfit);
    if (pl->h_out) cudaFreeHost(pl->h_out);
    if (pl->h_status) cudaFreeHost(pl->h_status);
    if (pl->h_pack) cudaFreeHost(pl->h_pack);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    delete pl;
}

int jb_plan_create(const jb_plan_desc* d, jb_plan** out) {
    if (!d || !out) return fail(JB_ERR_BAD_ARG, "null descriptor");
    *out = nullptr;
    if (d->abi_version != JB_ABI_VERSION) return fail(JB_ERR_BAD_ARG, "ABI version %d, library is %d", d->abi_version, JB_ABI_VERSION);
    if (d->n_rows <= 0 || d->n_pix <= 0 || !d->xs) return fail(JB_ERR_BAD_ARG, "empty data block");
    if (d->n_components < 1 || d->n_components > jb::kMaxComp || !d->components)
        return fail(JB_ERR_UNSUPPORTED, "%d additive components (1..%d supported)", d->n_components, jb::kMaxComp);
    if ((d->ys == nullptr) != (d->yivar == nullptr)) return fail(JB_ERR_BAD_ARG, "ys and yivar must both be given or both be NULL");
    if (d->n_rows > INT_MAX / 2) return fail(JB_ERR_BAD_ARG, "too many rows");
    for (int c = 0; c < d->n_components; ++c) {
        const jb_component_desc& cd = d->components[c];
        if (cd.p_val < 1 || cd.p_val > 3) return fail(JB_ERR_UNSUPPORTED, "component %d: spline degree %d (1..3 supported)", c, cd.p_val);
        if (cd.kind != JB_COMP_SPLINE && cd.kind != JB_COMP_NORM_SPLINE) return fail(JB_ERR_BAD_ARG, "component %d: kind %d", c, cd.kind);
        if (cd.n_knots < cd.p_val + 1 || !(cd.dx > 0) || !std::isfinite(cd.xp0)) return fail(JB_ERR_BAD_ARG, "component %d: bad knot grid", c);
        const int64_t nctrl = cd.kind == JB_COMP_SPLINE ? cd.n_knots : cd.n_knots * cd.n_ctrl_keys;
        if (cd.ctrl_offset < 0 || cd.ctrl_offset + nctrl > d->n_params) return fail(JB_ERR_BAD_ARG, "component %d: control points outside params_all", c);
        if (cd.kind == JB_COMP_NORM_SPLINE && cd.n_ctrl_keys <= 0) return fail(JB_ERR_BAD_ARG, "component %d: n_ctrl_keys", c);
        if (cd.shift_offset >= 0 && (cd.n_shift <= 0 || cd.shift_offset + cd.n_shift > d->n_params)) return fail(JB_ERR_BAD_ARG, "component %d: shift slot", c);
        if (cd.stretch_offset >= 0 && (cd.n_stretch <= 0 || cd.stretch_offset + cd.n_stretch > d->n_params)) return fail(JB_ERR_BAD_ARG, "component %d: stretch slot", c);
        if (cd.n_knots > INT_MAX / 2) return fail(JB_ERR_BAD_ARG, "component %d: too many knots", c);
    }
    int ndev = jb_device_count();
    if (ndev <= 0) return fail(JB_ERR_CUDA, "no CUDA device: this library has no CPU path");
    if (d->device < 0 || d->device >= ndev) return fail(JB_ERR_BAD_ARG, "device %d of %d", d->device, ndev);
    JB_CUDA(cudaSetDevice(d->device));
    {   // the stream-ordered allocations of forward / grid search / curvature keep their memory between
        // calls (the default pool hands it back to the driver at every synchronisation: milliseconds per call)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d->device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }

    jb_plan* pl = new jb_plan;
    struct Guard { jb_plan* p; ~Guard() { if (p) jb_plan_destroy(p); } } guard{pl};
    pl->device = d->device;
    JB_CUDA(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    const int64_t E = d->n_rows, P = d->n_pix;
    pl->E = E; pl->P = P;
    pl->ppad = (P + jb::kTile - 1) / jb::kTile * jb::kTile;
    pl->n_tiles = (int)(pl->ppad / jb::kTile);
    if (pl->n_tiles > jb::kRedBases)
        return fail(JB_ERR_UNSUPPORTED, "%lld pixels per row (at most %d)", (long long)pl->P, jb::kRedBases * jb::kTile);
    pl->has_data = d->ys != nullptr;
    pl->coef = d->coefficient;
    pl->general_only = (d->flags & JB_PLAN_GENERAL_KERNEL) != 0;
    pl->strip_allowed = (d->flags & JB_PLAN_NO_STRIP) == 0;
    pl->s_hint = d->rows_per_group;
    pl->f32 = (d->flags & JB_PLAN_FLOAT32) != 0;
    if (pl->f32 && (pl->general_only || !d->ys)) return fail(JB_ERR_BAD_ARG, "a float32 plan needs ys / yivar and cannot ask for the general kernel");
    pl->n_comp = d->n_components;
    // store the components globals first, the per-key (normalisation) ones last
    std::vector<jb_component_desc> comps;
    for (int pass = 0; pass < 2; ++pass)
        for (int c = 0; c < d->n_components; ++c)
            if ((d->components[c].kind == JB_COMP_NORM_SPLINE) == (pass == 1)) {
                pl->ext_of[comps.size()] = c;
                comps.push_back(d->components[c]);
                if (pass == 0) pl->n_glob++;
            }
    pl->has_norm = pl->n_comp - pl->n_glob == 1 ? 1 : 0;
    pl->n_params = d->n_params;
    const int64_t ppad = pl->ppad;

    // ---- canonical copy of the block: every row ascending in x, tile-interleaved (x | y | ivar per
    // (row, tile)), mask folded in: a pixel that carries no weight (masked, padded, or ivar == 0)
    // gets ivar = 0, y = 0 and an x between its weighted neighbours (its own x when that is finite
    // and monotone, else the x of the nearest weighted pixel before it; the first weighted x for
    // leading ones), so it is evaluated on knots the data already touch and contributes exactly
    // zero -- ChiSquare's where(~mask, ., 0), loss.py:168-173, for value and cotangent.
    const int NT = pl->n_tiles;
    PhaseTimer phase("create: host staging alloc");
    HostBuf<double> hdata((size_t)E * NT * jb::kTileDoubles);
    HostBuf<double> hx((size_t)E * ppad);     // original x, canonical orientation
    HostBuf<double> hxf((size_t)E * ppad);    // filled x (what the fused kernel sees)
    HostBuf<uint8_t> hm((size_t)E * ppad);    // 1 = carries no weight
    if (!hdata.data() || !hx.data() || !hxf.data() || !hm.data()) return fail(JB_ERR_BAD_ARG, "out of host memory for the canonical block");
    pl->h_row_rev.assign(E, 0);
    pl->h_xfirst.assign(E, INFINITY);
    std::vector<double> txmin((size_t)E * NT, INFINITY), txmax((size_t)E * NT, -INFINITY);
    std::vector<uint8_t> dense((size_t)E * NT, 0);
    std::vector<uint32_t> live((size_t)E * NT, 0);
    double need = 0.0;   // knots two lanes of one class must be apart, in x units, 1 % margin
    for (int c = 0; c < pl->n_glob; ++c)
        need = std::max(need, (comps[c].p_val + 1.01) * comps[c].dx);
    double smooth = INFINITY;   // two pixels closer than this are at most one knot interval apart
    for (int c = 0; c < pl->n_glob; ++c) smooth = std::min(smooth, 0.99 * comps[c].dx);
    double norm_dx = INFINITY;  // a lane's 4 pixels must span less than one knot of every normalisation grid
    for (int c = pl->n_glob; c < pl->n_comp; ++c) norm_dx = std::min(norm_dx, 0.99 * comps[c].dx);
    double any_x = 0.0;
    for (int64_t i = 0; i < E * P && any_x == 0.0; ++i)
        if ((!d->mask || !d->mask[i]) && std::isfinite(d->xs[i])) any_x = d->xs[i];
    // (rows are independent: the loop runs on all host cores)
    phase.lap("create: canonical rows");
    bool norm_smooth_all = true, f2_ok_all = true;
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, bad_row = -1, bad_pix = -1;
#pragma omp parallel for schedule(dynamic, 8) reduction(&& : norm_smooth_all, f2_ok_all) reduction(+ : tk0, tk1, tk2, tk3)
    for (int64_t r = 0; r < E; ++r) {
        const double* xr = d->xs + r * P;
        const uint8_t* mr = d->mask ? d->mask + r * P : nullptr;
        auto weighted = [&](int64_t p) {
            if (mr && mr[p]) return false;
            return !pl->has_data || d->yivar[r * P + p] != 0.0;
        };
        int64_t first = -1, last = -1;
        for (int64_t p = 0; p < P; ++p)
            if (weighted(p)) { if (first < 0) first = p; last = p; }
        const bool rev = first >= 0 && xr[last] < xr[first];
        pl->h_row_rev[r] = rev;
        double fill = first >= 0 ? xr[rev ? last : first] : any_x;
        // next weighted x after each canonical position (+inf behind the last one)
        std::vector<double> nextw(ppad + 1, INFINITY);
        for (int64_t p = P - 1; p >= 0; --p) {
            const int64_t src = rev ? P - 1 - p : p;
            nextw[p] = weighted(src) ? xr[src] : nextw[p + 1];
        }
        for (int64_t p = 0; p < ppad; ++p) {
            const int64_t src = rev ? P - 1 - p : p;
            const int t = (int)(p / jb::kTile);
            const int pt = (int)(p % jb::kTile);   // lane-interleaved position inside the record
            double* rec = hdata.data() + ((size_t)r * NT + t) * jb::kTileDoubles +
                          (pt % jb::kPixPerLane) * 32 + pt / jb::kPixPerLane;
            const bool w = p < P && weighted(src);
            hx[(size_t)r * ppad + p] = (p < P && std::isfinite(xr[src])) ? xr[src] : 0.0;
            hm[(size_t)r * ppad + p] = w ? 0 : 1;
            rec[jb::kTile] = 0.0; rec[2 * jb::kTile] = 0.0;
            if (w) {
                if (!std::isfinite(xr[src])) {
#pragma omp critical
                    if (bad_row < 0) { bad_row = r; bad_pix = src; }
                }
                fill = xr[src];
                if (pl->has_data) { rec[jb::kTile] = d->ys[r * P + src]; rec[2 * jb::kTile] = d->yivar[r * P + src]; }
                pl->h_xfirst[r] = std::min(pl->h_xfirst[r], fill);
            }
            else if (p < P && std::isfinite(xr[src]) && xr[src] >= fill && xr[src] <= nextw[p] && std::isfinite(nextw[p]))
                fill = xr[src];   // an interior masked pixel keeps its own x when it sits between its weighted neighbours
            rec[0] = fill;
            hxf[(size_t)r * ppad + p] = fill;
        }
        // tile extents, over every pixel the kernel will walk (filled ones included), and density:
        // lanes l and l+4 (13 pixels apart) share an accumulator copy, which is safe when every pixel
        // of lane l lies at least p_val+1 knots above every pixel of lanes l-4, l-8, ...  Knot
        // distances inside a row do not depend on the shift, so this is decided here, from x alone.
        for (int t = 0; t < NT; ++t) {
            const size_t ti = (size_t)r * NT + t;
            bool any_w = false;
            double cmax[jb::kClasses], cmax2[jb::kClasses2];
            for (int q = 0; q < jb::kClasses; ++q) cmax[q] = -INFINITY;
            for (int q = 0; q < jb::kClasses2; ++q) cmax2[q] = -INFINITY;
            // k_fused2 freezes a lane that owns no weighted pixel in a row (its filled x values sit on
            // its neighbours' knots), so its own tests only look at the other lanes
            bool bad = false, jumpy = false, bad2 = false, jumpy2 = false;
            unsigned lv = 0;
            for (int l = 0; l < 32; ++l) {
                double lmin = INFINITY, lmax = -INFINITY;
                bool lane_w = false, lane_jumpy = false;
                for (int j = 0; j < jb::kPixPerLane; ++j) {
                    const size_t i = (size_t)r * ppad + (size_t)t * jb::kTile + l * jb::kPixPerLane + j;
                    lmin = std::min(lmin, hxf[i]); lmax = std::max(lmax, hxf[i]);
                    any_w |= !hm[i];
                    lane_w |= !hm[i];
                    // inside a lane's run: a step of less than one knot of the finest grid, never backwards
                    if (j > 0 && !(hxf[i] >= hxf[i - 1] && hxf[i] - hxf[i - 1] < smooth)) lane_jumpy = true;
                }
                jumpy |= lane_jumpy;
                if (any_w && !(lmax - lmin < norm_dx)) norm_smooth_all = false;
                double& cm = cmax[l % jb::kClasses];
                if (!(lmin - cm >= need)) bad = true;   // also catches non-monotone rows
                cm = std::max(cm, lmax);
                if (lane_w) {
                    lv |= 1u << l;
                    jumpy2 |= lane_jumpy;
                    double& cm2 = cmax2[l % jb::kClasses2];
                    if (!(lmin - cm2 >= need)) bad2 = true;
                    cm2 = std::max(cm2, lmax);
                }
                txmin[ti] = std::min(txmin[ti], lmin);
                txmax[ti] = std::max(txmax[ti], lmax);
            }
            live[ti] = lv;
            if (lv && (bad2 || jumpy2)) f2_ok_all = false;
            // flags: 1 = lanes of a class may collide (walk lane by lane), 2 = nothing weighted (skip),
            //        4 = consecutive pixels a knot or more apart somewhere (general per-pixel path)
            dense[ti] = !any_w ? 2 : ((bad ? 1 : 0) | (jumpy ? 4 : 0));
            if (!any_w) { txmin[ti] = INFINITY; txmax[ti] = -INFINITY; }
            const int kind = !any_w ? 3 : bad ? 2 : jumpy ? 1 : 0;
            tk0 += kind == 0; tk1 += kind == 1; tk2 += kind == 2; tk3 += kind == 3;
        }
    }
    if (bad_row >= 0) return fail(JB_ERR_BAD_ARG, "row %lld pixel %lld: unmasked x is not finite", bad_row, bad_pix);
    pl->norm_smooth = pl->norm_smooth && norm_smooth_all;
    pl->f2_ok = pl->f2_ok && f2_ok_all;
    pl->tiles_kind[0] += tk0; pl->tiles_kind[1] += tk1; pl->tiles_kind[2] += tk2; pl->tiles_kind[3] += tk3;
    phase.lap("create: row offsets");
    // Position of every row relative to one reference row (k_strip, jb_strip.cuh): x[r][p] - x[ref][p]
    // at the row's first weighted pixel whose reference pixel carries an x of its own.  Unlike the
    // first weighted x this does not jump by a pixel when a row's first pixel is masked.
    {
        std::vector<int64_t> nw(E, 0);
        for (int64_t r = 0; r < E; ++r)
            for (int64_t p = 0; p < P; ++p) nw[r] += hm[(size_t)r * ppad + p] ? 0 : 1;
        const int64_t ref = std::max_element(nw.begin(), nw.end()) - nw.begin();
        pl->h_xrel.assign(E, INFINITY);
        for (int64_t r = 0; r < E && nw[ref] > 0; ++r) {
            if (!nw[r]) continue;
            double o = NAN;
            for (int64_t p = 0; p < P; ++p) {
                const size_t i = (size_t)r * ppad + p, j = (size_t)ref * ppad + p;
                if (!hm[i] && (!hm[j] || hxf[j] == hx[j])) { o = hxf[i] - hxf[j]; break; }
            }
            pl->h_xrel[r] = o;
            if (!std::isfinite(o)) { pl->h_xrel.clear(); break; }      // a row without a common pixel: no column walk
        }
    }
    int rc;
    pl->h_txmin = txmin; pl->h_txmax = txmax; pl->h_flags = dense; pl->h_live = live;
    phase.lap("create: uploads");
    if ((rc = dev_alloc(pl, &pl->d_tinfo, (size_t)NT * E))) return rc;
    {
        char* rec = nullptr;
        if ((rc = dev_alloc(pl, &rec, (size_t)E * sizeof(jb::RowRec<2, 1>)))) return rc;
        pl->d_grouprec = rec;
    }
    if (pl->f32) {
        // float records: (x - x_ref) / dx, x_ref = the first x of the (row, tile) -- TileInfo.xmin, what
        // the kernel adds back in double -- split into a knot count (a byte) and a fraction (a float),
        // then y and ivar; same lane interleave
        for (int c = 1; c < pl->n_glob; ++c)
            if (comps[c].dx != comps[0].dx) return fail(JB_ERR_UNSUPPORTED, "a float32 plan needs one knot spacing for all spline components");
        const double inv_dx0 = 1.0 / comps[0].dx;
        std::vector<unsigned char> h32((size_t)E * NT * jb::kTileBytes32, 0);
        for (int64_t r = 0; r < E; ++r)
            for (int t = 0; t < NT; ++t) {
                const size_t ti = (size_t)r * NT + t;
                const double* rec = hdata.data() + ti * jb::kTileDoubles;
                float* out = reinterpret_cast<float*>(h32.data() + ti * jb::kTileBytes32);
                unsigned char* kout = h32.data() + ti * jb::kTileBytes32 + jb::kTileFloats * 4;
                for (int q = 0; q < jb::kTile; ++q) {
                    double off = std::isfinite(txmin[ti]) ? (rec[q] - txmin[ti]) * inv_dx0 : 0.0;
                    if (!(off >= 0.0)) off = 0.0;
                    if (off > 255.0) return fail(JB_ERR_UNSUPPORTED, "a float32 plan needs tiles narrower than 255 knots");
                    const double k = std::floor(off);
                    kout[q] = (unsigned char)k;
                    out[q] = (float)(off - k);
                    out[jb::kTile + q] = (float)rec[jb::kTile + q];
                    out[2 * jb::kTile + q] = (float)rec[2 * jb::kTile + q];
                }
            }
        {
            unsigned char* dptr = nullptr;
            if ((rc = dev_upload(pl, &dptr, h32))) return rc;
            pl->d_data32 = reinterpret_cast<float*>(dptr);
        }
    } else if ((rc = dev_upload(pl, &pl->d_data, hdata))) return rc;
    if ((rc = dev_upload(pl, &pl->d_x, hx))) return rc;
    if ((rc = dev_upload(pl, &pl->d_row_rev, pl->h_row_rev))) return rc;
    std::vector<int> perm(E);
    std::iota(perm.begin(), perm.end(), 0);
    if ((rc = dev_upload(pl, &pl->d_row_perm, perm))) return rc;

    // ---- components ---------------------------------------------------------------------------
    phase.lap("create: components, scratch");
    if ((rc = dev_alloc(pl, &pl->d_params, (size_t)d->n_params))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_grad_all, (size_t)d->n_params))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_fisher_all, (size_t)d->n_params))) return rc;
    JB_CUDA(cudaMemset(pl->d_params, 0, d->n_params * sizeof(double)));
    JB_CUDA(cudaMemset(pl->d_grad_all, 0, d->n_params * sizeof(double)));
    JB_CUDA(cudaMemset(pl->d_fisher_all, 0, d->n_params * sizeof(double)));
    pl->h_params.assign(d->n_params, 0.0);
    std::vector<int> ident(E);
    std::iota(ident.begin(), ident.end(), 0);
    for (int c = 0; c < pl->n_comp; ++c) {
        const jb_component_desc& cd = comps[c];
        pl->desc[c] = cd;
        pl->desc[c].ctrl_key = pl->desc[c].shift_key = pl->desc[c].stretch_key = nullptr;  // host pointers are not retained
        jb::CompDev& cv = pl->comp[c];
        memset(&cv, 0, sizeof cv);
        cv.kind = cd.kind; cv.p_val = cd.p_val; cv.n_knots = (int)cd.n_knots;
        cv.xp0 = cd.xp0; cv.inv_dx = 1.0 / cd.dx;
        cv.ctrl = pl->d_params + cd.ctrl_offset;
        if (cd.kind == JB_COMP_SPLINE) pl->max_knots = std::max<int64_t>(pl->max_knots, cd.n_knots);
        const int32_t* keys[3] = {cd.shift_key, cd.stretch_key, cd.ctrl_key};
        const int64_t nkeys[3] = {cd.shift_offset >= 0 ? cd.n_shift : 0, cd.stretch_offset >= 0 ? cd.n_stretch : 0,
                                  cd.kind == JB_COMP_NORM_SPLINE ? cd.n_ctrl_keys : 0};
        for (int s = 0; s < 3; ++s) {
            if (nkeys[s] == 0) continue;
            std::vector<int> k(E);
            for (int64_t r = 0; r < E; ++r) {
                k[r] = keys[s] ? keys[s][r] : (int)r;
                if (k[r] < 0 || k[r] >= nkeys[s]) return fail(JB_ERR_BAD_ARG, "component %d: key %d of row %lld outside [0,%lld)", c, k[r], (long long)r, (long long)nkeys[s]);
            }
            if ((rc = dev_upload(pl, &pl->d_keys[c][s], k))) return rc;
            if (s == 0) pl->h_shift_key[c] = k;
            {   // CSR key -> rows, rows ascending: the fixed order of the per-key sums
                std::vector<int> ptr(nkeys[s] + 1, 0), rows(E);
                for (int64_t r = 0; r < E; ++r) ptr[k[r] + 1]++;
                for (int64_t i = 0; i < nkeys[s]; ++i) ptr[i + 1] += ptr[i];
                std::vector<int> fillp(ptr.begin(), ptr.end() - 1);
                for (int64_t r = 0; r < E; ++r) rows[fillp[k[r]]++] = (int)r;
                if ((rc = dev_upload(pl, &pl->d_key_ptr[c][s], ptr))) return rc;
                if ((rc = dev_upload(pl, &pl->d_key_rows[c][s], rows))) return rc;
            }
        }
        if (cd.shift_offset >= 0) { cv.shift = pl->d_params + cd.shift_offset; cv.shift_key = pl->d_keys[c][0]; }
        if (cd.stretch_offset >= 0) { cv.stretch = pl->d_params + cd.stretch_offset; cv.stretch_key = pl->d_keys[c][1]; }
        if (cd.kind == JB_COMP_NORM_SPLINE) cv.ctrl_key = pl->d_keys[c][2];
    }

    // ---- scratch ------------------------------------------------------------------------------
    // on-chip window: the widest tile, in knots of the finest grid, plus the spline support and room
    // for the rows of a group to sit a few knots apart; grown on demand (eval_sync)
    double max_knots_tile = 0.0;
    for (size_t i = 0; i < txmin.size(); ++i)
        if (txmin[i] <= txmax[i])
            for (int c = 0; c < pl->n_glob; ++c)
                max_knots_tile = std::max(max_knots_tile, (txmax[i] - txmin[i]) / comps[c].dx + comps[c].p_val + 2);
    pl->wincap = (int)std::min<double>(jb::kWinCapMax, std::max(64.0, std::ceil((max_knots_tile + 8) / 32.0) * 32.0));
    const int NQ = jb::row_sum_stride(pl->n_glob, pl->has_norm);
    if (pl->has_norm) {
        pl->norm_stride = (int)std::min<int64_t>(32, (comps[pl->n_glob].n_knots + 3) / 4 * 4);
        if ((rc = dev_alloc(pl, &pl->d_normpart, (size_t)E * pl->n_tiles * pl->norm_stride, true))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_normrow, (size_t)E * pl->norm_stride, true))) return rc;
    }
    pl->G = choose_rows_per_group(E, pl->n_tiles, d->rows_per_group);
    if ((rc = dev_alloc(pl, &pl->d_rowpart, (size_t)E * pl->n_tiles * NQ, true))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_rowsum, (size_t)E * NQ, true))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_tmp, (size_t)pl->n_glob * pl->n_tiles * pl->max_knots, true))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_tile_rng, (size_t)2 * pl->n_glob * pl->n_tiles, true))) return rc;
    if ((rc = alloc_scratch(pl))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_status, 2))) return rc;
    JB_CUDA(cudaMemset(pl->d_status, 0, 2 * sizeof(int)));
    JB_CUDA(cudaMallocHost(&pl->h_status, 2 * sizeof(int)));
    if ((rc = dev_alloc(pl, &pl->d_pack, 1))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_off, 2 * K_COUNT))) return rc;
    JB_CUDA(cudaMallocHost(&pl->h_pack, sizeof(jb_plan::Pack)));
    // a fit map of length zero until jb_plan_set_fit is called
    if ((rc = dev_alloc(pl, &pl->d_fit_map, 1))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_pfit, 1))) return rc;
    if ((rc = dev_alloc(pl, &pl->d_out, 1))) return rc;
    JB_CUDA(cudaMallocHost(&pl->h_pfit, sizeof(double)));
    JB_CUDA(cudaMallocHost(&pl->h_out, sizeof(double)));
    pl->fit_bufs = true;

    guard.p = nullptr;
    *out = pl;
    return JB_OK;
}

int jb_plan_info(const jb_plan* pl, jb_plan_info_t* info) {
    if (!pl || !info) return fail(JB_ERR_BAD_ARG, "null argument");
    memset(info, 0, sizeof *info);
    info->n_rows = pl->E; info->n_pix = pl->P; info->n_pix_padded = pl->ppad;
    info->n_params = pl->n_params; info->n_fit = pl->n_fit;
    const bool strip = pl->variant >= 0 && (pl->variant & 128);
    info->n_tasks = strip ? pl->s_ntasks : pl->n_tasks;
    info->rows_per_group = strip ? pl->sK * pl->sBPT : pl->G;
    info->tile_pixels = strip ? jb::kStripPx : jb::kTile;
    info->device_bytes = pl->device_bytes; info->scratch_bytes = pl->scratch_bytes;
    // x, y, ivar (8 B each) + 1-byte mask per pixel; read S,T and write their
    // gradients; per row read shift/stretch and write their gradients, Fisher and row loss
    int64_t knots = 0;
    for (int c = 0; c < pl->n_comp; ++c)
        knots += pl->desc[c].kind == JB_COMP_SPLINE ? pl->desc[c].n_knots : pl->desc[c].n_knots * pl->desc[c].n_ctrl_keys;
    info->algorithmic_bytes = pl->E * pl->P * (pl->f32 ? 13 : 25) + 2 * knots * 8 + pl->E * 7 * 8;
    info->kernel_launches = pl->launches; info->evaluations = pl->evaluations;
    info->serial_rows = pl->serial_rows;
    info->tiles_smooth = pl->tiles_kind[0]; info->tiles_general = pl->tiles_kind[1];
    info->tiles_serial = pl->tiles_kind[2]; info->tiles_empty = pl->tiles_kind[3];
    info->window_knots = strip ? kStripWin : pl->wincap;
    info->fused_variant = pl->variant;
    info->strip_reason = pl->strip_reason;
    info->strip_rows_per_block = strip ? pl->sK : 0;
    return JB_OK;
}

int jb_plan_set_params(jb_plan* pl, const double* params_all, int64_t n) {
    if (!pl || !params_all) return fail(JB_ERR_BAD_ARG, "null argument");
    if (n != pl->n_params) return fail(JB_ERR_BAD_ARG, "n_params %lld, plan has %lld", (long long)n, (long long)pl->n_params);
    JB_CUDA(cudaSetDevice(pl->device));
    pl->h_params.assign(params_all, params_all + n);
    JB_CUDA(cudaMemcpyAsync(pl->d_params, pl->h_params.data(), n * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    JB_CUDA(cudaStreamSynchronize(pl->stream));
    // The rows stay in the order (and the strip layout) they were sorted into: that order only has to keep
    // the rows of a group near one another, and an evaluation says so when it no longer does (the knot
    // window overflows: eval_sync / batch_settle re-sort with the parameters of the moment).  The first
    // upload, and one that moves a shift by more than half a knot, sort right away.
    bool moved = !pl->params_set;
    if (!moved)
        for (int c = 0; c < pl->n_comp && !moved; ++c) {
            const jb_component_desc& cd = pl->desc[c];
            if (cd.shift_offset < 0) continue;
            for (int64_t i = 0; i < cd.n_shift && !moved; ++i)
                if ((int64_t)pl->sorted_params.size() != n ||
                    std::fabs(params_all[cd.shift_offset + i] - pl->sorted_params[cd.shift_offset + i]) > 0.5 * cd.dx) moved = true;
        }
    pl->params_set = true;
    if (moved) pl->sort_dirty = true;
    return JB_OK;
}

int jb_plan_set_fit(jb_plan* pl, const int64_t* fit_index, int64_t n_fit) {
    if (!pl || n_fit < 0 || (n_fit > 0 && !fit_index)) return fail(JB_ERR_BAD_ARG, "bad fit map");
    for (int64_t i = 0; i < n_fit; ++i)
        if (fit_index[i] < 0 || fit_index[i] >= pl->n_params) return fail(JB_ERR_BAD_ARG, "fit index %lld outside params_all", (long long)fit_index[i]);
    JB_CUDA(cudaSetDevice(pl->device));
    JB_CUDA(cudaStreamSynchronize(pl->stream));
    int rc;
    if (n_fit != pl->n_fit || !pl->fit_bufs) {      // same length as before: the buffers are reused
        dev_free(pl, pl->d_fit_map); dev_free(pl, pl->d_pfit); dev_free(pl, pl->d_out);
        if ((rc = dev_alloc(pl, &pl->d_fit_map, (size_t)std::max<int64_t>(n_fit, 1)))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_pfit, (size_t)std::max<int64_t>(n_fit, 1)))) return rc;
        if ((rc = dev_alloc(pl, &pl->d_out, (size_t)n_fit + 1))) return rc;
        cudaFreeHost(pl->h_pfit); cudaFreeHost(pl->h_out);
        pl->h_pfit = pl->h_out = nullptr;
        JB_CUDA(cudaMallocHost(&pl->h_pfit, std::max<int64_t>(n_fit, 1) * sizeof(double)));
        JB_CUDA(cudaMallocHost(&pl->h_out, (n_fit + 1) * sizeof(double)));
        pl->fit_bufs = true;
    }
    if (n_fit) JB_CUDA(cudaMemcpy(pl->d_fit_map, fit_index, n_fit * sizeof(int64_t), cudaMemcpyHostToDevice));
    pl->h_fit_map.assign(fit_index, fit_index + n_fit);
    pl->n_fit = n_fit;
    for (int c = 0; c < pl->n_comp; ++c) {
        pl->fit_shift[c] = false;
        const jb_component_desc& cd = pl->desc[c];
        if (cd.shift_offset < 0) continue;
        for (int64_t i = 0; i < n_fit; ++i)
            if (fit_index[i] >= cd.shift_offset && fit_index[i] < cd.shift_offset + cd.n_shift) { pl->fit_shift[c] = true; break; }
    }
    pl->pack_dirty = true;
    return JB_OK;
}

int jb_plan_eval_device(jb_plan* pl, const double* d_p_fit, double* d_out, void* stream) {
    if (!pl || !d_out) return fail(JB_ERR_BAD_ARG, "null argument");
    JB_CUDA(cudaSetDevice(pl->device));
    return enqueue_eval(pl, d_p_fit, d_out, stream ? (cudaStream_t)stream : pl->stream);
}

int jb_plan_check(jb_plan* pl) {
    if (!pl) return fail(JB_ERR_BAD_ARG, "null plan");
    JB_CUDA(cudaSetDevice(pl->device));
    JB_CUDA(cudaDeviceSynchronize());
    return read_status(pl, pl->stream, nullptr);
}

int jb_plan_eval(jb_plan* pl, const double* p_fit, int64_t n_fit, double* loss, double* grad) {
    if (!pl || !loss) return fail(JB_ERR_BAD_ARG, "null argument");
    if (n_fit != pl->n_fit) return fail(JB_ERR_BAD_ARG, "n_fit %lld, plan expects %lld", (long long)n_fit, (long long)pl->n_fit);
    if (n_fit > 0 && !p_fit) return fail(JB_ERR_BAD_ARG, "null p_fit");
    JB_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = pl->stream;
    if (n_fit) {
        memcpy(pl->h_pfit, p_fit, n_fit * sizeof(double));
        JB_CUDA(cudaMemcpyAsync(pl->d_pfit, pl->h_pfit, n_fit * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    int rc = eval_sync(pl, n_fit ? pl->d_pfit : nullptr);
    if (rc) return rc;
    *loss = pl->h_out[0];
    if (grad && n_fit) memcpy(grad, pl->h_out + 1, n_fit * sizeof(double));
    if (!std::isfinite(*loss)) return fail(JB_ERR_NONFINITE, "loss is not finite");
    return JB_OK;
}

int jb_plan_forward(jb_plan* pl, const double* p_fit, int64_t n_fit, double* model_out) {
    if (!pl || !model_out) return fail(JB_ERR_BAD_ARG, "null argument");
    if (n_fit != 0 && n_fit != pl->n_fit) return fail(JB_ERR_BAD_ARG, "n_fit %lld, plan expects %lld", (long long)n_fit, (long long)pl->n_fit);
    if (!pl->params_set) return fail(JB_ERR_BAD_ARG, "jb_plan_set_params has not been called");
    JB_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = pl->stream;
    if (n_fit) {
        if (!p_fit) return fail(JB_ERR_BAD_ARG, "null p_fit");
        memcpy(pl->h_pfit, p_fit, n_fit * sizeof(double));
        JB_CUDA(cudaMemcpyAsync(pl->d_pfit, pl->h_pfit, n_fit * sizeof(double), cudaMemcpyHostToDevice, st));
        int rc = refresh_pack(pl);
        if (rc) return rc;
        jb::k_scatter_params<<<(unsigned)pl->nb[K_SCATTER], 256, 0, st>>>(&pl->d_pack->sc, pl->d_off + 2 * K_SCATTER, 1, pl->d_pfit);
        pl->launches++;
    }
    double* d_model = nullptr;
    const size_t n = (size_t)pl->E * pl->P;
    JB_CUDA(cudaMallocAsync(&d_model, n * sizeof(double), st));
    jb::ForwardParams fp;
    memset(&fp, 0, sizeof fp);
    fp.x = pl->d_x; fp.row_rev = pl->d_row_rev; fp.out = d_model;
    fp.n_rows = pl->E; fp.n_pix = pl->P; fp.ppad = pl->ppad; fp.n_comp = pl->n_comp;
    for (int c = 0; c < pl->n_comp; ++c) fp.comp[c] = pl->comp[c];
    const int64_t blocks = std::min<int64_t>((int64_t)((n + 255) / 256), 148LL * 16);
    jb::k_forward<<<(unsigned)blocks, 256, 0, st>>>(fp);
    pl->launches++;
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaMemcpyAsync(model_out, d_model, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaFreeAsync(d_model, st));
    JB_CUDA(cudaStreamSynchronize(st));
    return JB_OK;
}

int jb_plan_fisher(jb_plan* pl, int32_t component, double* f_out, int64_t n_shift) {
    if (!pl || !f_out) return fail(JB_ERR_BAD_ARG, "null argument");
    if (component < 0 || component >= pl->n_comp) return fail(JB_ERR_BAD_ARG, "component %d", component);
    int ci = 0;
    while (pl->ext_of[ci] != component) ++ci;
    if (ci >= pl->n_glob) return fail(JB_ERR_UNSUPPORTED, "Fisher information of a NormalizationModel shift is not computed");
    const jb_component_desc& cd = pl->desc[ci];
    if (cd.shift_offset < 0) return fail(JB_ERR_BAD_ARG, "component %d has no ShiftingModel", component);
    if (n_shift != cd.n_shift) return fail(JB_ERR_BAD_ARG, "n_shift %lld, component has %lld", (long long)n_shift, (long long)cd.n_shift);
    JB_CUDA(cudaSetDevice(pl->device));
    pl->force_der = true; pl->pack_dirty = true;     // the Fisher sums of every shift, whatever is in fit mode
    int rc = eval_sync(pl, nullptr);
    pl->force_der = false; pl->pack_dirty = true;
    if (rc) return rc;
    JB_CUDA(cudaMemcpy(f_out, pl->d_fisher_all + cd.shift_offset, n_shift * sizeof(double), cudaMemcpyDeviceToHost));
    return JB_OK;
}

int jb_plan_grid_search(jb_plan* pl, int32_t component, const double* grid, int64_t n_shift, int64_t n_grid,
                        double* loss_out) {
    if (!pl || !grid || !loss_out) return fail(JB_ERR_BAD_ARG, "null argument");
    if (component < 0 || component >= pl->n_comp) return fail(JB_ERR_BAD_ARG, "component %d", component);
    if (!pl->has_data) return fail(JB_ERR_BAD_ARG, "plan was created without ys/yivar (forward only)");
    if (!pl->params_set) return fail(JB_ERR_BAD_ARG, "jb_plan_set_params has not been called");
    int ci = 0;
    while (pl->ext_of[ci] != component) ++ci;
    const jb_component_desc& cd = pl->desc[ci];
    if (pl->f32) return fail(JB_ERR_UNSUPPORTED, "grid search is a float64 path: not available on a float32 plan");
    if (cd.kind != JB_COMP_SPLINE) return fail(JB_ERR_UNSUPPORTED, "grid search over the shift of a NormalizationModel component");
    if (cd.shift_offset < 0) return fail(JB_ERR_BAD_ARG, "component %d has no ShiftingModel", component);
    if (n_shift != cd.n_shift) return fail(JB_ERR_BAD_ARG, "n_shift %lld, component has %lld", (long long)n_shift, (long long)cd.n_shift);
    if (n_grid <= 0 || n_grid > (1 << 20)) return fail(JB_ERR_BAD_ARG, "n_grid %lld", (long long)n_grid);
    JB_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = pl->stream;
    double *d_grid = nullptr, *d_out = nullptr;
    JB_CUDA(cudaMallocAsync(&d_grid, (size_t)n_shift * n_grid * sizeof(double), st));
    JB_CUDA(cudaMallocAsync(&d_out, (size_t)pl->E * n_grid * sizeof(double), st));
    JB_CUDA(cudaMemcpyAsync(d_grid, grid, (size_t)n_shift * n_grid * sizeof(double), cudaMemcpyHostToDevice, st));
    jb::GridParams gp;
    memset(&gp, 0, sizeof gp);
    gp.data = pl->d_data; gp.n_rows = pl->E; gp.n_tiles = pl->n_tiles; gp.n_comp = pl->n_comp;
    for (int c = 0; c < pl->n_comp; ++c) gp.comp[c] = pl->comp[c];
    gp.search = ci; gp.grid = d_grid; gp.M = (int)n_grid; gp.out = d_out; gp.status = pl->d_status;
    const dim3 blocks((unsigned)pl->E, (unsigned)((n_grid + jb::kTile - 1) / jb::kTile));
    switch (cd.p_val) {
        case 1: jb::k_grid_search<1><<<blocks, jb::kTile, 0, st>>>(gp); break;
        case 2: jb::k_grid_search<2><<<blocks, jb::kTile, 0, st>>>(gp); break;
        default: jb::k_grid_search<3><<<blocks, jb::kTile, 0, st>>>(gp); break;
    }
    pl->launches++;
    JB_CUDA(cudaGetLastError());
    std::vector<double> rows((size_t)pl->E * n_grid);
    JB_CUDA(cudaMemcpyAsync(rows.data(), d_out, rows.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaFreeAsync(d_grid, st));
    JB_CUDA(cudaFreeAsync(d_out, st));
    int rc = read_status(pl, st, nullptr);     // synchronises the stream
    if (rc) return rc;
    // rows -> keys, in row order (model.py:777-781); ChiSquare: coefficient * 0.5 * ivar * res^2
    std::fill(loss_out, loss_out + n_shift * n_grid, 0.0);
    const std::vector<int>& key = pl->h_shift_key[ci];
    for (int64_t r = 0; r < pl->E; ++r)
        for (int64_t j = 0; j < n_grid; ++j) loss_out[(int64_t)key[r] * n_grid + j] += rows[(size_t)r * n_grid + j];
    for (int64_t i = 0; i < n_shift * n_grid; ++i) loss_out[i] *= 0.5 * pl->coef;
    return JB_OK;
}

int jb_plan_curvature(jb_plan* pl, int32_t component, double* h_out, int64_t n_shift) {
    if (!pl || !h_out) return fail(JB_ERR_BAD_ARG, "null argument");
    if (component < 0 || component >= pl->n_comp) return fail(JB_ERR_BAD_ARG, "component %d", component);
    if (!pl->has_data) return fail(JB_ERR_BAD_ARG, "plan was created without ys/yivar (forward only)");
    if (!pl->params_set) return fail(JB_ERR_BAD_ARG, "jb_plan_set_params has not been called");
    int ci = 0;
    while (pl->ext_of[ci] != component) ++ci;
    const jb_component_desc& cd = pl->desc[ci];
    if (pl->f32) return fail(JB_ERR_UNSUPPORTED, "the curvature is a float64 path: not available on a float32 plan");
    if (cd.kind != JB_COMP_SPLINE) return fail(JB_ERR_UNSUPPORTED, "curvature of the shift of a NormalizationModel component");
    if (cd.shift_offset < 0) return fail(JB_ERR_BAD_ARG, "component %d has no ShiftingModel", component);
    if (n_shift != cd.n_shift) return fail(JB_ERR_BAD_ARG, "n_shift %lld, component has %lld", (long long)n_shift, (long long)cd.n_shift);
    JB_CUDA(cudaSetDevice(pl->device));
    cudaStream_t st = pl->stream;
    double* d_out = nullptr;
    JB_CUDA(cudaMallocAsync(&d_out, (size_t)pl->E * sizeof(double), st));
    jb::GridParams gp;
    memset(&gp, 0, sizeof gp);
    gp.data = pl->d_data; gp.n_rows = pl->E; gp.n_tiles = pl->n_tiles; gp.n_comp = pl->n_comp;
    for (int c = 0; c < pl->n_comp; ++c) gp.comp[c] = pl->comp[c];
    gp.search = ci; gp.out = d_out; gp.status = pl->d_status;
    switch (cd.p_val) {
        case 1: jb::k_row_curvature<1><<<(unsigned)pl->E, jb::kTile, 0, st>>>(gp); break;
        case 2: jb::k_row_curvature<2><<<(unsigned)pl->E, jb::kTile, 0, st>>>(gp); break;
        default: jb::k_row_curvature<3><<<(unsigned)pl->E, jb::kTile, 0, st>>>(gp); break;
    }
    pl->launches++;
    JB_CUDA(cudaGetLastError());
    std::vector<double> rows((size_t)pl->E);
    JB_CUDA(cudaMemcpyAsync(rows.data(), d_out, rows.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaFreeAsync(d_out, st));
    int rc = read_status(pl, st, nullptr);     // synchronises the stream
    if (rc) return rc;
    std::fill(h_out, h_out + n_shift, 0.0);
    const std::vector<int>& key = pl->h_shift_key[ci];
    for (int64_t r = 0; r < pl->E; ++r) h_out[key[r]] += rows[(size_t)r];
    for (int64_t i = 0; i < n_shift; ++i) h_out[i] *= pl->coef;
    return JB_OK;
}

int jb_plan_stream(jb_plan* pl, void** stream_out) {
    if (!pl || !stream_out) return fail(JB_ERR_BAD_ARG, "null argument");
    *stream_out = (void*)pl->stream;
    return JB_OK;
}

int jb_plan_profile(jb_plan* pl, int enable) {
    if (!pl) return fail(JB_ERR_BAD_ARG, "null plan");
    pl->profiling = enable != 0;
    return JB_OK;
}

int jb_plan_profile_read(jb_plan* pl, double* fused_ms_total, int64_t* n_launches) {
    if (!pl || !fused_ms_total || !n_launches) return fail(JB_ERR_BAD_ARG, "null argument");
    JB_CUDA(cudaSetDevice(pl->device));
    if (pl->prof_stream) JB_CUDA(cudaStreamSynchronize(pl->prof_stream));
    double total = 0.0;
    for (size_t i = 0; i < pl->prof_used; ++i) {
        float ms = 0.f;
        JB_CUDA(cudaEventElapsedTime(&ms, pl->prof_events[i].first, pl->prof_events[i].second));
        total += ms;
    }
    *fused_ms_total = total;
    *n_launches = (int64_t)pl->prof_used;
    pl->prof_used = 0;
    return JB_OK;
}

// would an evaluation of this batch enqueue exactly what the previous one did?
static bool batch_is_clean(const jb_batch* b) {
    if (b->dirty || b->plans.empty()) return false;
    for (const jb_plan* pl : b->plans)
        if (pl->sort_dirty || pl->pack_dirty || pl->profiling || pl->wincap != b->plans[0]->wincap) return false;
    return true;
}

int jb_batch_create(jb_plan* const* plans, int32_t n, jb_batch** out) {
    if (!plans || n <= 0 || !out) return fail(JB_ERR_BAD_ARG, "bad batch");
    *out = nullptr;
    jb_plan* p0 = plans[0];
    for (int i = 0; i < n; ++i) {
        jb_plan* pl = plans[i];
        if (!pl) return fail(JB_ERR_BAD_ARG, "null plan in batch");
        int rc = check_supported_for_eval(pl);
        if (rc) return rc;
        if (pl->device != p0->device) return fail(JB_ERR_BAD_ARG, "plans of a batch must live on one device");
        if (pl->f32 != p0->f32) return fail(JB_ERR_UNSUPPORTED, "float32 and float64 plans cannot share a batch");
        if (pl->n_glob != p0->n_glob || pl->has_norm != p0->has_norm || pl->desc[0].p_val != p0->desc[0].p_val ||
            (pl->has_norm && pl->desc[pl->n_glob].p_val != p0->desc[p0->n_glob].p_val))
            return fail(JB_ERR_UNSUPPORTED, "plans of a batch must share the model shape (components, spline degree)");
        if (!pl->params_set) return fail(JB_ERR_BAD_ARG, "plan %d: jb_plan_set_params has not been called", i);
    }
    JB_CUDA(cudaSetDevice(p0->device));
    jb_batch* b = new jb_batch;
    b->plans.assign(plans, plans + n);
    b->device = p0->device;
    JB_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    b->dirty = true;
    *out = b;
    return JB_OK;
}

void jb_batch_destroy(jb_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->gstream) cudaStreamSynchronize(b->gstream);
    drop_graph(b->gexec);
    for (void* q : b->dev) cudaFree(q);
    if (b->d_flat_p) cudaFree(b->d_flat_p);
    if (b->d_flat_out) cudaFree(b->d_flat_out);
    if (b->h_flat_p) cudaFreeHost(b->h_flat_p);
    if (b->h_flat_out) cudaFreeHost(b->h_flat_out);
    if (b->d_stat_ptrs) cudaFree(b->d_stat_ptrs);
    if (b->d_stat_flat) cudaFree(b->d_stat_flat);
    if (b->h_stat_flat) cudaFreeHost(b->h_stat_flat);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int jb_batch_bind(jb_batch* b, const double* const* d_p_fit, double* const* d_out) {
    if (!b || !d_out) return fail(JB_ERR_BAD_ARG, "null argument");
    b->p_fit.assign(b->plans.size(), nullptr);
    if (d_p_fit) b->p_fit.assign(d_p_fit, d_p_fit + b->plans.size());
    b->out.assign(d_out, d_out + b->plans.size());
    b->dirty = true;
    b->host_bound = false;
    return JB_OK;
}

int jb_batch_eval_device(jb_batch* b, void* stream);

static int batch_status_buffers(jb_batch* b) {
    if (b->d_stat_ptrs) return JB_OK;
    const size_t n = b->plans.size();
    std::vector<int*> ptrs;
    for (jb_plan* pl : b->plans) ptrs.push_back(pl->d_status);
    JB_CUDA(cudaMalloc(&b->d_stat_ptrs, n * sizeof(int*)));
    JB_CUDA(cudaMemcpy(b->d_stat_ptrs, ptrs.data(), n * sizeof(int*), cudaMemcpyHostToDevice));
    JB_CUDA(cudaMalloc(&b->d_stat_flat, 2 * n * sizeof(int)));
    JB_CUDA(cudaMallocHost(&b->h_stat_flat, 2 * n * sizeof(int)));
    return JB_OK;
}

// After an evaluation was enqueued on st: wait, and give every plan whose knot window overflowed
// what jb_plan_eval gives it (eval_sync: first a re-sort of its rows with the current parameters,
// then a wider window / shorter row groups) and evaluate the whole batch again, until it fits.
// copy_out: re-enqueue the device->host copy of the flat output buffer behind each evaluation.
static int batch_settle(jb_batch* b, cudaStream_t st, bool copy_out) {
    const size_t n = b->plans.size();
    int rc;
    for (int attempt = 0;; ++attempt) {
        jb::k_gather_status<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b->d_stat_ptrs, (int)n, b->d_stat_flat);
        JB_CUDA(cudaMemcpyAsync(b->h_stat_flat, b->d_stat_flat, 2 * n * sizeof(int), cudaMemcpyDeviceToHost, st));
        JB_CUDA(cudaStreamSynchronize(st));
        bool again = false;
        for (size_t i = 0; i < n; ++i) {
            if (!b->h_stat_flat[2 * i] && !b->h_stat_flat[2 * i + 1]) continue;
            jb_plan* pl = b->plans[i];
            bool window = false;
            rc = read_status(pl, st, &window);
            if (rc == JB_OK) continue;
            if (rc == JB_ERR_OUT_OF_RANGE && b->soft_range) continue;     // its loss came back NaN (k_finish)
            if (!window || attempt >= 14) return rc;
            const std::string msg = g_err;
            bool retry = false;
            int rc2 = window_retry(pl, attempt, &retry);
            if (rc2) return rc2;
            if (!retry) { g_err = msg; return rc; }
            again = true;
        }
        if (!again) return JB_OK;
        if ((rc = jb_batch_eval_device(b, st))) return rc;
        if (copy_out)
            JB_CUDA(cudaMemcpyAsync(b->h_flat_out, b->d_flat_out, b->off_out[n] * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
}

int jb_batch_eval_device_checked(jb_batch* b, void* stream) {
    if (!b) return fail(JB_ERR_BAD_ARG, "null batch");
    JB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    int rc = batch_status_buffers(b);
    if (rc) return rc;
    if ((rc = jb_batch_eval_device(b, st))) return rc;
    return batch_settle(b, st, false);
}

namespace {
// (re)build the flat device buffers and their pinned host staging for the plans' current fit sizes
int batch_host_bind(jb_batch* b) {
    const size_t n = b->plans.size();
    int rc0;
    if (b->host_bound)
        for (size_t i = 0; i < n; ++i)
            if (b->off_p[i + 1] - b->off_p[i] != b->plans[i]->n_fit) b->host_bound = false;   // a fit vector changed size
    if (b->host_bound) return JB_OK;
    JB_CUDA(cudaStreamSynchronize(b->stream));
    if (b->d_flat_p) { cudaFree(b->d_flat_p); cudaFree(b->d_flat_out); cudaFreeHost(b->h_flat_p); cudaFreeHost(b->h_flat_out); }
    b->d_flat_p = b->d_flat_out = b->h_flat_p = b->h_flat_out = nullptr;
    if ((rc0 = batch_status_buffers(b))) return rc0;
    b->off_p.assign(n + 1, 0); b->off_out.assign(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
        b->off_p[i + 1] = b->off_p[i] + b->plans[i]->n_fit;
        b->off_out[i + 1] = b->off_out[i] + b->plans[i]->n_fit + 1;
    }
    const size_t np = (size_t)std::max<int64_t>(b->off_p[n], 1), no = (size_t)b->off_out[n];
    JB_CUDA(cudaMalloc(&b->d_flat_p, np * sizeof(double)));
    JB_CUDA(cudaMalloc(&b->d_flat_out, no * sizeof(double)));
    JB_CUDA(cudaMallocHost(&b->h_flat_p, np * sizeof(double)));
    JB_CUDA(cudaMallocHost(&b->h_flat_out, no * sizeof(double)));
    b->p_fit.resize(n); b->out.resize(n);
    for (size_t i = 0; i < n; ++i) { b->p_fit[i] = b->d_flat_p + b->off_p[i]; b->out[i] = b->d_flat_out + b->off_out[i]; }
    b->dirty = true;
    b->host_bound = true;
    return JB_OK;
}

// pinned fit vectors -> device, one evaluation of every plan, [loss, gradient] -> pinned (all asynchronous)
int batch_host_launch(jb_batch* b) {
    const size_t n = b->plans.size();
    cudaStream_t st = b->stream;
    if (b->off_p[n]) JB_CUDA(cudaMemcpyAsync(b->d_flat_p, b->h_flat_p, b->off_p[n] * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = jb_batch_eval_device(b, st);
    if (rc) return rc;
    JB_CUDA(cudaMemcpyAsync(b->h_flat_out, b->d_flat_out, b->off_out[n] * sizeof(double), cudaMemcpyDeviceToHost, st));
    return JB_OK;
}
}  // namespace

int jb_batch_eval_host_begin(jb_batch* b, const double* const* p_fit_host) {
    if (!b || !p_fit_host) return fail(JB_ERR_BAD_ARG, "null argument");
    JB_CUDA(cudaSetDevice(b->device));
    const size_t n = b->plans.size();
    int rc = batch_host_bind(b);
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) {
        if (b->plans[i]->n_fit && !p_fit_host[i]) return fail(JB_ERR_BAD_ARG, "null p_fit for plan %zu", i);
        if (b->plans[i]->n_fit) memcpy(b->h_flat_p + b->off_p[i], p_fit_host[i], b->plans[i]->n_fit * sizeof(double));
    }
    if ((rc = batch_host_launch(b))) return rc;
    b->host_pending = true;
    return JB_OK;
}

int jb_batch_host_buffers(jb_batch* b, double** p_fit_pinned, double** out_pinned, int64_t* off_p, int64_t* off_out) {
    if (!b || !p_fit_pinned || !out_pinned || !off_p || !off_out) return fail(JB_ERR_BAD_ARG, "null argument");
    if (b->host_pending) return fail(JB_ERR_BAD_ARG, "an evaluation is pending (jb_batch_eval_host_end)");
    JB_CUDA(cudaSetDevice(b->device));
    int rc = batch_host_bind(b);
    if (rc) return rc;
    const size_t n = b->plans.size();
    *p_fit_pinned = b->h_flat_p; *out_pinned = b->h_flat_out;
    for (size_t i = 0; i <= n; ++i) { off_p[i] = b->off_p[i]; off_out[i] = b->off_out[i]; }
    return JB_OK;
}

int jb_batch_eval_pinned_begin(jb_batch* b) {
    if (!b) return fail(JB_ERR_BAD_ARG, "null batch");
    if (b->host_pending) return fail(JB_ERR_BAD_ARG, "an evaluation is pending (jb_batch_eval_pinned_end / jb_batch_eval_host_end)");
    if (!b->host_bound || !b->h_flat_p) return fail(JB_ERR_BAD_ARG, "jb_batch_host_buffers has not been called");
    JB_CUDA(cudaSetDevice(b->device));
    const size_t n = b->plans.size();
    for (size_t i = 0; i < n; ++i)
        if (b->off_p[i + 1] - b->off_p[i] != b->plans[i]->n_fit)
            return fail(JB_ERR_BAD_ARG, "the fit vector of plan %zu changed size: call jb_batch_host_buffers again", i);
    int rc = batch_host_launch(b);
    if (rc) return rc;
    b->host_pending = true;
    return JB_OK;
}

int jb_batch_eval_pinned_end(jb_batch* b) {
    if (!b) return fail(JB_ERR_BAD_ARG, "null batch");
    if (!b->host_pending) return fail(JB_ERR_BAD_ARG, "jb_batch_eval_pinned_begin has not been called");
    JB_CUDA(cudaSetDevice(b->device));
    const size_t n = b->plans.size();
    b->host_pending = false;
    int rc = batch_settle(b, b->stream, true);
    if (rc) return rc;
    if (!b->soft_range)
        for (size_t i = 0; i < n; ++i)
            if (!std::isfinite(b->h_flat_out[b->off_out[i]])) return fail(JB_ERR_NONFINITE, "loss of plan %zu is not finite", i);
    return JB_OK;
}

int jb_batch_eval_pinned(jb_batch* b) {
    int rc = jb_batch_eval_pinned_begin(b);
    if (rc) return rc;
    return jb_batch_eval_pinned_end(b);
}

int jb_batch_eval_host_end(jb_batch* b, double* const* out_host) {
    if (!b || !out_host) return fail(JB_ERR_BAD_ARG, "null argument");
    if (!b->host_pending) return fail(JB_ERR_BAD_ARG, "jb_batch_eval_host_begin has not been called");
    JB_CUDA(cudaSetDevice(b->device));
    const size_t n = b->plans.size();
    cudaStream_t st = b->stream;
    b->host_pending = false;
    int rc = batch_settle(b, st, true);
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) {
        if (!out_host[i]) return fail(JB_ERR_BAD_ARG, "null output for plan %zu", i);
        memcpy(out_host[i], b->h_flat_out + b->off_out[i], (b->plans[i]->n_fit + 1) * sizeof(double));
        if (!std::isfinite(out_host[i][0]) && !b->soft_range) return fail(JB_ERR_NONFINITE, "loss of plan %zu is not finite", i);
    }
    return JB_OK;
}

int jb_batch_soft_range(jb_batch* b, int enable) {
    if (!b) return fail(JB_ERR_BAD_ARG, "null batch");
    b->soft_range = enable != 0;
    return JB_OK;
}

int jb_batch_eval_host(jb_batch* b, const double* const* p_fit_host, double* const* out_host) {
    if (!out_host) return fail(JB_ERR_BAD_ARG, "null argument");
    int rc = jb_batch_eval_host_begin(b, p_fit_host);
    if (rc) return rc;
    return jb_batch_eval_host_end(b, out_host);
}

int jb_batch_eval_device(jb_batch* b, void* stream) {
    if (!b) return fail(JB_ERR_BAD_ARG, "null batch");
    if (b->out.size() != b->plans.size()) return fail(JB_ERR_BAD_ARG, "jb_batch_bind has not been called");
    JB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    const int n = (int)b->plans.size();
    int rc;
    bool regroup = b->dirty;
    for (jb_plan* pl : b->plans)
        if (pl->sort_dirty) { if ((rc = resort_rows(pl))) return rc; }
    {   // one launch -> one kernel: plans that disagree on k_strip all leave it (their rows per block may
        // differ: the launch sizes its shared memory for the longest); plans that then still disagree on the variant all take the general kernel
        auto all_same = [&]() {
            bool same = true;
            for (jb_plan* pl : b->plans) {
                const int v = plan_variant(pl), v0 = plan_variant(b->plans[0]);
                same = same && v == v0;
            }
            return same;
        };
        std::vector<std::pair<bool, bool>> old;
        for (jb_plan* pl : b->plans) { old.emplace_back(pl->force_nostrip, pl->force_v1); pl->force_nostrip = pl->force_v1 = false; }
        bool want_nostrip = false, want_v1 = false;
        if (!all_same()) {
            want_nostrip = true;
            for (jb_plan* pl : b->plans) pl->force_nostrip = true;
            want_v1 = !all_same();
        }
        for (size_t i = 0; i < b->plans.size(); ++i) {
            jb_plan* pl = b->plans[i];
            pl->force_nostrip = want_nostrip; pl->force_v1 = want_v1;
            if (old[i].first != want_nostrip || old[i].second != want_v1) pl->pack_dirty = true;
        }
    }
    for (jb_plan* pl : b->plans) {
        if (pl->pack_dirty) regroup = true;
        if ((rc = refresh_pack(pl))) return rc;
        if (pl->wincap != b->plans[0]->wincap) {   // one kernel launch -> one shared-memory layout
            int w = 0;
            for (jb_plan* q : b->plans) w = std::max(w, q->wincap);
            for (jb_plan* q : b->plans)
                if (q->wincap != w) { q->wincap = w; if ((rc = alloc_scratch(q))) return rc; if ((rc = refresh_pack(q))) return rc; }
            regroup = true;
        }
    }
    if (regroup) {
        JB_CUDA(cudaStreamSynchronize(st));
        if (b->gstream && b->gstream != st) JB_CUDA(cudaStreamSynchronize(b->gstream));
        drop_graph(b->gexec);
        b->steady = 0;
        for (void* q : b->dev) cudaFree(q);
        b->dev.clear();
        std::vector<jb::ScatterParams> sc(n); std::vector<jb::PrepareParams> pr(n);
        std::vector<jb::FusedParams> fu(n); std::vector<jb::ReduceParams> re(n);
        std::vector<jb::StripParams> stv(n); std::vector<jb::StripPrepParams> spv(n);
        std::vector<int> off((size_t)K_COUNT * (n + 1), 0);
        b->scatter = true;
        for (int i = 0; i < n; ++i) {
            const jb_plan::Pack& k = *b->plans[i]->h_pack;
            sc[i] = k.sc; pr[i] = k.pr; fu[i] = k.fu; re[i] = k.re; stv[i] = k.st; spv[i] = k.sp;
            sc[i].p_fit = b->p_fit.empty() ? nullptr : b->p_fit[i];
            re[i].out = b->out[i];
            for (int q = 0; q < K_COUNT; ++q) off[(size_t)q * (n + 1) + i + 1] = off[(size_t)q * (n + 1) + i] + b->plans[i]->nb[q];
        }
        for (int q = 0; q < K_COUNT; ++q) b->total[q] = off[(size_t)q * (n + 1) + n];
        auto up = [&](const void* h, size_t bytes, void** d) -> int {
            JB_CUDA(cudaMalloc(d, bytes));
            b->dev.push_back(*d);
            JB_CUDA(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
            return JB_OK;
        };
        if ((rc = up(sc.data(), n * sizeof sc[0], (void**)&b->d_sc))) return rc;
        if ((rc = up(pr.data(), n * sizeof pr[0], (void**)&b->d_pr))) return rc;
        if ((rc = up(fu.data(), n * sizeof fu[0], (void**)&b->d_fu))) return rc;
        if ((rc = up(re.data(), n * sizeof re[0], (void**)&b->d_re))) return rc;
        if ((rc = up(stv.data(), n * sizeof stv[0], (void**)&b->d_st))) return rc;
        if ((rc = up(spv.data(), n * sizeof spv[0], (void**)&b->d_sp))) return rc;
        if ((rc = up(off.data(), off.size() * sizeof(int), (void**)&b->d_off))) return rc;
        b->dirty = false;
    }
    LaunchSet L;
    L.sc = b->d_sc; L.pr = b->d_pr; L.fu = b->d_fu; L.re = b->d_re; L.st = b->d_st; L.sp = b->d_sp;
    L.strip_K = 0;
    for (jb_plan* q : b->plans) L.strip_K = std::max(L.strip_K, q->sK);
    for (int q = 0; q < K_COUNT; ++q) { L.off[q] = b->d_off + (size_t)q * (n + 1); L.total[q] = b->total[q]; }
    L.n = n; L.n_comp = b->plans[0]->n_glob; L.p_val = b->plans[0]->desc[0].p_val; L.wincap = b->plans[0]->wincap;
    L.pvn = b->plans[0]->has_norm ? b->plans[0]->desc[b->plans[0]->n_glob].p_val : 0;
    L.max_knots = 0;
    for (jb_plan* q : b->plans) L.max_knots = std::max(L.max_knots, q->max_knots);
    L.variant = b->plans[0]->variant;
    L.scatter = !b->p_fit.empty();
    bool replayed = false;
    if (graphs_enabled() && b->graphs_ok && !b->plans[0]->profiling && !stream_is_capturing(st)) {
        if (b->gexec && b->gstream != st) { JB_CUDA(cudaStreamSynchronize(b->gstream)); drop_graph(b->gexec); b->steady = 0; }
        if (!b->gexec && ++b->steady >= 2) {
            // nothing changed since the previous call: record the sequence, replay it from now on
            if (!record_graph(st, &b->gexec, &rc, [&] { return launch_all(L, st, nullptr, nullptr, nullptr); })) b->graphs_ok = false;
            if (rc) return rc;
            b->gstream = st;
        }
        if (b->gexec) { JB_CUDA(cudaGraphLaunch(b->gexec, st)); replayed = true; }
    }
    if (!replayed && (rc = launch_all(L, st, nullptr, nullptr, b->plans[0]))) return rc;
    b->launches += L.scatter ? 6 : 5;
    for (jb_plan* pl : b->plans) pl->evaluations++;
    return JB_OK;
}

int jb_batch_check(jb_batch* b) {
    if (!b) return fail(JB_ERR_BAD_ARG, "null batch");
    JB_CUDA(cudaSetDevice(b->device));
    JB_CUDA(cudaDeviceSynchronize());
    for (jb_plan* pl : b->plans) {
        int rc = read_status(pl, pl->stream, nullptr);
        if (rc) return rc;
    }
    return JB_OK;
}

int64_t jb_batch_launches(const jb_batch* b) { return b ? b->launches : 0; }

int jb_plan_grad_all(jb_plan* pl, double* grad_all, int64_t n) {
    if (!pl || !grad_all) return fail(JB_ERR_BAD_ARG, "null argument");
    if (n != pl->n_params) return fail(JB_ERR_BAD_ARG, "n_params mismatch");
    JB_CUDA(cudaSetDevice(pl->device));
    JB_CUDA(cudaStreamSynchronize(pl->stream));
    JB_CUDA(cudaMemcpy(grad_all, pl->d_grad_all, n * sizeof(double), cudaMemcpyDeviceToHost));
    return JB_OK;
}

}  // extern "C"

#include "jb_group.cuh"
#include "jb_lbfgs.cuh"
