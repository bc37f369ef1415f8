// k_strip: the fused loss + gradient + Fisher kernel in "column-walk" form (round 2).
//
// k_fused2 (jb_fused2.cuh) walks ALONG a row: a lane's pixels cross a knot every 1.6 pixels, so its
// register window slides all the time (selects, a shared-memory read-modify-write per pixel) and the
// basis weights are rebuilt per pixel.  k_strip walks DOWN the rows instead:
//
//   * the rows of a chunk are stored in sorted order (by first x - shift), in strips of 32 pixel
//     columns: record (strip, row) = x[32] | y[32] | ivar[32], 768 contiguous bytes, consecutive rows
//     of a strip contiguous -- one bulk async copy (TMA engine) brings kStageRows rows;
//   * a warp owns one strip for a long run of rows, lane l owns pixel column l.  Down the sorted rows
//     the knot coordinate of a column moves slowly: over a block of K rows it stays inside TWO
//     adjacent knot intervals [I, I+2) (checked, see below).  On that span the spline is ONE
//     polynomial plus one truncated power:
//         S(g) = P0 + P1 g + .. + Pn g^n + d * max(g - 1/2, 0)^n,     g = t - (I + 1/2) in [-1/2, 3/2)
//     (P = the piece of interval I in its centred coordinate, d = the (n+1)-th difference of the
//     control points = the jump of the n-th derivative at the knot between I and I+1).  So per pixel:
//     no floor, no basis weights, no window slide -- a Horner evaluation that also yields S';
//   * the adjoint of the control points is accumulated in the same basis, as moments
//         M_j += q g^j (j = 0..n),   Md += q max(g - 1/2, 0)^n,        q = ivar (y - m) a
//     in registers that never move inside a block.  Every K rows ("checkpoint") a lane whose base
//     interval changed adds its moments to a table in shared memory and loads the polynomial of the
//     new base; at the end of the task the table is converted to per-knot sums
//         dL/dc[klo+k] += sum_j A[j][k] M_j + D[k] Md
//     and published as one gradient window per task, the format the reducers already take;
//   * per-row sums (shift gradient, stretch gradient, Fisher) cross the 32 lanes once per row: the
//     lanes park their terms in a swizzled shared-memory tile and every 8 rows the warp adds them up
//     transposed (fixed order), one store per row and sum.
//
// Preconditions, decided per plan on the host and re-checked by the kernel for every weighted pixel
// (a violation raises kStatStrip and the caller falls back to k_fused2 / k_fused):
//   (1) separable grid: x[row][p] - xref[row] is the same for the rows of a block to within
//       kStripTol knots (true for a spectrograph whose wavelength solution is stable; shifts and
//       per-epoch offsets are in xref - shift);
//   (2) the keys xref[row] - shift_c[row] of the rows of a block span less than 1 - 2 kStripTol knots
//       for every component (rows sorted by the key of the first shifted component);
//   (3) columns NCLS apart are more than a knot apart (lanes l and l + NCLS share a copy of the table).
// Reference semantics per pixel: cardinal_vmap_model (jabble/model.py:951-989), ShiftingModel /
// StretchingModel (model.py:924-946), Additive / Composite (model.py:670-697), ChiSquare
// (loss.py:164-173); adjoints as in SURVEY.md section 3.3; Fisher sums model.py:893-897.
// No float atomics; every sum has one fixed order.
#pragma once
#include <type_traits>
#include "jb_fused2.cuh"

namespace jb {

constexpr int kStripPx = 32;                       // pixel columns of a strip = lanes
constexpr int kStripRec = 3 * kStripPx;            // doubles per (strip, row) record
constexpr int kStripRecBytes = kStripRec * 8;      // 768
#ifndef JB_STRIP_STAGE_ROWS
#define JB_STRIP_STAGE_ROWS 4
#endif
#ifndef JB_STRIP_STAGES
#define JB_STRIP_STAGES 2
#endif
constexpr int kStageRows = JB_STRIP_STAGE_ROWS;    // rows per bulk copy
constexpr int kStripStages = JB_STRIP_STAGES;      // copies in flight per warp
constexpr int kBatchRows = 8;                      // rows per transposed row-sum reduction
constexpr int kStripWs = 32;                       // base intervals a task's table holds (lane j builds entry j)
constexpr int kStripMaxK = 32;                     // rows per block, at most (multiple of kBatchRows)
constexpr double kStripTol = 0.05;                 // knots: the most precondition (1) allows; a plan's own tolerance
                                                   // (StripParams::tol) is what its pixel grid needs, usually far less
constexpr int kStatStrip = 8;                      // status bit: a weighted pixel left its two-interval span
constexpr int kStripNoBase = INT_MIN;
#ifndef JB_STRIP_WARPS
#define JB_STRIP_WARPS 4
#endif
constexpr int kStripWarps = JB_STRIP_WARPS;        // independent tasks per block

// per sorted row, refreshed every evaluation (k_prepare_strip)
template <int NC> struct alignas(16) StripRow { double sh[NC]; double st[NC]; };
// per block of K sorted rows and component: range of key = xref - shift over the block's rows
struct alignas(16) StripBlk { double kmin[2], kmax[2]; int wide; int pad_[3]; };   // wide: the range breaks precondition (2)

struct StripParams {
    const double* data;      // [n_strips][E][3][32]  x | y | ivar, rows in sorted order, mask folded (ivar = 0)
    const double* xoff;      // [n_strips][n_blocks][32]  column offset x - xref of the block (-inf: column has no weighted pixel)
    const double* xoff_rng;  // [n_rgroups][n_strips][2]  smallest / largest finite xoff of the task (min > max: none)
    const void* rowc;        // [E] StripRow<NC>, sorted order
    const StripBlk* blk;     // [n_blocks]
    const int* rowid;        // [E] caller's row of sorted position
    int64_t n_rows;
    int n_strips, K, blocks_per_task, n_blocks, n_rgroups, ncls;   // ncls: lanes this far apart never share a base interval
    int64_t n_tasks;
    int wincap;              // doubles per published gradient window (>= kStripWs + PV + 1, multiple of 32)
    double tol;              // knots: how far a row's pixel grid may deviate from the block's column offsets
    int n_comp;
    CompDev comp[kMaxComp];
    double* part;            // [n_comp][n_tasks][wincap]
    int* part_base;          // [n_comp][n_tasks]
    double* rowpart;         // [E][n_strips][NS] (NS = f2_nsums, not padded), rows in SORTED order (k_reduce_stage1 maps them back through rowid)
    double* losspart;        // [n_tasks]
    int* status;
};

struct StripPrepParams {
    const int* rowid; const double* xref;      // [E] sorted order; xref = +inf for a row without weighted pixels
    void* rowc; StripBlk* blk;
    int64_t n_rows; int K, n_blocks, n_comp;
    double tol;                                 // as StripParams::tol
    CompDev comp[kMaxComp];
    int* status;
};

// One warp per block of K rows: row records in sorted order and the key range of the block.
template <int NC>
__global__ void k_prepare_strip(const StripPrepParams* __restrict__ arr, const int* __restrict__ off, int n) {
    const int pi = find_plan(off, n, blockIdx.x);
    const StripPrepParams& a = arr[pi];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = (blockIdx.x - off[pi]) * (blockDim.x >> 5) + warp;
    if (b == 0 && lane == 0 && a.status) { a.status[0] = 0; a.status[1] = 0; }
    if (b >= a.n_blocks) return;
    const int64_t pos = (int64_t)b * a.K + lane;
    const bool have = lane < a.K && pos < a.n_rows;
    StripRow<NC> rec;
    double kmin[NC], kmax[NC];
    const int r = have ? a.rowid[pos] : 0;
    const double xr = have ? a.xref[pos] : INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const CompDev& cd = a.comp[c];
        rec.sh[c] = (have && cd.shift) ? cd.shift[cd.shift_key[r]] : 0.0;
        rec.st[c] = (have && cd.stretch) ? cd.stretch[cd.stretch_key[r]] : 1.0;
        const double key = xr - rec.sh[c];
        const bool fin = have && xr < 1.0e300;
        kmin[c] = fin ? key : INFINITY;
        kmax[c] = fin ? key : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin[c] = fmin(kmin[c], __shfl_xor_sync(0xffffffffu, kmin[c], o));
            kmax[c] = fmax(kmax[c], __shfl_xor_sync(0xffffffffu, kmax[c], o));
        }
    }
    if (have) reinterpret_cast<StripRow<NC>*>(a.rowc)[pos] = rec;
    if (lane == 0) {
        bool wide = false;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (kmin[c] <= kmax[c] && !((kmax[c] - kmin[c]) * a.comp[c].inv_dx < 1.0 - 2.0 * a.tol - 1.0e-6)) wide = true;

        StripBlk bk;
#pragma unroll
        for (int c = 0; c < 2; ++c) { bk.kmin[c] = c < NC ? kmin[c < NC ? c : 0] : 0.0; bk.kmax[c] = c < NC ? kmax[c < NC ? c : 0] : 0.0; }
        bk.wide = wide ? 1 : 0; bk.pad_[0] = bk.pad_[1] = bk.pad_[2] = 0;
        a.blk[b] = bk;
    }
}

// ---- the spline in the two-interval basis ------------------------------------------------------
// Coordinate h = t - (knot between interval I and I+1), in [-1, 1).  Piece of interval I (Basis<PV>,
// jb_device.cuh) expanded in powers of h:  P_j = sum_k A[j][k] c[klo + k];  the next piece differs by
// d h^PV, d = sum_k D[k] c[klo + k] (k = 0..PV+1: the (PV+1)-th difference of the control points).
// knots(): the transposed map, moments of the adjoint -> per-knot sums a[k] = sum_j A[j][k] M[j] + D[k] Md.
template <int PV> struct StripBasis;
template <> struct StripBasis<1> {
    __device__ __forceinline__ static void coeffs(const double (&c)[3], double (&P)[2], double& d) {
        P[0] = c[1];
        P[1] = c[1] - c[0];
        d = (c[0] - 2.0 * c[1]) + c[2];
    }
    __device__ __forceinline__ static void knots(const double (&M)[2], double Md, double (&a)[3]) {
        a[0] = Md - M[1];
        a[1] = (M[0] + M[1]) - 2.0 * Md;
        a[2] = Md;
    }
};
template <> struct StripBasis<2> {
    __device__ __forceinline__ static void coeffs(const double (&c)[4], double (&P)[3], double& d) {
        P[0] = c[1] + c[2];
        P[1] = 2.0 * (c[2] - c[1]);
        P[2] = (c[0] - 2.0 * c[1]) + c[2];
        d = (c[3] - c[0]) + 3.0 * (c[1] - c[2]);
    }
    __device__ __forceinline__ static void knots(const double (&M)[3], double Md, double (&a)[4]) {
        a[0] = M[2] - Md;
        a[1] = (M[0] - 2.0 * (M[1] + M[2])) + 3.0 * Md;
        a[2] = (M[0] + 2.0 * M[1] + M[2]) - 3.0 * Md;
        a[3] = Md;
    }
};
template <> struct StripBasis<3> {
    __device__ __forceinline__ static void coeffs(const double (&c)[5], double (&P)[4], double& d) {
        P[0] = (c[1] + c[3]) + 4.0 * c[2];
        P[1] = 3.0 * (c[3] - c[1]);
        P[2] = 3.0 * ((c[1] + c[3]) - 2.0 * c[2]);
        P[3] = (c[3] - c[0]) + 3.0 * (c[1] - c[2]);
        d = ((c[0] + c[4]) - 4.0 * (c[1] + c[3])) + 6.0 * c[2];
    }
    __device__ __forceinline__ static void knots(const double (&M)[4], double Md, double (&a)[5]) {
        a[0] = Md - M[3];
        a[1] = (M[0] - 3.0 * M[1]) + 3.0 * (M[2] + M[3]) - 4.0 * Md;
        a[2] = (4.0 * M[0] - 6.0 * M[2]) - 3.0 * M[3] + 6.0 * Md;
        a[3] = (M[0] + 3.0 * (M[1] + M[2]) + M[3]) - 4.0 * Md;
        a[4] = Md;
    }
};

// per lane and component: the polynomial of the base interval and the moments of its adjoint
template <int PV> struct StripLane {
    double P[PV + 1], d, dd;       // dd = PV * d (the derivative of the truncated power)
    double M[PV + 1], Md;
    double ngofs;        // -(base + 1) for odd degrees, -(base + 1/2) for even ones: h = (xs - xp0) / dx + ngofs
    int base;            // kStripNoBase: none (column without weighted pixels)
};

// max(h, 0) without the FP64 pipe: clear the value when its sign bit is set
__device__ __forceinline__ double relu_bits(double h) {
    const int hi = __double2hiint(h), lo = __double2loint(h);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, lo & keep);
}

// One pixel.  sums: compact per-row terms in f2_slot order (jb_fused2.cuh).
template <int NC, int PV, int FL>
__device__ __forceinline__ void strip_pixel(double x, double y, double iv, const StripRow<NC>& rc, const double (&xp0)[NC],
                                            const double (&inv_dx)[NC], StripLane<PV> (&L)[NC], double& loss, int& bad,
                                            double (&sums)[f2_pad(f2_nsums(NC, FL))]) {
    double h[NC], hpn[NC], S[NC], dS[NC];
    double m = 0.0;
    int out = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double xs = __dsub_rn(x, rc.sh[c]);                  // ShiftingModel.call, model.py:927 (operation order is part of parity)
        const double dd = __dsub_rn(xs, xp0[c]);                   // cardinal_vmap_model, model.py:979
        h[c] = __fma_rn(dd, inv_dx[c], L[c].ngofs);
#ifdef JB_STRIP_CHECK   // checking builds: every weighted pixel verifies that it sits in [I, I+2)
        out |= ((__double2hiint(h[c]) & 0x7fffffff) >= 0x3ff00000) ? 1 : 0;
#endif
        const double hp = relu_bits(h[c]);
        // Horner for S and S' together
        double s, ds;
        if (PV == 3) {
            const double b2 = __fma_rn(L[c].P[3], h[c], L[c].P[2]);
            const double b1 = __fma_rn(b2, h[c], L[c].P[1]);
            s = __fma_rn(b1, h[c], L[c].P[0]);
            const double hp2 = hp * hp;
            hpn[c] = hp2 * hp;
            s = __fma_rn(L[c].d, hpn[c], s);
            if (f2_der(FL, c)) {
                const double c2 = __fma_rn(L[c].P[3], h[c], b2);
                ds = __fma_rn(c2, h[c], b1);
                ds = __fma_rn(L[c].dd, hp2, ds);
            }
        } else if (PV == 2) {
            const double b1 = __fma_rn(L[c].P[2], h[c], L[c].P[1]);
            s = __fma_rn(b1, h[c], L[c].P[0]);
            hpn[c] = hp * hp;
            s = __fma_rn(L[c].d, hpn[c], s);
            if (f2_der(FL, c)) {
                ds = __fma_rn(L[c].P[2], h[c], b1);
                ds = __fma_rn(L[c].dd, hp, ds);
            }
        } else {
            s = __fma_rn(L[c].P[1], h[c], L[c].P[0]);
            hpn[c] = hp;
            s = __fma_rn(L[c].d, hp, s);
            if (f2_der(FL, c)) ds = L[c].P[1] + (__double2hiint(h[c]) >= 0 ? L[c].d : 0.0);   // right-continuous, as floor()
        }
        S[c] = s;
        if (f2_der(FL, c)) dS[c] = ds;
        if (c == 0) m = f2_str(FL, c) ? rc.st[c] * s : s;
        else m = f2_str(FL, c) ? __fma_rn(rc.st[c], s, m) : m + s;       // StretchingModel / AdditiveModel
    }
#ifdef JB_STRIP_CHECK
    bad |= (__double2hiint(iv) != 0) ? out : 0;                            // unweighted pixels may sit anywhere
#else
    (void)out;
#endif
    const double res = y - m;                                              // ChiSquare, loss.py:169-171
    const double wr = iv * res;
    loss = __fma_rn(wr, res, loss);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double q = f2_str(FL, c) ? wr * rc.st[c] : wr;
        double t = q;
        L[c].M[0] += q;
#pragma unroll
        for (int j = 1; j <= PV; ++j) { t *= h[c]; L[c].M[j] += t; }
        L[c].Md = __fma_rn(q, hpn[c], L[c].Md);
        if (f2_der(FL, c)) {
            const int s0 = f2_slot(NC, FL, c, 0), s2 = f2_slot(NC, FL, c, 2);
            sums[s0] = wr * dS[c];                                         // d/d shift   (stretch, 1/PV, PV/dx: later)
            sums[s2] = (iv * dS[c]) * dS[c];                               // Fisher
        }
        if (f2_str(FL, c)) sums[f2_slot(NC, FL, c, 1)] = wr * S[c];        // d/d stretch
    }
}

// NR consecutive rows of the strip as independent instruction streams (the compiler interleaves them:
// the FP64 dependency chain of one pixel is ~16 operations long).  Returns the row terms in sums.
template <int NC, int PV, int FL, int NR>
__device__ __forceinline__ void strip_rows(const double* px, const StripRow<NC>* rowc, const double (&xp0)[NC],
                                           const double (&inv_dx)[NC], StripLane<PV> (&L)[NC], double& loss, int& bad,
                                           double (&sums)[NR][f2_pad(f2_nsums(NC, FL))]) {
    double x[NR], y[NR], iv[NR];
    StripRow<NC> rc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        x[i] = px[i * kStripRec]; y[i] = px[i * kStripRec + 32]; iv[i] = px[i * kStripRec + 64];
        rc[i] = rowc[i];
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) strip_pixel<NC, PV, FL>(x[i], y[i], iv[i], rc[i], xp0, inv_dx, L, loss, bad, sums[i]);
}

// Shared memory of one warp (bytes), K rows per block.
//   ring   kStripStages x (kStageRows x 768 B): bulk-copy ring
//   tile   [NS][kStageRows][32] row terms of a stage, parked by the lanes for the transposed reduction
//   cwin   [NC][kStripCw] control points of the task's knot window
//   mtab   [NC][kStripWs][PV + 2] moments per base interval
//   hdr    [2] x { StripRow[K] | xoff[32] | StripBlk }   row records, column offsets, key range of a block
//   bars   kStripStages + 2 mbarriers
constexpr int kStripCw = 40;      // >= kStripWs + 3 + 1 knots
template <int NC, int PV, int FL> struct StripLay {
    static constexpr int NS = f2_nsums(NC, FL), NSP = f2_pad(NS);
    static constexpr int TE = PV + 2;                                     // doubles per table entry: M[PV+1], Md
    static constexpr uint32_t slot_bytes = kStageRows * kStripRecBytes;
    static constexpr uint32_t ring = 0;
    static constexpr uint32_t ring_bytes = kStripStages * slot_bytes;
    static constexpr uint32_t tile = ring + ring_bytes;                  // 256-byte aligned (XOR swizzle on addresses)
    static constexpr uint32_t tile_bytes = NS * kStageRows * 256;
    static constexpr uint32_t cwin = tile + tile_bytes;
    static constexpr uint32_t cwin_bytes = NC * kStripCw * 8;
    static constexpr uint32_t mtab = cwin + cwin_bytes;
    static constexpr uint32_t mtab_bytes = NC * kStripWs * TE * 8;
    static constexpr uint32_t hdr = mtab + mtab_bytes;
    __host__ __device__ static constexpr uint32_t hdr_slot(int K) { return (uint32_t)(K * sizeof(StripRow<NC>) + 32 * 8 + sizeof(StripBlk)); }
    __host__ __device__ static constexpr uint32_t bars(int K) { return hdr + 2 * hdr_slot(K); }
    __host__ __device__ static constexpr uint32_t total(int K) { return (bars(K) + (kStripStages + 2) * 8 + 255) / 256 * 256; }
    __host__ __device__ static constexpr uint32_t launch_bytes(int K) { return kStripWarps * total(K); }
};

__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void lds_v2f64(uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef JB_STRIP_MINB
#define JB_STRIP_MINB 4       // blocks per SM the register budget is set for: 4 x 4 warps x 128 registers
#endif
#ifndef JB_STRIP_ILP
#define JB_STRIP_ILP 4
#endif
constexpr int kStripIlp = JB_STRIP_ILP;       // rows of a stage computed as independent streams (divides kStageRows)

template <int NC, int PV, int FL>
__global__ void __launch_bounds__(kStripWarps * 32, JB_STRIP_MINB) k_strip(const StripParams* __restrict__ arr, const int* __restrict__ off, int n, int kmax) {
    // No static shared memory in this kernel: the dynamic block then starts at a 1 KB boundary, which
    // the XOR-addressed row-term tile relies on (256 bytes; checked below).  The plan's parameters
    // are read from global memory (constant for the launch, L1-resident).
    extern __shared__ __align__(256) unsigned char strip_smem[];
    using Lay = StripLay<NC, PV, FL>;
    constexpr int NS = Lay::NS, NSP = Lay::NSP, TE = Lay::TE;
    static_assert(kStageRows == 4 && kStripIlp >= 1 && kStageRows % kStripIlp == 0 && kStripStages == 2, "stage shape");
    const int pi = find_plan(off, n, blockIdx.x);
    const StripParams& a = arr[pi];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)(blockIdx.x - off[pi]) * kStripWarps + warp;
    if (task >= a.n_tasks) return;
    const int strip = (int)(task % a.n_strips);
    const int rgroup = (int)(task / a.n_strips);
    const int K = a.K;
    const int spb = K >> 2;                        // stages per block (kStageRows = 4)
    const int blk0 = rgroup * a.blocks_per_task;
    const int nblk = min(a.blocks_per_task, a.n_blocks - blk0);
    const int64_t row0 = (int64_t)blk0 * K;
    const int nrow = (int)min((int64_t)nblk * K, a.n_rows - row0);
    const int nstage = (nrow + kStageRows - 1) / kStageRows;

    // (kmax: the longest block among the plans of the launch -- they share one shared-memory layout)
    unsigned char* wb = strip_smem + (size_t)warp * Lay::total(kmax);
    const uint32_t wb32 = smem_u32(wb);
    if ((wb32 & 255u) != 0) { if (lane == 0) atomicOr(a.status, kStatBounds); return; }     // never: see above
    const uint32_t bar0 = wb32 + Lay::bars(kmax);               // stage barriers, then the two header barriers
    const uint32_t hbar0 = bar0 + 8 * kStripStages;
    const uint32_t hslot = Lay::hdr_slot(kmax);
    double* cwin = reinterpret_cast<double*>(wb + Lay::cwin);
    double* mtab = reinterpret_cast<double*>(wb + Lay::mtab);

    const double* strip_data = a.data + ((int64_t)strip * a.n_rows + row0) * kStripRec;
    auto issue_stage = [&](int s) {       // lane 0
        const int rows = min(kStageRows, nrow - s * kStageRows);
        const uint32_t bar = bar0 + 8 * (s % kStripStages);
        mbar_expect_tx(bar, rows * kStripRecBytes);
        tma_load_1d(wb32 + Lay::ring + (s % kStripStages) * Lay::slot_bytes,
                    strip_data + (int64_t)s * kStageRows * kStripRec, rows * kStripRecBytes, bar);
    };
    auto issue_hdr = [&](int b) {         // lane 0: row records, column offsets and key range of block b of the task
        const int rows = min(K, nrow - b * K);
        const uint32_t bar = hbar0 + 8 * (b & 1);
        const uint32_t dst = wb32 + Lay::hdr + (b & 1) * hslot;
        mbar_expect_tx(bar, rows * (uint32_t)sizeof(StripRow<NC>) + 32 * 8 + (uint32_t)sizeof(StripBlk));
        tma_load_1d(dst, reinterpret_cast<const StripRow<NC>*>(a.rowc) + row0 + (int64_t)b * K, rows * (uint32_t)sizeof(StripRow<NC>), bar);
        tma_load_1d(dst + K * (uint32_t)sizeof(StripRow<NC>), a.xoff + ((int64_t)strip * a.n_blocks + blk0 + b) * 32, 32 * 8, bar);
        tma_load_1d(dst + K * (uint32_t)sizeof(StripRow<NC>) + 32 * 8, a.blk + blk0 + b, (uint32_t)sizeof(StripBlk), bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStripStages + 2; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
        issue_hdr(0);
        for (int s = 0; s < kStripStages && s < nstage; ++s) issue_stage(s);
        if (nblk > 1) issue_hdr(1);
    }

    double xp0[NC], inv_dx[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { xp0[c] = a.comp[c].xp0; inv_dx[c] = a.comp[c].inv_dx; }
    const double tol = a.tol;
    constexpr double kHalf = (PV & 1) ? 0.5 : 0.0;       // piece I covers u = t - kHalf in [I - 1/2, I + 1/2)
    constexpr double kKnot = (PV & 1) ? 1.0 : 0.5;       // t of the knot between piece I and piece I + 1, minus I
    // base interval of a column over a block: the piece that holds its smallest coordinate minus the
    // tolerance.  The block's key range is less than 1 - 2 tol knots wide (k_prepare_strip raises
    // kStatWindow otherwise), so every row of the block then lies in [base, base + 2).
    auto predict = [&](double xo, double kmin, int c, int& base) {
        const double umin = ((xo + kmin) - xp0[c]) * inv_dx[c] - kHalf;
        const double fl = floor(umin - tol + 0.5);
        base = fabs(fl) < 1.0e9 ? (int)fl : INT_MAX / 2;
    };
    // does the column reach into piece base + 1 during the block?
    auto needs_next = [&](double xo, double kmax, int c, int base) {
        const double umax = ((xo + kmax) - xp0[c]) * inv_dx[c] - kHalf;
        return umax + tol >= (double)base + 0.5;
    };

    // ---- 1. window of the task: bases of its smallest and largest column offset over the key range
    // of its blocks (lane b takes block b) ----------------------------------------------------------
    int lo[NC];
    bool err_window = false, empty;
    {
        const double2 xr = reinterpret_cast<const double2*>(a.xoff_rng)[(int64_t)rgroup * a.n_strips + strip];   // {min, max}; min > max: nothing weighted
        double kmin[NC], kmax[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { kmin[c] = INFINITY; kmax[c] = -INFINITY; }
        for (int b = lane; b < nblk; b += 32) {
            const StripBlk bk = a.blk[blk0 + b];
            err_window |= bk.wide != 0;
#pragma unroll
            for (int c = 0; c < NC; ++c) { kmin[c] = fmin(kmin[c], bk.kmin[c]); kmax[c] = fmax(kmax[c], bk.kmax[c]); }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                kmin[c] = fmin(kmin[c], __shfl_xor_sync(0xffffffffu, kmin[c], o));
                kmax[c] = fmax(kmax[c], __shfl_xor_sync(0xffffffffu, kmax[c], o));
            }
        empty = !(xr.x <= xr.y) || !(kmin[0] <= kmax[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int b_lo = 0, b_hi = 0;
            predict(xr.x, kmin[c], c, b_lo);
            predict(xr.y, kmax[c], c, b_hi);
            lo[c] = b_lo;
            if (!empty && (b_lo == INT_MAX / 2 || b_hi == INT_MAX / 2 || b_hi - b_lo >= kStripWs)) err_window = true;
        }
    }
    err_window = __any_sync(0xffffffffu, err_window) && !empty;
    constexpr int KOFF = (PV + 1) / 2 - PV;       // klo = I + KOFF (knot_lo)
    if (lane == 0) {
        if (err_window) atomicOr(a.status, kStatWindow);
#pragma unroll
        for (int c = 0; c < NC; ++c) a.part_base[(int64_t)c * a.n_tasks + task] = (empty || err_window) ? INT_MAX : lo[c] + KOFF;
        if (empty || err_window) a.losspart[task] = 0.0;
    }
    if (empty || err_window) {
        // nothing weighted (or an error): the row sums of this strip are zero
        if (NS > 0)
            for (int i = lane; i < nrow * NS; i += 32)
                a.rowpart[((row0 + i / (NS > 0 ? NS : 1)) * a.n_strips + strip) * NS + i % (NS > 0 ? NS : 1)] = 0.0;
        // copies are in flight into this warp's shared memory: let them land before leaving
        mbar_wait(hbar0, 0);
        if (nblk > 1) mbar_wait(hbar0 + 8, 0);
        for (int s = 0; s < kStripStages && s < nstage; ++s) mbar_wait(bar0 + 8 * s, 0);
        return;
    }

    // ---- 2. control points of the window; zeroed moment table ---------------------------------------
    unsigned okP[NC], okD[NC];       // bit j: piece lo+j inside the knot grid / piece lo+j+1 as well
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int M = a.comp[c].n_knots;
        for (int k = lane; k < kStripCw; k += 32) {
            const int kn = lo[c] + KOFF + k;
            cwin[c * kStripCw + k] = (kn >= 0 && kn < M) ? __ldg(a.comp[c].ctrl + kn) : 0.0;
        }
        const int klo = lo[c] + lane + KOFF;
        okP[c] = __ballot_sync(0xffffffffu, klo >= 0 && klo + PV < M);
        okD[c] = __ballot_sync(0xffffffffu, klo >= 0 && klo + PV + 1 < M);
    }
    for (int i = lane; i < NC * kStripWs * TE; i += 32) mtab[i] = 0.0;
    __syncwarp();

    StripLane<PV> L[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int k = 0; k <= PV; ++k) { L[c].P[k] = 0.0; L[c].M[k] = 0.0; }
        L[c].d = 0.0; L[c].dd = 0.0; L[c].Md = 0.0; L[c].ngofs = 0.0; L[c].base = kStripNoBase;
    }
    const int ncls = a.ncls;
    // moments of the lanes whose base changes -> table.  Lanes ncls apart never share an entry
    // (precondition 3), so the lanes take turns by lane % ncls.
    auto flush = [&](const bool (&mv)[NC]) {
        for (int ph = 0; ph < ncls; ++ph) {
            if (lane % ncls == ph) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (mv[c] && L[c].base != kStripNoBase) {
                        double* e = mtab + ((size_t)c * kStripWs + (L[c].base - lo[c])) * TE;
#pragma unroll
                        for (int k = 0; k <= PV; ++k) e[k] += L[c].M[k];
                        e[PV + 1] += L[c].Md;
                    }
            }
            __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (mv[c]) {
#pragma unroll
                for (int k = 0; k <= PV; ++k) L[c].M[k] = 0.0;
                L[c].Md = 0.0;
            }
    };

    double loss = 0.0;
    int bad = 0, err_range = 0;
    const unsigned char* hd = wb + Lay::hdr;

    // ---- 3. stages of kStageRows rows; a checkpoint at the start of every block of K rows --------------
    // Row terms: element `lane` of vector (sum q, row i) is parked at [lane ^ 4i] of the vector's 32
    // doubles (its 16-byte chunks XOR-swizzled by 2i); lane (v = lane >> 1, half = lane & 1) then adds
    // the chunks of parity `half` of vector v (2 (j ^ i) + half, j = 0..7) and the two halves meet in
    // one shuffle: conflict-free both ways, one fixed order, one XOR per load.
    const int rsel = (lane >> 1) & (kStageRows - 1);             // row of the stage whose sums this lane adds up
    const int vsel = lane >> 1;                                    // vector (q * kStageRows + row) of a pass
    const uint32_t tile32 = wb32 + Lay::tile;
    const uint32_t rd_addr = tile32 + (uint32_t)(vsel * 256 + (lane & 1) * 16 + rsel * 32);
    const uint32_t park_addr = tile32 + (uint32_t)(lane * 8);
    // where this lane stores the sum it adds up (v0 = 0 round; later rounds 16 / kStageRows sums further), per stage
    double* rowpart_p = a.rowpart + ((row0 + rsel) * a.n_strips + strip) * NS + vsel / kStageRows;
    const int64_t rowpart_step = (int64_t)kStageRows * a.n_strips * NS;
    // scale of that sum: the reducers expect S' / PV (jb_fused2.cuh); a stretched component's derivative
    // sums carry the row's stretch (only then does the scale depend on the row)
    constexpr bool kRowScale = (f2_der(FL, 0) && f2_str(FL, 0)) || (NC > 1 && f2_der(FL, 1) && f2_str(FL, 1));
    auto sum_scale = [&](int q, const StripRow<NC>* rc) {
        constexpr double ipv = 1.0 / PV;
        double sc = 1.0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (f2_der(FL, c)) {
                const double st = f2_str(FL, c) ? rc->st[c] : 1.0;
                if (q == f2_slot(NC, FL, c, 0)) sc = st * ipv;
                if (q == f2_slot(NC, FL, c, 2)) sc = (st * ipv) * (st * ipv);
            }
        return sc;
    };
    StripRow<NC> unit_row;
#pragma unroll
    for (int c = 0; c < NC; ++c) { unit_row.sh[c] = 0.0; unit_row.st[c] = 1.0; }
    double sc_lane[NS > 0 ? (NS * kStageRows + 15) / 16 : 1];
#pragma unroll
    for (int v0 = 0; v0 < NS * kStageRows; v0 += 16) sc_lane[v0 / 16] = sum_scale((v0 + vsel) / kStageRows, &unit_row);
    uint32_t ring_off = 0, bar_off = 0;              // slot of the current stage: toggled, not computed
    const double* src_next = strip_data + (int64_t)kStripStages * kStageRows * kStripRec;
    const int rows_last = nrow - (nstage - 1) * kStageRows;
    const int nfull = rows_last == kStageRows ? nstage : nstage - 1;      // only the last stage of a task can be ragged
    int sb = 0, b = 0;

    auto checkpoint = [&]() {
        hd = wb + Lay::hdr + (b & 1) * hslot;
        mbar_wait(hbar0 + 8 * (b & 1), (uint32_t)((b >> 1) & 1));
        const double xo = reinterpret_cast<const double*>(hd + K * sizeof(StripRow<NC>))[lane];
        const StripBlk bk = *reinterpret_cast<const StripBlk*>(hd + K * sizeof(StripRow<NC>) + 32 * 8);
        const bool live = xo > -1.0e300 && bk.kmin[0] <= bk.kmax[0];
        int nb[NC];
        bool mv[NC], any = false;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int base = kStripNoBase;
            if (live) {
                predict(xo, bk.kmin[c], c, base);
                const unsigned j = (unsigned)(base - lo[c]);
                if (j >= (unsigned)kStripWs) { bad = 1; base = kStripNoBase; }
                else if (!((okD[c] >> j) & 1u)) {      // at the end of the knot grid: is piece base + 1 needed?
                    if (!((okP[c] >> j) & 1u) || needs_next(xo, bk.kmax[c], c, base)) { err_range = 1; base = kStripNoBase; }
                }
            }
            nb[c] = base;
            mv[c] = base != L[c].base;
            any |= mv[c];
        }
        if (__any_sync(0xffffffffu, any)) {
            flush(mv);
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (mv[c]) {
                    if (nb[c] != kStripNoBase) {
                        const double* cp = cwin + c * kStripCw + (nb[c] - lo[c]);
                        double cc[PV + 2];
#pragma unroll
                        for (int k = 0; k < PV + 2; ++k) cc[k] = cp[k];
                        StripBasis<PV>::coeffs(cc, L[c].P, L[c].d);
                        L[c].dd = (double)PV * L[c].d;
                        L[c].ngofs = -((double)nb[c] + kKnot);
                    } else {
#pragma unroll
                        for (int k = 0; k <= PV; ++k) L[c].P[k] = 0.0;
                        L[c].d = 0.0; L[c].dd = 0.0; L[c].ngofs = 0.0;
                    }
                    L[c].base = nb[c];
                }
        }
    };

    // one stage; FULL: all kStageRows rows present
    auto stage_body = [&](int s, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int rows = FULL ? kStageRows : rows_last;
        mbar_wait(bar0 + bar_off, (uint32_t)((s / kStripStages) & 1));
        const double* px = reinterpret_cast<const double*>(wb + Lay::ring + ring_off) + lane;
        const StripRow<NC>* rowc = reinterpret_cast<const StripRow<NC>*>(hd) + sb * kStageRows;
        constexpr int G = FULL ? kStripIlp : 1;          // rows computed together
#pragma unroll(G == kStageRows ? kStageRows : 1)
        for (int i0 = 0; i0 < kStageRows; i0 += G) {
            if (FULL || i0 < rows) {
                double sums[G][NSP];
                strip_rows<NC, PV, FL, G>(px + i0 * kStripRec, rowc + i0, xp0, inv_dx, L, loss, bad, sums);
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const uint32_t pa = (park_addr ^ (uint32_t)(32 * (i0 + i))) + (uint32_t)((i0 + i) * 256);
#pragma unroll
                    for (int q = 0; q < NS; ++q) sts_f64(pa + q * kStageRows * 256, sums[i][q]);
                }
            }
        }
        __syncwarp();                                  // rows consumed, row terms parked
        if (lane == 0 && s + kStripStages < nstage) {
            const int rows_n = (s + kStripStages + 1 < nstage) ? kStageRows : rows_last;
            mbar_expect_tx(bar0 + bar_off, rows_n * kStripRecBytes);
            tma_load_1d(wb32 + Lay::ring + ring_off, src_next, rows_n * kStripRecBytes, bar0 + bar_off);
        }
#pragma unroll
        for (int v0 = 0; v0 < NS * kStageRows; v0 += 16) {          // 16 vectors per round (two lanes each)
            const int nv = NS * kStageRows - v0 < 16 ? NS * kStageRows - v0 : 16;
            double tot = 0.0;
            if (vsel < nv) {
                double tx, ty;
                lds_v2f64(rd_addr + v0 * 256, tx, ty);
#pragma unroll
                for (int j = 1; j < 8; ++j) {
                    double wx, wy;
                    lds_v2f64((rd_addr + v0 * 256) ^ (uint32_t)(32 * j), wx, wy);
                    tx += wx; ty += wy;
                }
                tot = tx + ty;
            }
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            if ((lane & 1) == 0 && vsel < nv && (FULL || rsel < rows)) {
                const double sc = kRowScale ? sum_scale((v0 + vsel) / kStageRows, rowc + rsel) : sc_lane[v0 / 16];
                rowpart_p[v0 / kStageRows] = tot * sc;
            }
        }
        if (NS > 0) __syncwarp();                          // the tile is free for the next stage
    };

    for (int s = 0; s < nfull; ++s) {
        if (sb == 0) checkpoint();
        stage_body(s, std::true_type{});
        rowpart_p += rowpart_step;
        src_next += kStageRows * kStripRec;
        ring_off ^= Lay::slot_bytes; bar_off ^= 8u;
        if (++sb == spb) {
            // last stage of a block: its header slot is free, fetch the block after next
            if (lane == 0 && b + 2 < nblk) issue_hdr(b + 2);
            sb = 0; ++b;
        }
    }
    if (nfull < nstage) {
        if (sb == 0) checkpoint();
        stage_body(nfull, std::false_type{});
    }

    // ---- 4. moments -> per-knot sums, published as the task's gradient window ----------------------
    {
        bool all[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) all[c] = true;
        flush(all);
    }
    double* scr = reinterpret_cast<double*>(wb + Lay::ring);      // [NC][kStripWs][TE] knot terms per entry (the ring is idle now)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double* e = mtab + ((size_t)c * kStripWs + lane) * TE;
        double M[PV + 1], kn[PV + 2];
#pragma unroll
        for (int i = 0; i <= PV; ++i) M[i] = e[i];
        StripBasis<PV>::knots(M, e[PV + 1], kn);
#pragma unroll
        for (int t = 0; t < PV + 2; ++t) scr[((size_t)c * kStripWs + lane) * TE + t] = kn[t];
    }
    __syncwarp();
    const int W = a.wincap;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double* dst = a.part + ((int64_t)c * a.n_tasks + task) * W;
        for (int k = lane; k < W; k += 32) {
            // knot lo + KOFF + k takes position t = k - j of the entries j = k - PV - 1 .. k
            double sk = 0.0;
#pragma unroll
            for (int t = PV + 1; t >= 0; --t) {
                const int j = k - t;
                if (j >= 0 && j < kStripWs) sk += scr[((size_t)c * kStripWs + j) * TE + t];
            }
            dst[k] = sk;
        }
    }
    loss = warp_sum(loss);
    bad = __any_sync(0xffffffffu, bad != 0);
    err_range = __any_sync(0xffffffffu, err_range != 0);
    if (lane == 0) {
        a.losspart[task] = loss;
        if (bad) atomicOr(a.status, kStatStrip);
        if (err_range) atomicOr(a.status, kStatOutOfRange);
    }
}

// ------------------------------------------------------------------------------------------
// Building the strip layout (on the device, from the plan's tile-interleaved canonical block).

struct StripLayoutParams {
    const double* tiled;     // [E][n_tiles][3][128] lane-interleaved (jb_kernels.cuh)
    const int* rowid;        // [E] caller's row of sorted position
    double* sdata;           // [n_strips][E][3][32]
    int64_t n_rows;
    int n_tiles;
};

// One warp per (sorted row, 128-pixel tile): read the 3 KB record as it lies, write its four strips.
__global__ void k_strip_layout(StripLayoutParams a) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= a.n_rows * a.n_tiles) return;
    const int64_t pos = w / a.n_tiles;
    const int tile = (int)(w - pos * a.n_tiles);
    const double* src = a.tiled + ((int64_t)a.rowid[pos] * a.n_tiles + tile) * kTileDoubles;
    const int n_strips = a.n_tiles * (kTile / kStripPx);
#pragma unroll
    for (int e0 = 0; e0 < kTileDoubles; e0 += 32) {
        const int e = e0 + lane;
        const int arr = e / kTile, q = e % kTile;                  // array (x, y, ivar), interleaved position
        const int p = (q % 32) * kPixPerLane + q / 32;             // pixel of the tile
        const int strip = tile * (kTile / kStripPx) + p / kStripPx;
        a.sdata[((int64_t)strip * a.n_rows + pos) * kStripRec + arr * kStripPx + p % kStripPx] = src[e];
    }
    (void)n_strips;
}

struct StripXoffParams {
    const double* sdata; const double* xref;      // xref [E] sorted order: first weighted x of the row (+inf: none)
    double* xoff;                                  // [n_strips][n_blocks][32]
    int64_t n_rows; int n_strips, K, n_blocks;
    double tol_x;                                  // kStripTol knots of the finest grid, in x
    unsigned long long* stats;                     // [0] rows that break precondition (1)  [1] widest strip (x, as bits)  [2] largest deviation (x, as bits)
                                                   // [2 + n] smallest x[p + n] - x[p], n = 1..8
};

// One warp per (strip, block): the column offsets of the block and the checks of preconditions (1), (3).
__global__ void k_strip_xoff(StripXoffParams a) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)a.n_strips * a.n_blocks) return;
    const int strip = (int)(w / a.n_blocks), b = (int)(w - (int64_t)strip * a.n_blocks);
    const int64_t r0 = (int64_t)b * a.K;
    const int rows = (int)min((int64_t)a.K, a.n_rows - r0);
    const double* rec = a.sdata + ((int64_t)strip * a.n_rows + r0) * kStripRec + lane;
    double xo = -INFINITY, dev = 0.0;
    bool have = false;
    for (int i = 0; i < rows; ++i) {
        const double x = rec[(int64_t)i * kStripRec], iv = rec[(int64_t)i * kStripRec + 2 * kStripPx];
        if (iv != 0.0) {
            const double o = x - a.xref[r0 + i];
            if (!have) { xo = o; have = true; }
            else dev = fmax(dev, fabs(o - xo));
        }
    }
    a.xoff[((int64_t)strip * a.n_blocks + b) * 32 + lane] = xo;
    const bool bad = dev > a.tol_x || (have && !(fabs(xo) < 1.0e300));
    const unsigned nbad = __popc(__ballot_sync(0xffffffffu, bad));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    double xmin = have ? xo : INFINITY, xmax = have ? xo : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    }
    if (lane == 0) {
        if (nbad) atomicAdd(a.stats, (unsigned long long)nbad);
        if (xmin <= xmax) atomicMax(a.stats + 1, (unsigned long long)__double_as_longlong(xmax - xmin));   // >= 0: bits order like values
        if (dev > 0.0 && dev < 1.0e300) atomicMax(a.stats + 2, (unsigned long long)__double_as_longlong(dev));
    }
    // smallest distance between weighted columns n lanes apart (negative: not monotone -> 0)
#pragma unroll
    for (int nn = 1; nn <= 8; ++nn) {
        const double other = __shfl_down_sync(0xffffffffu, xo, nn);
        const bool oh = __shfl_down_sync(0xffffffffu, have ? 1 : 0, nn) != 0;
        double gap = (have && oh && lane + nn < 32) ? fmax(other - xo, 0.0) : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gap = fmin(gap, __shfl_xor_sync(0xffffffffu, gap, o));
        if (lane == 0 && gap < 1.0e300) atomicMin(a.stats + 2 + nn, (unsigned long long)__double_as_longlong(gap));
    }
}

struct StripRngParams { const double* xoff; double* rng; int n_strips, n_blocks, blocks_per_task, n_rgroups; };

// One warp per task (row group, strip): range of the finite column offsets of its blocks.
__global__ void k_strip_rng(StripRngParams a) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)a.n_rgroups * a.n_strips) return;
    const int g = (int)(w / a.n_strips), strip = (int)(w - (int64_t)g * a.n_strips);
    const int b0 = g * a.blocks_per_task, b1 = min(b0 + a.blocks_per_task, a.n_blocks);
    double xmin = INFINITY, xmax = -INFINITY;
    for (int b = b0; b < b1; ++b) {
        const double xo = a.xoff[((int64_t)strip * a.n_blocks + b) * 32 + lane];
        if (xo > -1.0e300) { xmin = fmin(xmin, xo); xmax = fmax(xmax, xo); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    }
    if (lane == 0) { a.rng[2 * w] = xmin; a.rng[2 * w + 1] = xmax; }
}

}  // namespace jb
