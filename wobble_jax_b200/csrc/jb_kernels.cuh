// Kernels of libjabble_b200 (sm_100a).  See DESIGN.md for the decomposition.
//
//   k_scatter_params   p_fit -> params_all[fit_map]
//   k_prepare_rows     per-row shift/stretch records in sorted-row order
//   k_forward          model(p, x, meta) per pixel                       (Model.__call__)
//   k_fused<NC,PV,PVN> warp->knot, basis weights, gather, residual, chi-square, adjoints of the
//                      control points (deterministic, no float atomics), of shift/stretch, and the
//                      per-row Fisher sums -- ONE pass over x, y, ivar, mask.  The general kernel:
//                      any pixel spacing, NormalizationModel component.  Smooth plans run on
//                      k_fused2 (jb_fused2.cuh), their float32 tier on k_fused32 (jb_fused32.cuh)
//   k_grid_search<PV>  chi-square of every row at M candidate shifts    (EpochSpecificModel.grid_search)
//   k_row_curvature<PV> d^2 chi-square / d shift^2 per row              (quickplay.get_RV_error_2d)
//   k_reduce_stage1/2  fixed-order reduction of the per-task partial windows and per-row sums
//   k_finish           loss, gather of the fit-mode gradient
#pragma once
#include "jb_device.cuh"

namespace jb {

// ------------------------------------------------------------------------------------------
struct ScatterParams {
    const double* p_fit;
    const int64_t* fit_map;
    double* params;
    int64_t n_fit;
};

// Every evaluation kernel takes an ARRAY of per-plan parameter blocks plus the first block index of
// each plan (n+1 entries): one launch serves one plan (n = 1) or a whole batch of independent
// chunks (jb_batch), so that small per-chunk grids do not leave the GPU half empty.
__device__ __forceinline__ int find_plan(const int* __restrict__ off, int n, int blk) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= blk) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void k_scatter_params(const ScatterParams* __restrict__ arr, const int* __restrict__ off, int n,
                                 const double* p_fit_override) {
    const int pi = find_plan(off, n, blockIdx.x);
    const ScatterParams a = arr[pi];
    const double* p_fit = p_fit_override ? p_fit_override : a.p_fit;
    const int64_t i = (int64_t)(blockIdx.x - off[pi]) * blockDim.x + threadIdx.x;
    if (p_fit && i < a.n_fit) a.params[a.fit_map[i]] = p_fit[i];
}

// ------------------------------------------------------------------------------------------
// Per-row parameters in sorted-row order, refreshed every evaluation (the shifts are fit parameters):
// grouprec[pos] = { shift_c, stretch_c (c < NC), row id } for the row at sorted position pos.
// HN = 1 when the plan has a per-key (NormalizationModel) spline component: its shift, stretch and
// key ride in the same record.
template <int NC, int HN> struct alignas(16) RowRec {
    double sh[NC + HN]; double st[NC + HN]; int row; int dense; int nkey; int pad_;
};
constexpr int kNormMaxKnots = 32;   // knots per key of a normalisation spline handled on chip

struct PrepareParams {
    const int* row_perm;
    void* grouprec;
    int64_t n_rows;
    int n_comp;
    CompDev comp[kMaxComp];
    int* status;      // cleared here, at the start of every evaluation: what it holds afterwards is this evaluation's
};

template <int NC, int HN>
__global__ void k_prepare_rows(const PrepareParams* __restrict__ arr, const int* __restrict__ off, int n) {
    const int pi = find_plan(off, n, blockIdx.x);
    const PrepareParams& a = arr[pi];
    const int64_t pos = (int64_t)(blockIdx.x - off[pi]) * blockDim.x + threadIdx.x;
    if (pos == 0 && a.status) { a.status[0] = 0; a.status[1] = 0; }
    if (pos >= a.n_rows) return;
    const int r = a.row_perm[pos];
    RowRec<NC, HN> rec;
#pragma unroll
    for (int c = 0; c < NC + HN; ++c) {
        const CompDev& cd = a.comp[c];
        rec.sh[c] = cd.shift ? cd.shift[cd.shift_key[r]] : 0.0;
        rec.st[c] = cd.stretch ? cd.stretch[cd.stretch_key[r]] : 1.0;
    }
    rec.row = r; rec.dense = 0; rec.pad_ = 0;
    rec.nkey = HN ? a.comp[NC].ctrl_key[r] : 0;
    reinterpret_cast<RowRec<NC, HN>*>(a.grouprec)[pos] = rec;
}

// ------------------------------------------------------------------------------------------
// Data layout in HBM ("tile-interleaved"): for every (row r, tile t) one contiguous record of
//   x[kTile] | y[kTile] | ivar[kTile]      (3 * kTile doubles = 3072 B)
// at offset (r * n_tiles + t) * kTileDoubles; inside each array pixel p of the tile is stored at
// [(p % 4) * 32 + p / 4], i.e. lane-interleaved: lane l of the warp that owns the tile walks pixels
// 4l..4l+3 and finds them at l, 32 + l, 64 + l, 96 + l -- one bank per lane.  The mask is folded into the data at upload: a masked
// or padded pixel gets ivar = 0, y = 0 and the x of the nearest weighted pixel before it in the row
// (so it is evaluated on knots the lane already holds and contributes exactly 0 to every sum),
// which is what where(~mask, ., 0) of ChiSquare does to value and cotangent.  The forward kernel
// reads the original x from a separate row-major copy.
constexpr int kTileDoubles = 3 * kTile;
constexpr int kTileBytes = kTileDoubles * 8;

struct ForwardParams {
    const double* x;       // [E][ppad] canonical orientation, original x of every pixel
    const uint8_t* row_rev;
    double* out;           // [E][n_pix], caller orientation
    int64_t n_rows, n_pix, ppad;
    int n_comp;
    CompDev comp[kMaxComp];
};

template <int PV>
__device__ __forceinline__ double spline_value(const CompDev& c, const double* ctrl, double x, double shift,
                                               bool& in_range) {
    double g; int I;
    knot_coord<PV>(x, shift, c.xp0, c.inv_dx, g, I);
    // recompute the range test on the un-wrapped coordinate: the magic-number trick is only valid
    // for |u| < 2^31
    const double t = ((x - shift) - c.xp0) * c.inv_dx;
    const int klo = knot_lo<PV>(I);
    if (!(t > -1.0e9 && t < 1.0e9) || klo < 0 || klo + PV > c.n_knots - 1) { in_range = false; return 0.0; }
    double w[PV + 1], dw[PV > 0 ? PV : 1];
    Basis<PV>::eval(g, w, dw);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j <= PV; ++j) s = __fma_rn(__ldg(ctrl + klo + j), w[j], s);
    return s;
}

__global__ void k_forward(ForwardParams a) {
    const int64_t total = a.n_rows * a.n_pix;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / a.n_pix;
        const int64_t p = idx - r * a.n_pix;
        const int64_t pc = a.row_rev[r] ? (a.n_pix - 1 - p) : p;   // canonical (ascending) position
        const double x = a.x[r * a.ppad + pc];
        double m = 0.0;
        bool ok = true;
        for (int ci = 0; ci < a.n_comp; ++ci) {
            const CompDev& c = a.comp[ci];
            const double shift = c.shift ? c.shift[c.shift_key[r]] : 0.0;
            const double* ctrl = c.kind == 1 ? c.ctrl + (int64_t)c.ctrl_key[r] * c.n_knots : c.ctrl;
            double s;
            if (c.p_val == 3) s = spline_value<3>(c, ctrl, x, shift, ok);
            else if (c.p_val == 2) s = spline_value<2>(c, ctrl, x, shift, ok);
            else s = spline_value<1>(c, ctrl, x, shift, ok);
            if (c.stretch) s *= c.stretch[c.stretch_key[r]];
            m += s;
        }
        a.out[idx] = ok ? m : __longlong_as_double(0x7ff8000000000000LL);
    }
}

// ------------------------------------------------------------------------------------------
// EpochSpecificModel.grid_search (jabble/model.py:738-786) for the ShiftingModel of one additive
// component: the chi-square of every row at M candidate values of that row's shift, everything
// else as stored.  Only the searched component depends on the candidate, so a block first forms
// y - (all other components) for a tile of pixels, once, and then every thread walks the tile for
// its own candidate: out[row][j] = sum_pix ivar * (y - m_j)^2.
struct GridParams {
    const double* data;            // tile-interleaved x | y | ivar (mask folded: ivar = 0)
    int64_t n_rows;
    int n_tiles;
    int n_comp;
    CompDev comp[kMaxComp];
    int search;                    // stored index of the component whose shift is searched
    const double* grid;            // [n_keys][M] candidate shifts
    int M;
    double* out;                   // [n_rows][M]
    int* status;
};

template <int PV>
__global__ void k_grid_search(GridParams a) {
    __shared__ double sx[kTile], sr[kTile], sw[kTile];
    const int64_t r = blockIdx.x;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;      // candidate of this thread
    const CompDev& cs = a.comp[a.search];
    const double st = cs.stretch ? cs.stretch[cs.stretch_key[r]] : 1.0;
    const double sh = j < a.M ? a.grid[(int64_t)cs.shift_key[r] * a.M + j] : 0.0;
    const double* ctrl = cs.ctrl;
    double acc = 0.0;
    bool ok = true;
    for (int t = 0; t < a.n_tiles; ++t) {
        const double* rec = a.data + (r * a.n_tiles + t) * kTileDoubles;
        const int p = threadIdx.x;                             // blockDim.x == kTile: one pixel each
        const int pos = (p % kPixPerLane) * 32 + p / kPixPerLane;
        const double x = rec[pos], y = rec[kTile + pos], w = rec[2 * kTile + pos];
        double mo = 0.0;
        if (w != 0.0) {
            for (int ci = 0; ci < a.n_comp; ++ci) {
                if (ci == a.search) continue;
                const CompDev& c = a.comp[ci];
                const double shift = c.shift ? c.shift[c.shift_key[r]] : 0.0;
                const double* cc = c.kind == 1 ? c.ctrl + (int64_t)c.ctrl_key[r] * c.n_knots : c.ctrl;
                double s;
                if (c.p_val == 3) s = spline_value<3>(c, cc, x, shift, ok);
                else if (c.p_val == 2) s = spline_value<2>(c, cc, x, shift, ok);
                else s = spline_value<1>(c, cc, x, shift, ok);
                if (c.stretch) s *= c.stretch[c.stretch_key[r]];
                mo += s;
            }
        }
        __syncthreads();       // the previous tile has been consumed
        sx[p] = x; sr[p] = y - mo; sw[p] = w;
        __syncthreads();
        if (j < a.M) {
            for (int q = 0; q < kTile; ++q) {
                const double wq = sw[q];
                if (wq == 0.0) continue;                       // same for every thread of the block
                double g; int I;
                knot_coord<PV>(sx[q], sh, cs.xp0, cs.inv_dx, g, I);
                const double tt = ((sx[q] - sh) - cs.xp0) * cs.inv_dx;
                const int klo = knot_lo<PV>(I);
                if (!(tt > -1.0e9 && tt < 1.0e9) || klo < 0 || klo + PV > cs.n_knots - 1) { ok = false; continue; }
                double wgt[PV + 1], dw[PV > 0 ? PV : 1];
                Basis<PV>::eval(g, wgt, dw);
                double sv = 0.0;
#pragma unroll
                for (int k = 0; k <= PV; ++k) sv = __fma_rn(__ldg(ctrl + klo + k), wgt[k], sv);
                const double res = sr[q] - st * sv;
                acc = __fma_rn(wq * res, res, acc);
            }
        }
    }
    if (j < a.M) a.out[r * a.M + j] = acc;
    if (!ok) atomicOr(a.status, kStatOutOfRange);
}

// ------------------------------------------------------------------------------------------
// Diagonal of the Hessian of the chi-square with respect to the shift of one component
// (quickplay.get_RV_error_2d, jabble/quickplay.py:455-482: jacrev(grad(loss)) [key, key] per row):
//   d^2/d shift^2 of 0.5 ivar (y - m)^2 = ivar ((a S')^2 - (y - m) a S''),   m = a S(x - shift) + others
// S' and S'' are the derivatives of the piecewise polynomial (what differentiating the reference's
// floor / mod arithmetic gives).  One block per row; out[row] = the row's sum.
template <int PV>
__global__ void k_row_curvature(GridParams a) {
    __shared__ double red[kTile];
    const int64_t r = blockIdx.x;
    const CompDev& cs = a.comp[a.search];
    const double st = cs.stretch ? cs.stretch[cs.stretch_key[r]] : 1.0;
    const double sh = cs.shift[cs.shift_key[r]];
    double acc = 0.0;
    bool ok = true;
    for (int t = 0; t < a.n_tiles; ++t) {
        const double* rec = a.data + (r * a.n_tiles + t) * kTileDoubles;
        const int p = threadIdx.x;
        const int pos = (p % kPixPerLane) * 32 + p / kPixPerLane;
        const double x = rec[pos], y = rec[kTile + pos], w = rec[2 * kTile + pos];
        if (w == 0.0) continue;
        double mo = 0.0;
        for (int ci = 0; ci < a.n_comp; ++ci) {
            if (ci == a.search) continue;
            const CompDev& c = a.comp[ci];
            const double shift = c.shift ? c.shift[c.shift_key[r]] : 0.0;
            const double* cc = c.kind == 1 ? c.ctrl + (int64_t)c.ctrl_key[r] * c.n_knots : c.ctrl;
            double s;
            if (c.p_val == 3) s = spline_value<3>(c, cc, x, shift, ok);
            else if (c.p_val == 2) s = spline_value<2>(c, cc, x, shift, ok);
            else s = spline_value<1>(c, cc, x, shift, ok);
            if (c.stretch) s *= c.stretch[c.stretch_key[r]];
            mo += s;
        }
        double g; int I;
        knot_coord<PV>(x, sh, cs.xp0, cs.inv_dx, g, I);
        const double tt = ((x - sh) - cs.xp0) * cs.inv_dx;
        const int klo = knot_lo<PV>(I);
        if (!(tt > -1.0e9 && tt < 1.0e9) || klo < 0 || klo + PV > cs.n_knots - 1) { ok = false; continue; }
        double cv[PV + 1], wgt[PV + 1], dw[PV > 0 ? PV : 1];
#pragma unroll
        for (int k = 0; k <= PV; ++k) cv[k] = __ldg(cs.ctrl + klo + k);
        Basis<PV>::eval(g, wgt, dw);
        double S = 0.0, dS = 0.0, d2S = 0.0;
#pragma unroll
        for (int k = 0; k <= PV; ++k) S = __fma_rn(cv[k], wgt[k], S);
#pragma unroll
        for (int k = 0; k < PV; ++k) dS = __fma_rn(cv[k + 1] - cv[k], dw[k], dS);
        // second derivative: PV (PV-1) sum_j (c[j+2] - 2 c[j+1] + c[j]) * [(PV-2)!-scaled B-spline of degree PV-2]
        if (PV == 3) d2S = ((cv[2] - 2.0 * cv[1]) + cv[0]) * (0.5 - g) + ((cv[3] - 2.0 * cv[2]) + cv[1]) * (0.5 + g);
        else if (PV == 2) d2S = (cv[2] - 2.0 * cv[1]) + cv[0];
        const double s1 = PV * cs.inv_dx * dS * st;                          // a S'
        const double s2 = PV * (PV - 1) * cs.inv_dx * cs.inv_dx * d2S * st;  // a S''
        const double res = y - (mo + st * S);
        acc += w * (s1 * s1 - res * s2);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kTile / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.out[r] = red[0];
    if (!ok) atomicOr(a.status, kStatOutOfRange);
}

// ------------------------------------------------------------------------------------------
// Fused loss + gradient + Fisher.
//
// Work item ("task") = one warp, one tile of kTile pixel columns, one group of <= 32 rows.
//
// Data movement: lane 0 issues ONE 3 KB bulk async copy (TMA engine) per row into a kStages-deep
// ring in shared memory, completion counted on an mbarrier, so kStages-1 rows are in flight per
// warp independent of occupancy.  The control points the task can touch (a window of <= wincap
// knots per component) are copied to shared memory once per task and read from there per pixel.
//
// Arithmetic: lane l owns pixel columns [4l, 4l+4) of the tile for every row, so its pixels walk
// through the knot intervals monotonically and the control-point adjoint of the current interval
// lives in registers: p_val+1 accumulators that slide (with selects, no branches) when the lane
// moves to the next interval; the knot that leaves the window is added, with a plain
// shared-memory read-modify-write, to the warp's gradient window.  Rows are walked in alternating
// direction (pixels 0..3, then 3..0) so the accumulators carry over from row to row and nothing is
// flushed at a row end.  Lanes whose knot ranges could overlap never share an accumulator copy:
// there are kClasses copies and lane l uses copy l % kClasses; lanes l and l+4 are 13 pixels apart,
// and (row, tile) pairs where that is not enough knots ("dense", decided at upload from x alone:
// knot distances inside a row do not depend on the shift) are walked lane by lane.
// Every sum has one fixed order: no float atomics, bit-reproducible.
// Per (tile, sorted row): extent of the pixels the kernel walks and the path flags
// (1 = walk lane by lane, 2 = nothing weighted, 4 = general per-pixel path).
struct alignas(16) TileInfo { double xmin, xmax; int flags; unsigned live; int pad_[2]; };   // live: lanes with a weighted pixel

struct FusedParams {
    const double* data;                         // tile-interleaved x | y | ivar
    const void* grouprec;                       // [E] RowRec<NC> in sorted-row order (k_prepare_rows, every evaluation)
    const TileInfo* tinfo;                      // [n_tiles][E] in sorted-row order (host, at (re)sort time)
    int64_t n_rows;
    int n_tiles, rows_per_group, n_groups;
    int wincap;          // knots per on-chip window (multiple of 32, <= kWinCapMax)
    int64_t n_tasks;
    int n_comp;
    CompDev comp[kMaxComp];
    double* part;        // [n_comp][n_tasks][wincap] per-task gradient windows
    int* part_base;      // [n_comp][n_tasks] first knot of the window (INT_MAX: empty)
    double* rowpart;     // [E][n_tiles][NQP]   per-row sums: (d/dshift, d/dstretch, Fisher) per component
    double* losspart;    // [n_tasks] sum of ivar*res^2 of the task
    double* normpart;    // [E][n_tiles][norm_stride] gradient of the row's normalisation control points
    int norm_stride;     // knots per key rounded up to 4 (<= 32)
    int* status;         // [0] error bits  [1] rows that took the serial path
};

// per-row sums: (d/dshift, d/dstretch, Fisher) per global component, (d/dshift, d/dstretch) for the
// normalisation component; padded to 4 or 8 for the multi-value warp reduction
template <int NC, int HN> struct RowQ { static constexpr int NQ = NC * kRowQ + HN * 2; static constexpr int NQP = NQ <= 4 ? 4 : 8; };
__host__ __device__ constexpr int row_sum_stride(int n_global, int has_norm) { return (n_global * kRowQ + has_norm * 2) <= 4 ? 4 : 8; }

template <int NC, int HN>
__host__ __device__ constexpr size_t fused_warp_smem(int wincap) {
    return ((size_t)NC * kClasses * wincap * 8 + (size_t)NC * wincap * 8 + (size_t)kStages * kTileBytes +
            (size_t)kMaxGroup * sizeof(RowRec<NC, HN>) + (size_t)HN * 3 * kNormMaxKnots * 8 + kStages * 8 + 127) / 128 * 128;
}

// ---- normalisation component (NormalizationModel over a CardinalSplineMixture, model.py:1110-1143)
// Few knots, hundreds of pixels per knot interval: a lane's 4 pixels sit in one interval or straddle
// one knot (enforced at upload), so the lane keeps PVN+2 control points and PVN+2 adjoint sums for
// the whole row in registers, indexed from the interval of its lowest pixel; after the row the lanes
// are summed by knot with warp shuffles (lanes grouped by base knot, in lane order).
template <int PVN>
struct NormLane {
    static constexpr int NS = PVN + 2;
    double c[PVN + 2];
    double a[PVN + 2];
    int kbase;
};

template <int PVN>
__device__ __forceinline__ void norm_begin(NormLane<PVN>& N, double x_low, double sh, const CompDev& cd, const double* nctrl, int& bad) {
    double g; int I;
    knot_coord<PVN>(x_low, sh, cd.xp0, cd.inv_dx, g, I);
    N.kbase = knot_lo<PVN>(I);
    const double t = ((x_low - sh) - cd.xp0) * cd.inv_dx;
    if (!(t > -1.0e9 && t < 1.0e9) || N.kbase < 0 || N.kbase + PVN + 1 > cd.n_knots) { bad = 1; N.kbase = 0; }
#pragma unroll
    for (int s = 0; s < PVN + 2; ++s) {
        N.c[s] = nctrl[min(N.kbase + s, kNormMaxKnots - 1)];
        N.a[s] = 0.0;
    }
}

// value S and derivative dS (per knot, before the PVN/dx factor) at x; ws = basis weights placed on
// the lane's PVN+2 slots
template <int PVN>
__device__ __forceinline__ void norm_eval(const NormLane<PVN>& N, double x, double sh, const CompDev& cd,
                                          double (&ws)[PVN + 2], double& S, double& dS, int& bad) {
    double g; int I;
    knot_coord<PVN>(x, sh, cd.xp0, cd.inv_dx, g, I);
    const int o = knot_lo<PVN>(I) - N.kbase;            // 0 or 1 by construction of the plan
    if (o < 0 || o > 1 || (o == 1 && N.kbase + PVN + 2 > cd.n_knots)) bad = 1;
    const bool up = o >= 1;
    double w[PVN + 1], dw[PVN];
    Basis<PVN>::eval(g, w, dw);
    double e[PVN + 1];                                   // dS = sum_k c[k] * e[k]
    e[0] = -dw[0];
#pragma unroll
    for (int k = 1; k < PVN; ++k) e[k] = dw[k - 1] - dw[k];
    e[PVN] = dw[PVN - 1];
    S = 0.0; dS = 0.0;
#pragma unroll
    for (int s = 0; s < PVN + 2; ++s) {
        const double wl = s <= PVN ? w[s] : 0.0, wh = s >= 1 ? w[s - 1] : 0.0;
        const double el = s <= PVN ? e[s] : 0.0, eh = s >= 1 ? e[s - 1] : 0.0;
        ws[s] = up ? wh : wl;
        S = __fma_rn(N.c[s], ws[s], S);
        dS = __fma_rn(N.c[s], up ? eh : el, dS);
    }
}

// sum the lanes' adjoints by knot into rowg[0..K) (shared memory, this warp), lanes grouped by base knot
template <int PVN>
__device__ __forceinline__ void norm_reduce(const NormLane<PVN>& N, int lane, double* rowg) {
    unsigned remaining = 0xffffffffu;
    while (remaining) {
        const int cur = __shfl_sync(0xffffffffu, N.kbase, __ffs(remaining) - 1);
        const bool mine = N.kbase == cur;
        double v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) v[s] = (s < PVN + 2 && mine) ? N.a[s < PVN + 2 ? s : 0] : 0.0;
        const double tot = warp_sum_multi<8>(v, lane);
        const int s = lane >> 2;
        if ((lane & 3) == 0 && s < PVN + 2 && cur + s < kNormMaxKnots) rowg[cur + s] += tot;
        remaining &= ~__ballot_sync(0xffffffffu, mine);
        __syncwarp();
    }
}


template <int PV>
struct Acc {
    double a[PV + 1];   // adjoint sums of knots klo .. klo+PV accumulated by this lane
    double c[PV + 1];   // control points of those knots (fast path only; refreshed at every row start)
    int klo;
};

template <int PV>
__device__ __forceinline__ void acc_flush(Acc<PV>& A, double* accw, int base) {
#pragma unroll
    for (int j = 0; j <= PV; ++j) { accw[A.klo + j - base] += A.a[j]; A.a[j] = 0.0; }
}

// Move the lane's window to the knots of its next pixel.  DIR = +1: the lane walks its pixels
// upwards in x, -1: downwards.  The common case (same interval, or one interval further in the
// walking direction) is branch-free: the knot that leaves the window is retired to shared memory
// and the accumulators slide by one.
template <int PV, int DIR>
__device__ __forceinline__ void acc_move(Acc<PV>& A, int kl, double* accw, int base) {
    const int d = kl - A.klo;
    const bool step = d == DIR;
    if (d != 0 && !step) {            // turn-around mismatch or a jump over several knots: rare
        acc_flush<PV>(A, accw, base);
        A.klo = kl;
    }
    if (step) {
        if (DIR > 0) accw[A.klo - base] += A.a[0];
        else accw[A.klo + PV - base] += A.a[PV];
    }
    if (DIR > 0) {
#pragma unroll
        for (int j = 0; j < PV; ++j) A.a[j] = step ? A.a[j + 1] : A.a[j];
        A.a[PV] = step ? 0.0 : A.a[PV];
    } else {
#pragma unroll
        for (int j = PV; j > 0; --j) A.a[j] = step ? A.a[j - 1] : A.a[j];
        A.a[0] = step ? 0.0 : A.a[0];
    }
    A.klo += step ? DIR : 0;
}

template <int NC, int PV, int PVN, int DIR>
__device__ __forceinline__ void fused_pixel(double x, double y, double iv, const double* sh, const double* st,
                                            const CompDev* comp, Acc<PV> (&acc)[NC], const double* ctrl_warp,
                                            double* acc_cls, const int W, const int (&lo)[NC], double& loss,
                                            double (&sums)[RowQ<NC, (PVN > 0)>::NQP], NormLane<PVN>& nl, int& bad) {
    double w[NC][PV + 1], S[NC], dS[NC];
    double m = 0.0;
    double nws[PVN + 2], nS = 0.0, ndS = 0.0;
    (void)nws; (void)ndS;
    if constexpr (PVN > 0) {
        norm_eval<PVN>(nl, x, sh[NC], comp[NC], nws, nS, ndS, bad);
        m = st[NC] * nS;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double g; int I;
        knot_coord<PV>(x, sh[c], comp[c].xp0, comp[c].inv_dx, g, I);
        acc_move<PV, DIR>(acc[c], knot_lo<PV>(I), acc_cls + c * kClasses * W, lo[c]);
        const double* cw = ctrl_warp + c * W + (acc[c].klo - lo[c]);
        double cv[PV + 1];
#pragma unroll
        for (int k = 0; k <= PV; ++k) cv[k] = cw[k];
        double dw[PV];
        Basis<PV>::eval(g, w[c], dw);
        double sv = cv[0] * w[c][0];
        double d = (cv[1] - cv[0]) * dw[0];
#pragma unroll
        for (int k = 1; k <= PV; ++k) sv = __fma_rn(cv[k], w[c][k], sv);
#pragma unroll
        for (int k = 1; k < PV; ++k) d = __fma_rn(cv[k + 1] - cv[k], dw[k], d);
        S[c] = sv; dS[c] = d;
        m = __fma_rn(st[c], sv, m);                // StretchingModel / AdditiveModel
    }
    const double res = y - m;                      // ChiSquare, loss.py:169-171
    const double wr = iv * res;
    loss = __fma_rn(wr, res, loss);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double q = wr * st[c];
#pragma unroll
        for (int k = 0; k <= PV; ++k) acc[c].a[k] = __fma_rn(q, w[c][k], acc[c].a[k]);
        sums[c * kRowQ + 0] = __fma_rn(q, dS[c], sums[c * kRowQ + 0]);              // d/d shift
        sums[c * kRowQ + 1] = __fma_rn(wr, S[c], sums[c * kRowQ + 1]);              // d/d stretch
        sums[c * kRowQ + 2] = __fma_rn(iv * dS[c], dS[c], sums[c * kRowQ + 2]);     // Fisher
    }
    if constexpr (PVN > 0) {
        const double q = wr * st[NC];
#pragma unroll
        for (int k = 0; k < PVN + 2; ++k) nl.a[k] = __fma_rn(q, nws[k], nl.a[k]);
        sums[NC * kRowQ + 0] = __fma_rn(q, ndS, sums[NC * kRowQ + 0]);
        sums[NC * kRowQ + 1] = __fma_rn(wr, nS, sums[NC * kRowQ + 1]);
    }
}

// ---- fast path: rows whose consecutive pixels are less than one knot apart ("smooth" rows) ------
// Inside such a row a lane's window either stays or moves by exactly one knot in the walking
// direction, so the per-pixel code needs no branch: the slide is folded into the accumulation
// (the addend of the FMA is the neighbour's accumulator when the window moved).
template <int PV, int DIR>
__device__ __forceinline__ void acc_reposition(Acc<PV>& A, int kl, const double* ctrlw, double* accw, int base) {
    // row start: bring the window to the first pixel of the row (usually already there) and take
    // its control points into registers
    if (kl != A.klo) {
        acc_flush<PV>(A, accw, base);
        A.klo = kl;
    }
#pragma unroll
    for (int k = 0; k <= PV; ++k) A.c[k] = ctrlw[kl + k - base];
}

template <int NC, int PV, int PVN, int DIR>
__device__ __forceinline__ void fused_pixel_fast(double x, double y, double iv, const double* sh, const double* st,
                                                 const CompDev* comp, Acc<PV> (&acc)[NC], const double* ctrl_warp,
                                                 double* acc_cls, const int W, const int (&lo)[NC], double& loss,
                                                 double (&sums)[RowQ<NC, (PVN > 0)>::NQP], NormLane<PVN>& nl, int& bad) {
    double w[NC][PV + 1], S[NC], dS[NC];
    bool step[NC];
    double m = 0.0;
    double nws[PVN + 2], nS = 0.0, ndS = 0.0;
    (void)nws; (void)ndS;
    if constexpr (PVN > 0) {
        norm_eval<PVN>(nl, x, sh[NC], comp[NC], nws, nS, ndS, bad);
        m = st[NC] * nS;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double g; int I;
        knot_coord<PV>(x, sh[c], comp[c].xp0, comp[c].inv_dx, g, I);
        const int kl = knot_lo<PV>(I);
        Acc<PV>& A = acc[c];
        step[c] = kl != A.klo;
        double* accw = acc_cls + c * kClasses * W;
        const double* ctrlw = ctrl_warp + c * W;
        // the knot that enters the window / the one that leaves it
        double cnew = DIR > 0 ? A.c[PV] : A.c[0];
        if (step[c]) {
            if (DIR > 0) { accw[A.klo - lo[c]] += A.a[0]; cnew = ctrlw[kl + PV - lo[c]]; }
            else { accw[A.klo + PV - lo[c]] += A.a[PV]; cnew = ctrlw[kl - lo[c]]; }
        }
        A.klo = kl;
        if (DIR > 0) {
#pragma unroll
            for (int k = 0; k < PV; ++k) A.c[k] = step[c] ? A.c[k + 1] : A.c[k];
            A.c[PV] = cnew;
        } else {
#pragma unroll
            for (int k = PV; k > 0; --k) A.c[k] = step[c] ? A.c[k - 1] : A.c[k];
            A.c[0] = cnew;
        }
        double dw[PV];
        Basis<PV>::eval(g, w[c], dw);
        double sv = A.c[0] * w[c][0];
        double d = (A.c[1] - A.c[0]) * dw[0];
#pragma unroll
        for (int k = 1; k <= PV; ++k) sv = __fma_rn(A.c[k], w[c][k], sv);
#pragma unroll
        for (int k = 1; k < PV; ++k) d = __fma_rn(A.c[k + 1] - A.c[k], dw[k], d);
        S[c] = sv; dS[c] = d;
        m = __fma_rn(st[c], sv, m);
    }
    const double res = y - m;
    const double wr = iv * res;
    loss = __fma_rn(wr, res, loss);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double q = wr * st[c];
        if (DIR > 0) {
#pragma unroll
            for (int k = 0; k < PV; ++k) acc[c].a[k] = __fma_rn(q, w[c][k], step[c] ? acc[c].a[k + 1] : acc[c].a[k]);
            acc[c].a[PV] = __fma_rn(q, w[c][PV], step[c] ? 0.0 : acc[c].a[PV]);
        } else {
#pragma unroll
            for (int k = PV; k > 0; --k) acc[c].a[k] = __fma_rn(q, w[c][k], step[c] ? acc[c].a[k - 1] : acc[c].a[k]);
            acc[c].a[0] = __fma_rn(q, w[c][0], step[c] ? 0.0 : acc[c].a[0]);
        }
        sums[c * kRowQ + 0] = __fma_rn(q, dS[c], sums[c * kRowQ + 0]);
        sums[c * kRowQ + 1] = __fma_rn(wr, S[c], sums[c * kRowQ + 1]);
        sums[c * kRowQ + 2] = __fma_rn(iv * dS[c], dS[c], sums[c * kRowQ + 2]);
    }
    if constexpr (PVN > 0) {
        const double q = wr * st[NC];
#pragma unroll
        for (int k = 0; k < PVN + 2; ++k) nl.a[k] = __fma_rn(q, nws[k], nl.a[k]);
        sums[NC * kRowQ + 0] = __fma_rn(q, ndS, sums[NC * kRowQ + 0]);
        sums[NC * kRowQ + 1] = __fma_rn(wr, nS, sums[NC * kRowQ + 1]);
    }
}

// the four pixels of a lane in walking order; px = the lane's column in the staged record, which
// is stored pixel-slot-major: pixel 4*lane + j of the tile sits at [j * 32 + lane] (conflict-free
// shared-memory reads), x | y | ivar kTile doubles apart
template <int NC, int PV, int PVN, int DIR>
__device__ __forceinline__ void fused_run_fast(const double* px, const double* sh, const double* st,
                                               const CompDev* comp, Acc<PV> (&acc)[NC], const double* ctrl_warp,
                                               double* acc_cls, const int W, const int (&lo)[NC], double& loss,
                                               double (&sums)[RowQ<NC, (PVN > 0)>::NQP], NormLane<PVN>& nl, int& bad) {
    constexpr int j0 = DIR > 0 ? 0 : 3;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double g; int I;
        knot_coord<PV>(px[32 * j0], sh[c], comp[c].xp0, comp[c].inv_dx, g, I);
        acc_reposition<PV, DIR>(acc[c], knot_lo<PV>(I), ctrl_warp + c * W, acc_cls + c * kClasses * W, lo[c]);
    }
#ifndef JB_PIXEL_UNROLL
#define JB_PIXEL_UNROLL 1
#endif
    constexpr int kPixelUnroll = JB_PIXEL_UNROLL;
#pragma unroll kPixelUnroll
    for (int jj = 0; jj < 4; ++jj) {
        const int j = DIR > 0 ? jj : 3 - jj;
        fused_pixel_fast<NC, PV, PVN, DIR>(px[32 * j], px[kTile + 32 * j], px[2 * kTile + 32 * j], sh, st, comp, acc,
                                           ctrl_warp, acc_cls, W, lo, loss, sums, nl, bad);
    }
}

template <int NC, int PV, int PVN>
__global__ void __maxnreg__(JB_MAXREG) k_fused(const FusedParams* __restrict__ arr, const int* __restrict__ off, int n) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ FusedParams a;            // this block's plan
    const int pi = find_plan(off, n, blockIdx.x);
    {
        const int* src = reinterpret_cast<const int*>(arr + pi);
        int* dst = reinterpret_cast<int*>(&a);
        for (int i = threadIdx.x; i < (int)(sizeof(FusedParams) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)(blockIdx.x - off[pi]) * kWarpsPerBlock + warp;
    if (task >= a.n_tasks) return;
    const int tile = (int)(task % a.n_tiles);
    const int group = (int)(task / a.n_tiles);
    const int row0 = group * a.rows_per_group;
    const int nrow = min(a.rows_per_group, (int)a.n_rows - row0);
    constexpr int HN = PVN > 0 ? 1 : 0;
    constexpr int NQP = RowQ<NC, HN>::NQP;
    using Rec = RowRec<NC, HN>;
    const int W = a.wincap;

    unsigned char* wbase = smem_raw + (size_t)warp * fused_warp_smem<NC, HN>(W);
    double* acc_warp = reinterpret_cast<double*>(wbase);                     // [NC][kClasses][W]
    double* ctrl_warp = acc_warp + NC * kClasses * W;                        // [NC][W]
    unsigned char* ring = reinterpret_cast<unsigned char*>(ctrl_warp + NC * W);   // [kStages][kTileBytes]
    Rec* recs = reinterpret_cast<Rec*>(ring + (size_t)kStages * kTileBytes);
    double* nctrl = reinterpret_cast<double*>(recs + kMaxGroup);             // [2][kNormMaxKnots] control points of the row's key
    double* rowg = nctrl + 2 * kNormMaxKnots * HN;                            // [kNormMaxKnots] gradient of the row's key
    const uint32_t bar0 = smem_u32(rowg + kNormMaxKnots * HN);               // [kStages] mbarriers

    // ---- 0. lane i fetches the record of row i of the group: two independent loads ------------
    int lo[NC], hi[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { lo[c] = INT_MAX; hi[c] = INT_MIN; }
    bool has_work = false;
    if (lane < nrow) {
        Rec rec = reinterpret_cast<const Rec*>(a.grouprec)[row0 + lane];
        const TileInfo ti = a.tinfo[(int64_t)tile * a.n_rows + row0 + lane];
        rec.dense = ti.flags;
        has_work = ti.flags != 2;
        if (has_work) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const CompDev& cd = a.comp[c];
                double g; int I0, I1;
                knot_coord<PV>(ti.xmin, rec.sh[c], cd.xp0, cd.inv_dx, g, I0);
                knot_coord<PV>(ti.xmax, rec.sh[c], cd.xp0, cd.inv_dx, g, I1);
                // guard the magic-number trick: coordinates must be far inside +-2^31
                const double t0 = ((ti.xmin - rec.sh[c]) - cd.xp0) * cd.inv_dx, t1 = ((ti.xmax - rec.sh[c]) - cd.xp0) * cd.inv_dx;
                if (!(t0 > -1.0e9 && t1 < 1.0e9)) { I0 = -(1 << 30); I1 = (1 << 30); }
                lo[c] = knot_lo<PV>(I0);
                hi[c] = knot_lo<PV>(I1) + PV;
            }
        }
        recs[lane] = rec;
    }
    const bool empty = !__any_sync(0xffffffffu, has_work);
    auto issue_row = [&](int i, int r) {      // lane 0 only
        const int s = i % kStages;
        const uint32_t bar = bar0 + 8 * s;
        mbar_expect_tx(bar, kTileBytes);
        tma_load_1d(smem_u32(ring + (size_t)s * kTileBytes), a.data + ((int64_t)r * a.n_tiles + tile) * kTileDoubles,
                    kTileBytes, bar);
    };
    // ---- 1. start the first copies as early as possible ------------------------------------------
    __syncwarp();
    if (lane == 0 && !empty) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
        for (int i = 0; i < kStages && i < nrow; ++i) issue_row(i, recs[i].row);
    }
    // ---- 2. knot window of the task ---------------------------------------------------------
    bool bad_range = false, bad_window = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (!empty) {
            if (lo[c] < 0 || hi[c] > a.comp[c].n_knots - 1) bad_range = true;
            if (hi[c] - lo[c] + 1 > W) bad_window = true;
        }
    }
    const bool skip = empty || bad_range || bad_window;
    if (lane == 0) {
        if (bad_range) atomicOr(a.status, kStatOutOfRange);
        else if (bad_window) atomicOr(a.status, kStatWindow);
#pragma unroll
        for (int c = 0; c < NC; ++c) a.part_base[(int64_t)c * a.n_tasks + task] = skip ? INT_MAX : lo[c];
        if (skip) a.losspart[task] = 0.0;
    }
    if (skip) {   // nothing weighted here (or an error): the row sums of this tile are zero
        for (int i = lane; i < nrow * NQP; i += 32)
            a.rowpart[((int64_t)recs[i / NQP].row * a.n_tiles + tile) * NQP + i % NQP] = 0.0;
        if (HN)
            for (int i = lane; i < nrow * a.norm_stride; i += 32)
                a.normpart[((int64_t)recs[i / a.norm_stride].row * a.n_tiles + tile) * a.norm_stride + i % a.norm_stride] = 0.0;
        if (!empty)   // copies are in flight into this warp's ring: let them land before leaving
            for (int i = 0; i < kStages && i < nrow; ++i) mbar_wait(bar0 + 8 * (i % kStages), 0);
        return;
    }

    // ---- 3. accumulators and control-point window ------------------------------------------------
    {
        double2* z = reinterpret_cast<double2*>(acc_warp);
        for (int i = lane; i < NC * kClasses * W / 2; i += 32) z[i] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int n = hi[c] - lo[c] + 1;
        for (int k = lane; k < W; k += 32) ctrl_warp[c * W + k] = k < n ? __ldg(a.comp[c].ctrl + lo[c] + k) : 0.0;
    }
    __syncwarp();

    double* acc_cls = acc_warp + (lane & (kClasses - 1)) * W;   // + c * kClasses * W per component
    Acc<PV> acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        acc[c].klo = lo[c];            // a valid (empty) window until the first pixel moves it
#pragma unroll
        for (int k = 0; k <= PV; ++k) acc[c].a[k] = 0.0;
    }
    int serial_rows = 0;
    double loss = 0.0;
    int bad = 0;
    NormLane<PVN> nl;
    const int nK = HN ? a.comp[NC].n_knots : 0;
    double nnext = 0.0;                  // control point `lane` of the NEXT row's key, fetched one row ahead
    if (HN) {
        if (lane < kNormMaxKnots) { rowg[lane] = 0.0; nctrl[lane] = lane < nK ? __ldg(a.comp[NC].ctrl + (int64_t)recs[0].nkey * nK + lane) : 0.0; }
        __syncwarp();
    }

    // ---- 4. stream the rows of the group ------------------------------------------------------
    for (int i = 0; i < nrow; ++i) {
        const int s = i % kStages;
        const Rec rec = recs[i];
        if (HN && i + 1 < nrow && lane < nK) nnext = __ldg(a.comp[NC].ctrl + (int64_t)recs[i + 1].nkey * nK + lane);
        mbar_wait(bar0 + 8 * s, (uint32_t)((i / kStages) & 1));
        if (rec.dense == 2) {          // no weighted pixel in this row of the tile
            if (lane < NQP) a.rowpart[((int64_t)rec.row * a.n_tiles + tile) * NQP + lane] = 0.0;
            if (HN) {
                if (lane < a.norm_stride) a.normpart[((int64_t)rec.row * a.n_tiles + tile) * a.norm_stride + lane] = 0.0;
                if (lane < kNormMaxKnots) nctrl[((i + 1) & 1) * kNormMaxKnots + lane] = nnext;
            }
            __syncwarp();
            if (lane == 0 && i + kStages < nrow) issue_row(i + kStages, recs[i + kStages].row);
            continue;
        }
        const double* pl4 = reinterpret_cast<const double*>(ring + (size_t)s * kTileBytes) + lane;
        if constexpr (HN != 0) norm_begin<PVN>(nl, pl4[0], rec.sh[NC], a.comp[NC], nctrl + (i & 1) * kNormMaxKnots, bad);

        double sums[NQP];
#pragma unroll
        for (int q = 0; q < NQP; ++q) sums[q] = 0.0;

        if (rec.dense == 0) {
            // smooth row: branch-free sliding
            if ((i & 1) == 0) fused_run_fast<NC, PV, PVN, +1>(pl4, rec.sh, rec.st, a.comp, acc, ctrl_warp, acc_cls, W, lo, loss, sums, nl, bad);
            else fused_run_fast<NC, PV, PVN, -1>(pl4, rec.sh, rec.st, a.comp, acc, ctrl_warp, acc_cls, W, lo, loss, sums, nl, bad);
        } else {
            // general row (pixels a knot or more apart, non-monotone x), lane by lane when the lanes
            // of one class could touch the same knots (rec.dense & 1)
            const bool serial = (rec.dense & 1) != 0;
            if (serial) ++serial_rows;
            const int nturn = serial ? 32 / kClasses : 1;
            for (int turn = 0; turn < nturn; ++turn) {
                if (!serial || (lane / kClasses) == turn) {
#pragma unroll 1
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = (i & 1) ? 3 - jj : jj;
                        const double* pj = reinterpret_cast<const double*>(ring + (size_t)s * kTileBytes) + lane + 32 * j;
                        const double xj = pj[0], yj = pj[kTile], wj = pj[2 * kTile];
                        fused_pixel<NC, PV, PVN, +1>(xj, yj, wj, rec.sh, rec.st, a.comp, acc, ctrl_warp, acc_cls, W, lo, loss, sums, nl, bad);
                    }
                    if (serial) {   // lanes take turns: nothing may stay in registers across turns
#pragma unroll
                        for (int c = 0; c < NC; ++c) acc_flush<PV>(acc[c], acc_cls + c * kClasses * W, lo[c]);
                    }
                }
                __syncwarp();
            }
        }

        // per-row sums of this tile (fixed order); lane L ends up with value q = L >> (NQP == 8 ? 2 : 3)
        const double tot = warp_sum_multi<NQP>(sums, lane);
        constexpr int kShift = (NQP == 8) ? 2 : 3;
        if ((lane & ((1 << kShift) - 1)) == 0) {
            const int q = lane >> kShift;
            double s2 = 1.0;                       // Fisher sums carry the squared stretch of the row
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (q == c * kRowQ + 2) s2 = rec.st[c] * rec.st[c];
            a.rowpart[((int64_t)rec.row * a.n_tiles + tile) * NQP + q] = tot * s2;
        }
        if constexpr (HN != 0) {   // gradient of the row's own control points: lanes summed by knot, then one store per knot
            norm_reduce<PVN>(nl, lane, rowg);
            if (lane < a.norm_stride) a.normpart[((int64_t)rec.row * a.n_tiles + tile) * a.norm_stride + lane] = lane < kNormMaxKnots ? rowg[lane] : 0.0;
            if (lane < kNormMaxKnots) { rowg[lane] = 0.0; nctrl[((i + 1) & 1) * kNormMaxKnots + lane] = nnext; }
        }
        // every lane has consumed stage s (the sums above depend on all of its data): refill it
        __syncwarp();
        if (lane == 0 && i + kStages < nrow) issue_row(i + kStages, recs[i + kStages].row);
    }

    // ---- 5. flush, fold the accumulator copies in a fixed order, publish the window ------------
    // All windows now sit on pixels of the same (last walked) row, so lanes of one class are on
    // disjoint knots (or hold zeros after a lane-by-lane row): one parallel flush.
#pragma unroll
    for (int c = 0; c < NC; ++c) acc_flush<PV>(acc[c], acc_cls + c * kClasses * W, lo[c]);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double* ac = acc_warp + c * kClasses * W;
        double* dst = a.part + ((int64_t)c * a.n_tasks + task) * W;
        for (int k = lane; k < W; k += 32) {
            double sacc = ac[k];
#pragma unroll
            for (int q = 1; q < kClasses; ++q) sacc += ac[q * W + k];
            dst[k] = sacc;
        }
    }
    loss = warp_sum(loss);
    bad = __any_sync(0xffffffffu, bad != 0);
    if (lane == 0) {
        a.losspart[task] = loss;
        if (serial_rows) atomicAdd(a.status + 1, serial_rows);
        if (bad) atomicOr(a.status, kStatOutOfRange);
    }
}

// ------------------------------------------------------------------------------------------
// Reduction of the partials, every sum in a fixed order.
struct ReduceParams {
    int n_comp;
    CompDev comp[kMaxComp];
    int64_t ctrl_off[kMaxComp], shift_off[kMaxComp], stretch_off[kMaxComp];
    int n_shift[kMaxComp], n_stretch[kMaxComp];
    const int* shift_rowptr[kMaxComp];   const int* shift_rows[kMaxComp];     // key -> rows (CSR)
    const int* stretch_rowptr[kMaxComp]; const int* stretch_rows[kMaxComp];
    int64_t n_rows, n_tasks;
    int n_tiles, n_groups, wincap;
    int nb_ctrl;            // blocks of k_reduce_stage1 that sum gradient windows (n_comp * n_groups)
    const double* part; const int* part_base; const double* rowpart; const double* losspart;
    // normalisation component (index n_comp, when has_norm): per-row gradients of its control points
    int has_norm, norm_stride, norm_knots, n_norm_keys;
    int64_t norm_ctrl_off;
    const double* normpart; double* normrow;           // [E][n_tiles][norm_stride] -> [E][norm_stride]
    const int* norm_rowptr; const int* norm_rows;      // key -> rows (CSR)
    int nb_stage1_rows;                                 // blocks of the per-row tile sums in stage 1
    // layout of a row of rowpart / rowsum: nq values; rq[c][kind] = position of component c's
    // shift-gradient (0), stretch-gradient (1), Fisher (2) sum, -1 when the fused kernel of this
    // plan does not compute it
    int nq;
    int rq[kMaxComp][3];
    double* tilepart;   // [n_comp][n_tiles][max_knots]: per-tile sums over the row groups (k_reduce_stage1)
    int* tile_rng;      // [n_comp][n_tiles][2]: the knots [lo, hi) of tilepart that are valid
    int64_t max_knots;
    double* rowsum;     // [E][NQ]
    const int* row_of_pos;   // k_strip plans: rowpart row -> caller's row (nullptr: the same)
    double coef;
    double* grad_all; double* fisher_all;
    const int64_t* fit_map; int64_t n_fit;
    double* out;        // [1 + n_fit]
    int* status;
};

// Gradient of the control points.  A task (row group g, tile t) leaves a window of `wincap` knot sums
// in `part`; the windows of one tile start at nearly the same knot in every group (the groups differ
// by their rows' shifts only).  Blocks [0, nb_ctrl) of k_reduce_stage1, one per (component, tile): a
// thread owns one knot of the union of the tile's windows and adds, in group order, the element of
// every group's window that lies on it -- one predicated load per (knot, group), all independent,
// every element of `part` read exactly once.  The result, tilepart[c][t][knot] on the knots
// [tile_rng[c][t][0], tile_rng[c][t][1]), is a few hundred doubles per tile; k_reduce_stage2 adds
// the (one to three) tiles that cover a knot in tile order.  One fixed order throughout.
constexpr int kRedSlab = 256;      // threads per block of the reducers
constexpr int kRedBases = 2048;    // window bases of one tile held in shared memory at a time

// blocks [0, nb_ctrl): tilepart (above).  blocks [nb_ctrl, ...): rowsum[r][q] = sum over tiles of rowpart.
__global__ void k_reduce_stage1(const ReduceParams* __restrict__ arr, const int* __restrict__ off, int n) {
    const int pi = find_plan(off, n, blockIdx.x);
    const ReduceParams& a = arr[pi];
    const int blk = blockIdx.x - off[pi];
    const int nb_ctrl = a.nb_ctrl;       // = n_comp * n_tiles
    const int NQ = a.nq;
    if (blk < nb_ctrl) {
        __shared__ int sb[kRedBases];
        __shared__ int s_lo, s_hi;
        const int NT = a.n_tiles, W = a.wincap, n_groups = a.n_groups;
        const int c = blk / NT, t = blk - c * NT;
        const int M = a.comp[c].n_knots;
        const int* __restrict__ bases = a.part_base + (int64_t)c * a.n_tasks + t;      // group g at [g * NT]
        const double* __restrict__ part = a.part + ((int64_t)c * a.n_tasks + t) * W;   // group g at [g * NT * W]
        const int64_t gstride = (int64_t)NT * W;
        if (threadIdx.x == 0) { s_lo = INT_MAX; s_hi = -1; }
        __syncthreads();
        {   // the union of the tile's windows over the groups
            int lo = INT_MAX, hi = -1;
            for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
                const int base = bases[(int64_t)g * NT];
                if (base != INT_MAX) { lo = min(lo, base); hi = max(hi, base); }
            }
            if (hi >= 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
        }
        __syncthreads();
        const int k_lo = s_lo, k_hi = s_hi < 0 ? -1 : min(s_hi + W, M);    // knots [k_lo, k_hi)
        if (threadIdx.x == 0) {
            a.tile_rng[2 * blk] = s_hi < 0 ? 0 : k_lo;
            a.tile_rng[2 * blk + 1] = s_hi < 0 ? 0 : k_hi;
        }
        if (s_hi < 0) return;
        double* __restrict__ dst = a.tilepart + (int64_t)blk * a.max_knots;
        for (int kb = k_lo; kb < k_hi; kb += blockDim.x) {          // normally one pass
            const int k = kb + (int)threadIdx.x;
            double s = 0.0;
            for (int g0 = 0; g0 < n_groups; g0 += kRedBases) {      // normally one pass
                const int ng = min(kRedBases, n_groups - g0);
                __syncthreads();
                for (int g = threadIdx.x; g < ng; g += blockDim.x) sb[g] = bases[(int64_t)(g0 + g) * NT];
                __syncthreads();
                const double* __restrict__ pg = part + (int64_t)g0 * gstride;
                int g = 0;
                for (; g + 7 < ng; g += 8) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const unsigned w = (unsigned)(k - sb[g + u]);       // no window (INT_MAX): wraps far above W
                        v[u] = w < (unsigned)W ? __ldg(pg + (int64_t)(g + u) * gstride + w) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) s += v[u];
                }
                for (; g < ng; ++g) {
                    const unsigned w = (unsigned)(k - sb[g]);
                    s += w < (unsigned)W ? __ldg(pg + (int64_t)g * gstride + w) : 0.0;
                }
            }
            if (k < k_hi) dst[k] = s;
        }
    } else if (blk < nb_ctrl + a.nb_stage1_rows) {
        // per-row sums over the tiles: one WARP per row, lane l adds the tiles l, l + 32, .. in tile order,
        // then the lanes are added in a fixed butterfly (the same tree for every row and every run)
        const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
        const int64_t r = (int64_t)(blk - nb_ctrl) * warps_per_block + (threadIdx.x >> 5);
        if (r >= a.n_rows) return;
        const int64_t rr = a.row_of_pos ? (int64_t)a.row_of_pos[r] : r;       // k_strip numbers the rows in sorted order
        const double* __restrict__ src = a.rowpart + r * a.n_tiles * NQ;
        for (int q = 0; q < NQ; ++q) {
            double s = 0.0;
            for (int t = lane; t < a.n_tiles; t += 32) s += src[(int64_t)t * NQ + q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) a.rowsum[rr * NQ + q] = s;
        }
    } else {   // normalisation control points: per row, sum of the tile partials
        const int64_t i = (int64_t)(blk - nb_ctrl - a.nb_stage1_rows) * blockDim.x + threadIdx.x;
        if (i >= a.n_rows * a.norm_stride) return;
        const int64_t r = i / a.norm_stride;
        const int k = (int)(i - r * a.norm_stride);
        double s = 0.0;
        for (int t = 0; t < a.n_tiles; ++t) s += a.normpart[(r * a.n_tiles + t) * a.norm_stride + k];
        a.normrow[i] = s;
    }
}

// One thread per parameter of params_all that the kernels differentiate.
//   d loss / d ctrl[k]    = -coef * sum wr*a*B_k                 
//   d loss / d shift[key] = +coef * PV/dx * sum wr*a*S'          (d/d shift of S(x - shift) = -S')
//   d loss / d stretch[k] = -coef * sum wr*S
//   F[key]                = (PV/dx)^2 * sum ivar*(a*S')^2        (model.py:893-897; no coefficient)
__global__ void k_reduce_stage2(const ReduceParams* __restrict__ arr, const int* __restrict__ off, int n) {
    const int pi = find_plan(off, n, blockIdx.x);
    const ReduceParams& a = arr[pi];
    const int NQ = a.nq;
    int64_t i = (int64_t)(blockIdx.x - off[pi]) * blockDim.x + threadIdx.x;
    for (int c = 0; c < a.n_comp + a.has_norm; ++c) {
        const bool is_norm = c == a.n_comp;
        const int M = is_norm ? a.n_norm_keys * a.norm_knots : a.comp[c].n_knots;
        const int Mp = (M + 31) & ~31;              // every segment is padded to whole warps
        if (i < Mp) {
            if (is_norm) {
                if (i >= M) return;   // control point k of key: sum over the key's rows, in row order
                const int key = (int)(i / a.norm_knots), k = (int)(i % a.norm_knots);
                double s = 0.0;
                for (int e = a.norm_rowptr[key]; e < a.norm_rowptr[key + 1]; ++e)
                    s += a.normrow[(int64_t)a.norm_rows[e] * a.norm_stride + k];
                a.grad_all[a.norm_ctrl_off + i] = -a.coef * s;
                return;
            }
            // control point i of template c: the tiles whose windows cover it, in tile order
            const int NT = a.n_tiles;
            const int* __restrict__ rng = a.tile_rng + 2 * c * NT;
            const double* __restrict__ tp = a.tilepart + (int64_t)c * NT * a.max_knots + i;
            // the warp's 32 knots share the scan of the tiles: every lane tests one tile per round, the
            // tiles whose range meets the warp's knots are then added in tile order (the segments of this
            // kernel's index space are padded to 32, so a warp never straddles two of them)
            const int lane = threadIdx.x & 31;
            const int i0 = (int)i & ~31;
            double s = 0.0;
            for (int t0 = 0; t0 < NT; t0 += 32) {
                int2 r = make_int2(0, 0);
                if (t0 + lane < NT) r = *reinterpret_cast<const int2*>(rng + 2 * (t0 + lane));
                unsigned m = __ballot_sync(0xffffffffu, r.y > r.x && r.x < i0 + 32 && r.y > i0);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int rx = __shfl_sync(0xffffffffu, r.x, b), ry = __shfl_sync(0xffffffffu, r.y, b);
                    if ((int)i >= rx && (int)i < ry && i < M) s += tp[(int64_t)(t0 + b) * a.max_knots];
                }
            }
            if (i >= M) return;
            a.grad_all[a.ctrl_off[c] + i] = -a.coef * s;
            return;
        }
        i -= Mp;
        if (i < ((a.n_shift[c] + 31) & ~31)) {
            if (i >= a.n_shift[c]) return;
            double gs = 0.0, fs = 0.0;
            const int qg = a.rq[c][0], qf = a.rq[c][2];
            for (int e = a.shift_rowptr[c][i]; e < a.shift_rowptr[c][i + 1]; ++e) {
                const int64_t r = a.shift_rows[c][e];
                if (qg >= 0) gs += a.rowsum[r * NQ + qg];
                if (qf >= 0) fs += a.rowsum[r * NQ + qf];
            }
            const double sc = a.comp[c].p_val * a.comp[c].inv_dx;
            a.grad_all[a.shift_off[c] + i] = a.coef * sc * gs;
            if (!is_norm) a.fisher_all[a.shift_off[c] + i] = sc * sc * fs;
            return;
        }
        i -= (a.n_shift[c] + 31) & ~31;
        if (i < ((a.n_stretch[c] + 31) & ~31)) {
            if (i >= a.n_stretch[c]) return;
            double gs = 0.0;
            const int qs = a.rq[c][1];
            for (int e = a.stretch_rowptr[c][i]; e < a.stretch_rowptr[c][i + 1]; ++e)
                if (qs >= 0) gs += a.rowsum[(int64_t)a.stretch_rows[c][e] * NQ + qs];
            a.grad_all[a.stretch_off[c] + i] = -a.coef * gs;
            return;
        }
        i -= (a.n_stretch[c] + 31) & ~31;
    }
}

// Block 0: loss = 0.5 * coef * sum_tasks losspart, fixed-shape tree.  Other blocks: gather.
__global__ void k_finish(const ReduceParams* __restrict__ arr, const int* __restrict__ off, int n, double* out_override) {
    const int pi = find_plan(off, n, blockIdx.x);
    const ReduceParams& a = arr[pi];
    const int blk = blockIdx.x - off[pi];
    double* out = out_override ? out_override : a.out;
    if (blk == 0) {
        __shared__ double sh[1024];
        double s = 0.0;
        for (int64_t r = threadIdx.x; r < a.n_tasks; r += blockDim.x) s += a.losspart[r];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        // an evaluation that ran off the knot grid or out of its window has no value: NaN, so that a
        // device-side line search rejects the trial (host callers read the status word and get the error)
        if (threadIdx.x == 0) out[0] = a.status[0] ? __longlong_as_double(0x7ff8000000000000LL) : 0.5 * a.coef * sh[0];
    } else {
        const int64_t i = (int64_t)(blk - 1) * blockDim.x + threadIdx.x;
        if (i < a.n_fit) out[1 + i] = a.grad_all[a.fit_map[i]];
    }
}

// status words of the plans of a batch -> one flat array (one copy instead of one per plan)
__global__ void k_gather_status(int* const* __restrict__ ptrs, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[2 * i] = ptrs[i][0]; out[2 * i + 1] = ptrs[i][1]; }
}

}  // namespace jb
