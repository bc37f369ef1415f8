// Kernels of libjabble_b200 (sm_100a).  See DESIGN.md for the decomposition.
//
//   k_scatter_params   p_fit -> params_all[fit_map]
//   k_forward          model(p, x, meta) per pixel                       (Model.__call__)
//   k_fused<NC,PV>     warp->knot, basis weights, gather, residual, chi-square, adjoints of the
//                      control points (deterministic, no float atomics), of shift/stretch, and the
//                      per-row Fisher sums -- ONE pass over x, y, ivar, mask
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

__global__ void k_scatter_params(ScatterParams a) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < a.n_fit) a.params[a.fit_map[i]] = a.p_fit[i];
}

// ------------------------------------------------------------------------------------------
struct ForwardParams {
    const double* x;       // [E][ppad], canonical orientation
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
// Fused loss + gradient + Fisher.
//
// Work item ("task") = one warp, one tile of kTile pixel columns, one group of <= 32 rows.
//
// Data movement: lane 0 issues, per row, four 1-D bulk async copies (TMA engine; x, y, ivar: 1 KB
// each, mask: 128 B) into a kStages-deep ring in shared memory, completion counted on an mbarrier,
// so kStages-1 rows are in flight per warp independent of occupancy.  The control points the task
// can touch (a window of <= kWinCap knots per component) are copied to shared memory once.
//
// Arithmetic: lane l owns pixel columns [4l, 4l+4) of the tile for every row, so its pixels walk
// through the knot intervals monotonically and the control-point adjoint of the current interval
// lives in registers (a sliding window of p_val+1 accumulators).  A knot leaves the window when
// the lane moves on; its sum is then added, with a plain shared-memory read-modify-write, to the
// warp's gradient window.  Lanes whose knot ranges could overlap never share an accumulator:
// there are kClasses copies of the window and lane l uses copy l % kClasses; lanes l and l+4 are
// 13 pixels apart, and the plan marks (row, tile) pairs where that is not enough knots ("dense")
// -- those rows run lane by lane.  Every sum has one fixed order: no float atomics,
// bit-reproducible.
struct FusedParams {
    const double* x; const double* y; const double* ivar; const uint8_t* mask;   // [E][ppad]
    const double* txmin; const double* txmax;   // [E][n_tiles] extent of the unmasked pixels of a tile
    const uint8_t* dense;                       // [E][n_tiles] 1 = lanes of one class may share knots
    const int* row_perm;                        // rows sorted so that a group spans few knots
    int64_t n_rows, ppad;
    int n_tiles, rows_per_group, n_groups;
    int64_t n_tasks;
    int n_comp;
    CompDev comp[kMaxComp];
    double* part;        // [n_comp][n_tasks][kWinCap] per-task gradient windows
    int* part_base;      // [n_comp][n_tasks] first knot of the window (INT_MAX: empty)
    double* rowpart;     // [E][n_tiles][NQP]   NQP = 4 (1 component) or 8 (2 components)
    int* status;         // [0] error bits  [1] rows that took the serial path
};

template <int NC> struct RowQ { static constexpr int NQP = (NC == 1) ? 4 : 8; };

template <int NC>
__host__ __device__ constexpr size_t fused_warp_smem() {
    return (size_t)NC * kClasses * kWinCap * 8 + (size_t)NC * kWinCap * 8 + (size_t)kStages * kStageBytes + kStages * 8;
}

template <int PV>
struct Run {
    double c[PV + 1];     // control points klo .. klo+PV
    double acc[PV + 1];   // adjoint sums of those knots accumulated by this lane
    int klo;
    bool valid;
};

template <int PV>
__device__ __forceinline__ void run_flush(Run<PV>& R, double* accw, int base) {
    if (R.valid) {
#pragma unroll
        for (int j = 0; j <= PV; ++j) accw[R.klo + j - base] += R.acc[j];
    }
    R.valid = false;
}

// ctrlw: this component's control-point window in shared memory (index = knot - base)
template <int PV>
__device__ __forceinline__ void run_advance(Run<PV>& R, int klo_new, const double* ctrlw, double* accw, int base) {
    if (R.valid && klo_new == R.klo + 1) {
        accw[R.klo - base] += R.acc[0];
#pragma unroll
        for (int j = 0; j < PV; ++j) { R.c[j] = R.c[j + 1]; R.acc[j] = R.acc[j + 1]; }
        R.c[PV] = ctrlw[klo_new + PV - base];
        R.acc[PV] = 0.0;
    } else {
        run_flush<PV>(R, accw, base);
#pragma unroll
        for (int j = 0; j <= PV; ++j) { R.c[j] = ctrlw[klo_new + j - base]; R.acc[j] = 0.0; }
        R.valid = true;
    }
    R.klo = klo_new;
}

template <int NC, int PV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2) k_fused(const __grid_constant__ FusedParams a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (task >= a.n_tasks) return;
    const int tile = (int)(task % a.n_tiles);
    const int group = (int)(task / a.n_tiles);
    const int row0 = group * a.rows_per_group;
    const int nrow = min(a.rows_per_group, (int)a.n_rows - row0);
    constexpr int NQP = RowQ<NC>::NQP;

    unsigned char* wbase = smem_raw + (size_t)warp * fused_warp_smem<NC>();
    double* acc_warp = reinterpret_cast<double*>(wbase);                     // [NC][kClasses][kWinCap]
    double* ctrl_warp = acc_warp + NC * kClasses * kWinCap;                  // [NC][kWinCap]
    unsigned char* ring = reinterpret_cast<unsigned char*>(ctrl_warp + NC * kWinCap);   // [kStages][kStageBytes]
    const uint32_t bar0 = smem_u32(ring + (size_t)kStages * kStageBytes);    // [kStages] mbarriers

    // ---- 0. per-lane row record: lane i keeps row i of the group ------------------------------
    int my_row = 0;
    double my_sh[NC], my_st[NC];
    uint32_t my_dense = 0;
    int lo[NC], hi[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { lo[c] = INT_MAX; hi[c] = INT_MIN; my_sh[c] = 0.0; my_st[c] = 1.0; }
    if (lane < nrow) {
        const int r = a.row_perm[row0 + lane];
        my_row = r;
        my_dense = a.dense[(int64_t)r * a.n_tiles + tile];
        const double xmin = a.txmin[(int64_t)r * a.n_tiles + tile];
        const double xmax = a.txmax[(int64_t)r * a.n_tiles + tile];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const CompDev& cd = a.comp[c];
            if (cd.shift) my_sh[c] = cd.shift[cd.shift_key[r]];
            if (cd.stretch) my_st[c] = cd.stretch[cd.stretch_key[r]];
            if (xmin <= xmax) {
                double g; int I0, I1;
                knot_coord<PV>(xmin, my_sh[c], cd.xp0, cd.inv_dx, g, I0);
                knot_coord<PV>(xmax, my_sh[c], cd.xp0, cd.inv_dx, g, I1);
                // guard the magic-number trick: coordinates must be far inside +-2^31
                const double t0 = ((xmin - my_sh[c]) - cd.xp0) * cd.inv_dx, t1 = ((xmax - my_sh[c]) - cd.xp0) * cd.inv_dx;
                if (!(t0 > -1.0e9 && t1 < 1.0e9)) { I0 = -(1 << 30); I1 = (1 << 30); }
                lo[c] = knot_lo<PV>(I0);
                hi[c] = knot_lo<PV>(I1) + PV;
            }
        }
    }
    // ---- 1. knot window of the task ---------------------------------------------------------
    bool bad_range = false, bad_window = false, empty = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lo[c] > hi[c]) empty = true;
        else {
            if (lo[c] < 0 || hi[c] > a.comp[c].n_knots - 1) bad_range = true;
            if (hi[c] - lo[c] + 1 > kWinCap) bad_window = true;
        }
    }
    const bool skip = empty || bad_range || bad_window;
    if (lane == 0) {
        if (!empty && bad_range) atomicOr(a.status, kStatOutOfRange);
        else if (!empty && bad_window) atomicOr(a.status, kStatWindow);
#pragma unroll
        for (int c = 0; c < NC; ++c) a.part_base[(int64_t)c * a.n_tasks + task] = skip ? INT_MAX : lo[c];
    }
    if (skip) {   // nothing unmasked here (or an error): the row sums of this tile are zero
        for (int i = lane; i < nrow * NQP; i += 32) {
            const int r = a.row_perm[row0 + i / NQP];
            a.rowpart[((int64_t)r * a.n_tiles + tile) * NQP + i % NQP] = 0.0;
        }
        return;
    }

    // ---- 2. barriers, first copies, accumulators, control-point window ------------------------
    const int64_t col = (int64_t)tile * kTile;
    auto issue_row = [&](int i, int r) {      // lane 0 only
        const int s = i % kStages;
        const uint32_t bar = bar0 + 8 * s;
        const uint32_t dst = smem_u32(ring + (size_t)s * kStageBytes);
        const int64_t off = (int64_t)r * a.ppad + col;
        mbar_expect_tx(bar, kStageBytes);
        tma_load_1d(dst, a.x + off, kTile * 8, bar);
        tma_load_1d(dst + kTile * 8, a.y + off, kTile * 8, bar);
        tma_load_1d(dst + kTile * 16, a.ivar + off, kTile * 8, bar);
        tma_load_1d(dst + kTile * 24, a.mask + off, kTile, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
        const int r = __shfl_sync(0xffffffffu, my_row, i);
        if (lane == 0 && i < nrow) issue_row(i, r);
    }
    for (int i = lane; i < NC * kClasses * kWinCap; i += 32) acc_warp[i] = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int n = hi[c] - lo[c] + 1;
        for (int k = lane; k < n; k += 32) ctrl_warp[c * kWinCap + k] = __ldg(a.comp[c].ctrl + lo[c] + k);
    }
    __syncwarp();

    const int cls = lane & (kClasses - 1);
    Run<PV> run[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) run[c].valid = false;
    int serial_rows = 0;

    // ---- 3. stream the rows of the group ------------------------------------------------------
    for (int i = 0; i < nrow; ++i) {
        const int s = i % kStages;
        const int r = __shfl_sync(0xffffffffu, my_row, i);
        const bool dense = __shfl_sync(0xffffffffu, my_dense, i) != 0;
        double sh[NC], st[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            sh[c] = __shfl_sync(0xffffffffu, my_sh[c], i);
            st[c] = __shfl_sync(0xffffffffu, my_st[c], i);
        }
        mbar_wait(bar0 + 8 * s, (uint32_t)((i / kStages) & 1));

        const unsigned char* stg = ring + (size_t)s * kStageBytes;
        const double2* px = reinterpret_cast<const double2*>(stg) + lane * 2;
        const double2* py = reinterpret_cast<const double2*>(stg + kTile * 8) + lane * 2;
        const double2* pi = reinterpret_cast<const double2*>(stg + kTile * 16) + lane * 2;
        const double2 x01 = px[0], x23 = px[1], y01 = py[0], y23 = py[1], i01 = pi[0], i23 = pi[1];
        const uint32_t mk = reinterpret_cast<const uint32_t*>(stg + kTile * 24)[lane];
        const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
        const double yv[4] = {y01.x, y01.y, y23.x, y23.y};
        const double iv[4] = {i01.x, i01.y, i23.x, i23.y};

        double sums[NQP];
#pragma unroll
        for (int q = 0; q < NQP; ++q) sums[q] = 0.0;
        if (dense) ++serial_rows;

        const int nturn = dense ? 32 / kClasses : 1;
        for (int turn = 0; turn < nturn; ++turn) {
            if (!dense || (lane / kClasses) == turn) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((mk >> (8 * j)) & 0xffu) continue;       // masked pixel: a hole
                    double w[NC][PV + 1], S[NC], dS[NC];
                    double m = 0.0;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        double g; int I;
                        knot_coord<PV>(xv[j], sh[c], a.comp[c].xp0, a.comp[c].inv_dx, g, I);
                        const int kl = knot_lo<PV>(I);
                        double* accw = acc_warp + ((size_t)c * kClasses + cls) * kWinCap;
                        if (!run[c].valid || kl != run[c].klo)
                            run_advance<PV>(run[c], kl, ctrl_warp + c * kWinCap, accw, lo[c]);
                        double dw[PV];
                        Basis<PV>::eval(g, w[c], dw);
                        double sv = run[c].c[0] * w[c][0];
                        double d = (run[c].c[1] - run[c].c[0]) * dw[0];
#pragma unroll
                        for (int k = 1; k <= PV; ++k) sv = __fma_rn(run[c].c[k], w[c][k], sv);
#pragma unroll
                        for (int k = 1; k < PV; ++k) d = __fma_rn(run[c].c[k + 1] - run[c].c[k], dw[k], d);
                        S[c] = sv; dS[c] = d;
                        m = __fma_rn(st[c], sv, m);               // StretchingModel / AdditiveModel
                    }
                    const double res = yv[j] - m;                 // ChiSquare, loss.py:169-171
                    const double wr = iv[j] * res;
                    sums[0] = __fma_rn(wr, res, sums[0]);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const double q = wr * st[c];
#pragma unroll
                        for (int k = 0; k <= PV; ++k) run[c].acc[k] = __fma_rn(q, w[c][k], run[c].acc[k]);
                        sums[1 + c * kRowQ + 0] = __fma_rn(q, dS[c], sums[1 + c * kRowQ + 0]);              // d/d shift
                        sums[1 + c * kRowQ + 1] = __fma_rn(wr, S[c], sums[1 + c * kRowQ + 1]);              // d/d stretch
                        sums[1 + c * kRowQ + 2] = __fma_rn(iv[j] * dS[c], dS[c], sums[1 + c * kRowQ + 2]);  // Fisher
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    run_flush<PV>(run[c], acc_warp + ((size_t)c * kClasses + cls) * kWinCap, lo[c]);
            }
            __syncwarp();
        }

        // per-row sums of this tile (fixed order); lane L ends up with value q = L >> (NQP == 8 ? 2 : 3)
        const double tot = warp_sum_multi<NQP>(sums, lane);
        constexpr int kShift = (NQP == 8) ? 2 : 3;
        if ((lane & ((1 << kShift) - 1)) == 0) {
            const int q = lane >> kShift;
            double s2 = 1.0;                       // Fisher sums carry the squared stretch of the row
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (q == 1 + c * kRowQ + 2) s2 = st[c] * st[c];
            const double v = tot * s2;
            a.rowpart[((int64_t)r * a.n_tiles + tile) * NQP + q] = v;
        }
        // every lane has consumed stage s (the sums above depend on all of its data): refill it
        __syncwarp();
        if (i + kStages < nrow) {
            const int rn = __shfl_sync(0xffffffffu, my_row, i + kStages);
            if (lane == 0) issue_row(i + kStages, rn);
        }
    }

    // ---- 4. fold the accumulator copies in a fixed order and publish the window ----------------
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double* ac = acc_warp + (size_t)c * kClasses * kWinCap;
        double* dst = a.part + ((int64_t)c * a.n_tasks + task) * kWinCap;
        for (int k = lane; k < kWinCap; k += 32) {
            double sacc = ac[k];
#pragma unroll
            for (int q = 1; q < kClasses; ++q) sacc += ac[q * kWinCap + k];
            dst[k] = sacc;
        }
    }
    if (lane == 0 && serial_rows) atomicAdd(a.status + 1, serial_rows);
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
    int n_tiles, n_groups;
    const double* part; const int* part_base; const double* rowpart;
    double* tmp;        // [n_comp][n_groups][max_knots]
    double* rowsum;     // [E][NQ]
    int64_t max_knots;
    double coef;
    double* grad_all; double* fisher_all;
    const int64_t* fit_map; int64_t n_fit;
    double* out;        // [1 + n_fit]
    int* status;
};

// blocks [0, nb_ctrl): tmp[c][g][k] = sum over tiles of the group, in tile order, of the windows
// covering knot k.  blocks [nb_ctrl, ...): rowsum[r][q] = sum over tiles of rowpart.
__global__ void k_reduce_stage1(ReduceParams a, int nb_knot, int nb_ctrl) {
    const int NQ = a.n_comp == 1 ? 4 : 8;
    if ((int)blockIdx.x < nb_ctrl) {
        const int kb = blockIdx.x % nb_knot;
        const int g = (blockIdx.x / nb_knot) % a.n_groups;
        const int c = blockIdx.x / (nb_knot * a.n_groups);
        const int k = kb * blockDim.x + threadIdx.x;
        if (k >= a.comp[c].n_knots) return;
        double s = 0.0;
        for (int t = 0; t < a.n_tiles; ++t) {
            const int64_t task = (int64_t)g * a.n_tiles + t;
            const int base = a.part_base[(int64_t)c * a.n_tasks + task];
            const int w = k - base;
            if (base != INT_MAX && w >= 0 && w < kWinCap)
                s += a.part[((int64_t)c * a.n_tasks + task) * kWinCap + w];
        }
        a.tmp[((int64_t)c * a.n_groups + g) * a.max_knots + k] = s;
    } else {
        const int64_t i = (int64_t)(blockIdx.x - nb_ctrl) * blockDim.x + threadIdx.x;
        if (i >= a.n_rows * NQ) return;
        const int64_t r = i / NQ;
        const int q = (int)(i - r * NQ);
        double s = 0.0;
        for (int t = 0; t < a.n_tiles; ++t) s += a.rowpart[(r * a.n_tiles + t) * NQ + q];
        a.rowsum[i] = s;
    }
}

// One thread per parameter of params_all that the kernels differentiate.
//   d loss / d ctrl[k]    = -coef * sum wr*a*B_k                 
//   d loss / d shift[key] = +coef * PV/dx * sum wr*a*S'          (d/d shift of S(x - shift) = -S')
//   d loss / d stretch[k] = -coef * sum wr*S
//   F[key]                = (PV/dx)^2 * sum ivar*(a*S')^2        (model.py:893-897; no coefficient)
__global__ void k_reduce_stage2(ReduceParams a) {
    const int NQ = a.n_comp == 1 ? 4 : 8;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int c = 0; c < a.n_comp; ++c) {
        const int M = a.comp[c].n_knots;
        if (i < M) {
            double s = 0.0;
            for (int g = 0; g < a.n_groups; ++g) s += a.tmp[((int64_t)c * a.n_groups + g) * a.max_knots + i];
            a.grad_all[a.ctrl_off[c] + i] = -a.coef * s;
            return;
        }
        i -= M;
        if (i < a.n_shift[c]) {
            double gs = 0.0, fs = 0.0;
            for (int e = a.shift_rowptr[c][i]; e < a.shift_rowptr[c][i + 1]; ++e) {
                const int64_t r = a.shift_rows[c][e];
                gs += a.rowsum[r * NQ + 1 + c * kRowQ + 0];
                fs += a.rowsum[r * NQ + 1 + c * kRowQ + 2];
            }
            const double sc = a.comp[c].p_val * a.comp[c].inv_dx;
            a.grad_all[a.shift_off[c] + i] = a.coef * sc * gs;
            a.fisher_all[a.shift_off[c] + i] = sc * sc * fs;
            return;
        }
        i -= a.n_shift[c];
        if (i < a.n_stretch[c]) {
            double gs = 0.0;
            for (int e = a.stretch_rowptr[c][i]; e < a.stretch_rowptr[c][i + 1]; ++e)
                gs += a.rowsum[(int64_t)a.stretch_rows[c][e] * NQ + 1 + c * kRowQ + 1];
            a.grad_all[a.stretch_off[c] + i] = -a.coef * gs;
            return;
        }
        i -= a.n_stretch[c];
    }
}

// Block 0: loss = 0.5 * coef * sum_rows rowsum[r][0], fixed-shape tree.  Other blocks: gather.
__global__ void k_finish(ReduceParams a) {
    const int NQ = a.n_comp == 1 ? 4 : 8;
    if (blockIdx.x == 0) {
        __shared__ double sh[1024];
        double s = 0.0;
        for (int64_t r = threadIdx.x; r < a.n_rows; r += blockDim.x) s += a.rowsum[r * NQ];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) a.out[0] = 0.5 * a.coef * sh[0];
    } else {
        const int64_t i = (int64_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x;
        if (i < a.n_fit) a.out[1 + i] = a.grad_all[a.fit_map[i]];
    }
}

}  // namespace jb
