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
// Work item ("task") = one warp, one tile of kTile pixel columns, one group of rows.  Lane l owns
// pixel columns [4l, 4l+4) of the tile for every row of the group, so its pixels walk through the
// knot intervals monotonically and the control-point adjoint of the current interval lives in
// registers (a sliding window of p_val+1 accumulators).  A knot leaves the window when the lane
// moves on; its sum is then added, with a plain shared-memory read-modify-write, to the warp's
// on-chip gradient window.  Lanes whose knot ranges could overlap never share an accumulator:
// there are kClasses copies of the window and lane l uses copy l % kClasses; the ranges of lanes
// l and l+4 are checked to be disjoint per row (13 pixels apart), else the row takes the serial
// path.  Every sum therefore has one fixed order: no float atomics, bit-reproducible.
struct FusedParams {
    const double* x; const double* y; const double* ivar; const uint8_t* mask;   // [E][ppad]
    const double* txmin; const double* txmax;   // [E][n_tiles] extent of the unmasked pixels of a tile
    const int* row_perm;                        // rows sorted so that a group spans few knots
    int64_t n_rows, ppad;
    int n_tiles, rows_per_group, n_groups;
    int64_t n_tasks;
    int n_comp;
    CompDev comp[kMaxComp];
    double* part;        // [n_comp][n_tasks][kWinCap] per-task gradient windows
    int* part_base;      // [n_comp][n_tasks] first knot of the window (INT_MAX: empty)
    double* rowpart;     // [E][n_tiles][1 + kRowQ*n_comp]
    int* status;         // [0] error bits  [1] rows that took the serial path
};

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

template <int PV>
__device__ __forceinline__ void run_advance(Run<PV>& R, int klo_new, const double* __restrict__ ctrl,
                                            double* accw, int base) {
    if (R.valid && klo_new == R.klo + 1) {
        accw[R.klo - base] += R.acc[0];
#pragma unroll
        for (int j = 0; j < PV; ++j) { R.c[j] = R.c[j + 1]; R.acc[j] = R.acc[j + 1]; }
        R.c[PV] = __ldg(ctrl + klo_new + PV);
        R.acc[PV] = 0.0;
    } else {
        run_flush<PV>(R, accw, base);
#pragma unroll
        for (int j = 0; j <= PV; ++j) { R.c[j] = __ldg(ctrl + klo_new + j); R.acc[j] = 0.0; }
        R.valid = true;
    }
    R.klo = klo_new;
}

template <int NC, int PV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3) k_fused(const __grid_constant__ FusedParams a) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (task >= a.n_tasks) return;
    const int tile = (int)(task % a.n_tiles);
    const int group = (int)(task / a.n_tiles);
    const int row0 = group * a.rows_per_group;
    const int nrow = min(a.rows_per_group, (int)a.n_rows - row0);
    constexpr int NQ = 1 + kRowQ * NC;

    // this warp's accumulator copies: [NC][kClasses][kWinCap]
    double* acc_warp = smem + (size_t)warp * NC * kClasses * kWinCap;

    // ---- 1. knot window of the task: min/max interval over the group's rows -------------------
    int lo[NC], hi[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { lo[c] = INT_MAX; hi[c] = INT_MIN; }
    for (int i = lane; i < nrow; i += 32) {
        const int r = a.row_perm[row0 + i];
        const double xmin = a.txmin[(int64_t)r * a.n_tiles + tile];
        const double xmax = a.txmax[(int64_t)r * a.n_tiles + tile];
        if (xmin <= xmax) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const CompDev& cd = a.comp[c];
                const double sh = cd.shift ? cd.shift[cd.shift_key[r]] : 0.0;
                double g; int I0, I1;
                knot_coord<PV>(xmin, sh, cd.xp0, cd.inv_dx, g, I0);
                knot_coord<PV>(xmax, sh, cd.xp0, cd.inv_dx, g, I1);
                // guard the magic-number trick: coordinates must be far inside +-2^31
                const double t0 = ((xmin - sh) - cd.xp0) * cd.inv_dx, t1 = ((xmax - sh) - cd.xp0) * cd.inv_dx;
                if (!(t0 > -1.0e9 && t1 < 1.0e9)) { I0 = -(1 << 30); I1 = (1 << 30); }
                lo[c] = min(lo[c], knot_lo<PV>(I0));
                hi[c] = max(hi[c], knot_lo<PV>(I1) + PV);
            }
        }
    }
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
        for (int i = lane; i < nrow * NQ; i += 32) {
            const int r = a.row_perm[row0 + i / NQ];
            a.rowpart[((int64_t)r * a.n_tiles + tile) * NQ + i % NQ] = 0.0;
        }
        return;
    }

    // ---- 2. zero the accumulator copies -------------------------------------------------------
    for (int i = lane; i < NC * kClasses * kWinCap; i += 32) acc_warp[i] = 0.0;
    __syncwarp();

    const int cls = lane & (kClasses - 1);
    Run<PV> run[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) run[c].valid = false;

    // ---- 3. stream the rows of the group ------------------------------------------------------
    const int64_t col = (int64_t)tile * kTile + lane * kPixPerLane;
    double xv[4], yv[4], iv[4];
    uint32_t mk;
    {
        const int64_t off = (int64_t)a.row_perm[row0] * a.ppad + col;
        ldg256(a.x + off, xv); ldg256(a.y + off, yv); ldg256(a.ivar + off, iv);
        mk = ldg32_stream(a.mask + off);
    }
    int serial_rows = 0;
    for (int i = 0; i < nrow; ++i) {
        const int r = a.row_perm[row0 + i];
        // software prefetch of the next row of the group
        double xn[4], yn[4], in_[4];
        uint32_t mn = 0;
        if (i + 1 < nrow) {
            const int64_t off = (int64_t)a.row_perm[row0 + i + 1] * a.ppad + col;
            ldg256(a.x + off, xn); ldg256(a.y + off, yn); ldg256(a.ivar + off, in_);
            mn = ldg32_stream(a.mask + off);
        }

        // per-row parameters
        double sh[NC], st[NC];
        const double* ctrl[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const CompDev& cd = a.comp[c];
            sh[c] = cd.shift ? cd.shift[cd.shift_key[r]] : 0.0;
            st[c] = cd.stretch ? cd.stretch[cd.stretch_key[r]] : 1.0;
            ctrl[c] = cd.ctrl;
        }

        // knot coordinates of this lane's pixels and its knot range in this row
        double g[NC][4];
        int kl[NC][4];
        int rlo[NC], rhi[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            rlo[c] = INT_MAX; rhi[c] = INT_MIN;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int I;
                knot_coord<PV>(xv[j], sh[c], a.comp[c].xp0, a.comp[c].inv_dx, g[c][j], I);
                kl[c][j] = knot_lo<PV>(I);
                if (!((mk >> (8 * j)) & 0xffu)) { rlo[c] = min(rlo[c], kl[c][j]); rhi[c] = max(rhi[c], kl[c][j] + PV); }
            }
        }
        // lanes l-4, l-8, ... share this lane's accumulator copy: their knot ranges must lie
        // strictly below this lane's range
        bool ok = true;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int v = rhi[c];
#pragma unroll
            for (int o = kClasses; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v = max(v, t);
            }
            int prev = __shfl_up_sync(0xffffffffu, v, kClasses);
            if (lane < kClasses) prev = INT_MIN;
            if (rlo[c] <= rhi[c] && rlo[c] <= prev) ok = false;
        }
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) ++serial_rows;

        double s_loss = 0.0, s_q[NC][kRowQ];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int q = 0; q < kRowQ; ++q) s_q[c][q] = 0.0;

        const int nturn = ok ? 1 : 32 / kClasses;
        for (int turn = 0; turn < nturn; ++turn) {
            if (ok || (lane / kClasses) == turn) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((mk >> (8 * j)) & 0xffu) continue;       // masked pixel: a hole
                    double w[NC][PV + 1], S[NC], dS[NC];
                    double m = 0.0;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        double* accw = acc_warp + ((size_t)c * kClasses + cls) * kWinCap;
                        if (!run[c].valid || kl[c][j] != run[c].klo)
                            run_advance<PV>(run[c], kl[c][j], ctrl[c], accw, lo[c]);
                        double dw[PV];
                        Basis<PV>::eval(g[c][j], w[c], dw);
                        double s = run[c].c[0] * w[c][0];
                        double d = (run[c].c[1] - run[c].c[0]) * dw[0];
#pragma unroll
                        for (int k = 1; k <= PV; ++k) s = __fma_rn(run[c].c[k], w[c][k], s);
#pragma unroll
                        for (int k = 1; k < PV; ++k) d = __fma_rn(run[c].c[k + 1] - run[c].c[k], dw[k], d);
                        S[c] = s; dS[c] = d;
                        m = __fma_rn(st[c], s, m);                // StretchingModel / AdditiveModel
                    }
                    const double res = yv[j] - m;                 // ChiSquare, loss.py:169-171
                    const double wr = iv[j] * res;
                    s_loss = __fma_rn(wr, res, s_loss);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const double q = wr * st[c];
#pragma unroll
                        for (int k = 0; k <= PV; ++k) run[c].acc[k] = __fma_rn(q, w[c][k], run[c].acc[k]);
                        s_q[c][0] = __fma_rn(q, dS[c], s_q[c][0]);              // -> d/d shift
                        s_q[c][1] = __fma_rn(wr, S[c], s_q[c][1]);              // -> d/d stretch
                        s_q[c][2] = __fma_rn(iv[j] * dS[c], dS[c], s_q[c][2]);  // -> Fisher
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    run_flush<PV>(run[c], acc_warp + ((size_t)c * kClasses + cls) * kWinCap, lo[c]);
            }
            __syncwarp();
        }

        // per-row sums of this tile (fixed butterfly order)
        s_loss = warp_sum(s_loss);
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int q = 0; q < kRowQ; ++q) s_q[c][q] = warp_sum(s_q[c][q]);
        if (lane == 0) {
            double* rp = a.rowpart + ((int64_t)r * a.n_tiles + tile) * NQ;
            rp[0] = s_loss;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                rp[1 + c * kRowQ + 0] = s_q[c][0];
                rp[1 + c * kRowQ + 1] = s_q[c][1];
                rp[1 + c * kRowQ + 2] = s_q[c][2] * st[c] * st[c];
            }
        }

#pragma unroll
        for (int j = 0; j < 4; ++j) { xv[j] = xn[j]; yv[j] = yn[j]; iv[j] = in_[j]; }
        mk = mn;
    }

    // ---- 4. fold the accumulator copies in a fixed order and publish the window ----------------
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double* ac = acc_warp + (size_t)c * kClasses * kWinCap;
        double* dst = a.part + ((int64_t)c * a.n_tasks + task) * kWinCap;
        for (int k = lane; k < kWinCap; k += 32) {
            double s = ac[k];
#pragma unroll
            for (int q = 1; q < kClasses; ++q) s += ac[q * kWinCap + k];
            dst[k] = s;
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
    const int NQ = 1 + kRowQ * a.n_comp;
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
    const int NQ = 1 + kRowQ * a.n_comp;
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
    const int NQ = 1 + kRowQ * a.n_comp;
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
