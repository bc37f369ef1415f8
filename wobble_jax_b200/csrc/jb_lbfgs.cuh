// jb_lbfgs: L-BFGS-B for many independent chunks at once, with every vector on the device.
// Included at the end of jb_plan.cu.
//
// The reference fits a chunk with scipy.optimize.fmin_l_bfgs_b(func, x0, fprime, **options)
// (jabble/model.py:226-232) and echelle orders in a multiprocessing.Pool, one optimiser per process
// (notebooks/02-51peg.py:66-78).  No bounds are ever passed, so what runs is the unconstrained
// branch of L-BFGS-B 3.0 (Zhu, Byrd, Lu, Nocedal 1997; Morales, Nocedal 2011): the quasi-Newton
// direction -H g of the limited-memory BFGS matrix with H0 = (s'y / y'y) I, then the
// More'-Thuente line search (MINPACK-2 dcsrch / dcstep, ftol 1e-3, gtol 0.9, xtol 0.1), the first
// step of a fit scaled to 1 / |d|, at most maxls trial points per search, a BFGS pair skipped when
// s'y <= eps * (-g'd * stp), the memory dropped and the iteration restarted from steepest descent when
// a search fails, and the two stopping tests  max|g| <= pgtol  and
// (f_k - f_k+1) / max(|f_k|, |f_k+1|, 1) <= factr * eps  -- restated here from the published
// algorithm with scipy's status codes, so that a caller gets the (x, f, d) it got from scipy.
//
// One round = ONE batched evaluation of every chunk still iterating (jb_batch: one launch per kernel
// for all of them) followed by ONE launch of k_lbfgs_advance, a block per chunk, which takes the new
// [loss, gradient] through that chunk's state machine (line search decision; on an accepted step the
// BFGS update, the two-loop recursion over the stored pairs, the next trial point) and leaves the next
// x in place.  The host reads back one int per chunk per round (who is done) and nothing else.
//
// The direction -H g comes from the compact representation of the inverse L-BFGS matrix (Byrd, Nocedal,
// Schnabel 1994): H = gamma I + [S gamma Y] [[R^-T (D + gamma Y'Y) R^-1, -R^-T], [-R^-1, 0]] [S gamma Y]',
// R = upper triangle of S'Y, so that a block reads the stored pairs in TWO passes of independent loads
// (all 2 m dot products S'g, Y'g at once -- fused with the dots S'y, Y'y that extend R and Y'Y by the new
// pair --, then the combination) instead of the 2 m dependent passes of the two-loop recursion; the m x m
// algebra in between is one thread's work.
//
// Differences from scipy that a caller can see: L-BFGS-B applies the compact form of B and a subspace
// step where this applies the compact form of H (the same matrix, different rounding), so iterates agree
// to rounding for the first iterations and then drift like any two builds of the same optimiser.
#include <cfloat>
#include <mutex>

namespace jb {

constexpr int kLbfgsMaxM = 32;
constexpr int kLbfgsThreads = 512;       // 128 registers per thread: the scalar state of the search lives in registers

enum { kLbFirst = 0, kLbSearch = 1, kLbDone = 2, kLbRepeat = 3 };
// scipy/optimize/_lbfgsb_py.py: status_messages / task_messages
enum { kLbTaskConvergence = 4, kLbTaskStop = 5, kLbTaskError = 7, kLbTaskAbnormal = 8 };
enum { kLbMsgNone = 0, kLbMsgPgtol = 401, kLbMsgFactr = 402, kLbMsgMaxfun = 502, kLbMsgMaxiter = 504, kLbMsgOutOfRange = 799 };

struct LbfgsState {
    double f, fold, stp, dnorm, gd, gdold, theta, sbgnrm, f_inside;
    double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;     // dcsrch
    long long iter, funcalls;
    int brackt, stage, col, head, ifun, iback, phase, task, msg, warnflag, n_oor, nskip, nrestart, pad_;
};
struct LbfgsReg { const long long* idx; long long n; double weight; double constant; int kind; int pad_; };   // kind 0: L2, 1: L1
struct LbfgsProblem {
    long long n;
    double *x, *out, *xold, *gold, *d, *z, *S, *Y;     // out = [f, g[n]] of the evaluation; S, Y: [m][n]
    LbfgsState* st;
    double* sy;                                         // [m][m] by slot: sy[i][j] = s_i'y_j (used for i not newer than j), then
                                                        // [m][m]: yy[i][j] = y_i'y_j
    const LbfgsReg* regs;
    int n_regs, pad_;
};
struct LbfgsOpts { int m, maxls; long long maxiter, maxfun; double tol, pgtol; };

struct LbfgsRed { double buf[2][32]; int par; };

// Sum / max over the block in one fixed order; every thread gets the result.  One barrier per call
// (two alternating buffers: the barrier of call k + 1 separates the reads of call k from the writes
// of call k + 2).
__device__ __forceinline__ double lb_sum(double v, LbfgsRed& r, int& par) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) r.buf[par][warp] = v;
    __syncthreads();
    double t = lane < (int)(blockDim.x >> 5) ? r.buf[par][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    par ^= 1;
    return t;
}
__device__ __forceinline__ double lb_max(double v, LbfgsRed& r, int& par) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) r.buf[par][warp] = v;
    __syncthreads();
    double t = lane < (int)(blockDim.x >> 5) ? r.buf[par][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    par ^= 1;
    return t;
}

// MINPACK-2 dcstep: the safeguarded cubic / quadratic step of the More'-Thuente search and the update
// of the interval of uncertainty [stx, sty].
__device__ inline void lb_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                                 double fp, double dp, int& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf;
    if (fp > fx) {                       // higher value: the minimum is bracketed
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
        const double stpc = stx + r * (stp - stx);
        const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = fabs(stpc - stx) < fabs(stpq - stx) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {             // lower value, derivatives of opposite sign: bracketed
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
        const double stpc = stp + r * (stx - stp);
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {    // lower value, same sign, the derivative shrinks
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
        double stpc;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else stpc = stp > stx ? stpmax : stpmin;
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = fabs(stpc - stp) < fabs(stpq - stp) ? stpc : stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {                             // lower value, same sign, the derivative does not shrink
        if (brackt) {
            const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            const double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
            stpf = stp + r * (sty - stp);
        } else {
            stpf = stp > stx ? stpmax : stpmin;
        }
    }
    if (fp > fx) { sty = stp; fy = fp; dy = dp; }
    else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

constexpr double kLbFtol = 1.0e-3, kLbGtol = 0.9, kLbXtol = 0.1, kLbStpMin = 0.0, kLbStpMax = 1.0e10;

__device__ inline void lb_dcsrch_start(LbfgsState& s, double f, double g) {
    s.brackt = 0; s.stage = 1;
    s.finit = f; s.ginit = g; s.gtest = kLbFtol * g;
    s.width = kLbStpMax - kLbStpMin; s.width1 = s.width / 0.5;
    s.stx = 0.0; s.fx = f; s.gx = g; s.sty = 0.0; s.fy = f; s.gy = g;
    s.stmin = 0.0; s.stmax = s.stp + 4.0 * s.stp;
}
// One More'-Thuente step at the trial point (f, g = slope along d).  true: the search is over
// (converged, or one of dcsrch's warnings -- L-BFGS-B treats both as "step accepted").
__device__ inline bool lb_dcsrch(LbfgsState& s, double f, double g) {
    double stp = s.stp;
    const double ftest = s.finit + stp * s.gtest;
    if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
    bool over = false;
    if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) over = true;               // rounding errors prevent progress
    if (s.brackt && s.stmax - s.stmin <= kLbXtol * s.stmax) over = true;           // xtol test satisfied
    if (stp == kLbStpMax && f <= ftest && g <= s.gtest) over = true;
    if (stp == kLbStpMin && (f > ftest || g >= s.gtest)) over = true;
    if (f <= ftest && fabs(g) <= kLbGtol * -s.ginit) over = true;                  // strong Wolfe conditions hold
    if (over) return true;
    if (s.stage == 1 && f <= s.fx && f > ftest) {
        // a lower value without sufficient decrease: step on the modified function psi
        const double fm = f - stp * s.gtest;
        double fxm = s.fx - s.stx * s.gtest, fym = s.fy - s.sty * s.gtest;
        const double gm = g - s.gtest;
        double gxm = s.gx - s.gtest, gym = s.gy - s.gtest;
        lb_dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * s.gtest; s.fy = fym + s.sty * s.gtest;
        s.gx = gxm + s.gtest; s.gy = gym + s.gtest;
    } else {
        lb_dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
    }
    if (s.brackt) {
        if (fabs(s.sty - s.stx) >= 0.66 * s.width1) stp = s.stx + 0.5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
    }
    if (s.brackt) { s.stmin = fmin(s.stx, s.sty); s.stmax = fmax(s.stx, s.sty); }
    else { s.stmin = stp + 1.1 * (stp - s.stx); s.stmax = stp + 4.0 * (stp - s.stx); }
    stp = fmax(stp, kLbStpMin);
    stp = fmin(stp, kLbStpMax);
    if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) || (s.brackt && s.stmax - s.stmin <= kLbXtol * s.stmax)) stp = s.stx;
    s.stp = stp;
    return false;
}

// One block per chunk that is still iterating (active[blockIdx.x] names it).
// status[b]: status words of the plan evaluated for block b (jb_plan: bit kStatOutOfRange = the point is
// off a knot grid, its loss is NaN, handled below; otherwise any bit = the evaluation is not valid --
// knot window overflow: the host re-sorts the plan's rows and the chunk repeats the evaluation next round).
__global__ void __launch_bounds__(kLbfgsThreads) k_lbfgs_advance(const LbfgsProblem* __restrict__ probs, const int* __restrict__ active,
                                                                 LbfgsOpts o, int* __restrict__ flags, int* const* __restrict__ status) {
    __shared__ LbfgsRed red;
    __shared__ double sy[kLbfgsMaxM * kLbfgsMaxM], yy[kLbfgsMaxM * kLbfgsMaxM];     // by slot, row stride m
    __shared__ double dots[4][kLbfgsMaxM];          // S'y, Y'y, S'g, Y'g by age (0 = oldest pair)
    __shared__ double coef[2][kLbfgsMaxM];          // what multiplies S_k and gamma Y_k in H g
    __shared__ double wred[kLbfgsThreads / 32][16];
    const int pid = active[blockIdx.x];
    const LbfgsProblem P = probs[pid];
    const int tid = threadIdx.x, T = blockDim.x;
    const int bits = status[blockIdx.x][0];
    if (bits && !(bits & kStatOutOfRange)) {       // (off the grid wins, as in read_status: the point is refused, not repeated)
        if (tid == 0) flags[pid] = kLbRepeat;
        return;
    }
    LbfgsState s = *P.st;                  // every thread carries the (identical) scalar state; thread 0 stores it
    if (s.phase == kLbDone) return;
    for (int i = tid; i < o.m * o.m; i += T) { sy[i] = P.sy[i]; yy[i] = P.sy[o.m * o.m + i]; }
    __syncthreads();
    const long long n = P.n;
    int par = 0;
    double* g = P.out + 1;
    const int m = o.m;

    auto finish = [&](int task, int msg, int warn) {
        s.phase = kLbDone; s.task = task; s.msg = msg; s.warnflag = warn;
    };

    // ---- the evaluation: [loss, gradient] of the data term, then the regularisers -------------------
    double f = P.out[0];
    if (!isfinite(f)) {
        // the trial point carried an unmasked pixel off a knot grid (k_finish wrote NaN).  The reference's
        // gather clamps there and scipy sees a large finite loss; here the search is told the same: a
        // loss far above the last valid one, no slope.  At the starting point it is the caller's error.
        if (s.phase == kLbFirst) {
            finish(kLbTaskError, kLbMsgOutOfRange, 2);
            if (tid == 0) { *P.st = s; flags[pid] = kLbDone; }
            return;
        }
        f = 1.0e8 * fmax(fabs(s.f_inside), 1.0);
        for (long long i = tid; i < n; i += T) g[i] = 0.0;
        s.n_oor++;
    } else {
        s.f_inside = f;
    }
    if (P.n_regs) __syncthreads();
    for (int r = 0; r < P.n_regs; ++r) {
        const LbfgsReg R = P.regs[r];
        double part = 0.0;
        for (long long j = tid; j < R.n; j += T) {
            const long long i = R.idx[j];
            const double dl = P.x[i] - R.constant;
            if (R.kind == 0) { part += dl * dl; g[i] += R.weight * dl; }
            else { part += fabs(dl); g[i] += R.weight * (dl > 0.0 ? 1.0 : (dl < 0.0 ? -1.0 : 0.0)); }
        }
        const double tot = lb_sum(part, red, par);
        f += R.kind == 0 ? R.weight * 0.5 * tot : R.weight * tot;
    }
    if (P.n_regs) __syncthreads();
    s.funcalls++;

    bool start = false;          // begin a new iteration (direction + first trial point) from (x, f, g)
    int new_slot = -1;           // a pair was stored in this slot: its column of S'Y and Y'Y is formed with the direction
    double new_dr = 0.0;
    if (s.phase == kLbFirst) {
        double mx = 0.0;
        for (long long i = tid; i < n; i += T) mx = fmax(mx, fabs(g[i]));
        s.sbgnrm = lb_max(mx, red, par);
        s.f = f;
        if (s.sbgnrm <= o.pgtol) finish(kLbTaskConvergence, kLbMsgPgtol, 0);
        else start = true;
    } else {
        double part = 0.0;
        for (long long i = tid; i < n; i += T) part += g[i] * P.d[i];
        const double gd = lb_sum(part, red, par);
        s.gd = gd;
        if (!lb_dcsrch(s, f, gd)) {
            // another trial point, unless the search has used up its evaluations
            s.ifun++; s.iback = s.ifun - 1;
            if (s.iback >= o.maxls) {
                // give up the search: back to the iterate it started from
                for (long long i = tid; i < n; i += T) { P.x[i] = P.xold[i]; g[i] = P.gold[i]; }
                f = s.fold; s.f = f;
                if (s.col == 0) { s.iter++; finish(kLbTaskAbnormal, kLbMsgNone, 2); }
                else { s.col = 0; s.head = 0; s.theta = 1.0; s.nrestart++; start = true; }
            } else {
                const double stp = s.stp;
                for (long long i = tid; i < n; i += T) P.x[i] = stp * P.d[i] + P.xold[i];
            }
        } else {
            // the step is accepted: x is the new iterate
            s.iter++;
            s.f = f;
            double mx = 0.0;
            for (long long i = tid; i < n; i += T) mx = fmax(mx, fabs(g[i]));
            s.sbgnrm = lb_max(mx, red, par);
            const double dmax = fmax(fmax(fabs(s.fold), fabs(f)), 1.0);
            if (s.iter >= o.maxiter) finish(kLbTaskStop, kLbMsgMaxiter, 1);
            else if (s.funcalls > o.maxfun) finish(kLbTaskStop, kLbMsgMaxfun, 1);
            else if (s.sbgnrm <= o.pgtol) finish(kLbTaskConvergence, kLbMsgPgtol, 0);
            else if (s.fold - f <= o.tol * dmax) finish(kLbTaskConvergence, kLbMsgFactr, 0);
            else {
                // BFGS pair: s = stp d, y = g - g_old (left in gold); s'y from the search's slopes
                double dr, ddum;
                if (s.stp == 1.0) { dr = gd - s.gdold; ddum = -s.gdold; }
                else { dr = (gd - s.gdold) * s.stp; ddum = -s.gdold * s.stp; }
                const bool upd = !(dr <= DBL_EPSILON * ddum);
                int slot = 0;
                if (upd) {
                    if (s.col < m) { slot = (s.head + s.col) % m; s.col++; }
                    else { slot = s.head; s.head = (s.head + 1) % m; }
                } else {
                    s.nskip++;
                }
                double* Ss = P.S + (long long)slot * n;
                double* Ys = P.Y + (long long)slot * n;
                const double stp = s.stp;
                double rr = 0.0;
                for (long long i = tid; i < n; i += T) {
                    const double y = g[i] - P.gold[i];
                    P.gold[i] = y; rr += y * y;
                    if (upd) { Ss[i] = stp * P.d[i]; Ys[i] = y; }
                }
                rr = lb_sum(rr, red, par);
                if (upd) { s.theta = rr / dr; new_slot = slot; new_dr = dr; }
                start = true;
            }
        }
    }

    // ---- a new iteration: direction, set-up of the search, first trial point --------------------------
    while (start) {
        start = false;
        const int col = s.col;
        double* q = P.d;
        if (col == 0) {
            for (long long i = tid; i < n; i += T) q[i] = -g[i];
        } else {
            // (1) every dot product with the stored pairs in one sweep, four pairs (16 sums) at a time; a
            // thread owns the elements i = tid (mod T) of every vector, so only the sums cross threads
            for (int k0 = 0; k0 < col; k0 += 4) {
                const int np = col - k0 < 4 ? col - k0 : 4;
                const double *Sp[4], *Yp[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int slot = (s.head + k0 + (k < np ? k : 0)) % m;
                    Sp[k] = P.S + (long long)slot * n; Yp[k] = P.Y + (long long)slot * n;
                }
                double acc[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[k] = 0.0;
                for (long long i = tid; i < n; i += T) {
                    const double gi = g[i], yi = P.gold[i];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < np) {
                            const double sv = Sp[k][i], yv = Yp[k][i];
                            acc[4 * k] = __fma_rn(sv, yi, acc[4 * k]); acc[4 * k + 1] = __fma_rn(yv, yi, acc[4 * k + 1]);
                            acc[4 * k + 2] = __fma_rn(sv, gi, acc[4 * k + 2]); acc[4 * k + 3] = __fma_rn(yv, gi, acc[4 * k + 3]);
                        }
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    double v = acc[k];
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
                    if ((tid & 31) == 0) wred[tid >> 5][k] = v;
                }
                __syncthreads();
                if (tid < 4 * np) {
                    double t = 0.0;
                    for (int w = 0; w < (T >> 5); ++w) t += wred[w][tid];
                    dots[tid & 3][k0 + (tid >> 2)] = t;
                }
                __syncthreads();
            }
            // (2) the m x m algebra, one thread: the new pair's column of S'Y and Y'Y, then
            //   u = R^-1 S'g,  w = (D + gamma Y'Y) u - gamma Y'g,  H g = gamma g + S R^-T w - gamma Y u
            if (tid == 0) {
                const double gamma = 1.0 / s.theta;
                auto slot_of = [&](int k) { return (s.head + k) % m; };
                if (new_slot >= 0) {
                    for (int k = 0; k < col; ++k) {
                        const int sk = slot_of(k);
                        sy[sk * m + new_slot] = dots[0][k];
                        yy[sk * m + new_slot] = dots[1][k]; yy[new_slot * m + sk] = dots[1][k];
                    }
                    sy[new_slot * m + new_slot] = new_dr;          // s'y of the newest pair as the search measured it (L-BFGS-B's D)
                }
                double* u = coef[1];
                double* w = coef[0];
                for (int i = col - 1; i >= 0; --i) {
                    double t = dots[2][i];
                    for (int j = i + 1; j < col; ++j) t -= sy[slot_of(i) * m + slot_of(j)] * u[j];
                    u[i] = t / sy[slot_of(i) * m + slot_of(i)];
                }
                for (int i = 0; i < col; ++i) {
                    double t = 0.0;
                    for (int j = 0; j < col; ++j) t += yy[slot_of(i) * m + slot_of(j)] * u[j];
                    w[i] = sy[slot_of(i) * m + slot_of(i)] * u[i] + gamma * t - gamma * dots[3][i];
                }
                for (int i = 0; i < col; ++i) {
                    double t = w[i];
                    for (int j = 0; j < i; ++j) t -= sy[slot_of(j) * m + slot_of(i)] * w[j];
                    w[i] = t / sy[slot_of(i) * m + slot_of(i)];          // in place: w[j < i] already hold R^-T w
                }
                for (int i = 0; i < col; ++i) u[i] = -gamma * u[i];
            }
            __syncthreads();
            // (3) direction = -H g
            const double gamma = 1.0 / s.theta;
            for (int k0 = 0; k0 < col; k0 += 4) {
                const int np = col - k0 < 4 ? col - k0 : 4;
                const double *Sp[4], *Yp[4];
                double c1[4], c2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int slot = (s.head + k0 + (k < np ? k : 0)) % m;
                    Sp[k] = P.S + (long long)slot * n; Yp[k] = P.Y + (long long)slot * n;
                    c1[k] = k < np ? coef[0][k0 + k] : 0.0; c2[k] = k < np ? coef[1][k0 + k] : 0.0;
                }
                for (long long i = tid; i < n; i += T) {
                    double a = k0 == 0 ? gamma * g[i] : -q[i];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < np) a = __fma_rn(Sp[k][i], c1[k], __fma_rn(Yp[k][i], c2[k], a));
                    q[i] = -a;
                }
            }
            new_slot = -1;
        }
        // L-BFGS-B forms the point z = x + direction and searches along d = z - x
        double dtd = 0.0, gd = 0.0;
        for (long long i = tid; i < n; i += T) {
            const double xi = P.x[i], gi = g[i];
            const double zi = xi + q[i];
            const double di = zi - xi;
            P.z[i] = zi; q[i] = di;
            P.xold[i] = xi; P.gold[i] = gi;
            dtd += di * di; gd += gi * di;
        }
        dtd = lb_sum(dtd, red, par);
        gd = lb_sum(gd, red, par);
        s.dnorm = sqrt(dtd);
        s.stp = s.iter == 0 ? fmin(1.0 / s.dnorm, kLbStpMax) : 1.0;
        s.fold = s.f;
        s.ifun = 0; s.iback = 0;
        s.gd = gd; s.gdold = gd;
        if (!(gd < 0.0)) {
            // not a descent direction (rounding in a nearly singular memory): drop the memory
            if (col == 0) { s.iter++; finish(kLbTaskAbnormal, kLbMsgNone, 2); }
            else { s.col = 0; s.head = 0; s.theta = 1.0; s.nrestart++; start = true; }
            continue;
        }
        lb_dcsrch_start(s, s.f, gd);
        s.ifun = 1; s.iback = 0;
        const double stp = s.stp;
        if (stp == 1.0) { for (long long i = tid; i < n; i += T) P.x[i] = P.z[i]; }
        else { for (long long i = tid; i < n; i += T) P.x[i] = stp * q[i] + P.xold[i]; }
        s.phase = kLbSearch;
    }
    __syncthreads();
    for (int i = tid; i < o.m * o.m; i += T) { P.sy[i] = sy[i]; P.sy[o.m * o.m + i] = yy[i]; }
    if (tid == 0) { *P.st = s; flags[pid] = s.phase; }
}

}  // namespace jb

struct jb_lbfgs {
    std::vector<jb_plan*> plans;
    int device = 0;
    jb::LbfgsOpts opt{};
    std::vector<jb::LbfgsProblem> probs;
    std::vector<std::vector<jb::LbfgsReg>> regs;              // device index pointers inside
    std::vector<void*> dev;                                    // everything cudaMalloc'ed here
    jb::LbfgsProblem* d_probs = nullptr;
    jb::LbfgsState* d_states = nullptr;
    jb::LbfgsReg* d_regs = nullptr;
    int *d_flags = nullptr, *d_active = nullptr, *h_flags = nullptr;
    double *d_x = nullptr, *d_out = nullptr;                  // all chunks' x / [f, g], contiguous
    size_t n_x = 0, n_out = 0;
    std::vector<double> h_x, h_out;
    bool regs_dirty = true;
    int64_t rounds = 0, launches = 0, batches_built = 0;
};

namespace {
// device memory from the stream-ordered pool (the plans keep its release threshold at "never": a second
// optimiser finds the first one's memory there -- no driver call, nothing shared with other processes)
template <typename T>
int lb_alloc(jb_lbfgs* h, T** out, size_t count) {
    void* p = nullptr;
    JB_CUDA(cudaMallocAsync(&p, std::max<size_t>(count * sizeof(T), 16), h->plans[0]->stream));
    h->dev.push_back(p);
    *out = static_cast<T*>(p);
    return JB_OK;
}
// page-locked flag buffers of finished optimisers, kept for the next one
std::mutex g_lb_pinned_mutex;
std::vector<std::pair<int*, size_t>> g_lb_pinned;
int* lb_pinned_get(size_t n) {
    {
        std::lock_guard<std::mutex> lock(g_lb_pinned_mutex);
        for (size_t i = 0; i < g_lb_pinned.size(); ++i)
            if (g_lb_pinned[i].second >= n) {
                int* p = g_lb_pinned[i].first;
                g_lb_pinned.erase(g_lb_pinned.begin() + i);
                return p;
            }
    }
    int* p = nullptr;
    if (cudaMallocHost(&p, std::max<size_t>(n, 256) * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void lb_pinned_put(int* p, size_t n) {
    std::lock_guard<std::mutex> lock(g_lb_pinned_mutex);
    if (g_lb_pinned.size() < 8) g_lb_pinned.emplace_back(p, std::max<size_t>(n, 256));
    else cudaFreeHost(p);
}
}  // namespace

extern "C" {

int jb_lbfgs_create(jb_plan* const* plans, int32_t n, const jb_lbfgs_options* opt, jb_lbfgs** out) {
    if (!plans || n <= 0 || !opt || !out) return fail(JB_ERR_BAD_ARG, "bad arguments");
    *out = nullptr;
    if (opt->m <= 0 || opt->m > jb::kLbfgsMaxM) return fail(JB_ERR_BAD_ARG, "m must be in 1..%d", jb::kLbfgsMaxM);
    if (opt->maxls <= 0) return fail(JB_ERR_BAD_ARG, "maxls must be positive");
    if (opt->factr < 0 || opt->pgtol < 0) return fail(JB_ERR_BAD_ARG, "factr and pgtol must not be negative");
    for (int i = 0; i < n; ++i) {
        if (!plans[i]) return fail(JB_ERR_BAD_ARG, "null plan");
        if (plans[i]->device != plans[0]->device) return fail(JB_ERR_BAD_ARG, "the plans must live on one device");
        if (plans[i]->n_fit <= 0) return fail(JB_ERR_BAD_ARG, "plan %d has no parameter in fit mode", i);
    }
    JB_CUDA(cudaSetDevice(plans[0]->device));
    std::unique_ptr<jb_lbfgs, void (*)(jb_lbfgs*)> h(new jb_lbfgs, jb_lbfgs_destroy);
    h->plans.assign(plans, plans + n);
    h->device = plans[0]->device;
    h->opt.m = opt->m; h->opt.maxls = opt->maxls; h->opt.maxiter = opt->maxiter; h->opt.maxfun = opt->maxfun;
    h->opt.tol = opt->factr * DBL_EPSILON; h->opt.pgtol = opt->pgtol;
    h->probs.resize(n);
    h->regs.resize(n);
    int rc;
    if ((rc = lb_alloc(h.get(), &h->d_states, (size_t)n))) return rc;
    // one allocation for every vector of every chunk: all x | all out (1 + n each) -- so that starting
    // points go up and results come down in one copy each -- then per chunk
    // xold | gold | d | z | S (m n) | Y (m n) | S'Y, Y'Y (2 m m)
    size_t total = 0, nx = 0, nout = 0;
    for (int i = 0; i < n; ++i) {
        const size_t nn = (size_t)plans[i]->n_fit;
        nx += nn; nout += nn + 1;
        total += nn * (4 + 2 * (size_t)opt->m) + 2 * (size_t)opt->m * opt->m;
    }
    h->n_x = nx; h->n_out = nout;
    double* base = nullptr;
    if ((rc = lb_alloc(h.get(), &base, nx + nout + total))) return rc;
    h->d_x = base; h->d_out = base + nx;
    double* px = h->d_x;
    double* po = h->d_out;
    base += nx + nout;
    for (int i = 0; i < n; ++i) {
        jb::LbfgsProblem& P = h->probs[i];
        memset(&P, 0, sizeof P);
        P.n = plans[i]->n_fit;
        const size_t nn = (size_t)P.n;
        P.x = px; px += nn;
        P.out = po; po += nn + 1;
        P.xold = base; P.gold = P.xold + nn; P.d = P.gold + nn; P.z = P.d + nn;
        P.S = P.z + nn; P.Y = P.S + nn * opt->m;
        P.sy = P.Y + nn * opt->m;
        P.st = h->d_states + i;
        base = P.sy + 2 * (size_t)opt->m * opt->m;
    }
    h->h_x.resize(nx); h->h_out.resize(nout);
    if ((rc = lb_alloc(h.get(), &h->d_probs, (size_t)n))) return rc;
    if ((rc = lb_alloc(h.get(), &h->d_flags, (size_t)n))) return rc;
    if ((rc = lb_alloc(h.get(), &h->d_active, (size_t)n))) return rc;
    h->h_flags = lb_pinned_get((size_t)n);
    if (!h->h_flags) return fail(JB_ERR_CUDA, "no page-locked memory for the optimiser's flags");
    *out = h.release();
    return JB_OK;
}

void jb_lbfgs_destroy(jb_lbfgs* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStream_t st = h->plans.empty() ? nullptr : h->plans[0]->stream;
    for (void* p : h->dev) cudaFreeAsync(p, st);
    if (h->d_regs) cudaFreeAsync(h->d_regs, st);
    cudaStreamSynchronize(st);
    if (h->h_flags) lb_pinned_put(h->h_flags, h->plans.size());
    delete h;
}

int jb_lbfgs_add_reg(jb_lbfgs* h, int32_t problem, int32_t kind, double weight, double constant, const int64_t* indices, int64_t n_indices) {
    if (!h || problem < 0 || problem >= (int)h->plans.size() || (kind != 0 && kind != 1) || n_indices < 0 || (n_indices && !indices))
        return fail(JB_ERR_BAD_ARG, "bad regulariser");
    JB_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->probs[problem].n;
    std::vector<long long> idx((size_t)n_indices);
    for (int64_t j = 0; j < n_indices; ++j) {
        if (indices[j] < 0 || indices[j] >= n) return fail(JB_ERR_BAD_ARG, "regulariser index %lld outside the fit vector (%lld)", (long long)indices[j], (long long)n);
        idx[(size_t)j] = indices[j];
    }
    long long* d_idx = nullptr;
    int rc = lb_alloc(h, &d_idx, idx.size());
    if (rc) return rc;
    if (!idx.empty()) {
        JB_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(long long), cudaMemcpyHostToDevice, h->plans[0]->stream));
        JB_CUDA(cudaStreamSynchronize(h->plans[0]->stream));
    }
    jb::LbfgsReg R;
    R.idx = d_idx; R.n = n_indices; R.weight = weight; R.constant = constant; R.kind = kind; R.pad_ = 0;
    h->regs[problem].push_back(R);
    h->regs_dirty = true;
    return JB_OK;
}

int jb_lbfgs_run(jb_lbfgs* h, const double* const* x0, double* const* x_out, double* const* grad_out, jb_lbfgs_result* results) {
    if (!h || !x0 || !x_out || !results) return fail(JB_ERR_BAD_ARG, "null argument");
    JB_CUDA(cudaSetDevice(h->device));
    const int n = (int)h->plans.size();
    cudaStream_t st = h->plans[0]->stream;
    int rc;
    for (int i = 0; i < n; ++i) {
        if (!x0[i] || !x_out[i]) return fail(JB_ERR_BAD_ARG, "null vector for chunk %d", i);
        if (h->plans[i]->n_fit != h->probs[i].n) return fail(JB_ERR_BAD_ARG, "chunk %d: the fit vector changed size since jb_lbfgs_create", i);
    }
    if (h->regs_dirty) {
        if (h->d_regs) { cudaFreeAsync(h->d_regs, st); h->d_regs = nullptr; }
        std::vector<jb::LbfgsReg> flat;
        for (int i = 0; i < n; ++i) flat.insert(flat.end(), h->regs[i].begin(), h->regs[i].end());
        if (!flat.empty()) {
            JB_CUDA(cudaMallocAsync(&h->d_regs, flat.size() * sizeof(jb::LbfgsReg), st));
            JB_CUDA(cudaMemcpyAsync(h->d_regs, flat.data(), flat.size() * sizeof(jb::LbfgsReg), cudaMemcpyHostToDevice, st));
            JB_CUDA(cudaStreamSynchronize(st));
        }
        size_t o = 0;
        for (int i = 0; i < n; ++i) {
            h->probs[i].regs = h->d_regs ? h->d_regs + o : nullptr;
            h->probs[i].n_regs = (int)h->regs[i].size();
            o += h->regs[i].size();
        }
        JB_CUDA(cudaMemcpyAsync(h->d_probs, h->probs.data(), (size_t)n * sizeof(jb::LbfgsProblem), cudaMemcpyHostToDevice, st));
        JB_CUDA(cudaStreamSynchronize(st));
        h->regs_dirty = false;
    }
    // starting points and fresh states
    std::vector<jb::LbfgsState> hs((size_t)n);
    memset(hs.data(), 0, hs.size() * sizeof(jb::LbfgsState));
    for (int i = 0; i < n; ++i) {
        hs[i].theta = 1.0; hs[i].phase = jb::kLbFirst;
        memcpy(h->h_x.data() + (h->probs[i].x - h->d_x), x0[i], (size_t)h->probs[i].n * sizeof(double));
    }
    JB_CUDA(cudaMemcpyAsync(h->d_x, h->h_x.data(), h->n_x * sizeof(double), cudaMemcpyHostToDevice, st));
    JB_CUDA(cudaMemcpyAsync(h->d_states, hs.data(), hs.size() * sizeof(jb::LbfgsState), cudaMemcpyHostToDevice, st));
    JB_CUDA(cudaStreamSynchronize(st));

    std::vector<int> active((size_t)n);
    for (int i = 0; i < n; ++i) active[i] = i;
    jb_batch* batch = nullptr;
    cudaGraphExec_t round_exec = nullptr;
    int steady = 0;
    bool round_graphs_ok = true;
    auto drop_batch = [&]() {
        if (round_exec) { cudaStreamSynchronize(st); drop_graph(round_exec); }
        steady = 0;
        if (batch) { jb_batch_destroy(batch); batch = nullptr; }
    };
    rc = JB_OK;
    std::vector<int> repeats((size_t)n, 0);
    while (!active.empty()) {
        if (!batch) {
            std::vector<jb_plan*> pl; std::vector<const double*> px; std::vector<double*> po;
            for (int i : active) { pl.push_back(h->plans[i]); px.push_back(h->probs[i].x); po.push_back(h->probs[i].out); }
            if ((rc = jb_batch_create(pl.data(), (int)pl.size(), &batch))) break;
            if ((rc = jb_batch_bind(batch, px.data(), po.data()))) break;
            jb_batch_soft_range(batch, 1);
            if ((rc = batch_status_buffers(batch))) break;
            if (cudaMemcpyAsync(h->d_active, active.data(), active.size() * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = fail(JB_ERR_CUDA, "cudaMemcpyAsync failed"); break; }
            h->batches_built++;
        }
        // one round, one wait: the evaluation, the optimiser's step, the status words and the flags --
        // recorded as ONE CUDA graph once two rounds in a row enqueue the same thing (same chunks still
        // iterating, no plan re-laid out), replayed until that changes
        const size_t na = active.size();
        auto enqueue_round = [&]() -> int {
            int r = jb_batch_eval_device(batch, st);
            if (r) return r;
            jb::k_lbfgs_advance<<<(unsigned)na, jb::kLbfgsThreads, 0, st>>>(h->d_probs, h->d_active, h->opt, h->d_flags, batch->d_stat_ptrs);
            jb::k_gather_status<<<(unsigned)((na + 127) / 128), 128, 0, st>>>(batch->d_stat_ptrs, (int)na, batch->d_stat_flat);
            if (cudaMemcpyAsync(batch->h_stat_flat, batch->d_stat_flat, 2 * na * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaMemcpyAsync(h->h_flags, h->d_flags, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return fail(JB_ERR_CUDA, "L-BFGS round: %s", cudaGetErrorString(cudaGetLastError()));
            return JB_OK;
        };
        bool replayed = false;
        if (graphs_enabled() && round_graphs_ok && batch_is_clean(batch)) {
            if (!round_exec && ++steady >= 2) {
                const int64_t launches0 = batch->launches;
                std::vector<int64_t> ev0;
                for (jb_plan* pl : batch->plans) ev0.push_back(pl->evaluations);
                int rrc = JB_OK;
                if (!record_graph(st, &round_exec, &rrc, enqueue_round) || rrc) { round_graphs_ok = false; drop_graph(round_exec); }
                batch->launches = launches0;
                for (size_t i = 0; i < ev0.size(); ++i) batch->plans[i]->evaluations = ev0[i];
            }
            if (round_exec) {
                if (cudaGraphLaunch(round_exec, st) != cudaSuccess) { rc = fail(JB_ERR_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(cudaGetLastError())); break; }
                batch->launches += batch->p_fit.empty() ? 5 : 6;
                for (jb_plan* pl : batch->plans) pl->evaluations++;
                replayed = true;
            }
        } else {
            if (round_exec) { cudaStreamSynchronize(st); drop_graph(round_exec); }
            steady = 0;
        }
        if (!replayed && (rc = enqueue_round())) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            rc = fail(JB_ERR_CUDA, "L-BFGS round failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        h->rounds++; h->launches += 2;
        // plans that raised a status word: an off-grid trial point is the optimiser's business (its loss is
        // NaN); a knot-window overflow is settled here (rows re-sorted with the current shifts, or the plan
        // leaves k_strip) and the chunk, which did not advance, evaluates the same point again
        for (size_t k = 0; k < na && !rc; ++k) {
            const bool repeat = h->h_flags[active[k]] == jb::kLbRepeat;
            if (!batch->h_stat_flat[2 * k] && !batch->h_stat_flat[2 * k + 1]) {
                if (repeat) rc = fail(JB_ERR_CUDA, "chunk %d: evaluation refused without a status word", active[k]);
                continue;
            }
            jb_plan* pl = h->plans[active[k]];
            bool window = false;
            const int rs = read_status(pl, st, &window);
            if (rs == JB_OK || rs == JB_ERR_OUT_OF_RANGE) {
                if (repeat) rc = fail(JB_ERR_CUDA, "chunk %d: evaluation refused with status %d", active[k], batch->h_stat_flat[2 * k]);
                continue;
            }
            if (!window || repeats[active[k]] >= 14) { rc = rs; break; }
            const std::string msg = g_err;
            bool retry = false;
            if ((rc = window_retry(pl, repeats[active[k]]++, &retry))) break;
            if (!retry) { g_err = msg; rc = rs; break; }
        }
        if (rc) break;
        std::vector<int> next;
        for (int i : active) if (h->h_flags[i] != jb::kLbDone) next.push_back(i);
        if (next.size() != active.size()) { active.swap(next); drop_batch(); }
    }
    if (batch) h->launches += jb_batch_launches(batch);
    drop_batch();
    if (rc) return rc;
    JB_CUDA(cudaMemcpyAsync(hs.data(), h->d_states, hs.size() * sizeof(jb::LbfgsState), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaMemcpyAsync(h->h_x.data(), h->d_x, h->n_x * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad_out) JB_CUDA(cudaMemcpyAsync(h->h_out.data(), h->d_out, h->n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) {
        const jb::LbfgsProblem& P = h->probs[i];
        memcpy(x_out[i], h->h_x.data() + (P.x - h->d_x), (size_t)P.n * sizeof(double));
        if (grad_out && grad_out[i]) memcpy(grad_out[i], h->h_out.data() + (P.out - h->d_out) + 1, (size_t)P.n * sizeof(double));
        jb_lbfgs_result& R = results[i];
        R.f = hs[i].f; R.pg_norm = hs[i].sbgnrm; R.nit = hs[i].iter; R.funcalls = hs[i].funcalls;
        R.warnflag = hs[i].warnflag; R.task = hs[i].task; R.task_msg = hs[i].msg;
        R.n_out_of_range = hs[i].n_oor; R.n_skipped = hs[i].nskip; R.n_restarts = hs[i].nrestart;
    }
    return JB_OK;
}

int64_t jb_lbfgs_rounds(const jb_lbfgs* h) { return h ? h->rounds : 0; }

}  // extern "C"
