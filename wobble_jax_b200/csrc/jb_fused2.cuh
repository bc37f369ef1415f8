// k_fused2: the fused loss + gradient + Fisher kernel for plans whose every (row, tile) is "smooth"
// (consecutive pixels less than one knot apart, lanes of one class never on the same knots; decided
// at upload, jb_plan.cu) and that have no per-key (NormalizationModel) component.  Same work
// decomposition, data layout and outputs as k_fused (jb_kernels.cuh); what differs is the code per
// pixel:
//   * the four pixels of a lane and the two walking directions are written out (no pixel loop, no
//     direction branch: rows are taken two at a time, one walked up, the next walked down), so the
//     window registers are renamed statically and never copied;
//   * no branch per pixel: the knot that leaves a lane's window and the control point that enters
//     it move under a predicate (one predicated shared-memory read-modify-write and one predicated
//     load), the sliding of the adjoint sums rides in the addend of their FMA;
//   * at a row start the window follows the first pixel by at most one knot either way (rows of a
//     group are sorted to land on the same knots) -- a three-way select, no flush;
//   * what is not asked for is not computed: FL says per component whether the shift is in fit mode
//     (derivative of the spline, shift gradient, Fisher sum) and whether a StretchingModel
//     multiplies it.
// Reference semantics per pixel: cardinal_vmap_model (jabble/model.py:951-989), ShiftingModel /
// StretchingModel (model.py:924-946), Additive / Composite (model.py:670-697), ChiSquare
// (loss.py:164-173); adjoints as in SURVEY.md section 3.3; Fisher sums model.py:893-897.
#pragma once
#include "jb_kernels.cuh"

namespace jb {

// FL bits: c (c < 2): derivative sums of component c;  2 + c: component c has a stretch factor
__host__ __device__ constexpr bool f2_der(int FL, int c) { return (FL >> c) & 1; }
__host__ __device__ constexpr bool f2_str(int FL, int c) { return (FL >> (2 + c)) & 1; }
// compact per-row sums, in this order: per component [shift gradient, Fisher] if der, [stretch gradient] if str
__host__ __device__ constexpr int f2_nsums(int NC, int FL) {
    int n = 0;
    for (int c = 0; c < NC; ++c) n += (f2_der(FL, c) ? 2 : 0) + (f2_str(FL, c) ? 1 : 0);
    return n;
}
__host__ __device__ constexpr int f2_pad(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : 8; }
// index of a sum in the compact row (-1: not computed); kind 0 shift gradient, 1 stretch gradient, 2 Fisher
__host__ __device__ constexpr int f2_slot(int NC, int FL, int comp, int kind) {
    int n = 0;
    for (int c = 0; c < NC; ++c) {
        if (f2_der(FL, c)) { if (c == comp && kind == 0) return n; if (c == comp && kind == 2) return n + 1; n += 2; }
        if (f2_str(FL, c)) { if (c == comp && kind == 1) return n; n += 1; }
    }
    return -1;
}

// Sum NV (1, 2, 4 or 8) per-lane values over the warp; afterwards the total of value q sits on the
// lanes with lane / (32 / NV) == q.  Fixed order.
template <int NV>
__device__ __forceinline__ double warp_sum_n(double (&v)[NV], int lane) {
    static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16, "NV");
    int bit = 16;
#pragma unroll
    for (int n = NV / 2; n >= 1; n >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const double send = up ? v[i] : v[i + n];
            const double keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        bit >>= 1;
    }
    for (; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
    return v[0];
}

// lane l accumulates into copy l % kClasses2 of the gradient window: lanes l and l + kClasses2 are
// 4 * kClasses2 pixels apart, which the plan checks is more than a spline support (jb_plan.cu)
#ifndef JB_F2_PAIRSUM
#define JB_F2_PAIRSUM 1
#endif
#ifndef JB_F2_CLASSES
#define JB_F2_CLASSES 3
#endif
constexpr int kClasses2 = JB_F2_CLASSES;

template <int PV>
struct Win2 {
    double c[PV + 1];   // control points of knots klo .. klo+PV, klo = I + KOFF
    double a[PV + 1];   // adjoint sums of those knots
    int I;              // interval the window sits on
    uint32_t pa;        // shared-memory address of the class copy's entry for knot klo
#ifdef JB_DEBUG_BOUNDS  // self-check builds (compute-sanitizer is not available on the GPU pool): the
    uint32_t dbg_lo, dbg_hi;   // first / last legal value of pa; violations raise kStatWindow
    int* dbg_status;
#endif
};

template <int PV>
__device__ __forceinline__ void win_check(const Win2<PV>& w, uint32_t pa) {
#ifdef JB_DEBUG_BOUNDS
    if (pa < w.dbg_lo || pa > w.dbg_hi) atomicOr(w.dbg_status, kStatBounds);
#endif
}

// predicated shared-memory helpers (no branch, no memory clobber: the windows are only touched
// through these between the __syncwarp()s that fence the prologue and the epilogue)
__device__ __forceinline__ void lds_if_ne(double& v, uint32_t addr, int i0, int i1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, %3;\n\t@p ld.shared.f64 %0, [%1];\n\t}"
                 : "+d"(v) : "r"(addr), "r"(i0), "r"(i1));
}
__device__ __forceinline__ void rmw_if(uint32_t addr, double v, int pred) {
    // the load and the add are unconditional (the address is always inside the window): a temporary
    // written under a predicate would stay live, in a register of its own, for the whole kernel
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f64 t;\n\tsetp.ne.s32 p, %2, 0;\n\tld.shared.f64 t, [%0];\n\t"
                 "add.rn.f64 t, t, %1;\n\t@p st.shared.f64 [%0], t;\n\t}"
                 :: "r"(addr), "d"(v), "r"(pred));
}
// one step of the window: if (i0 != i1) { cnew = smem[ca]; smem[ra] += rv; }
__device__ __forceinline__ void step_if_ne(double& cnew, uint32_t ca, uint32_t ra, double rv, int i0, int i1) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f64 t;\n\tsetp.ne.s32 p, %4, %5;\n\t@p ld.shared.f64 %0, [%1];\n\t"
                 "ld.shared.f64 t, [%2];\n\tadd.rn.f64 t, t, %3;\n\t@p st.shared.f64 [%2], t;\n\t}"
                 : "+d"(cnew) : "r"(ca), "r"(ra), "d"(rv), "r"(i0), "r"(i1));
}

// Row start: bring the window to interval In.  One knot up or down slides (the knot that leaves is
// retired to the class copy, the other sums move over); further away (first row of a task, rows of
// a group far apart) everything is retired.  The control points are re-read when the window moved.
template <int PV>
__device__ __forceinline__ void win_turn(Win2<PV>& w, int In, uint32_t abase, uint32_t cdelta, bool resync, int lane,
                                         bool first_row) {
    const int d = In - w.I;
    const bool up = d == 1, dn = d == -1;
    // resync: the lane sat out the previous row (no weighted pixel), so its window is where an
    // earlier row left it and the per-row disjointness of the lanes of a class says nothing about it
    const bool far = (unsigned)(d + 1) > 2u || (resync && d != 0);
    if (!first_row && __any_sync(0xffffffffu, far)) {      // (first row of a task: the sums are still zero)
        // windows from different rows may overlap inside a class: one lane of the class at a time
#pragma unroll 1
        for (int m = 0; m < (32 + kClasses2 - 1) / kClasses2; ++m) {
            const int mine = (far && lane / kClasses2 == m) ? 1 : 0;
#pragma unroll
            for (int k = 0; k <= PV; ++k) rmw_if(w.pa + 8u * k, w.a[k], mine);
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k <= PV; ++k) w.a[k] = far ? 0.0 : w.a[k];
    }
    rmw_if(w.pa + (dn ? 8u * PV : 0u), dn ? w.a[PV] : w.a[0], (!far && (up || dn)) ? 1 : 0);
    // slide in place: one pass upwards, one downwards (no temporaries)
    const bool su = up && !far, sd = dn && !far;
#pragma unroll
    for (int k = 0; k < PV; ++k) w.a[k] = su ? w.a[k + 1] : w.a[k];
    w.a[PV] = su ? 0.0 : w.a[PV];
#pragma unroll
    for (int k = PV; k > 0; --k) w.a[k] = sd ? w.a[k - 1] : w.a[k];
    w.a[0] = sd ? 0.0 : w.a[0];
    const uint32_t pa = abase + 8u * (uint32_t)In;
    win_check<PV>(w, pa);
    const int iold = resync ? In + 1 : w.I;      // after sitting out (and at the first row): always re-read
#pragma unroll
    for (int k = 0; k <= PV; ++k) lds_if_ne(w.c[k], pa + cdelta + 8u * k, In, iold);
    w.I = In;
    w.pa = pa;
}

// Next pixel of a row walked in direction DIR: the window stays or moves one knot that way.
// Returns whether it moved (the adjoint sums slide inside their FMA, see f2_pixel).
template <int PV, int DIR>
__device__ __forceinline__ bool win_step(Win2<PV>& w, int In, uint32_t abase, uint32_t cdelta) {
    const uint32_t pa = abase + 8u * (uint32_t)In;
    win_check<PV>(w, pa);
    const bool step = In != w.I;
    const double rv = DIR > 0 ? w.a[0] : w.a[PV];
    if (DIR > 0) {
#pragma unroll
        for (int k = 0; k < PV; ++k) w.c[k] = step ? w.c[k + 1] : w.c[k];
        step_if_ne(w.c[PV], pa + cdelta + 8u * PV, w.pa, rv, In, w.I);
    } else {
#pragma unroll
        for (int k = PV; k > 0; --k) w.c[k] = step ? w.c[k - 1] : w.c[k];
        step_if_ne(w.c[0], pa + cdelta, w.pa + 8u * PV, rv, In, w.I);
    }
    w.I = In;
    w.pa = pa;
    return step;
}

#ifdef JB_F2_SAMEGRID   // tuning builds: every component on the knot grid of component 0
#define JB_F2_GRID(c) 0
#else
#define JB_F2_GRID(c) (c)
#endif
template <int NC> struct RowConst { double sh[NC], st[NC], xp0[NC], inv_dx[NC]; };

// One pixel, windows already in place.  SLIDE = 0: the windows did not move (first pixel of a row);
// +1 / -1: window c moved one knot in that direction when step[c].
template <int NC, int PV, int FL, int SLIDE>
__device__ __forceinline__ void f2_pixel(const double (&g)[NC], double y, double iv, const RowConst<NC>& rc,
                                         Win2<PV> (&win)[NC], const bool (&step)[NC], double& loss,
                                         double (&sums)[f2_pad(f2_nsums(NC, FL))]) {
    double w[NC][PV + 1], S[NC], dS[NC];
    double m = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double dw[PV > 0 ? PV : 1];
        Basis<PV>::eval(g[c], w[c], dw);
        const Win2<PV>& W = win[c];
        double sv = W.c[0] * w[c][0];
#pragma unroll
        for (int k = 1; k <= PV; ++k) sv = __fma_rn(W.c[k], w[c][k], sv);
        S[c] = sv;
        if (f2_der(FL, c)) {
            double d = (W.c[1] - W.c[0]) * dw[0];
#pragma unroll
            for (int k = 1; k < PV; ++k) d = __fma_rn(W.c[k + 1] - W.c[k], dw[k], d);
            dS[c] = f2_str(FL, c) ? rc.st[c] * d : d;
        }
        if (c == 0) m = f2_str(FL, c) ? rc.st[c] * sv : sv;
        else m = f2_str(FL, c) ? __fma_rn(rc.st[c], sv, m) : m + sv;     // StretchingModel / AdditiveModel
    }
    const double res = y - m;                                             // ChiSquare, loss.py:169-171
    const double wr = iv * res;
    loss = __fma_rn(wr, res, loss);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const double q = f2_str(FL, c) ? wr * rc.st[c] : wr;
        Win2<PV>& W = win[c];
        if (SLIDE == 0) {
#pragma unroll
            for (int k = 0; k <= PV; ++k) W.a[k] = __fma_rn(q, w[c][k], W.a[k]);
        } else if (SLIDE > 0) {
#pragma unroll
            for (int k = 0; k < PV; ++k) W.a[k] = __fma_rn(q, w[c][k], step[c] ? W.a[k + 1] : W.a[k]);
            W.a[PV] = __fma_rn(q, w[c][PV], step[c] ? 0.0 : W.a[PV]);
        } else {
#pragma unroll
            for (int k = PV; k > 0; --k) W.a[k] = __fma_rn(q, w[c][k], step[c] ? W.a[k - 1] : W.a[k]);
            W.a[0] = __fma_rn(q, w[c][0], step[c] ? 0.0 : W.a[0]);
        }
        if (f2_der(FL, c)) {
            const int s0 = f2_slot(NC, FL, c, 0), s2 = f2_slot(NC, FL, c, 2);
            sums[s0] = __fma_rn(wr, dS[c], sums[s0]);                      // d/d shift (times PV/dx later)
            sums[s2] = __fma_rn(iv * dS[c], dS[c], sums[s2]);              // Fisher
        }
        if (f2_str(FL, c)) {
            const int s1 = f2_slot(NC, FL, c, 1);
            sums[s1] = __fma_rn(wr, S[c], sums[s1]);                       // d/d stretch
        }
    }
}

// One row of the task walked in direction DIR (+1: pixels 0..3 of the lane, -1: 3..0).
// px = the lane's column of the staged record (pixel j at [32 * j], x | y | ivar kTile doubles apart).
template <int NC, int PV, int FL, int DIR, bool ALL>
__device__ __forceinline__ void f2_row(const double* px, const RowConst<NC>& rc, Win2<PV> (&win)[NC],
                                       const uint32_t (&abase)[NC], uint32_t cdelta, double& loss,
                                       double (&sums)[f2_pad(f2_nsums(NC, FL))], bool live, bool resync, int lane,
                                       bool first_row) {
    // live = the lane owns a weighted pixel in this row.  A lane that does not keeps its windows
    // where they are (its filled pixels carry ivar = 0 and an x on its neighbours' knots: every
    // sum gets exactly zero from them, wherever the window sits).  ALL: every lane of the row is live.
    if (ALL) live = true;
    double g[NC];
    bool step[NC];
    {
        constexpr int j = DIR > 0 ? 0 : 3;
        const double x = px[32 * j], y = px[kTile + 32 * j], iv = px[2 * kTile + 32 * j];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int In;
            knot_coord<PV>(x, rc.sh[c], rc.xp0[JB_F2_GRID(c)], rc.inv_dx[JB_F2_GRID(c)], g[c], In);
            In = live ? In : win[c].I;
            win_turn<PV>(win[c], In, abase[c], cdelta, resync, lane, first_row);
            step[c] = false;
        }
        f2_pixel<NC, PV, FL, 0>(g, y, iv, rc, win, step, loss, sums);
    }
#ifndef JB_F2_UNROLL
#define JB_F2_UNROLL 3
#endif
    constexpr int kUnroll = JB_F2_UNROLL;
#pragma unroll kUnroll
    for (int jj = 1; jj < 4; ++jj) {
        const int j = DIR > 0 ? jj : 3 - jj;
        const double x = px[32 * j], y = px[kTile + 32 * j], iv = px[2 * kTile + 32 * j];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int In;
            knot_coord<PV>(x, rc.sh[c], rc.xp0[JB_F2_GRID(c)], rc.inv_dx[JB_F2_GRID(c)], g[c], In);
            In = live ? In : win[c].I;
            step[c] = win_step<PV, DIR>(win[c], In, abase[c], cdelta);
        }
        f2_pixel<NC, PV, FL, DIR>(g, y, iv, rc, win, step, loss, sums);
    }
}

__shared__ FusedParams g_f2_plan;     // the block's plan

// Per-warp shared memory: [ring kStages x kTileBytes][recs kMaxGroup][mbarriers kStages][pad to 16]
// at fixed offsets, then the control-point windows [NC][W] and the class copies [kClasses2][NC][W].
template <int NC> struct F2Lay {
    static constexpr uint32_t ring = 0;
    static constexpr uint32_t recs = kStages * kTileBytes;
    static constexpr uint32_t bars = recs + kMaxGroup * sizeof(RowRec<NC, 0>);
    static constexpr uint32_t ctrl = (bars + kStages * 8 + 15) / 16 * 16;
};
// what a row needs besides the windows
template <int NC> struct F2Ctx {
    unsigned char* wb;       // the warp's shared memory
    int nrow, tile, lane;
};

// Row i of the group: wait for its record in the ring, walk it, reduce and store its per-row sums,
// refill the ring stage.
// DEFER: the per-row sums are left in `sums` for the caller, which reduces the sums of an (up, down) pair
// of rows in one butterfly (half as many dependent shuffle rounds per row).
template <int NC, int PV, int FL, int DIR, bool DEFER>
__device__ __forceinline__ void f2_do_row(int i, const F2Ctx<NC>& cx, RowConst<NC>& rc, Win2<PV> (&win)[NC],
                                          const uint32_t (&abase)[NC], uint32_t cdelta, double& loss, bool& was_live,
                                          double (&sums)[f2_pad(f2_nsums(NC, FL))]) {
    constexpr int NS = f2_nsums(NC, FL), NSP = f2_pad(NS);
    using L = F2Lay<NC>;
    const FusedParams& a = g_f2_plan;
    const int s = i % kStages;
    const RowRec<NC, 0>* recs = reinterpret_cast<const RowRec<NC, 0>*>(cx.wb + L::recs);
    const RowRec<NC, 0>& rec = recs[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) { rc.sh[c] = rec.sh[c]; rc.st[c] = rec.st[c]; }
    const unsigned live_mask = (unsigned)rec.nkey;      // lanes that own a weighted pixel in this row of the tile
    const uint32_t bar = smem_u32(cx.wb + L::bars) + 8 * s;
    mbar_wait(bar, (uint32_t)((i / kStages) & 1));
#pragma unroll
    for (int q = 0; q < NSP; ++q) sums[q] = 0.0;
    if (live_mask != 0) {      // some weighted pixel in this row of the tile
        const double* px = reinterpret_cast<const double*>(cx.wb + L::ring + (size_t)s * kTileBytes) + cx.lane;
        const bool live = (live_mask >> cx.lane) & 1u;
        if (live_mask == 0xffffffffu)     // the usual row: no lane sits out (the frozen-lane selects fold away)
            f2_row<NC, PV, FL, DIR, true>(px, rc, win, abase, cdelta, loss, sums, true, !was_live, cx.lane, DIR > 0 && i == 0);
        else
            f2_row<NC, PV, FL, DIR, false>(px, rc, win, abase, cdelta, loss, sums, live, !was_live, cx.lane, DIR > 0 && i == 0);
        was_live = live;
    } else was_live = false;
    // every lane has consumed stage s: refill it
    __syncwarp();
    if (cx.lane == 0 && i + kStages < cx.nrow) {
        mbar_expect_tx(bar, kTileBytes);
        tma_load_1d(smem_u32(cx.wb + L::ring + (size_t)s * kTileBytes),
                    a.data + ((int64_t)recs[i + kStages].row * a.n_tiles + cx.tile) * kTileDoubles, kTileBytes, bar);
    }
    if (NS > 0 && !DEFER) {
        // per-row sums of this tile (fixed order); the total of value q sits on lanes q*(32/NSP)..
        const double tot = warp_sum_n<NSP>(sums, cx.lane);
        if ((cx.lane & (32 / NSP - 1)) == 0)
            a.rowpart[((int64_t)recs[i].row * a.n_tiles + cx.tile) * NSP + cx.lane / (32 / NSP)] = tot;
    }
}

// the sums of rows i (su) and i + 1 (sd) in one butterfly: value q of row i ends on lanes q * (32 / 2NSP) ..,
// value q of row i + 1 behind them; the same addition tree per value as warp_sum_n<NSP>
template <int NC, int FL>
__device__ __forceinline__ void f2_store_pair(int i, const F2Ctx<NC>& cx, const double (&su)[f2_pad(f2_nsums(NC, FL))],
                                              const double (&sd)[f2_pad(f2_nsums(NC, FL))]) {
    constexpr int NSP = f2_pad(f2_nsums(NC, FL));
    using L = F2Lay<NC>;
    const FusedParams& a = g_f2_plan;
    const RowRec<NC, 0>* recs = reinterpret_cast<const RowRec<NC, 0>*>(cx.wb + L::recs);
    double v[2 * NSP];
#pragma unroll
    for (int q = 0; q < NSP; ++q) { v[q] = su[q]; v[NSP + q] = sd[q]; }
    const double tot = warp_sum_n<2 * NSP>(v, cx.lane);
    constexpr int LPV = 32 / (2 * NSP);        // lanes per value
    if ((cx.lane & (LPV - 1)) == 0) {
        const int q = cx.lane / LPV;
        const int row = recs[i + (q >= NSP ? 1 : 0)].row;
        a.rowpart[((int64_t)row * a.n_tiles + cx.tile) * NSP + (q & (NSP - 1))] = tot;
    }
}

template <int NC>
__host__ __device__ constexpr size_t fused2_warp_smem(int wincap) {
    return (F2Lay<NC>::ctrl + (size_t)NC * wincap * 8 + (size_t)NC * kClasses2 * wincap * 8 + 15) / 16 * 16;   // 16 B: bulk-copy destination, double2 stores
}

template <int NC, int PV, int FL>
__global__ void __maxnreg__(JB_MAXREG) k_fused2(const FusedParams* __restrict__ arr, const int* __restrict__ off, int n) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedParams& a = g_f2_plan;
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
    constexpr int NS = f2_nsums(NC, FL), NSP = f2_pad(NS);
    using Rec = RowRec<NC, 0>;
    using L = F2Lay<NC>;
    const int W = a.wincap;

    unsigned char* wbase = smem_raw + (size_t)warp * fused2_warp_smem<NC>(W);
    unsigned char* ring = wbase + L::ring;                                   // [kStages][kTileBytes]
    Rec* recs = reinterpret_cast<Rec*>(wbase + L::recs);
    const uint32_t bar0 = smem_u32(wbase + L::bars);                         // [kStages] mbarriers
    double* ctrl_warp = reinterpret_cast<double*>(wbase + L::ctrl);          // [NC][W]
    double* acc_warp = ctrl_warp + NC * W;                                   // [kClasses2][NC][W]

    // ---- 0. lane i fetches the record of row i of the group ------------------------------------
    int lo[NC], hi[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { lo[c] = INT_MAX; hi[c] = INT_MIN; }
    bool has_work = false;
    if (lane < nrow) {
        Rec rec = reinterpret_cast<const Rec*>(a.grouprec)[row0 + lane];
        const TileInfo ti = a.tinfo[(int64_t)tile * a.n_rows + row0 + lane];
        rec.dense = ti.flags;
        rec.nkey = (int)ti.live;
        has_work = ti.live != 0;
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
        for (int i = lane; i < nrow * NSP; i += 32)
            a.rowpart[((int64_t)recs[i / NSP].row * a.n_tiles + tile) * NSP + i % NSP] = 0.0;
        if (!empty)   // copies are in flight into this warp's ring: let them land before leaving
            for (int i = 0; i < kStages && i < nrow; ++i) mbar_wait(bar0 + 8 * (i % kStages), 0);
        return;
    }

    // ---- 3. accumulators and control-point window ------------------------------------------------
    {
        double2* z = reinterpret_cast<double2*>(acc_warp);
        for (int i = lane; i < NC * kClasses2 * W / 2; i += 32) z[i] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int nk = hi[c] - lo[c] + 1;
        for (int k = lane; k < W; k += 32) ctrl_warp[c * W + k] = k < nk ? __ldg(a.comp[c].ctrl + lo[c] + k) : 0.0;
    }
    __syncwarp();

    constexpr int KOFF = (PV + 1) / 2 - PV;          // klo = I + KOFF (knot_lo)
    const int cls = lane % kClasses2;
    // address of the class copy's entry for knot klo of interval I: abase[c] + 8 * I
    uint32_t abase[NC];
    Win2<PV> win[NC];
    RowConst<NC> rc;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const uint32_t A = smem_u32(acc_warp + (size_t)(cls * NC + c) * W);
        abase[c] = A + 8u * (uint32_t)(KOFF - lo[c]);
        asm volatile("" : "+r"(abase[c]));   // one register, not base + offset
        win[c].I = lo[c] - KOFF;             // a valid (empty) window until the first pixel moves it
        win[c].pa = A;
#ifdef JB_DEBUG_BOUNDS
        win[c].dbg_lo = A; win[c].dbg_hi = A + 8u * (uint32_t)(W - PV - 1); win[c].dbg_status = a.status;
#endif
#pragma unroll
        for (int k = 0; k <= PV; ++k) { win[c].a[k] = 0.0; win[c].c[k] = 0.0; }
        rc.xp0[c] = a.comp[c].xp0; rc.inv_dx[c] = a.comp[c].inv_dx;
    }
    // control point of knot k of component c sits cdelta bytes behind the class copy's entry
    const uint32_t cdelta = smem_u32(ctrl_warp) - smem_u32(acc_warp + (size_t)cls * NC * W);

    double loss = 0.0;
    F2Ctx<NC> cx;
    cx.wb = wbase; cx.nrow = nrow; cx.tile = tile; cx.lane = lane;

    // ---- 4. stream the rows of the group, two at a time (up, down) ------------------------------
    bool was_live = false;
    {
        constexpr int NSPr = f2_pad(f2_nsums(NC, FL));
        constexpr bool kPair = JB_F2_PAIRSUM != 0 && f2_nsums(NC, FL) > 0 && 2 * NSPr <= 16;
        int i = 0;
        for (; i + 1 < nrow; i += 2) {
            double su[NSPr], sd[NSPr];
            f2_do_row<NC, PV, FL, +1, kPair>(i, cx, rc, win, abase, cdelta, loss, was_live, su);
            f2_do_row<NC, PV, FL, -1, kPair>(i + 1, cx, rc, win, abase, cdelta, loss, was_live, sd);
            if (kPair) f2_store_pair<NC, FL>(i, cx, su, sd);
        }
        if (i < nrow) {
            double su[NSPr];
            f2_do_row<NC, PV, FL, +1, false>(i, cx, rc, win, abase, cdelta, loss, was_live, su);
        }
    }

    // ---- 5. flush, fold the accumulator copies in a fixed order, publish the window ------------
    if (__all_sync(0xffffffffu, was_live)) {
        // every window sits on the last row: lanes of a class are on disjoint knots
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k <= PV; ++k) rmw_if(win[c].pa + 8u * k, win[c].a[k], 1);
    } else {
        // windows from different rows may overlap inside a class: one lane of the class at a time
#pragma unroll 1
        for (int m = 0; m < (32 + kClasses2 - 1) / kClasses2; ++m) {
            const int mine = lane / kClasses2 == m ? 1 : 0;
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int k = 0; k <= PV; ++k) rmw_if(win[c].pa + 8u * k, win[c].a[k], mine);
            __syncwarp();
        }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double* dst = a.part + ((int64_t)c * a.n_tasks + task) * W;
        for (int k = lane; k < W; k += 32) {
            double sacc = acc_warp[(size_t)c * W + k];
#pragma unroll
            for (int q = 1; q < kClasses2; ++q) sacc += acc_warp[(size_t)(q * NC + c) * W + k];
            dst[k] = sacc;
        }
    }
    loss = warp_sum(loss);
    if (lane == 0) a.losspart[task] = loss;
}

}  // namespace jb
