// k_fused32: the float32 tier of the fused loss + gradient + Fisher kernel (north star: results
// within 1e-5 of the float64 reference).  Same decomposition, window logic and outputs as k_fused2
// (jb_fused2.cuh); what differs:
//   * records are float: f | y | ivar, 128 each, plus one byte K per pixel (1664 B per (row, tile),
//     13 B per pixel -- the byte the mask would have taken; the mask itself is folded into ivar).
//     Absolute log-wavelength does not fit a float (x ~ 8.5 has a float ulp of 0.2 knots), and
//     neither does the position inside a tile (80 knots at 24 bits: 5e-6 knots, which moves a
//     gradient by 1e-5).  So the upload forms (x - x_ref) / dx in double, x_ref = the tile's first x
//     (TileInfo.xmin), and stores its integer part K (a byte) and its fraction f (a float: 6e-8
//     knots); the knot coordinate of x_ref minus the row's shift is formed in double once per row
//     and component and split the same way (Ib, fb); a pixel is then u = fb + f in [0, 2), its
//     interval Ib + K + round(u).  All spline components must share the knot spacing dx.
//   * all per-pixel arithmetic, the register windows, the shared-memory windows and the per-row sums
//     are float; what leaves the kernel (gradient windows, row sums, loss) is double, so the
//     reducers are the float64 ones.
// Reference semantics as in jb_fused2.cuh.
#pragma once
#include "jb_fused2.cuh"

namespace jb {

constexpr int kTileFloats = 3 * kTile;
constexpr int kTileBytes32 = kTileFloats * 4 + kTile;      // f | y | ivar floats, then K bytes
constexpr float kMagic32 = 12582912.0f;          // 1.5 * 2^23: adding it rounds a float to the nearest integer

template <int NC> struct alignas(16) Rec32 { int Ib[NC]; float fb[NC]; float st[NC]; int row; unsigned live; };

template <int PV> struct BasisF;
template <> struct BasisF<1> {
    __device__ __forceinline__ static void eval(float g, float (&w)[2], float (&dw)[1]) {
        w[0] = 0.5f - g; w[1] = 0.5f + g; dw[0] = 1.0f;
    }
};
template <> struct BasisF<2> {
    __device__ __forceinline__ static void eval(float g, float (&w)[3], float (&dw)[2]) {
        const float a = 0.5f + g, b = 0.5f - g;
        w[0] = b * b; w[1] = __fmaf_rn(2.0f * a, b, 1.0f); w[2] = a * a;
        dw[0] = b; dw[1] = a;
    }
};
template <> struct BasisF<3> {
    __device__ __forceinline__ static void eval(float g, float (&w)[4], float (&dw)[3]) {
        const float a = 0.5f + g, b = 0.5f - g;
        const float ab = a * b;
        const float q3 = __fmaf_rn(3.0f, ab, 3.0f);
        const float a2 = a * a, b2 = b * b;
        w[0] = b2 * b; w[1] = __fmaf_rn(b, q3, 1.0f); w[2] = __fmaf_rn(a, q3, 1.0f); w[3] = a2 * a;
        dw[0] = b2; dw[1] = __fmaf_rn(2.0f, ab, 1.0f); dw[2] = a2;
    }
};

template <int NV>
__device__ __forceinline__ float warp_sum_nf(float (&v)[NV], int lane) {
    static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16, "NV");
    int bit = 16;
#pragma unroll
    for (int n = NV / 2; n >= 1; n >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? v[i] : v[i + n];
            const float keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        bit >>= 1;
    }
    for (; bit >= 1; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
    return v[0];
}

template <int PV>
struct Win32 {
    float c[PV + 1];
    float a[PV + 1];
    int I;              // absolute interval the window sits on
    uint32_t pa;        // shared-memory address of the class copy's entry for knot klo
};

__device__ __forceinline__ void lds_if_ne32(float& v, uint32_t addr, int i0, int i1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, %3;\n\t@p ld.shared.f32 %0, [%1];\n\t}"
                 : "+f"(v) : "r"(addr), "r"(i0), "r"(i1));
}
__device__ __forceinline__ void rmw_if32(uint32_t addr, float v, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tsetp.ne.s32 p, %2, 0;\n\tld.shared.f32 t, [%0];\n\t"
                 "add.rn.f32 t, t, %1;\n\t@p st.shared.f32 [%0], t;\n\t}"
                 :: "r"(addr), "f"(v), "r"(pred));
}
__device__ __forceinline__ void step_if_ne32(float& cnew, uint32_t ca, uint32_t ra, float rv, int i0, int i1) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tsetp.ne.s32 p, %4, %5;\n\t@p ld.shared.f32 %0, [%1];\n\t"
                 "ld.shared.f32 t, [%2];\n\tadd.rn.f32 t, t, %3;\n\t@p st.shared.f32 [%2], t;\n\t}"
                 : "+f"(cnew) : "r"(ca), "r"(ra), "f"(rv), "r"(i0), "r"(i1));
}

// abase + 4 * I = address of the class copy's entry for knot klo of absolute interval I
template <int PV>
__device__ __forceinline__ void win_turn32(Win32<PV>& w, int In, uint32_t abase, uint32_t cdelta, bool resync, int lane,
                                           bool first_row) {
    const int d = In - w.I;
    const bool up = d == 1, dn = d == -1;
    const bool far = (unsigned)(d + 1) > 2u || (resync && d != 0);
    if (!first_row && __any_sync(0xffffffffu, far)) {
#pragma unroll 1
        for (int m = 0; m < (32 + kClasses2 - 1) / kClasses2; ++m) {
            const int mine = (far && lane / kClasses2 == m) ? 1 : 0;
#pragma unroll
            for (int k = 0; k <= PV; ++k) rmw_if32(w.pa + 4u * k, w.a[k], mine);
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k <= PV; ++k) w.a[k] = far ? 0.0f : w.a[k];
    }
    rmw_if32(w.pa + (dn ? 4u * PV : 0u), dn ? w.a[PV] : w.a[0], (!far && (up || dn)) ? 1 : 0);
    const bool su = up && !far, sd = dn && !far;
#pragma unroll
    for (int k = 0; k < PV; ++k) w.a[k] = su ? w.a[k + 1] : w.a[k];
    w.a[PV] = su ? 0.0f : w.a[PV];
#pragma unroll
    for (int k = PV; k > 0; --k) w.a[k] = sd ? w.a[k - 1] : w.a[k];
    w.a[0] = sd ? 0.0f : w.a[0];
    const uint32_t pa = abase + 4u * (uint32_t)In;
    const int iold = resync ? In + 1 : w.I;
#pragma unroll
    for (int k = 0; k <= PV; ++k) lds_if_ne32(w.c[k], pa + cdelta + 4u * k, In, iold);
    w.I = In;
    w.pa = pa;
}

template <int PV, int DIR>
__device__ __forceinline__ bool win_step32(Win32<PV>& w, int In, uint32_t abase, uint32_t cdelta) {
    const uint32_t pa = abase + 4u * (uint32_t)In;
    const bool step = In != w.I;
    const float rv = DIR > 0 ? w.a[0] : w.a[PV];
    if (DIR > 0) {
#pragma unroll
        for (int k = 0; k < PV; ++k) w.c[k] = step ? w.c[k + 1] : w.c[k];
        step_if_ne32(w.c[PV], pa + cdelta + 4u * PV, w.pa, rv, In, w.I);
    } else {
#pragma unroll
        for (int k = PV; k > 0; --k) w.c[k] = step ? w.c[k - 1] : w.c[k];
        step_if_ne32(w.c[0], pa + cdelta, w.pa + 4u * PV, rv, In, w.I);
    }
    w.I = In;
    w.pa = pa;
    return step;
}

template <int NC> struct RowConst32 { int Ib[NC]; float fb[NC], st[NC]; };

// u = (fraction of the row's reference knot coordinate) + (fraction of the pixel's offset), in [0, 2)
__device__ __forceinline__ void knot_coord32(float f, float fb, float& g, int& Irel) {
    const float u = __fadd_rn(f, fb);
    const float v = __fadd_rn(u, kMagic32);
    Irel = __float_as_int(v) - 0x4B400000;
    g = __fsub_rn(u, __fsub_rn(v, kMagic32));
}

template <int NC, int PV, int FL, int SLIDE>
__device__ __forceinline__ void f32_pixel(const float (&g)[NC], float y, float iv, const RowConst32<NC>& rc,
                                          Win32<PV> (&win)[NC], const bool (&step)[NC], float& loss,
                                          float (&sums)[f2_pad(f2_nsums(NC, FL))]) {
    float w[NC][PV + 1], S[NC], dS[NC];
    float m = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float dw[PV > 0 ? PV : 1];
        BasisF<PV>::eval(g[c], w[c], dw);
        const Win32<PV>& W = win[c];
        float sv = W.c[0] * w[c][0];
#pragma unroll
        for (int k = 1; k <= PV; ++k) sv = __fmaf_rn(W.c[k], w[c][k], sv);
        S[c] = sv;
        if (f2_der(FL, c)) {
            float d = (W.c[1] - W.c[0]) * dw[0];
#pragma unroll
            for (int k = 1; k < PV; ++k) d = __fmaf_rn(W.c[k + 1] - W.c[k], dw[k], d);
            dS[c] = f2_str(FL, c) ? rc.st[c] * d : d;
        }
        if (c == 0) m = f2_str(FL, c) ? rc.st[c] * sv : sv;
        else m = f2_str(FL, c) ? __fmaf_rn(rc.st[c], sv, m) : m + sv;
    }
    const float res = y - m;
    const float wr = iv * res;
    loss = __fmaf_rn(wr, res, loss);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float q = f2_str(FL, c) ? wr * rc.st[c] : wr;
        Win32<PV>& W = win[c];
        if (SLIDE == 0) {
#pragma unroll
            for (int k = 0; k <= PV; ++k) W.a[k] = __fmaf_rn(q, w[c][k], W.a[k]);
        } else if (SLIDE > 0) {
#pragma unroll
            for (int k = 0; k < PV; ++k) W.a[k] = __fmaf_rn(q, w[c][k], step[c] ? W.a[k + 1] : W.a[k]);
            W.a[PV] = __fmaf_rn(q, w[c][PV], step[c] ? 0.0f : W.a[PV]);
        } else {
#pragma unroll
            for (int k = PV; k > 0; --k) W.a[k] = __fmaf_rn(q, w[c][k], step[c] ? W.a[k - 1] : W.a[k]);
            W.a[0] = __fmaf_rn(q, w[c][0], step[c] ? 0.0f : W.a[0]);
        }
        if (f2_der(FL, c)) {
            const int s0 = f2_slot(NC, FL, c, 0), s2 = f2_slot(NC, FL, c, 2);
            sums[s0] = __fmaf_rn(wr, dS[c], sums[s0]);
            sums[s2] = __fmaf_rn(iv * dS[c], dS[c], sums[s2]);
        }
        if (f2_str(FL, c)) {
            const int s1 = f2_slot(NC, FL, c, 1);
            sums[s1] = __fmaf_rn(wr, S[c], sums[s1]);
        }
    }
}

template <int NC, int PV, int FL, int DIR>
__device__ __forceinline__ void f32_row(const float* px, const uint8_t* pk, const RowConst32<NC>& rc, Win32<PV> (&win)[NC],
                                        const uint32_t (&abase)[NC], uint32_t cdelta, float& loss,
                                        float (&sums)[f2_pad(f2_nsums(NC, FL))], bool live, bool resync, int lane,
                                        bool first_row) {
    float g[NC];
    bool step[NC];
    {
        constexpr int j = DIR > 0 ? 0 : 3;
        const float x = px[32 * j], y = px[kTile + 32 * j], iv = px[2 * kTile + 32 * j];
        const int K = pk[32 * j];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int Irel;
            knot_coord32(x, rc.fb[c], g[c], Irel);
            const int In = live ? rc.Ib[c] + K + Irel : win[c].I;
            win_turn32<PV>(win[c], In, abase[c], cdelta, resync, lane, first_row);
            step[c] = false;
        }
        f32_pixel<NC, PV, FL, 0>(g, y, iv, rc, win, step, loss, sums);
    }
#pragma unroll
    for (int jj = 1; jj < 4; ++jj) {
        const int j = DIR > 0 ? jj : 3 - jj;
        const float x = px[32 * j], y = px[kTile + 32 * j], iv = px[2 * kTile + 32 * j];
        const int K = pk[32 * j];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int Irel;
            knot_coord32(x, rc.fb[c], g[c], Irel);
            const int In = live ? rc.Ib[c] + K + Irel : win[c].I;
            step[c] = win_step32<PV, DIR>(win[c], In, abase[c], cdelta);
        }
        f32_pixel<NC, PV, FL, DIR>(g, y, iv, rc, win, step, loss, sums);
    }
}

// Per-warp shared memory: [ring kStages x kTileBytes32][recs kMaxGroup][mbarriers][pad to 16], then
// the control-point windows [NC][W] and the class copies [kClasses2][NC][W], floats.
template <int NC> struct F32Lay {
    static constexpr uint32_t ring = 0;
    static constexpr uint32_t recs = kStages * kTileBytes32;
    static constexpr uint32_t bars = recs + kMaxGroup * sizeof(Rec32<NC>);
    static constexpr uint32_t ctrl = (bars + kStages * 8 + 15) / 16 * 16;
};
template <int NC>
__host__ __device__ constexpr size_t fused32_warp_smem(int wincap) {
    return (F32Lay<NC>::ctrl + (size_t)NC * wincap * 4 + (size_t)NC * kClasses2 * wincap * 4 + 15) / 16 * 16;
}

struct F32Ctx { unsigned char* wb; int nrow, tile, lane; };

template <int NC, int PV, int FL, int DIR, bool DEFER>
__device__ __forceinline__ void f32_do_row(int i, const F32Ctx& cx, RowConst32<NC>& rc, Win32<PV> (&win)[NC],
                                           const uint32_t (&abase)[NC], uint32_t cdelta, float& loss, bool& was_live,
                                           const unsigned char* data32, float (&sums)[f2_pad(f2_nsums(NC, FL))]) {
    constexpr int NS = f2_nsums(NC, FL), NSP = f2_pad(NS);
    using L = F32Lay<NC>;
    const FusedParams& a = g_f2_plan;
    const int s = i % kStages;
    const Rec32<NC>* recs = reinterpret_cast<const Rec32<NC>*>(cx.wb + L::recs);
    const Rec32<NC>& rec = recs[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) { rc.Ib[c] = rec.Ib[c]; rc.fb[c] = rec.fb[c]; rc.st[c] = rec.st[c]; }
    const unsigned live_mask = rec.live;
    const uint32_t bar = smem_u32(cx.wb + L::bars) + 8 * s;
    mbar_wait(bar, (uint32_t)((i / kStages) & 1));
#pragma unroll
    for (int q = 0; q < NSP; ++q) sums[q] = 0.0f;
    if (live_mask != 0) {
        const float* px = reinterpret_cast<const float*>(cx.wb + L::ring + (size_t)s * kTileBytes32) + cx.lane;
        const bool live = (live_mask >> cx.lane) & 1u;
        uint32_t ab[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) ab[c] = abase[c];
        const uint8_t* pk = cx.wb + L::ring + (size_t)s * kTileBytes32 + kTileFloats * 4 + cx.lane;
        f32_row<NC, PV, FL, DIR>(px, pk, rc, win, ab, cdelta, loss, sums, live, !was_live, cx.lane, DIR > 0 && i == 0);
        was_live = live;
    } else was_live = false;
    __syncwarp();
    if (cx.lane == 0 && i + kStages < cx.nrow) {
        mbar_expect_tx(bar, kTileBytes32);
        tma_load_1d(smem_u32(cx.wb + L::ring + (size_t)s * kTileBytes32),
                    data32 + ((int64_t)recs[i + kStages].row * a.n_tiles + cx.tile) * kTileBytes32, kTileBytes32, bar);
    }
    if (NS > 0 && !DEFER) {
        const float tot = warp_sum_nf<NSP>(sums, cx.lane);
        if ((cx.lane & (32 / NSP - 1)) == 0)
            a.rowpart[((int64_t)recs[i].row * a.n_tiles + cx.tile) * NSP + cx.lane / (32 / NSP)] = (double)tot;
    }
}

// the sums of rows i (su) and i + 1 (sd) in one butterfly (see f2_store_pair)
template <int NC, int FL>
__device__ __forceinline__ void f32_store_pair(int i, const F32Ctx& cx, const float (&su)[f2_pad(f2_nsums(NC, FL))],
                                               const float (&sd)[f2_pad(f2_nsums(NC, FL))]) {
    constexpr int NSP = f2_pad(f2_nsums(NC, FL));
    using L = F32Lay<NC>;
    const FusedParams& a = g_f2_plan;
    const Rec32<NC>* recs = reinterpret_cast<const Rec32<NC>*>(cx.wb + L::recs);
    float v[2 * NSP];
#pragma unroll
    for (int q = 0; q < NSP; ++q) { v[q] = su[q]; v[NSP + q] = sd[q]; }
    const float tot = warp_sum_nf<2 * NSP>(v, cx.lane);
    constexpr int LPV = 32 / (2 * NSP);
    if ((cx.lane & (LPV - 1)) == 0) {
        const int q = cx.lane / LPV;
        const int row = recs[i + (q >= NSP ? 1 : 0)].row;
        a.rowpart[((int64_t)row * a.n_tiles + cx.tile) * NSP + (q & (NSP - 1))] = (double)tot;
    }
}

// FusedParams.data points at the float records for a float32 plan.
#ifndef JB_MAXREG32
#define JB_MAXREG32 96      // 5 warps per scheduler (16384 registers each): 20 resident warps
#endif
template <int NC, int PV, int FL>
__global__ void __maxnreg__(JB_MAXREG32) k_fused32(const FusedParams* __restrict__ arr, const int* __restrict__ off, int n) {
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
    using L = F32Lay<NC>;
    const int W = a.wincap;
    const unsigned char* data32 = reinterpret_cast<const unsigned char*>(a.data);

    unsigned char* wbase = smem_raw + (size_t)warp * fused32_warp_smem<NC>(W);
    unsigned char* ring = wbase + L::ring;
    Rec32<NC>* recs = reinterpret_cast<Rec32<NC>*>(wbase + L::recs);
    const uint32_t bar0 = smem_u32(wbase + L::bars);
    float* ctrl_warp = reinterpret_cast<float*>(wbase + L::ctrl);           // [NC][W]
    float* acc_warp = ctrl_warp + NC * W;                                   // [kClasses2][NC][W]

    // ---- 0. lane i prepares the record of row i: knot coordinate of the tile's reference x -------
    int lo[NC], hi[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { lo[c] = INT_MAX; hi[c] = INT_MIN; }
    bool has_work = false;
    if (lane < nrow) {
        const RowRec<NC, 0> rr = reinterpret_cast<const RowRec<NC, 0>*>(a.grouprec)[row0 + lane];
        const TileInfo ti = a.tinfo[(int64_t)tile * a.n_rows + row0 + lane];
        Rec32<NC> rec;
        rec.row = rr.row; rec.live = ti.live;
        has_work = ti.live != 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const CompDev& cd = a.comp[c];
            rec.st[c] = (float)rr.st[c];
            rec.Ib[c] = 0; rec.fb[c] = 0.0f;
            if (has_work) {
                // u of a pixel = t - 1/2 (odd degree) or t (even), t = ((x - shift) - xp0) / dx: the part
                // of it that belongs to the tile's reference x, in double, split into knot + fraction
                const double tb = ((ti.xmin - rr.sh[c]) - cd.xp0) * cd.inv_dx - ((PV & 1) ? 0.5 : 0.0);
                const double te = ((ti.xmax - rr.sh[c]) - cd.xp0) * cd.inv_dx - ((PV & 1) ? 0.5 : 0.0);
                if (!(tb > -1.0e9 && te < 1.0e9)) { lo[c] = -(1 << 30); hi[c] = (1 << 30); }
                else {
                    const double fl = floor(tb);
                    rec.Ib[c] = (int)fl;
                    rec.fb[c] = (float)(tb - fl);
                    lo[c] = knot_lo<PV>((int)rint(tb));          // knots the row touches (exact arithmetic)
                    hi[c] = knot_lo<PV>((int)rint(te)) + PV;
                }
            }
        }
        recs[lane] = rec;
    }
    const bool empty = !__any_sync(0xffffffffu, has_work);
    __syncwarp();
    if (lane == 0 && !empty) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_fence_init();
        for (int i = 0; i < kStages && i < nrow; ++i) {
            const uint32_t bar = bar0 + 8 * (i % kStages);
            mbar_expect_tx(bar, kTileBytes32);
            tma_load_1d(smem_u32(ring + (size_t)(i % kStages) * kTileBytes32),
                        data32 + ((int64_t)recs[i].row * a.n_tiles + tile) * kTileBytes32, kTileBytes32, bar);
        }
    }
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
            // one knot of slack either side: the float rounding of u can put a pixel that sits on a
            // knot into the neighbouring interval (same value there, the spline is continuous)
            lo[c] = max(lo[c] - 1, 0);
            hi[c] = min(hi[c] + 1, a.comp[c].n_knots - 1);
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
    if (skip) {
        for (int i = lane; i < nrow * NSP; i += 32)
            a.rowpart[((int64_t)recs[i / NSP].row * a.n_tiles + tile) * NSP + i % NSP] = 0.0;
        if (!empty)
            for (int i = 0; i < kStages && i < nrow; ++i) mbar_wait(bar0 + 8 * (i % kStages), 0);
        return;
    }

    for (int i = lane; i < NC * kClasses2 * W; i += 32) acc_warp[i] = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int nk = hi[c] - lo[c] + 1;
        for (int k = lane; k < W; k += 32) ctrl_warp[c * W + k] = k < nk ? (float)__ldg(a.comp[c].ctrl + lo[c] + k) : 0.0f;
    }
    __syncwarp();

    constexpr int KOFF = (PV + 1) / 2 - PV;
    const int cls = lane % kClasses2;
    uint32_t abase[NC];
    Win32<PV> win[NC];
    RowConst32<NC> rc;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const uint32_t A = smem_u32(acc_warp + (size_t)(cls * NC + c) * W);
        abase[c] = A + 4u * (uint32_t)(KOFF - lo[c]);
        asm volatile("" : "+r"(abase[c]));
        win[c].I = lo[c] - KOFF;
        win[c].pa = A;
#pragma unroll
        for (int k = 0; k <= PV; ++k) { win[c].a[k] = 0.0f; win[c].c[k] = 0.0f; }
    }
    const uint32_t cdelta = smem_u32(ctrl_warp) - smem_u32(acc_warp + (size_t)cls * NC * W);

    float loss = 0.0f;
    F32Ctx cx;
    cx.wb = wbase; cx.nrow = nrow; cx.tile = tile; cx.lane = lane;
    bool was_live = false;
    for (int i = 0; i < nrow; i += 2) {
        constexpr int NSPr = f2_pad(f2_nsums(NC, FL));
        constexpr bool kPair = JB_F2_PAIRSUM != 0 && f2_nsums(NC, FL) > 0 && 2 * NSPr <= 16;
        float su[NSPr], sd[NSPr];
        if (i + 1 < nrow) {
            f32_do_row<NC, PV, FL, +1, kPair>(i, cx, rc, win, abase, cdelta, loss, was_live, data32, su);
            f32_do_row<NC, PV, FL, -1, kPair>(i + 1, cx, rc, win, abase, cdelta, loss, was_live, data32, sd);
            if (kPair) f32_store_pair<NC, FL>(i, cx, su, sd);
        } else {
            f32_do_row<NC, PV, FL, +1, false>(i, cx, rc, win, abase, cdelta, loss, was_live, data32, su);
        }
    }

    if (__all_sync(0xffffffffu, was_live)) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k <= PV; ++k) rmw_if32(win[c].pa + 4u * k, win[c].a[k], 1);
    } else {
#pragma unroll 1
        for (int m = 0; m < (32 + kClasses2 - 1) / kClasses2; ++m) {
            const int mine = lane / kClasses2 == m ? 1 : 0;
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int k = 0; k <= PV; ++k) rmw_if32(win[c].pa + 4u * k, win[c].a[k], mine);
            __syncwarp();
        }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double* dst = a.part + ((int64_t)c * a.n_tasks + task) * W;
        for (int k = lane; k < W; k += 32) {
            double sacc = (double)acc_warp[(size_t)c * W + k];
#pragma unroll
            for (int q = 1; q < kClasses2; ++q) sacc += (double)acc_warp[(size_t)(q * NC + c) * W + k];
            dst[k] = sacc;
        }
    }
    double dl = (double)loss;
    dl = warp_sum(dl);
    if (lane == 0) a.losspart[task] = dl;
}

}  // namespace jb
