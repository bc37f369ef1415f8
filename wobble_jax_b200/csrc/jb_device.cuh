// Device-side building blocks shared by the kernels of libjabble_b200 (sm_100a only).
//
// Spline evaluation restates, in local-interval form, what the reference computes with a
// (2a+3)-tap gather:  cardinal_vmap_model (jabble/model.py:951-989) with the basis of
// CardinalSplineKernel (jabble/cardinalspline.py:44-55).  Only the p_val+1 non-zero taps are
// evaluated, as closed-form polynomials of the in-interval coordinate.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jb {

constexpr int kMaxComp = 3;       // additive components per plan
constexpr int kTile = 128;        // pixels per warp tile
constexpr int kPixPerLane = 4;    // consecutive pixels owned by one lane (one 256-bit load per array)
constexpr int kWinCapMax = 256;   // knots in a warp's on-chip gradient window: chosen per plan (multiple of 32)
constexpr int kClasses = 4;       // lane residue classes with private accumulator copies
// Tuning knobs (measured on B200, profiles/README.md): 2 warps (tasks) per block, 128 registers,
// a 2-deep copy ring and a rolled pixel loop give 14 resident warps per SM; every larger-register /
// unrolled variant tried was slower.
#ifndef JB_WARPS
#define JB_WARPS 2
#endif
#ifndef JB_MAXREG
#define JB_MAXREG 128
#endif
#ifndef JB_UNROLLED
#define JB_ROLLED 1
#endif
constexpr int kWarpsPerBlock = JB_WARPS;   // tasks per block
constexpr int kRowQ = 3;          // per-component per-row sums: d/dshift, d/dstretch, Fisher
#ifndef JB_STAGES
#define JB_STAGES 2
#endif
constexpr int kStages = JB_STAGES;        // rows in flight per warp (TMA bulk-copy ring)
constexpr int kStageBytes = kTile * 8 * 3 + kTile;   // x, y, ivar (f64) + mask (u8) of one row tile
constexpr int kMaxGroup = 32;     // rows per group: lane i keeps the parameters of row i

// status bits written by kernels into plan->d_status[0]
constexpr int kStatOutOfRange = 1;
constexpr int kStatWindow = 2;
constexpr int kStatBounds = 4;      // raised only by JB_DEBUG_BOUNDS builds: a window address left its class copy

struct CompDev {
    int kind;                 // jb_comp_kind
    int p_val;
    int n_knots;              // M, or K per key
    int pad_;
    double xp0;
    double inv_dx;
    const double* ctrl;       // control points (params_all + ctrl_offset)
    const int* ctrl_key;      // per-key splines only
    const double* shift;      // nullptr when the component has no ShiftingModel
    const int* shift_key;
    const double* stretch;    // nullptr when no StretchingModel
    const int* stretch_key;
};

// 1.5 * 2^52: adding it rounds a double to the nearest integer (ties to even) and leaves that
// integer in the low 32 bits of the mantissa.
constexpr double kMagic = 6755399441055744.0;

// Warped knot coordinate of one pixel.
//   xs = x - shift                     ShiftingModel.call        model.py:927
//   t  = (xs - xp[0]) / dx             cardinal_vmap_model       model.py:979-981
// The reference splits t into index = floor and inputs = t % 1 and then multiplies 2a+3 taps,
// of which p_val+1 are non-zero.  Here: odd p_val -> I = floor(t), even p_val -> I = round(t)
// (the interval of the piecewise polynomial), g = in-interval coordinate - 1/2 in [-1/2, 1/2].
// The subtraction order (x - shift first, then - xp0) is the reference's: it decides the
// last-bit rounding of xs, which is worth 2e-10 knots at x ~ 8.5.
template <int PV>
__device__ __forceinline__ void knot_coord(double x, double shift, double xp0, double inv_dx,
                                           double& g, int& I) {
    const double xs = __dsub_rn(x, shift);
    const double d = __dsub_rn(xs, xp0);
    const double u = (PV & 1) ? __fma_rn(d, inv_dx, -0.5) : __dmul_rn(d, inv_dx);
    const double v = __dadd_rn(u, kMagic);
    I = __double2loint(v);
    g = __dsub_rn(u, __dsub_rn(v, kMagic));
}

// Non-zero basis weights of the p_val! * Irwin-Hall B-spline on interval I, ordered by knot:
// w[j] multiplies control point klo + j, klo = I + (PV+1)/2 - PV.  dw[j] multiplies the
// difference c[j+1]-c[j]; the derivative w.r.t. the knot coordinate is PV * sum_j dw[j]*(c[j+1]-c[j]).
template <int PV> struct Basis;

template <> struct Basis<1> {
    __device__ __forceinline__ static void eval(double g, double (&w)[2], double (&dw)[1]) {
        w[0] = 0.5 - g;
        w[1] = 0.5 + g;
        dw[0] = 1.0;
    }
};

template <> struct Basis<2> {
    __device__ __forceinline__ static void eval(double g, double (&w)[3], double (&dw)[2]) {
        const double a = 0.5 + g, b = 0.5 - g;
        w[0] = b * b;
        w[1] = __fma_rn(2.0 * a, b, 1.0);
        w[2] = a * a;
        dw[0] = b;
        dw[1] = a;
    }
};

template <> struct Basis<3> {
    __device__ __forceinline__ static void eval(double g, double (&w)[4], double (&dw)[3]) {
        const double a = 0.5 + g, b = 0.5 - g;
        const double ab = a * b;
        const double q3 = __fma_rn(3.0, ab, 3.0);
        const double a2 = a * a, b2 = b * b;
        w[0] = b2 * b;
        w[1] = __fma_rn(b, q3, 1.0);
        w[2] = __fma_rn(a, q3, 1.0);
        w[3] = a2 * a;
        dw[0] = b2;
        dw[1] = __fma_rn(2.0, ab, 1.0);
        dw[2] = a2;
    }
};

template <int PV>
__device__ __forceinline__ int knot_lo(int I) { return I + (PV + 1) / 2 - PV; }

// 256-bit streaming loads: read-only path, do not allocate in L1 (each byte is used once).
__device__ __forceinline__ void ldg256(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ uint32_t ldg32_stream(const void* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine), shared::cta addresses as 32-bit ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "JB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra JB_DONE;\n"
        "bra JB_WAIT;\n"
        "JB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Sum NV (4 or 8) per-lane values over the warp with NV+1 shuffles instead of 5*NV: each step
// halves the number of values a lane carries.  Afterwards lane L holds the total of value
// q(L) = bits 4..(5-log2 NV) of L (replicated on the other lanes); fixed order -> deterministic.
template <int NV>
__device__ __forceinline__ double warp_sum_multi(double (&v)[NV], int lane) {
    static_assert(NV == 4 || NV == 8, "NV");
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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace jb
