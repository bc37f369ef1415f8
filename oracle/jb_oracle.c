/*
 * jb_oracle.c -- plain-C restatement of the reference's loss / gradient / Fisher path.
 *
 * TEST INFRASTRUCTURE ONLY (checker and CPU baseline): only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.  Nothing under wobble_jax_b200/
 * links or calls it.
 *
 * It follows the reference literally where rounding matters:
 *   jabble/model.py:924-927     ShiftingModel.call      xs = x - p[key]
 *   jabble/model.py:974-985     cardinal_vmap_model     dx = xp[1]-xp[0]; inputs = ((x-xp0)/dx) % 1;
 *                                                        index = (x-xp0)//dx; taps floor(arange(-a-1,a+2))
 *   jabble/cardinalspline.py:44-55  CardinalSplineKernel.__call__  ks = floor(t+(n+1)/2), Horner over
 *                                                        alphas[::-1,ks], zero outside 0<=ks<=n
 *   jabble/cardinalspline.py:6-18   _alpha_recursion
 *   jabble/model.py:943-946     StretchingModel.call    p[key] * x
 *   jabble/model.py:690-697     AdditiveModel.call      sum of components
 *   jabble/loss.py:164-173      ChiSquare               coef*where(~mask, 0.5*ivar*(y-m)^2, 0)
 *   jabble/loss.py:47-75        loss_all                sum over rows
 *   jabble/model.py:228         jax.grad(loss_all)      -> analytic adjoint of exactly these ops
 *   jabble/model.py:857-901     f_info                  sum ivar*(d model/d shift)^2, unmasked
 * The gradient is the analytic derivative of the piecewise polynomial (floor/astype carry no
 * cotangent, % 1 has unit derivative, where() blocks the masked branch: SURVEY.md section 3.3);
 * tests/test_oracle_c.py checks it against the torch-autograd restatement (oracle/jabble_oracle.py).
 *
 * Parallelism: OpenMP over rows, per-thread gradient buffers summed in thread order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t kind, p_val;          /* 0 global spline, 1 per-key spline */
    double xp0, dx;
    int64_t n_knots, ctrl_offset, n_ctrl_keys;
    const int32_t *ctrl_key;
    int64_t shift_offset, n_shift;
    const int32_t *shift_key;
    int64_t stretch_offset, n_stretch;
    const int32_t *stretch_key;
} jbo_comp;

static double comb(int n, int k) {
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return floor(r + 0.5);
}

/* cardinalspline.py:6-18 */
static double alpha_recursion(int j, int k, int n) {
    if (k == 0) return (j < n - 1) ? 0.0 : 1.0;
    double sign = ((n + k - j - 1) % 2 == 0) ? 1.0 : -1.0;
    return alpha_recursion(j, k - 1, n) + sign * comb(n, k) * comb(n - 1, j) * pow((double)k, n - j - 1);
}

void jbo_alphas(int p_val, double *out /* (n+1)*(n+1), [j][k] */) {
    int n = p_val;
    for (int j = 0; j <= n; ++j)
        for (int k = 0; k <= n; ++k) out[j * (n + 1) + k] = alpha_recursion(j, k, n + 1);
}

/* Python-style floor division of floats (what jnp.floor_divide / numpy do) */
static double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0) != (mod < 0))) div -= 1.0;
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, a / b);
}

static double py_mod1(double t) {
    double m = fmod(t, 1.0);
    if (m != 0.0 && m < 0) m += 1.0;
    return m;
}

/* value and d/dt of CardinalSplineKernel at t (cardinalspline.py:44-55) */
static inline void basis(const double *A, int n, double t, double *val, double *der) {
    double s = t + (n + 1) / 2.0;
    double ksf = floor(s);
    if (!(ksf >= 0.0 && ksf <= (double)n)) { *val = 0.0; *der = 0.0; return; }
    int ks = (int)ksf;
    double v = 0.0, d = 0.0;
    for (int j = n; j >= 0; --j) {            /* Horner, highest power first; d is its derivative */
        d = d * s + v;
        v = v * s + A[j * (n + 1) + ks];
    }
    *val = v; *der = d;
}

/*
 * loss, gradient w.r.t. every parameter, Fisher information of every shift parameter.
 * grad_all / fisher_all: n_params doubles (may be NULL).  Returns 0, or -1 if a gather index of an
 * unmasked pixel leaves the knot grid (the reference would clamp silently).
 */
int jbo_eval(int64_t E, int64_t P, const double *xs, const double *ys, const double *yivar,
             const uint8_t *mask, int n_comp, const jbo_comp *comps, const double *params,
             int64_t n_params, double coef, int nthreads, double *loss_out, double *grad_all,
             double *fisher_all) {
    double A[3][25];
    for (int c = 0; c < n_comp; ++c) jbo_alphas(comps[c].p_val, A[c]);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double *gbuf = (double *)calloc((size_t)nthreads * (size_t)n_params * 2 + (size_t)nthreads, sizeof(double));
    if (!gbuf) return -2;
    double *lbuf = gbuf + (size_t)nthreads * n_params * 2;
    int bad = 0;
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *g = gbuf + (size_t)tid * n_params * 2;
        double *f = g + n_params;
        double lsum = 0.0;
#pragma omp for schedule(static)
        for (int64_t r = 0; r < E; ++r) {
            double rowloss = 0.0;
            for (int64_t p = 0; p < P; ++p) {
                const int64_t i = r * P + p;
                if (mask && mask[i]) continue;
                double S[3], dS[3], stv[3];
                int64_t idx[3]; double frac[3];
                double m = 0.0;
                for (int c = 0; c < n_comp; ++c) {
                    const jbo_comp *cd = &comps[c];
                    const int n = cd->p_val;
                    double x = xs[i];
                    if (cd->shift_offset >= 0) x = x - params[cd->shift_offset + (cd->shift_key ? cd->shift_key[r] : r)];
                    const double d = x - cd->xp0;
                    frac[c] = py_mod1(d / cd->dx);
                    idx[c] = (int64_t)py_floordiv(d, cd->dx);
                    const double *ap = params + cd->ctrl_offset +
                        (cd->kind == 1 ? (int64_t)(cd->ctrl_key ? cd->ctrl_key[r] : r) * cd->n_knots : 0);
                    const double a = (n + 1) / 2.0;
                    double s = 0.0, ds = 0.0;
                    for (double kk = -a - 1; kk < a + 2; kk += 1.0) {      /* model.py:977 */
                        const int k = (int)floor(kk);
                        double w, dw;
                        basis(A[c], n, frac[c] + k, &w, &dw);
                        if (w == 0.0 && dw == 0.0) continue;
                        const int64_t j = idx[c] - k;
                        if (j < 0 || j >= cd->n_knots) { bad = 1; continue; }
                        s += ap[j] * w; ds += ap[j] * dw;
                    }
                    S[c] = s; dS[c] = ds / cd->dx;
                    stv[c] = cd->stretch_offset >= 0 ? params[cd->stretch_offset + (cd->stretch_key ? cd->stretch_key[r] : r)] : 1.0;
                    m += stv[c] * s;
                }
                const double res = ys[i] - m;
                const double iv = yivar[i];
                rowloss += coef * 0.5 * iv * res * res;
                const double wr = coef * iv * res;
                for (int c = 0; c < n_comp; ++c) {
                    const jbo_comp *cd = &comps[c];
                    const int n = cd->p_val;
                    const double a = (n + 1) / 2.0;
                    double *gp = g + cd->ctrl_offset +
                        (cd->kind == 1 ? (int64_t)(cd->ctrl_key ? cd->ctrl_key[r] : r) * cd->n_knots : 0);
                    for (double kk = -a - 1; kk < a + 2; kk += 1.0) {
                        const int k = (int)floor(kk);
                        double w, dw;
                        basis(A[c], n, frac[c] + k, &w, &dw);
                        const int64_t j = idx[c] - k;
                        if (w == 0.0 || j < 0 || j >= cd->n_knots) continue;
                        gp[j] -= wr * stv[c] * w;
                    }
                    if (cd->shift_offset >= 0) {
                        const int64_t kidx = cd->shift_offset + (cd->shift_key ? cd->shift_key[r] : r);
                        g[kidx] += wr * stv[c] * dS[c];
                        f[kidx] += iv * (stv[c] * dS[c]) * (stv[c] * dS[c]);
                    }
                    if (cd->stretch_offset >= 0)
                        g[cd->stretch_offset + (cd->stretch_key ? cd->stretch_key[r] : r)] -= wr * S[c];
                }
            }
            lsum += rowloss;
        }
        lbuf[tid] = lsum;
    }
    double loss = 0.0;
    for (int t = 0; t < nthreads; ++t) loss += lbuf[t];
    *loss_out = loss;
    if (grad_all) {
        memset(grad_all, 0, (size_t)n_params * sizeof(double));
        for (int t = 0; t < nthreads; ++t)
            for (int64_t k = 0; k < n_params; ++k) grad_all[k] += gbuf[(size_t)t * n_params * 2 + k];
    }
    if (fisher_all) {
        memset(fisher_all, 0, (size_t)n_params * sizeof(double));
        for (int t = 0; t < nthreads; ++t)
            for (int64_t k = 0; k < n_params; ++k) fisher_all[k] += gbuf[(size_t)t * n_params * 2 + n_params + k];
    }
    free(gbuf);
    return bad ? -1 : 0;
}

int jbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
