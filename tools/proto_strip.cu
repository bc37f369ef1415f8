// Stand-alone check + timing of k_strip (wobble_jax_b200/csrc/jb_strip.cuh) on the benchmark shape,
// before it is wired into the library: synthetic chunk(s) of E x P pixels in strip layout, a naive
// one-thread-per-pixel kernel with atomics as the checker, CUDA-event timing of a C-chunk launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o gpurun_out/proto_strip tools/proto_strip.cu
//   gpurun_out/proto_strip [E] [P] [chunks] [K] [blocks_per_task]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../wobble_jax_b200/csrc/jb_strip.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

using namespace jb;

struct NaiveP {
    const double* data; const StripRow<2>* rowc; int64_t E; int n_strips; CompDev comp[2];
    double* g[2]; double* rows; double* loss;
};
__global__ void k_naive(NaiveP a) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = a.E * a.n_strips * 32;
    if (idx >= total) return;
    const int lane = idx % 32; const int64_t row = (idx / 32) % a.E; const int strip = idx / 32 / a.E;
    const double* rec = a.data + ((int64_t)strip * a.E + row) * kStripRec;
    const double x = rec[lane], y = rec[32 + lane], iv = rec[64 + lane];
    if (iv == 0.0) return;
    const StripRow<2> rc = a.rowc[row];
    double w[2][4], dw[3], S[2], dS = 0; int klo[2];
    for (int c = 0; c < 2; ++c) {
        double g; int I;
        knot_coord<3>(x, rc.sh[c], a.comp[c].xp0, a.comp[c].inv_dx, g, I);
        klo[c] = knot_lo<3>(I);
        double dwc[3];
        Basis<3>::eval(g, w[c], dwc);
        const double* cp = a.comp[c].ctrl + klo[c];
        S[c] = cp[0] * w[c][0] + cp[1] * w[c][1] + cp[2] * w[c][2] + cp[3] * w[c][3];
        if (c == 0) { for (int k = 0; k < 3; ++k) dw[k] = dwc[k]; dS = (cp[1] - cp[0]) * dw[0] + (cp[2] - cp[1]) * dw[1] + (cp[3] - cp[2]) * dw[2]; }
    }
    const double m = S[0] + rc.st[1] * S[1];
    const double res = y - m, wr = iv * res;
    atomicAdd(a.loss, wr * res);
    for (int k = 0; k < 4; ++k) { atomicAdd(a.g[0] + klo[0] + k, wr * w[0][k]); atomicAdd(a.g[1] + klo[1] + k, wr * rc.st[1] * w[1][k]); }
    atomicAdd(a.rows + row * 4 + 0, wr * dS);
    atomicAdd(a.rows + row * 4 + 1, iv * dS * dS);
    atomicAdd(a.rows + row * 4 + 2, wr * S[1]);
}

struct FoldP { const double* part; const int* base; const double* rowpart; const double* losspart; const int* rowid;
               int64_t n_tasks; int W; int n_strips; int64_t E; double* g[2]; double* rows; double* loss; };
__global__ void k_fold(FoldP a) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < 2 * a.n_tasks * a.W) {
        const int c = i / (a.n_tasks * a.W); const int64_t t = (i / a.W) % a.n_tasks; const int k = i % a.W;
        const int b = a.base[c * a.n_tasks + t];
        if (b != INT_MAX && a.part[i] != 0.0) atomicAdd(a.g[c] + b + k, a.part[i]);
    }
    if (i < a.E * a.n_strips * 3) { const int64_t r = i / (a.n_strips * 3); const int q = i % 3; atomicAdd(a.rows + r * 4 + q, a.rowpart[i]); }
    if (i < a.n_tasks) atomicAdd(a.loss, a.losspart[i]);
}

int main(int argc, char** argv) {
    const int64_t E = argc > 1 ? atoll(argv[1]) : 2000;
    const int64_t P = argc > 2 ? atoll(argv[2]) : 4096;
    const int C = argc > 3 ? atoi(argv[3]) : 8;
    const int K = argc > 4 ? atoi(argv[4]) : 16;
    const int BPT = argc > 5 ? atoi(argv[5]) : 8;
    const int reps = argc > 6 ? atoi(argv[6]) : 10;
    const int n_strips = (int)(P / 32);
    const double Delta = 2.7297e-6, dx = std::log(1.0 + 1.0 / 230000.0), cl = 299792458.0;
    std::mt19937_64 rng(20260101);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> N(0.0, 1.0);
    const double x0 = std::log(4000.0);
    const double pad = 100e3 / cl;           // vel_padding
    const double xp0 = x0 - pad - 2 * dx;
    const int M = (int)((P * Delta + 2 * pad) / dx) + 8;
    // rows: jitter, shifts; sorted by first x - shift
    std::vector<double> jit(E), sh0(E), am(E), key(E);
    for (int64_t e = 0; e < E; ++e) {
        jit[e] = Delta * (U(rng) - 0.5);
        const double v = 30e3 * std::sin(2 * M_PI * e / (double)E) + 400 * (U(rng) - 0.5);
        sh0[e] = 0.5 * std::log((1 + v / cl) / (1 - v / cl));
        am[e] = 1.0 + U(rng);
        key[e] = x0 + jit[e] - sh0[e];
    }
    std::vector<int> perm(E);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
    const int n_blocks = (int)((E + K - 1) / K);
    std::vector<double> hdata((size_t)n_strips * E * kStripRec), hxoff((size_t)n_strips * n_blocks * 32), hxref(E);
    for (int64_t pos = 0; pos < E; ++pos) hxref[pos] = x0 + jit[perm[pos]];
    for (int s = 0; s < n_strips; ++s) {
        for (int b = 0; b < n_blocks; ++b)
            for (int l = 0; l < 32; ++l) hxoff[((size_t)s * n_blocks + b) * 32 + l] = Delta * (s * 32 + l);
        for (int64_t pos = 0; pos < E; ++pos) {
            double* rec = hdata.data() + ((size_t)s * E + pos) * kStripRec;
            for (int l = 0; l < 32; ++l) {
                rec[l] = hxref[pos] + Delta * (s * 32 + l);
                const bool masked = U(rng) < 0.01;
                rec[32 + l] = masked ? 0.0 : -0.3 + 0.01 * N(rng);
                rec[64 + l] = masked ? 0.0 : 1.0e4;
            }
        }
    }
    std::vector<double> hparams(2 * M + 2 * E);
    for (int i = 0; i < 2 * M; ++i) hparams[i] = -0.15 + 0.05 * N(rng);
    for (int64_t e = 0; e < E; ++e) { hparams[2 * M + e] = sh0[e]; hparams[2 * M + E + e] = am[e]; }
    std::vector<int> ident(E);
    std::iota(ident.begin(), ident.end(), 0);

    const int ncls = 2, W = 64;
    const int n_rgroups = (n_blocks + BPT - 1) / BPT;
    const int64_t n_tasks = (int64_t)n_rgroups * n_strips;
    printf("E %lld P %lld chunks %d K %d blocks/task %d -> %lld tasks per chunk, M %d knots, strips %d\n", (long long)E, (long long)P, C, K, BPT, (long long)n_tasks, M, n_strips);

    std::vector<StripParams> sp(C); std::vector<StripPrepParams> pp(C);
    std::vector<int> off_strip(C + 1, 0), off_prep(C + 1, 0);
    double *d_params, *d_xoff, *d_xref; int *d_rowid, *d_key;
    CK(cudaMalloc(&d_params, hparams.size() * 8)); CK(cudaMemcpy(d_params, hparams.data(), hparams.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_xoff, hxoff.size() * 8)); CK(cudaMemcpy(d_xoff, hxoff.data(), hxoff.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_xref, E * 8)); CK(cudaMemcpy(d_xref, hxref.data(), E * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_rowid, E * 4)); CK(cudaMemcpy(d_rowid, perm.data(), E * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_key, E * 4)); CK(cudaMemcpy(d_key, ident.data(), E * 4, cudaMemcpyHostToDevice));
    int* d_status; CK(cudaMalloc(&d_status, 8)); CK(cudaMemset(d_status, 0, 8));
    std::vector<double> hrng((size_t)n_rgroups * n_strips * 2);
    for (int g = 0; g < n_rgroups; ++g)
        for (int s = 0; s < n_strips; ++s) { hrng[((size_t)g * n_strips + s) * 2] = Delta * (s * 32); hrng[((size_t)g * n_strips + s) * 2 + 1] = Delta * (s * 32 + 31); }
    double* d_rng; CK(cudaMalloc(&d_rng, hrng.size() * 8)); CK(cudaMemcpy(d_rng, hrng.data(), hrng.size() * 8, cudaMemcpyHostToDevice));
    for (int i = 0; i < C; ++i) {
        double* d_data; CK(cudaMalloc(&d_data, hdata.size() * 8));
        CK(cudaMemcpy(d_data, hdata.data(), hdata.size() * 8, cudaMemcpyHostToDevice));
        StripParams& s = sp[i]; memset(&s, 0, sizeof s);
        s.data = d_data; s.xoff = d_xoff; s.xoff_rng = d_rng; s.rowid = d_rowid;
        void* rowc; CK(cudaMalloc(&rowc, E * sizeof(StripRow<2>))); s.rowc = rowc;
        StripBlk* blk; CK(cudaMalloc(&blk, n_blocks * sizeof(StripBlk))); s.blk = blk;
        s.n_rows = E; s.n_strips = n_strips; s.K = K; s.blocks_per_task = BPT; s.n_blocks = n_blocks; s.n_rgroups = n_rgroups; s.ncls = ncls;
        s.n_tasks = n_tasks; s.wincap = W; s.n_comp = 2; s.tol = 0.004;
        for (int c = 0; c < 2; ++c) {
            CompDev& cd = s.comp[c]; memset(&cd, 0, sizeof cd);
            cd.p_val = 3; cd.n_knots = M; cd.xp0 = xp0; cd.inv_dx = 1.0 / dx; cd.ctrl = d_params + c * M;
        }
        s.comp[0].shift = d_params + 2 * M; s.comp[0].shift_key = d_key;
        s.comp[1].stretch = d_params + 2 * M + E; s.comp[1].stretch_key = d_key;
        CK(cudaMalloc(&s.part, 2 * n_tasks * W * 8)); CK(cudaMalloc(&s.part_base, 2 * n_tasks * 4));
        CK(cudaMalloc(&s.rowpart, E * n_strips * 3 * 8)); CK(cudaMalloc(&s.losspart, n_tasks * 8));
        s.status = d_status;
        StripPrepParams& q = pp[i]; memset(&q, 0, sizeof q);
        q.rowid = d_rowid; q.xref = d_xref; q.rowc = rowc; q.blk = blk; q.n_rows = E; q.K = K; q.n_blocks = n_blocks; q.n_comp = 2;
        q.comp[0] = s.comp[0]; q.comp[1] = s.comp[1]; q.status = nullptr; q.tol = 0.004;
        off_strip[i + 1] = off_strip[i] + (int)((n_tasks + kStripWarps - 1) / kStripWarps);
        off_prep[i + 1] = off_prep[i] + (n_blocks + 3) / 4;
    }
    StripParams* d_sp; StripPrepParams* d_pp; int *d_off_s, *d_off_p;
    CK(cudaMalloc(&d_sp, C * sizeof(StripParams))); CK(cudaMemcpy(d_sp, sp.data(), C * sizeof(StripParams), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_pp, C * sizeof(StripPrepParams))); CK(cudaMemcpy(d_pp, pp.data(), C * sizeof(StripPrepParams), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_off_s, (C + 1) * 4)); CK(cudaMemcpy(d_off_s, off_strip.data(), (C + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_off_p, (C + 1) * 4)); CK(cudaMemcpy(d_off_p, off_prep.data(), (C + 1) * 4, cudaMemcpyHostToDevice));

    constexpr int FL = 9;
    using Lay = StripLay<2, 3, FL>;
    const size_t smem = Lay::launch_bytes(K);
    printf("shared memory per warp %u B (ring %u, control window %u, moment table %u)\n", Lay::total(K), Lay::ring_bytes, Lay::cwin_bytes, Lay::mtab_bytes);
    CK(cudaFuncSetAttribute(k_strip<2, 3, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_strip<2, 3, FL>, kStripWarps * 32, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_strip<2, 3, FL>));
    printf("registers %d, blocks per SM %d (%d warps)\n", fa.numRegs, occ, occ * kStripWarps);

    auto run = [&]() {
        k_prepare_strip<2><<<off_prep[C], 128>>>(d_pp, d_off_p, C);
        k_strip<2, 3, FL><<<off_strip[C], kStripWarps * 32, smem>>>(d_sp, d_off_s, C, K);
    };
    run();
    CK(cudaDeviceSynchronize());
    int hstat[2]; CK(cudaMemcpy(hstat, d_status, 8, cudaMemcpyDeviceToHost));
    printf("status bits %d\n", hstat[0]);

    // ---- check chunk 0 against the naive kernel -------------------------------------------------
    {
        double *g0, *g1, *rows, *loss, *n0, *n1, *nrows, *nloss;
        CK(cudaMalloc(&g0, M * 8)); CK(cudaMalloc(&g1, M * 8)); CK(cudaMalloc(&rows, E * 4 * 8)); CK(cudaMalloc(&loss, 8));
        CK(cudaMalloc(&n0, M * 8)); CK(cudaMalloc(&n1, M * 8)); CK(cudaMalloc(&nrows, E * 4 * 8)); CK(cudaMalloc(&nloss, 8));
        for (double* p : {g0, g1, n0, n1}) CK(cudaMemset(p, 0, M * 8));
        CK(cudaMemset(rows, 0, E * 32)); CK(cudaMemset(nrows, 0, E * 32)); CK(cudaMemset(loss, 0, 8)); CK(cudaMemset(nloss, 0, 8));
        FoldP f; f.part = sp[0].part; f.base = sp[0].part_base; f.rowpart = sp[0].rowpart; f.losspart = sp[0].losspart; f.rowid = d_rowid;
        f.n_tasks = n_tasks; f.W = W; f.n_strips = n_strips; f.E = E; f.g[0] = g0; f.g[1] = g1; f.rows = rows; f.loss = loss;
        const int64_t nf = std::max<int64_t>(2 * n_tasks * W, E * n_strips * 4);
        k_fold<<<(unsigned)((nf + 255) / 256), 256>>>(f);
        NaiveP np; np.data = sp[0].data; np.rowc = (const StripRow<2>*)sp[0].rowc; np.E = E; np.n_strips = n_strips;
        np.comp[0] = sp[0].comp[0]; np.comp[1] = sp[0].comp[1]; np.g[0] = n0; np.g[1] = n1; np.rows = nrows; np.loss = nloss;
        k_naive<<<(unsigned)((E * n_strips * 32 + 255) / 256), 256>>>(np);
        CK(cudaDeviceSynchronize());
        std::vector<double> a(M), b(M), ra(E * 4), rb(E * 4); double la, lb;
        CK(cudaMemcpy(&la, loss, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&lb, nloss, 8, cudaMemcpyDeviceToHost));
        printf("loss strip %.15e naive %.15e rel %.2e\n", la, lb, std::fabs(la - lb) / std::fabs(lb));
        for (int c = 0; c < 2; ++c) {
            CK(cudaMemcpy(a.data(), c ? g1 : g0, M * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), c ? n1 : n0, M * 8, cudaMemcpyDeviceToHost));
            double num = 0, den = 0;
            for (int k = 0; k < M; ++k) { num += (a[k] - b[k]) * (a[k] - b[k]); den += b[k] * b[k]; }
            printf("grad ctrl comp %d: norm-rel %.2e (norm %.3e)\n", c, std::sqrt(num / den), std::sqrt(den));
        }
        CK(cudaMemcpy(ra.data(), rows, E * 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rb.data(), nrows, E * 32, cudaMemcpyDeviceToHost));
        for (int q = 0; q < 3; ++q) {
            double num = 0, den = 0;
            for (int64_t pos = 0; pos < E; ++pos) {
                const double sc = 1.0;   // both in the S' / PV convention of the reducers
                const double va = ra[pos * 4 + q];      // k_strip numbers the rows in sorted order
                const double vb = rb[pos * 4 + q] * sc;
                num += (va - vb) * (va - vb); den += vb * vb;
            }
            printf("row sum %d: norm-rel %.2e (norm %.3e)\n", q, std::sqrt(num / den), std::sqrt(den));
        }
    }

    // ---- timing -----------------------------------------------------------------------------------
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) run();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    for (int i = 0; i < reps; ++i) {
        k_prepare_strip<2><<<off_prep[C], 128>>>(d_pp, d_off_p, C);
        CK(cudaEventRecord(e0));
        k_strip<2, 3, FL><<<off_strip[C], kStripWarps * 32, smem>>>(d_sp, d_off_s, C, K);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms); tot += ms;
    }
    const double bytes = (double)C * (E * P * 25.0 + 4.0 * M * 8 + E * 56.0);
    printf("k_strip: mean %.3f ms, best %.3f ms for %d chunks = %.1f us per chunk; %.0f GB/s algorithmic = %.3f of 6551\n",
           tot / reps, best, C, 1e3 * tot / reps / C, bytes / (tot / reps * 1e-3) / 1e9, bytes / (tot / reps * 1e-3) / 1e9 / 6551.0);
    CK(cudaMemcpy(hstat, d_status, 8, cudaMemcpyDeviceToHost));
    printf("status bits after timing %d\n", hstat[0]);
    return 0;
}
