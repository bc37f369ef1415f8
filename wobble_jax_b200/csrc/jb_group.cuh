// jb_group: one chunk whose epochs are sharded over the GPUs of a box, one process per GPU
// (BASELINE.json configs[4]).  Included at the end of jb_plan.cu.
//
// The reference has no multi-device path: its analogue is the epoch micro-batch loop of
// LossFunc.loss_all (jabble/loss.py:64-74), whose per-batch losses add up.  Here every rank holds
// the plans of its rows (templates replicated); loss and the gradient of the template control points
// are sums over rows, so one evaluation is, on ONE stream and without host work in between:
//     the rank's plans in one launch per kernel (jb_batch)
//  -> k_group_gather : [loss, gradient of every fit-mode control point] summed over the rank's plans
//  -> ncclAllReduce  : that vector (8 (1 + n_shared) bytes: 85 KB at configs[4]) summed over the ranks
//  -> k_group_scatter: the totals written back into every plan's [loss, gradient] output
// Per-epoch entries (shifts, airmass, their Fisher sums) belong to one rank and stay local.
// NCCL is loaded with dlopen at the first use (the library itself does not link it): the NCCL already
// in the process (torch's) is taken when there is one.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* env = getenv("JB_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (api.h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
            api.GetVersion = (decltype(api.GetVersion))dlsym(api.h, "ncclGetVersion");
            if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { dlclose(api.h); api.h = nullptr; }
        }
    }
    return api.h ? &api : nullptr;
}

#define JB_NCCL(api, call)                                                                         \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(JB_ERR_CUDA, "%s failed: %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r_) : "NCCL error"); \
    } while (0)

// red[0] = sum_i out_i[0];  red[1 + t] = sum_i out_i[pos_i[t]]   (plans in order: one fixed order)
__global__ void k_group_gather(double* const* __restrict__ outs, const int64_t* __restrict__ pos, int n_plans, int64_t n_shared,
                               double* __restrict__ red) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_shared) return;
    double s = 0.0;
    for (int i = 0; i < n_plans; ++i) s += outs[i][t == 0 ? 0 : pos[(int64_t)i * n_shared + t - 1]];
    red[t] = s;
}
__global__ void k_group_scatter(double* const* __restrict__ outs, const int64_t* __restrict__ pos, int n_plans, int64_t n_shared,
                                const double* __restrict__ red) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_shared) return;
    const double v = red[t];
    for (int i = 0; i < n_plans; ++i) outs[i][t == 0 ? 0 : pos[(int64_t)i * n_shared + t - 1]] = v;
}

}  // namespace

struct jb_group {
    jb_batch* batch = nullptr;
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    int64_t n_shared = 0;
    std::vector<int64_t> fit_sizes;      // n_fit of every plan when the layout was built
    int64_t* d_pos = nullptr;            // [n_plans][n_shared] position of shared entry t in plan i's output
    double** d_outs = nullptr;           // [n_plans] device pointers of the outputs
    double* d_red = nullptr;             // [1 + n_shared]
    bool layout_dirty = true;
    int64_t evals = 0;
    // the whole step (the plans' kernels, gather, all-reduce, scatter) as one CUDA graph once two calls in a
    // row enqueue the same thing: the host then issues ONE launch per step, and a rank whose host thread
    // is late by a few launches no longer holds the other ranks' all-reduce up
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t gstream = nullptr;
    int steady = 0;
    bool graphs_ok = true;
};

namespace {

// positions, in a plan's [loss, gradient] output, of the fit-mode control points of its global
// spline components -- the entries that are sums over ALL rows of the chunk
std::vector<int64_t> shared_positions(const jb_plan* pl) {
    std::vector<int64_t> pos;
    for (int64_t j = 0; j < pl->n_fit; ++j) {
        const int64_t q = pl->h_fit_map[j];
        for (int c = 0; c < pl->n_glob; ++c)
            if (q >= pl->desc[c].ctrl_offset && q < pl->desc[c].ctrl_offset + pl->desc[c].n_knots) { pos.push_back(1 + j); break; }
    }
    return pos;
}

int group_layout(jb_group* g) {
    const size_t n = g->batch->plans.size();
    bool same = !g->layout_dirty && g->fit_sizes.size() == n;
    for (size_t i = 0; same && i < n; ++i) same = g->fit_sizes[i] == g->batch->plans[i]->n_fit;
    if (same && g->d_outs) {
        std::vector<double*> outs(g->batch->out.begin(), g->batch->out.end());
        JB_CUDA(cudaMemcpy(g->d_outs, outs.data(), n * sizeof(double*), cudaMemcpyHostToDevice));
        return JB_OK;
    }
    std::vector<int64_t> all;
    int64_t ns = -1;
    for (size_t i = 0; i < n; ++i) {
        const std::vector<int64_t> pos = shared_positions(g->batch->plans[i]);
        if (ns < 0) ns = (int64_t)pos.size();
        if ((int64_t)pos.size() != ns)
            return fail(JB_ERR_BAD_ARG, "plans of a group must have the same template control points in fit mode (plan %zu has %zu, plan 0 has %lld)",
                        i, pos.size(), (long long)ns);
        all.insert(all.end(), pos.begin(), pos.end());
    }
    if (g->d_pos) cudaFree(g->d_pos);
    if (g->d_red) cudaFree(g->d_red);
    if (g->d_outs) cudaFree(g->d_outs);
    g->d_pos = nullptr; g->d_red = nullptr; g->d_outs = nullptr;
    g->n_shared = ns;
    JB_CUDA(cudaMalloc(&g->d_pos, std::max<size_t>(all.size(), 1) * sizeof(int64_t)));
    if (!all.empty()) JB_CUDA(cudaMemcpy(g->d_pos, all.data(), all.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    JB_CUDA(cudaMalloc(&g->d_red, (size_t)(1 + ns) * sizeof(double)));
    JB_CUDA(cudaMalloc(&g->d_outs, n * sizeof(double*)));
    std::vector<double*> outs(g->batch->out.begin(), g->batch->out.end());
    JB_CUDA(cudaMemcpy(g->d_outs, outs.data(), n * sizeof(double*), cudaMemcpyHostToDevice));
    g->fit_sizes.resize(n);
    for (size_t i = 0; i < n; ++i) g->fit_sizes[i] = g->batch->plans[i]->n_fit;
    g->layout_dirty = false;
    return JB_OK;
}

// gather -> all-reduce -> scatter, enqueued on st behind the evaluation of the rank's plans
int group_exchange(jb_group* g, cudaStream_t st) {
    const int n = (int)g->batch->plans.size();
    const unsigned nb = (unsigned)((g->n_shared + 1 + 255) / 256);
    k_group_gather<<<nb, 256, 0, st>>>(g->d_outs, g->d_pos, n, g->n_shared, g->d_red);
    if (g->world > 1) {
        NcclApi* api = nccl_api();
        JB_NCCL(api, api->AllReduce(g->d_red, g->d_red, (size_t)(1 + g->n_shared), ncclFloat64, ncclSum, g->comm, st));
    }
    k_group_scatter<<<nb, 256, 0, st>>>(g->d_outs, g->d_pos, n, g->n_shared, g->d_red);
    JB_CUDA(cudaGetLastError());
    g->batch->launches += 2;
    g->evals++;
    return JB_OK;
}

}  // namespace

extern "C" {

int jb_nccl_unique_id(void* id128) {
    if (!id128) return fail(JB_ERR_BAD_ARG, "null argument");
    NcclApi* api = nccl_api();
    if (!api) return fail(JB_ERR_CUDA, "NCCL could not be loaded (libnccl.so.2; set JB_NCCL_LIB)");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    JB_NCCL(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return JB_OK;
}

int jb_nccl_version(void) {
    NcclApi* api = nccl_api();
    int v = 0;
    if (!api || !api->GetVersion || api->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int jb_group_create(jb_plan* const* plans, int32_t n_plans, int32_t rank, int32_t world, const void* nccl_id128, jb_group** out) {
    if (!out) return fail(JB_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(JB_ERR_BAD_ARG, "rank %d of %d", rank, world);
    if (world > 1 && !nccl_id128) return fail(JB_ERR_BAD_ARG, "a group of %d ranks needs the NCCL id of jb_nccl_unique_id", world);
    jb_batch* b = nullptr;
    int rc = jb_batch_create(plans, n_plans, &b);
    if (rc) return rc;
    jb_group* g = new jb_group;
    g->batch = b; g->rank = rank; g->world = world; g->device = b->device;
    if (world > 1) {
        NcclApi* api = nccl_api();
        if (!api) { jb_batch_destroy(b); delete g; return fail(JB_ERR_CUDA, "NCCL could not be loaded (libnccl.so.2; set JB_NCCL_LIB)"); }
        ncclUniqueId id;
        memcpy(&id, nccl_id128, sizeof id);
        cudaSetDevice(g->device);
        ncclResult_t r = api->CommInitRank(&g->comm, world, id, rank);
        if (r != ncclSuccess) {
            jb_batch_destroy(b); delete g;
            return fail(JB_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
        }
    }
    *out = g;
    return JB_OK;
}

void jb_group_destroy(jb_group* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->batch && g->batch->stream) cudaStreamSynchronize(g->batch->stream);
    if (g->gstream) cudaStreamSynchronize(g->gstream);
    drop_graph(g->gexec);
    if (g->comm) { NcclApi* api = nccl_api(); if (api) api->CommDestroy(g->comm); }
    if (g->d_pos) cudaFree(g->d_pos);
    if (g->d_red) cudaFree(g->d_red);
    if (g->d_outs) cudaFree(g->d_outs);
    jb_batch_destroy(g->batch);
    delete g;
}

int jb_group_bind(jb_group* g, const double* const* d_p_fit, double* const* d_out) {
    if (!g) return fail(JB_ERR_BAD_ARG, "null group");
    int rc = jb_batch_bind(g->batch, d_p_fit, d_out);
    g->layout_dirty = true;
    return rc;
}

int jb_group_eval_device(jb_group* g, void* stream) {
    if (!g) return fail(JB_ERR_BAD_ARG, "null group");
    JB_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : g->batch->stream;
    int rc;
    for (jb_plan* pl : g->batch->plans)
        if ((int64_t)pl->h_fit_map.size() != pl->n_fit) return fail(JB_ERR_BAD_ARG, "jb_plan_set_fit has not been called on a plan of the group");
    bool stale = g->layout_dirty || g->fit_sizes.size() != g->batch->plans.size();
    for (size_t i = 0; !stale && i < g->fit_sizes.size(); ++i) stale = g->fit_sizes[i] != g->batch->plans[i]->n_fit;
    if (stale && (rc = group_layout(g))) return rc;
    static const bool group_graphs = [] { const char* e = getenv("JB_GROUP_GRAPH"); return !(e && e[0] == '0'); }();
    if (graphs_enabled() && group_graphs && g->graphs_ok && !stale && batch_is_clean(g->batch) && !stream_is_capturing(st)) {
        if (g->gexec && g->gstream != st) { JB_CUDA(cudaStreamSynchronize(g->gstream)); drop_graph(g->gexec); g->steady = 0; }
        if (!g->gexec && ++g->steady >= 2) {
            int rrc = JB_OK;
            const int64_t launches0 = g->batch->launches, evals0 = g->evals;
            std::vector<int64_t> ev0;
            for (jb_plan* pl : g->batch->plans) ev0.push_back(pl->evaluations);
            const bool ok = record_graph(st, &g->gexec, &rrc, [&] {
                int r = jb_batch_eval_device(g->batch, st);
                return r ? r : group_exchange(g, st);
            });
            // (recording counted a step that has not run yet)
            g->batch->launches = launches0; g->evals = evals0;
            for (size_t i = 0; i < ev0.size(); ++i) g->batch->plans[i]->evaluations = ev0[i];
            if (!ok || rrc) { g->graphs_ok = false; drop_graph(g->gexec); }
            else g->gstream = st;
        }
        if (g->gexec) {
            JB_CUDA(cudaGraphLaunch(g->gexec, st));
            g->batch->launches += (g->batch->p_fit.empty() ? 5 : 6) + 2;
            g->evals++;
            for (jb_plan* pl : g->batch->plans) pl->evaluations++;
            return JB_OK;
        }
    } else if (!batch_is_clean(g->batch) || stale) {
        if (g->gexec) { JB_CUDA(cudaStreamSynchronize(g->gstream)); drop_graph(g->gexec); }
        g->steady = 0;
    }
    if ((rc = jb_batch_eval_device(g->batch, st))) return rc;
    return group_exchange(g, st);
}

int jb_group_eval_device_checked(jb_group* g, void* stream) {
    if (!g) return fail(JB_ERR_BAD_ARG, "null group");
    JB_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : g->batch->stream;
    int rc;
    bool stale = g->layout_dirty || g->fit_sizes.size() != g->batch->plans.size();
    for (size_t i = 0; !stale && i < g->fit_sizes.size(); ++i) stale = g->fit_sizes[i] != g->batch->plans[i]->n_fit;
    if (stale && (rc = group_layout(g))) return rc;
    // settle the rank's own plans first (re-sort / wider windows as jb_plan_eval); every rank then
    // enqueues exactly one exchange, whatever it took locally
    const bool soft = g->batch->soft_range;
    g->batch->soft_range = true;                 // a rank off the knot grid must still join the all-reduce: its NaN loss tells everyone
    rc = jb_batch_eval_device_checked(g->batch, st);
    g->batch->soft_range = soft;
    if (rc) return rc;
    if ((rc = group_exchange(g, st))) return rc;
    JB_CUDA(cudaStreamSynchronize(st));
    return JB_OK;
}

int jb_group_info(const jb_group* g, int64_t* n_shared, int64_t* collective_bytes, int32_t* world) {
    if (!g) return fail(JB_ERR_BAD_ARG, "null group");
    if (n_shared) *n_shared = g->n_shared;
    if (collective_bytes) *collective_bytes = g->world > 1 ? (int64_t)sizeof(double) * (1 + g->n_shared) : 0;
    if (world) *world = g->world;
    return JB_OK;
}

int64_t jb_group_launches(const jb_group* g) { return g ? g->batch->launches : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------------
// The FP64 ceiling of the device, measured (in float64 this path sits at the
// roofline ridge, so the HBM fraction is reported next to the FP64 pipe's).  Every thread runs 8
// independent chains of dependent DFMAs; enough blocks to fill every SM several times over.
namespace {
__global__ void k_fp64_peak(double* out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (double)threadIdx.x * 1.0e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __fma_rn(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 12345.678) out[0] = s;        // never true: keeps the chains alive
}
}  // namespace

extern "C" int jb_double_peak(int32_t device, double* tflops) {
    if (!tflops) return fail(JB_ERR_BAD_ARG, "null argument");
    JB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    JB_CUDA(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    JB_CUDA(cudaMalloc(&d, sizeof(double)));
    const int blocks = prop.multiProcessorCount * 16, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    JB_CUDA(cudaEventCreate(&e0));
    JB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        JB_CUDA(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1.0e-9);
        JB_CUDA(cudaEventRecord(e1));
        JB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        JB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return JB_OK;
}

