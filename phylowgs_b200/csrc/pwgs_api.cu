// pwgs_api.cu — the C ABI of libpwgs_b200.so (include/pwgs_b200.h) over the sm_100a kernels.
// Host side only owns device memory, validates arguments and launches; there is no CPU compute path:
// every entry point that produces a likelihood or a sample runs a CUDA kernel or fails with PWGS_ECUDA.
#include "../../include/pwgs_b200.h"
#include "pwgs_kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <queue>
#include <functional>

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(PWGS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        n = count ? count : 1;
        return cudaMalloc((void **)&p, n * sizeof(T));
    }
    cudaError_t upload(const T *h, size_t count, cudaStream_t s)
    {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// page-locked host memory (DMA without the runtime's staging copy; copies from / to it are truly asynchronous)
struct PinnedBuf {
    char *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t bytes)
    {
        if (bytes <= n && p) return cudaSuccess;
        release();
        n = bytes ? bytes : 1;
        cudaError_t e = cudaHostAlloc((void **)&p, n, cudaHostAllocDefault);
        if (e != cudaSuccess) { p = nullptr; n = 0; }
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

struct pwgs_ctx {
    int device = 0, n_sms = 0;
    PinnedBuf stage;                                      // pwgs_set_tree / pwgs_move_data / row lists: host side of the uploads
    cudaEvent_t stage_ev = nullptr;                       // recorded after the last copy out of `stage`
    bool stage_busy = false;
    PinnedBuf user_buf;                                   // pwgs_host_buffer
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;             // around the last grid kernel (option "grid_kernel_ns")
    bool grid_timed = false;
    // ---- data set
    int n_ssm = 0, n_cnv = 0, T = 0, D = 0, Dp = 0, M = 0, L = 0;
    std::vector<int> ssm_of_m, pidx_host;                 // CNV-affected SSMs / plain data, ascending (the two work lists)
    int opt_grid_plain = 0;                               // 1: the grid kernel walks the data in file order (diagnostic: "grid_plain")
    DevBuf<int> a, d, cnv_ptr, cnv_idx, cnv_c1, cnv_c2;
    DevBuf<double> mu_r, mu_v, lbc;
    DevBuf<int> pidx, pa, pd, pnode, cidx, ca, cd;        // permuted work lists (plain first / CNV-affected SSMs)
    DevBuf<double> pmu_r, pmu_v, cmu_r, clbc;
    // error-rate classes of the plain data (opt-in sufficient-statistics path, K0d)
    int n_classes = 0, opt_collapse = 0;
    DevBuf<int> pcls;
    DevBuf<double> cl_mu_r, cl_mu_v;
    DevBuf<unsigned long long> cl_sums;
    double lbc_plain_sum = 0;
    // ---- tree
    bool tree_set = false;
    int K = 0;
    DevBuf<int> parent, depth, tin, tout, order, node_of_datum;
    DevBuf<double> phi, pi;
    DevBuf<int4> coef;                                    // dense [K][M] coefficients, built only for pwgs_get_data_states
    // sparse CNV work list of the MH kernel (rebuilt lazily after set_tree / when the CTA count changes)
    int maxL = 0, work_cpc = -1, work_collapse = -1, work_ncls = -1;
    int pchunk = 0;
    std::vector<int> plan_counts;                // [cpc][4] rounds per CTA: plain thirds, CNV rounds with 1 / 2 / 3 states (diagnostics)
    const int *pstart_dev = nullptr;
    int opt_w[4] = {100, 211, 482, 760};             // relative cost of a warp round: plain, CNV-affected with 1 / 2 / 3 merged states
    bool work_valid = false;
    DevBuf<int> code_tmp, enode_tmp, pos, w_a, w_d, w_code, w_enode, overflow, plan;
    DevBuf<int4> ecoef_tmp, w_ecoef;
    DevBuf<double> w_lbc, w_mu_r;
    CnvWork work;
    // ---- MH
    DevBuf<unsigned long long> sums;
    DevBuf<double> phisub;
    int root = 0;
    DevBuf<double> ring, phi_out, pi_out, llh_out, scratch;
    DevBuf<unsigned long long> counter, prof;
    bool opt_prof = false;
    unsigned long long prof_host[16 + 256] = {0};
    int prof_cpc = 0;
    DevBuf<int> acc_out, rows, cols, abort_flag;
    DevBuf<MhParams> plist;
    int opt_batch = PWGS_MAX_BATCH, opt_ctas = 0, opt_stage = 1, opt_dist = -1, opt_big = -1;
    DevBuf<double> gscratch;                               // mh_kernel<false, true>: the CTAs' node tables in global memory
    bool launched = false;
    int launched_iters = 0;
    long long n_launches = 0;                             // kernels of this library launched on behalf of the context
};

static const double kLogTinyMh = log(pow(10, -99));   // mh.hpp:100
static const double kLogTinyPy = log(1e-99);          // data.py:57
static const double kLogQuarter = log(0.25);          // mh.hpp:97

// Host-side preparation of a tree (pwgs_set_tree; also the host self-test of the grid likelihood): validation, depth, Euler
// tour (ancestor tests are two integer compares), children-first order, subtree sums of pi.
struct HostTree {
    int root = -1;
    std::vector<int> depth, tin, tout, order;
    std::vector<double> phisub;
};
static bool host_tree_prep(int K, int T, const int32_t *parent, const double *pi, HostTree &ht, std::string &err)
{
    // ---- validate: exactly one root, in-range parents, no cycles
    int root = -1;
    for (int k = 0; k < K; k++) {
        if (parent[k] < 0) { if (root >= 0) { err = "more than one root"; return false; } root = k; }
        else if (parent[k] >= K || parent[k] == k) { err = "parent index out of range at node " + std::to_string(k); return false; }
    }
    if (root < 0) { err = "no root (parent == -1) in tree"; return false; }
    std::vector<int> &depth = ht.depth;
    depth.assign(K, -1);
    depth[root] = 0;
    for (int k = 0; k < K; k++) {
        int v = k, steps = 0;
        std::vector<int> path;
        while (depth[v] < 0) { path.push_back(v); v = parent[v]; if (++steps > K) { err = "cycle in parent array"; return false; } }
        int dd = depth[v];
        for (int i = (int)path.size() - 1; i >= 0; i--) depth[path[i]] = ++dd;
    }
    // ---- Euler tour (children in increasing index) and the children-first accumulation order
    std::vector<std::vector<int>> kids(K);
    for (int k = 0; k < K; k++) if (parent[k] >= 0) kids[parent[k]].push_back(k);
    ht.tin.assign(K, 0); ht.tout.assign(K, 0);
    std::vector<int> stack{root}, it(K, 0);
    int clock = 0;
    ht.tin[root] = clock++;
    while (!stack.empty()) {
        int v = stack.back();
        if (it[v] < (int)kids[v].size()) { int ch = kids[v][it[v]++]; ht.tin[ch] = clock++; stack.push_back(ch); }
        else { ht.tout[v] = clock++; stack.pop_back(); }
    }
    ht.order.resize(K);
    for (int k = 0; k < K; k++) ht.order[k] = k;
    std::stable_sort(ht.order.begin(), ht.order.end(), [&](int x, int y) { return depth[x] > depth[y]; });
    // ---- subtree sums of pi, children before parents (deterministic; what the sparse grid sums multiply)
    ht.phisub.assign(pi, pi + (size_t)K * T);
    for (int i = 0; i < K; i++) {
        const int k = ht.order[i];
        if (parent[k] >= 0) for (int tp = 0; tp < T; tp++) ht.phisub[(size_t)parent[k] * T + tp] += ht.phisub[(size_t)k * T + tp];
    }
    ht.root = root;
    return true;
}

static DataView data_view(pwgs_ctx *c)
{
    DataView v;
    v.n_ssm = c->n_ssm; v.n_cnv = c->n_cnv; v.T = c->T;
    v.a = c->a.p; v.d = c->d.p; v.mu_r = c->mu_r.p; v.mu_v = c->mu_v.p; v.lbc = c->lbc.p;
    v.cnv_ptr = c->cnv_ptr.p; v.cnv_idx = c->cnv_idx.p; v.cnv_c1 = c->cnv_c1.p; v.cnv_c2 = c->cnv_c2.p;
    return v;
}
static TreeView tree_view(pwgs_ctx *c)
{
    TreeView v;
    v.K = c->K; v.T = c->T;
    v.parent = c->parent.p; v.depth = c->depth.p; v.tin = c->tin.p; v.tout = c->tout.p;
    v.phi = c->phi.p; v.pi = c->pi.p; v.node_of_datum = c->node_of_datum.p;
    v.phisub = c->phisub.p; v.root = c->root;
    return v;
}

extern "C" {

const char *pwgs_last_error(void) { return g_err.c_str(); }
int pwgs_version(void) { return 100; }

int pwgs_device_count(int *n)
{
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (n) *n = (e == cudaSuccess) ? cnt : 0;
    if (e != cudaSuccess) return fail(PWGS_ECUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return PWGS_OK;
}

int pwgs_destroy(pwgs_ctx *c)
{
    if (!c) return PWGS_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->a.release(); c->d.release(); c->cnv_ptr.release(); c->cnv_idx.release(); c->cnv_c1.release(); c->cnv_c2.release();
    c->mu_r.release(); c->mu_v.release(); c->lbc.release();
    c->pidx.release(); c->pa.release(); c->pd.release(); c->pnode.release(); c->cidx.release(); c->ca.release(); c->cd.release();
    c->pmu_r.release(); c->pmu_v.release(); c->cmu_r.release(); c->clbc.release();
    c->pcls.release(); c->cl_mu_r.release(); c->cl_mu_v.release(); c->cl_sums.release();
    c->parent.release(); c->depth.release(); c->tin.release(); c->tout.release(); c->order.release(); c->node_of_datum.release();
    c->phi.release(); c->pi.release(); c->coef.release();
    c->code_tmp.release(); c->enode_tmp.release(); c->pos.release(); c->w_a.release(); c->w_d.release(); c->w_code.release();
    c->w_enode.release(); c->overflow.release(); c->plan.release(); c->ecoef_tmp.release(); c->w_ecoef.release(); c->w_lbc.release(); c->w_mu_r.release();
    c->sums.release(); c->phisub.release(); c->ring.release(); c->phi_out.release(); c->pi_out.release(); c->llh_out.release(); c->scratch.release();
    c->counter.release(); c->prof.release(); c->acc_out.release(); c->abort_flag.release(); c->cols.release(); c->rows.release(); c->plist.release(); c->gscratch.release();
    c->stage.release(); c->user_buf.release();
    if (c->stage_ev) cudaEventDestroy(c->stage_ev);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PWGS_OK;
}

int pwgs_create(int device, int n_ssm, int n_cnv, int ntps,
                const int32_t *a, const int32_t *d, const double *mu_r, const double *mu_v,
                const int32_t *cnv_ptr, const int32_t *cnv_idx, const int32_t *cnv_c1, const int32_t *cnv_c2,
                pwgs_ctx **out)
{
    if (!out) return fail(PWGS_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_ssm < 0 || n_cnv < 0 || n_ssm + n_cnv <= 0) return fail(PWGS_EINVAL, "need at least one datum");
    if (ntps <= 0 || ntps > 65535) return fail(PWGS_EINVAL, "ntps must be in 1..65535");
    if (!a || !d || !mu_r || !mu_v) return fail(PWGS_EINVAL, "a, d, mu_r, mu_v must not be NULL");
    const int D = n_ssm + n_cnv, T = ntps;
    for (size_t i = 0; i < (size_t)D * T; i++)
        if (a[i] < 0 || d[i] < a[i]) return fail(PWGS_EINVAL, "need 0 <= a <= d for every datum and sample (index " + std::to_string(i) + ")");
    for (int j = 0; j < D; j++)
        if (!(mu_r[j] > 0.0 && mu_r[j] < 1.0 && mu_v[j] > 0.0 && mu_v[j] < 1.0))
            return fail(PWGS_EINVAL, "mu_r and mu_v must lie strictly inside (0,1) (datum " + std::to_string(j) + ")");
    int L = 0;
    if (cnv_ptr) {
        if (cnv_ptr[0] != 0) return fail(PWGS_EINVAL, "cnv_ptr[0] must be 0");
        for (int j = 0; j < n_ssm; j++) if (cnv_ptr[j + 1] < cnv_ptr[j]) return fail(PWGS_EINVAL, "cnv_ptr must be non-decreasing");
        L = cnv_ptr[n_ssm];
        if (L > 0 && (!cnv_idx || !cnv_c1 || !cnv_c2)) return fail(PWGS_EINVAL, "cnv_idx/cnv_c1/cnv_c2 missing");
        for (int l = 0; l < L; l++) {
            if (cnv_idx[l] < 0 || cnv_idx[l] >= n_cnv) return fail(PWGS_EINVAL, "cnv_idx out of range");
            if (cnv_c1[l] < 0 || cnv_c2[l] < 0) return fail(PWGS_EINVAL, "negative copy number");
            if ((long long)cnv_c1[l] + cnv_c2[l] > 32767) return fail(PWGS_ELIMIT, "copy numbers above 32767 are not supported");
        }
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PWGS_ECUDA, "no such CUDA device " + std::to_string(device));
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));

    pwgs_ctx *c = new pwgs_ctx;
    c->device = device; c->n_sms = prop.multiProcessorCount;
    c->n_ssm = n_ssm; c->n_cnv = n_cnv; c->T = T; c->D = D; c->L = L;
    if (cnv_ptr) for (int j = 0; j < n_ssm; j++) c->maxL = std::max(c->maxL, cnv_ptr[j + 1] - cnv_ptr[j]);
#define CREATE_TRY(expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            std::string m_ = std::string(#expr) + ": " + cudaGetErrorString(e_);                      \
            pwgs_destroy(c);                                                                          \
            return fail(PWGS_ECUDA, m_);                                                              \
        }                                                                                             \
    } while (0)
    CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreate(&c->ev0));
    CREATE_TRY(cudaEventCreate(&c->ev1));
    CREATE_TRY(cudaEventCreate(&c->ev2));
    CREATE_TRY(cudaEventCreate(&c->ev3));
    CREATE_TRY(cudaEventCreateWithFlags(&c->stage_ev, cudaEventDisableTiming));
    cudaStream_t s = c->stream;

    std::vector<int> zero_ptr(n_ssm + 1, 0), dummy(1, 0);
    const int *ptr = cnv_ptr ? cnv_ptr : zero_ptr.data();
    CREATE_TRY(c->a.upload(a, (size_t)D * T, s));
    CREATE_TRY(c->d.upload(d, (size_t)D * T, s));
    CREATE_TRY(c->mu_r.upload(mu_r, D, s));
    CREATE_TRY(c->mu_v.upload(mu_v, D, s));
    CREATE_TRY(c->cnv_ptr.upload(ptr, n_ssm + 1, s));
    CREATE_TRY(c->cnv_idx.upload(L ? cnv_idx : dummy.data(), L ? L : 1, s));
    CREATE_TRY(c->cnv_c1.upload(L ? cnv_c1 : dummy.data(), L ? L : 1, s));
    CREATE_TRY(c->cnv_c2.upload(L ? cnv_c2 : dummy.data(), L ? L : 1, s));
    CREATE_TRY(c->lbc.alloc((size_t)D * T));
    lbc_kernel<<<(D * T + 255) / 256, 256, 0, s>>>(D * T, c->a.p, c->d.p, c->lbc.p);
    c->n_launches++;
    CREATE_TRY(cudaGetLastError());

    // Work lists: "plain" data (mh.hpp:86-89: SSMs outside CNVs and the CNV data themselves) first, then the
    // CNV-affected SSMs (mh.hpp:90-161).  This split is a property of the inputs, fixed for the chain.
    std::vector<int> pidx;
    for (int j = 0; j < D; j++) {
        bool affected = j < n_ssm && ptr[j + 1] > ptr[j];
        if (affected) c->ssm_of_m.push_back(j); else pidx.push_back(j);
    }
    const int Dp = (int)pidx.size(), M = (int)c->ssm_of_m.size();
    c->Dp = Dp; c->M = M;
    std::vector<int> pa((size_t)std::max(Dp, 1) * T), pd((size_t)std::max(Dp, 1) * T), ca((size_t)std::max(M, 1) * T), cd((size_t)std::max(M, 1) * T);
    std::vector<double> pmr(std::max(Dp, 1)), pmv(std::max(Dp, 1)), cmr(std::max(M, 1));
    for (int i = 0; i < Dp; i++) {
        int j = pidx[i];
        pmr[i] = mu_r[j]; pmv[i] = mu_v[j];
        for (int tp = 0; tp < T; tp++) { pa[(size_t)tp * Dp + i] = a[(size_t)j * T + tp]; pd[(size_t)tp * Dp + i] = d[(size_t)j * T + tp]; }
    }
    for (int m = 0; m < M; m++) {
        int j = c->ssm_of_m[m];
        cmr[m] = mu_r[j];
        for (int tp = 0; tp < T; tp++) { ca[(size_t)tp * M + m] = a[(size_t)j * T + tp]; cd[(size_t)tp * M + m] = d[(size_t)j * T + tp]; }
    }
    {   // distinct (mu_r, mu_v) pairs of the plain data; more than 16 classes disables the collapsed path
        std::vector<int> pcls(std::max(Dp, 1), 0);
        std::vector<double> cr, cv;
        bool ok = true;
        for (int i = 0; i < Dp && ok; i++) {
            int f = -1;
            for (size_t q = 0; q < cr.size(); q++) if (cr[q] == pmr[i] && cv[q] == pmv[i]) { f = (int)q; break; }
            if (f < 0) { if (cr.size() >= 16) { ok = false; break; } cr.push_back(pmr[i]); cv.push_back(pmv[i]); f = (int)cr.size() - 1; }
            pcls[i] = f;
        }
        c->n_classes = ok ? (int)cr.size() : 0;
        if (c->n_classes > 0) {
            CREATE_TRY(c->pcls.upload(pcls.data(), Dp, s));
            CREATE_TRY(c->cl_mu_r.upload(cr.data(), cr.size(), s));
            CREATE_TRY(c->cl_mu_v.upload(cv.data(), cv.size(), s));
            CREATE_TRY(cudaStreamSynchronize(s));
        }
    }
    c->pidx_host = pidx;
    CREATE_TRY(c->pidx.upload(pidx.data(), Dp, s));
    CREATE_TRY(c->pa.upload(pa.data(), (size_t)Dp * T, s));
    CREATE_TRY(c->pd.upload(pd.data(), (size_t)Dp * T, s));
    CREATE_TRY(c->pmu_r.upload(pmr.data(), Dp, s));
    CREATE_TRY(c->pmu_v.upload(pmv.data(), Dp, s));
    CREATE_TRY(c->cidx.upload(c->ssm_of_m.data(), M, s));
    CREATE_TRY(c->ca.upload(ca.data(), (size_t)M * T, s));
    CREATE_TRY(c->cd.upload(cd.data(), (size_t)M * T, s));
    CREATE_TRY(c->cmu_r.upload(cmr.data(), M, s));
    CREATE_TRY(c->clbc.alloc((size_t)std::max(M, 1) * T));
    CREATE_TRY(c->pnode.alloc(std::max(Dp, 1)));
    CREATE_TRY(c->scratch.alloc(4));
    if (M > 0) {
        // clbc[tp][m] = lbc of the CNV-affected SSMs, recomputed in place by the same kernel
        lbc_kernel<<<(M * T + 255) / 256, 256, 0, s>>>(M * T, c->ca.p, c->cd.p, c->clbc.p);
        c->n_launches++;
        CREATE_TRY(cudaGetLastError());
    }
    sum_lbc_plain_kernel<<<1, 1024, 0, s>>>(Dp, T, c->pidx.p, c->lbc.p, c->scratch.p);
    c->n_launches++;
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaMemcpyAsync(&c->lbc_plain_sum, c->scratch.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    CREATE_TRY(cudaStreamSynchronize(s));   // host staging vectors go out of scope below
    CREATE_TRY(c->counter.alloc(1));
    CREATE_TRY(c->acc_out.alloc(1));
    CREATE_TRY(c->abort_flag.alloc(1));
    CREATE_TRY(c->llh_out.alloc(1));
    CREATE_TRY(c->plist.alloc(1));
#undef CREATE_TRY
    *out = c;
    return PWGS_OK;
}

// the staging buffer may be rewritten once the copies of the previous user have left it
static cudaError_t stage_acquire(pwgs_ctx *c, size_t bytes)
{
    if (c->stage_busy) { cudaError_t e = cudaEventSynchronize(c->stage_ev); if (e != cudaSuccess) return e; c->stage_busy = false; }
    return c->stage.alloc(bytes);
}
static cudaError_t stage_release(pwgs_ctx *c)
{
    c->stage_busy = true;
    return cudaEventRecord(c->stage_ev, c->stream);
}

int pwgs_set_tree(pwgs_ctx *c, int K, const int32_t *parent, const int32_t *node_of_datum,
                  const double *phi, const double *pi)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    if (K <= 0 || K > 65535) return fail(PWGS_EINVAL, "K must be in 1..65535");
    if (!parent || !phi || !pi) return fail(PWGS_EINVAL, "NULL array");
    if (!node_of_datum && (!c->tree_set || K < c->K))
        return fail(PWGS_EINVAL, "node_of_datum may only be NULL (= keep the device's assignments) after a tree with at most K nodes was set");
    HostTree ht;
    {
        std::string err;
        if (!host_tree_prep(K, c->T, parent, pi, ht, err)) return fail(PWGS_EINVAL, err);
    }
    for (size_t i = 0; i < (size_t)K * c->T; i++)
        if (!(phi[i] >= 0.0 && phi[i] <= 1.000001 && pi[i] >= 0.0 && pi[i] <= 1.000001))
            return fail(PWGS_EINVAL, "phi and pi must lie in [0,1] (node " + std::to_string(i / c->T) + ")");
    if (node_of_datum)
        for (int j = 0; j < c->D; j++)
            if (node_of_datum[j] < 0 || node_of_datum[j] >= K) return fail(PWGS_EINVAL, "node_of_datum out of range at datum " + std::to_string(j));

    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int T = c->T;
    const size_t KT = (size_t)K * T;
    // One page-locked staging area, plain asynchronous copies out of it, no synchronisation: an assignment sweep calls this
    // about a hundred times (every spawn event), and uploads from pageable memory synchronise the stream once per array.
    const size_t nd = 3 * KT, ni = 5 * (size_t)K + (node_of_datum ? (size_t)c->D : 0);
    CUDA_TRY(stage_acquire(c, nd * sizeof(double) + ni * sizeof(int)));
    double *hd = (double *)c->stage.p;
    int *hi = (int *)(hd + nd);
    memcpy(hd, phi, KT * sizeof(double)); memcpy(hd + KT, pi, KT * sizeof(double)); memcpy(hd + 2 * KT, ht.phisub.data(), KT * sizeof(double));
    memcpy(hi, parent, K * sizeof(int)); memcpy(hi + K, ht.depth.data(), K * sizeof(int)); memcpy(hi + 2 * K, ht.tin.data(), K * sizeof(int));
    memcpy(hi + 3 * K, ht.tout.data(), K * sizeof(int)); memcpy(hi + 4 * (size_t)K, ht.order.data(), K * sizeof(int));
    if (node_of_datum) memcpy(hi + 5 * (size_t)K, node_of_datum, (size_t)c->D * sizeof(int));
    c->K = K;
    CUDA_TRY(c->phi.upload(hd, KT, s));
    CUDA_TRY(c->pi.upload(hd + KT, KT, s));
    CUDA_TRY(c->phisub.upload(hd + 2 * KT, KT, s));
    CUDA_TRY(c->parent.upload(hi, K, s));
    CUDA_TRY(c->depth.upload(hi + K, K, s));
    CUDA_TRY(c->tin.upload(hi + 2 * K, K, s));
    CUDA_TRY(c->tout.upload(hi + 3 * K, K, s));
    CUDA_TRY(c->order.upload(hi + 4 * (size_t)K, K, s));
    if (node_of_datum) CUDA_TRY(c->node_of_datum.upload(hi + 5 * (size_t)K, c->D, s));
    CUDA_TRY(stage_release(c));
    c->root = ht.root;
    CUDA_TRY(c->phi_out.alloc(KT));
    CUDA_TRY(c->pi_out.alloc(KT));
    if (node_of_datum && c->Dp > 0) {
        gather_node_kernel<<<(c->Dp + 255) / 256, 256, 0, s>>>(c->Dp, c->pidx.p, c->node_of_datum.p, c->pnode.p);
        c->n_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    c->work_valid = false;
    c->tree_set = true;
    return PWGS_OK;
}

int pwgs_move_data(pwgs_ctx *c, int n, const int32_t *datum, const int32_t *node)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    if (!c->tree_set) return fail(PWGS_ESTATE, "pwgs_set_tree must be called first");
    if (n < 0 || (n > 0 && (!datum || !node))) return fail(PWGS_EINVAL, "pwgs_move_data: NULL array");
    for (int i = 0; i < n; i++)
        if (datum[i] < 0 || datum[i] >= c->D || node[i] < 0 || node[i] >= c->K) return fail(PWGS_EINVAL, "pwgs_move_data: index out of range");
    if (n == 0) return PWGS_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(stage_acquire(c, 2 * (size_t)n * sizeof(int)));
    int *h = (int *)c->stage.p;
    memcpy(h, datum, n * sizeof(int)); memcpy(h + n, node, n * sizeof(int));
    CUDA_TRY(c->rows.alloc(2 * (size_t)n));
    CUDA_TRY(cudaMemcpyAsync(c->rows.p, h, 2 * (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(stage_release(c));
    move_data_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(n, c->rows.p, c->rows.p + n, c->node_of_datum.p);
    CUDA_TRY(cudaGetLastError());
    c->n_launches++;
    if (c->Dp > 0) {
        gather_node_kernel<<<(c->Dp + 255) / 256, 256, 0, c->stream>>>(c->Dp, c->pidx.p, c->node_of_datum.p, c->pnode.p);
        CUDA_TRY(cudaGetLastError());
        c->n_launches++;
    }
    c->work_valid = false;
    return PWGS_OK;
}

int pwgs_host_buffer(pwgs_ctx *c, uint64_t bytes, void **ptr)
{
    if (!c || !ptr) return fail(PWGS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));          // nothing may still be writing into the buffer it replaces
    // page-locking is slow (~1 ms per MB): grow with headroom so that a tree that gains a few nodes does not pay it again
    if (bytes > c->user_buf.n || !c->user_buf.p) CUDA_TRY(c->user_buf.alloc(bytes + bytes / 2));
    *ptr = c->user_buf.p;
    return PWGS_OK;
}

} // extern "C"

// ------------------------------------------------------------------ sparse CNV work list (K0c) and the chain's work partition
// The MH kernel walks its data one warp round (32 data, one per lane) at a time and a round costs what its most expensive lane
// costs.  So the unit of the partition is the round: the CNV-affected SSMs, sorted by their number of merged states, are cut
// into rounds of 32, the plain data likewise, and whole rounds are dealt to the CTAs longest-processing-time-first (each
// round to the least loaded CTA).  Every CTA's lists are then multiples of 32 -- no partially filled rounds but the single
// tail of each list -- and the loads differ by less than one plain round.  (Dealing SSMs one by one gave every CTA a
// 5/32-filled and a mixed-state round: 80 % lane efficiency at config 3.)
struct Partition {
    std::vector<int> round_slot, cta_nc, pstart, counts;
    int cchunk = 0, pchunk = 0;
};
static Partition plan_partition(int cpc, int M, const int n_states[3] /* SSMs with 3, 2, 1 states */, int Dp_units_of, const int w[4],
                                int ncls /* proposal-group classes a full batch has: 1..MH_PCLASSES */)
{
    Partition P;
    const int U = (M + 31) / 32, PU = (Dp_units_of + 31) / 32;
    P.counts.assign((size_t)cpc * 4, 0);
    std::vector<int> n_cnv(cpc, 0);
    typedef std::pair<long long, int> Slot;
    std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> heap;
    for (int i = 0; i < cpc; i++) heap.push(Slot(0, i));
    std::vector<int> owner(U, 0), local(U, 0);
    for (int u = 0; u < U; u++) {                                         // rounds in sorted order = most expensive first
        const int r = 32 * u, cls = r < n_states[0] ? 3 : (r < n_states[0] + n_states[1] ? 2 : 1);
        Slot t = heap.top(); heap.pop();
        owner[u] = t.second; local[u] = n_cnv[t.second]++;
        P.counts[(size_t)t.second * 4 + cls]++;
        heap.push(Slot(t.first + w[cls], t.second));
    }
    // plain rounds in fractions: a full batch gives every warp ncls proposal groups in turn (classes 0..ncls-1; three at 128
    // proposals), and each class has its own partition of the plain list, so a CTA can take a round for some of the classes
    // only.  n3[c] = fractions (1/ncls of a round) of CTA c, levelled like the rounds above with loads in units of 1/ncls of a
    // weight; with S = running sum of n3, class i's boundary after CTA c is floor((S + i) / ncls) rounds: monotone, ending at PU
    // for every class, and the classes' counts of a CTA add up to n3[c] exactly (Hermite's identity).  The ranges of a CTA
    // differ by at most one round at either end.  Classes a full batch does not have share class 0's partition.
    {
        // identical units: water-filling in closed form instead of 3*PU heap operations.  Loads in units of w/3.
        std::vector<long long> L(cpc, 0);
        while (!heap.empty()) { Slot t = heap.top(); heap.pop(); L[t.second] = ncls * t.first; }
        const long long N = (long long)ncls * PU, wq = w[0];
        std::vector<int> n3(cpc, 0);
        if (N > 0) {
            long long lo = *std::min_element(L.begin(), L.end()), hi = *std::max_element(L.begin(), L.end()) + wq * (N / cpc + 2);
            auto filled = [&](long long lam) { long long t = 0; for (int c = 0; c < cpc; c++) if (lam > L[c]) t += (lam - L[c] + wq - 1) / wq; return t; };
            while (lo < hi) {                                             // smallest level whose fill covers N units
                const long long mid = lo + (hi - lo) / 2;
                if (filled(mid) >= N) hi = mid; else lo = mid + 1;
            }
            long long total = 0;
            for (int c = 0; c < cpc; c++) { n3[c] = lo > L[c] ? (int)((lo - L[c] + wq - 1) / wq) : 0; total += n3[c]; }
            // give back the surplus where the last unit started highest (what the greedy would have dealt last)
            std::vector<std::pair<long long, int>> last;
            for (int c = 0; c < cpc; c++) if (n3[c] > 0) last.push_back(std::make_pair(L[c] + (n3[c] - 1) * wq, c));
            std::sort(last.begin(), last.end(), std::greater<std::pair<long long, int>>());
            for (size_t k = 0; k < last.size() && total > N; k++) { n3[last[k].second]--; total--; }
        }
        for (int c = 0; c < cpc; c++) P.counts[(size_t)c * 4] = n3[c];
        P.pstart.assign((size_t)MH_PCLASSES * (cpc + 1), 0);
        long long S = 0;
        int maxp = 0;
        for (int c = 0; c < cpc; c++) {
            S += n3[c];
            int lo = Dp_units_of, hi = 0;
            for (int i = 0; i < MH_PCLASSES; i++) {
                const int e = (int)std::min<long long>(Dp_units_of, 32 * ((S + (i < ncls ? i : 0)) / ncls));
                P.pstart[(size_t)i * (cpc + 1) + c + 1] = e;
                lo = std::min(lo, P.pstart[(size_t)i * (cpc + 1) + c]); hi = std::max(hi, e);
            }
            maxp = std::max(maxp, hi - lo);
        }
        P.pchunk = maxp;
    }
    int maxc = 0;
    for (int i = 0; i < cpc; i++) maxc = std::max(maxc, n_cnv[i]);
    P.cchunk = 32 * maxc;
    P.round_slot.resize(std::max(U, 1));
    for (int u = 0; u < U; u++) P.round_slot[u] = owner[u] * P.cchunk + 32 * local[u];
    P.cta_nc.resize(cpc);
    for (int i = 0; i < cpc; i++) P.cta_nc[i] = 32 * n_cnv[i];
    if (U > 0 && M % 32) P.cta_nc[owner[U - 1]] -= 32 - M % 32;          // the tail round is the last (cheapest) one of its CTA
    return P;
}

// proposal-group classes of a full batch of B proposals: a warp's 1st, 2nd, 3rd group (see the large-batch loop of mh_kernel)
static int mh_group_classes(int B)
{
    // a full batch gives every warp whole proposal groups in turn, its i-th group being of class i: 256 proposals = 64 groups
    // over 11 warps = 6 classes, 128 = 3; smaller caps cut their groups into shares of the data rounds (MH_UNITS units again)
    // and use two classes of groups; tiny caps have a single group
    const int groups = B / MH_BU;
    if (groups >= MH_UNITS) return std::min(MH_PCLASSES, (groups + MH_CWARPS - 1) / MH_CWARPS);
    return B >= 2 * MH_BU ? 2 : 1;
}

static int prepare_cnv_work(pwgs_ctx *c, int cpc, int collapse, int ncls)
{
    if (c->work_valid && c->work_cpc == cpc && c->work_collapse == collapse && c->work_ncls == ncls) return PWGS_OK;
    const int M = c->M, T = c->T, K = c->K;
    CnvWork &w = c->work;
    w.Emax = std::max(1, std::min(K, c->maxL + 2));
    cudaStream_t s = c->stream;
    int h[4] = {0, 0, 0, 0};                                             // overflow flag, SSMs with 3 / 2 / 1 merged states
    if (M > 0) {
        CUDA_TRY(c->code_tmp.alloc(M)); CUDA_TRY(c->pos.alloc(M)); CUDA_TRY(c->overflow.alloc(4));
        CUDA_TRY(c->enode_tmp.alloc((size_t)M * w.Emax)); CUDA_TRY(c->ecoef_tmp.alloc((size_t)M * w.Emax));
        CUDA_TRY(cudaMemsetAsync(c->overflow.p, 0, 4 * sizeof(int), s));
        cnv_sparse_kernel<<<(M + 127) / 128, 128, 0, s>>>(data_view(c), tree_view(c), M, w.Emax, c->cidx.p, c->code_tmp.p, c->enode_tmp.p,
                                                          c->ecoef_tmp.p, c->overflow.p);
        CUDA_TRY(cudaGetLastError());
        cnv_sort_kernel<<<1, 1024, 0, s>>>(M, c->code_tmp.p, c->pos.p, c->overflow.p + 1);
        CUDA_TRY(cudaGetLastError());
        c->n_launches += 2;
        CUDA_TRY(cudaMemcpyAsync(h, c->overflow.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (h[0]) return fail(PWGS_ELIMIT, "internal: sparse CNV entry bound exceeded");
    }
    const Partition P = plan_partition(cpc, M, h + 1, collapse ? 0 : c->Dp, c->opt_w, ncls);
    w.cchunk = P.cchunk;
    w.Mp = std::max(1, w.cchunk * cpc);
    c->pchunk = P.pchunk;
    c->plan_counts = P.counts;
    // one upload: [round_slot | cta_nc | pstart]
    std::vector<int> tab(P.round_slot);
    tab.insert(tab.end(), P.cta_nc.begin(), P.cta_nc.end());
    tab.insert(tab.end(), P.pstart.begin(), P.pstart.end());
    CUDA_TRY(c->plan.alloc(tab.size()));
    CUDA_TRY(cudaMemcpyAsync(c->plan.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));   // pageable: staged before return
    const int *d_round_slot = c->plan.p, *d_cta_nc = c->plan.p + P.round_slot.size();
    c->pstart_dev = d_cta_nc + cpc;                                      // [MH_PCLASSES][cpc+1]
    CUDA_TRY(c->w_a.alloc((size_t)w.Mp * T)); CUDA_TRY(c->w_d.alloc((size_t)w.Mp * T)); CUDA_TRY(c->w_lbc.alloc((size_t)w.Mp * T));
    CUDA_TRY(c->w_mu_r.alloc(w.Mp)); CUDA_TRY(c->w_code.alloc(w.Mp));
    CUDA_TRY(c->w_enode.alloc((size_t)w.Mp * w.Emax)); CUDA_TRY(c->w_ecoef.alloc((size_t)w.Mp * w.Emax));
    w.a = c->w_a.p; w.d = c->w_d.p; w.lbc = c->w_lbc.p; w.mu_r = c->w_mu_r.p; w.code = c->w_code.p; w.enode = c->w_enode.p; w.ecoef = c->w_ecoef.p;
    w.cta_nc = d_cta_nc;
    if (M > 0) {
        cnv_scatter_kernel<<<(M + 127) / 128, 128, 0, s>>>(M, T, d_round_slot, w, c->pos.p, c->ca.p, c->cd.p, c->clbc.p, c->cmu_r.p, c->code_tmp.p,
                                                           c->enode_tmp.p, c->ecoef_tmp.p);
        CUDA_TRY(cudaGetLastError());
        c->n_launches += 1;
    }
    c->work_valid = true;
    c->work_cpc = cpc;
    c->work_collapse = collapse;
    c->work_ncls = ncls;
    return PWGS_OK;
}

// ------------------------------------------------------------------ MH
static size_t mh_smem_bytes(int K, int T, int B)
{
    return mh_smem_doubles(K, T, B) * sizeof(double) + mh_smem_ints(K) * sizeof(int) + 16;
}

// One chain's share of an MH launch: batch size, work partition for cpc CTAs, buffers, descriptor.  kind: 0 = data slice staged in
// shared memory, 1 = read in place, 2 = tree-sized tables in global memory (BIG).
struct MhPlan { MhParams p; size_t smem; int kind; };
static const void *mh_kernel_of(int kind)
{
    return kind == 0 ? (const void *)mh_kernel<true, false> : kind == 1 ? (const void *)mh_kernel<false, false> : (const void *)mh_kernel<false, true>;
}
static int mh_prepare(pwgs_ctx *c, int iters, double mh_std, uint64_t seed, uint64_t stream, int cpc, MhPlan &plan)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    if (!c->tree_set) return fail(PWGS_ESTATE, "pwgs_set_tree must be called before pwgs_mh_run");
    if (c->launched) return fail(PWGS_ESTATE, "an MH launch is pending: call pwgs_mh_wait first");
    if (iters < 0) return fail(PWGS_EINVAL, "iters must be >= 0");
    if (!(mh_std > 0)) return fail(PWGS_EINVAL, "mh_std must be > 0");
    CUDA_TRY(cudaSetDevice(c->device));
    int B = c->opt_batch;
    size_t smem = mh_smem_bytes(c->K, c->T, B);
    while (smem > 176 * 1024 && B > 1) { B >>= 1; smem = mh_smem_bytes(c->K, c->T, B); }
    // Trees of 48 nodes and more whose node tables do not leave room for a full batch in shared memory run with the tables in
    // global memory and full batches instead: measured faster from there on, by node count rather than by table size (5000
    // iterations, plain data collapsed; config 3 data, T = 1: K = 48 3.2 (shared) vs 3.3 ms (global), 56 3.0 vs 3.0, 64 3.5 vs 3.1,
    // 128 13.9 vs 3.6; config 4 data, T = 5: K = 26 4.1 vs 5.4, 34 4.8 vs 5.4, 48 6.9 vs 5.6, 96 12.5 vs 7.4 --
    // scripts/mh_big_tree.py, mh_big_tree_t5.py), and no limit short of device memory (the reference has none either).
    const bool big = c->opt_big > 0 || (c->opt_big < 0 && ((B < c->opt_batch && (c->K >= 48 || B <= 4)) || smem > 200 * 1024));
    size_t gstride = 0;
    if (big) {
        B = c->opt_batch;
        while (B > MH_BU && mh_tree_doubles(c->K, c->T, B) * sizeof(double) > ((size_t)8 << 20)) B >>= 1;   // <= 8 MiB per CTA
        gstride = (mh_tree_doubles(c->K, c->T, B) + 1) & ~(size_t)1;
        smem = mh_fixed_doubles(B) * sizeof(double) + mh_smem_ints(c->K) * sizeof(int) + 16;
        if (smem > 200 * 1024) return fail(PWGS_ELIMIT, "tree too large: the Euler-tour table of K nodes exceeds shared memory");
        CUDA_TRY(c->gscratch.alloc(gstride * cpc));
    }
    const int collapse = (c->opt_collapse && c->n_classes > 0 && c->Dp > 0) ? c->n_classes : 0;
    { int rc_ = prepare_cnv_work(c, cpc, collapse, mh_group_classes(B)); if (rc_ != PWGS_OK) return rc_; }
    // stage each CTA's slice of the data in shared memory when it fits beside the node tables
    const size_t stage_bytes = mh_stage_bytes(c->pchunk, c->work.cchunk, c->T, c->work.Emax) + 16;
    const int stage = (!big && c->opt_stage && smem + stage_bytes <= 224 * 1024) ? 1 : 0;
    if (stage) smem += stage_bytes;
    CUDA_TRY(c->sums.alloc((size_t)2 * PWGS_MAX_BATCH * MH_SUM_STRIDE));
    CUDA_TRY(cudaMemsetAsync(c->sums.p, 0, (size_t)2 * PWGS_MAX_BATCH * MH_SUM_STRIDE * sizeof(unsigned long long), c->stream));
    CUDA_TRY(c->ring.alloc((size_t)(MH_RING + 2) * mh_bufstride(c->K, c->T, B)));   // ring + the two epoch slots
    const int distributed = c->opt_dist < 0 ? (cpc >= MH_DIST_MIN_CTAS ? 1 : 0) : c->opt_dist;
    CUDA_TRY(cudaMemsetAsync(c->counter.p, 0, sizeof(unsigned long long), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->abort_flag.p, 0, sizeof(int), c->stream));

    if (collapse) {
        const size_t n = (size_t)c->K * collapse * c->T * 2;
        CUDA_TRY(c->cl_sums.alloc(n));
        CUDA_TRY(cudaMemsetAsync(c->cl_sums.p, 0, n * sizeof(unsigned long long), c->stream));
        collapse_kernel<<<(c->Dp + 255) / 256, 256, 0, c->stream>>>(c->Dp, c->T, collapse, c->pa.p, c->pd.p, c->pnode.p, c->pcls.p, c->cl_sums.p);
        CUDA_TRY(cudaGetLastError());
        c->n_launches++;
    }
    MhParams &p = plan.p;
    p.Dp = c->Dp; p.M = c->M; p.T = c->T; p.K = c->K;
    p.n_classes = collapse; p.cl_sums = c->cl_sums.p; p.cl_mu_r = c->cl_mu_r.p; p.cl_mu_v = c->cl_mu_v.p;
    p.pa = c->pa.p; p.pd = c->pd.p; p.pmu_r = c->pmu_r.p; p.pmu_v = c->pmu_v.p; p.pnode = c->pnode.p;
    p.cw = c->work;
    p.pstart = c->pstart_dev; p.pchunk = c->pchunk; p.ncls = c->work_ncls;
    p.tin = c->tin.p; p.tout = c->tout.p; p.phi0 = c->phi.p; p.pi0 = c->pi.p;
    p.phi_out = c->phi_out.p; p.pi_out = c->pi_out.p; p.acc_out = c->acc_out.p; p.llh_out = c->llh_out.p;
    p.iters = iters; p.max_batch = B; p.stage = stage; p.distributed = distributed; p.mh_std = (double)(float)mh_std;   // mh.hpp:28: MH_STD is a float
    p.seed = seed; p.stream = stream;
    p.sums = c->sums.p; p.ring = c->ring.p; p.counter = c->counter.p; p.abort_flag = c->abort_flag.p;
    p.lbc_plain_sum = c->lbc_plain_sum; p.log_tiny = kLogTinyMh;
    p.gscratch = big ? c->gscratch.p : nullptr; p.gstride = gstride;
    p.prof = nullptr;
    if (c->opt_prof) { CUDA_TRY(c->prof.alloc(16 + 256)); p.prof = c->prof.p; c->prof_cpc = std::min(cpc, 256); }
    plan.smem = smem;
    plan.kind = big ? 2 : (stage ? 0 : 1);
    return PWGS_OK;
}

// cooperative launch of n_chains * cpc CTAs on `s`; plist: device array of the chains' descriptors
static int mh_launch_grid(int device, int kind, const MhParams *plist, int n_chains, int cpc, size_t smem, cudaStream_t s)
{
    const void *kfn = mh_kernel_of(kind);
    // the attribute belongs to the function, not to the launch: chains running as threads of one process launch with
    // different sizes concurrently, so it is always set to the one value every launch fits under
    int smem_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaFuncAttributes kattr;
    CUDA_TRY(cudaFuncGetAttributes(&kattr, kfn));
    const int smem_dyn_max = smem_optin - (int)kattr.sharedSizeBytes;      // the limit covers static + dynamic
    if (smem > (size_t)smem_dyn_max) return fail(PWGS_ELIMIT, "shared-memory node table exceeds the device's per-block limit");
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dyn_max);
    if (e == cudaSuccess) {
        void *args[] = {(void *)&plist, (void *)&cpc};
        e = cudaLaunchCooperativeKernel(kfn, dim3(n_chains * cpc), dim3(MH_THREADS), args, smem, s);
    }
    if (e != cudaSuccess) {
        cudaFuncAttributes fa;
        int occ = -1;
        cudaGetLastError();
        cudaFuncGetAttributes(&fa, kfn);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, MH_THREADS, smem);
        return fail(PWGS_ECUDA, std::string("mh_kernel launch: ") + cudaGetErrorString(e) + " (grid " + std::to_string(n_chains * cpc) + " x " +
                    std::to_string(MH_THREADS) + " threads, " + std::to_string(smem) + " B dynamic smem, " + std::to_string(fa.numRegs) +
                    " regs, maxThreadsPerBlock " + std::to_string(fa.maxThreadsPerBlock) + ", static smem " + std::to_string(fa.sharedSizeBytes) +
                    ", local " + std::to_string(fa.localSizeBytes) + ", occupancy " + std::to_string(occ) + " CTA/SM)");
    }
    return PWGS_OK;
}

extern "C" {

int pwgs_mh_launch(pwgs_ctx *c, int iters, double mh_std, uint64_t seed, uint64_t stream)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    const int cpc = c->opt_ctas > 0 ? std::min(c->opt_ctas, c->n_sms) : c->n_sms;
    MhPlan plan;
    { int rc_ = mh_prepare(c, iters, mh_std, seed, stream, cpc, plan); if (rc_ != PWGS_OK) return rc_; }
    // pageable source: the runtime stages the bytes before returning, so the stack object is safe
    CUDA_TRY(cudaMemcpyAsync(c->plist.p, &plan.p, sizeof(MhParams), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    { int rc_ = mh_launch_grid(c->device, plan.kind, c->plist.p, 1, cpc, plan.smem, c->stream); if (rc_ != PWGS_OK) return rc_; }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->n_launches++;
    c->launched = true;
    c->launched_iters = iters;
    return PWGS_OK;
}

int pwgs_mh_launch_multi(int n_chains, pwgs_ctx *const *ctxs, int iters, const double *mh_std, const uint64_t *seed, const uint64_t *stream)
{
    if (n_chains < 1 || !ctxs || !mh_std || !seed || !stream) return fail(PWGS_EINVAL, "pwgs_mh_launch_multi: NULL argument or no chains");
    for (int i = 0; i < n_chains; i++) {
        if (!ctxs[i]) return fail(PWGS_EINVAL, "ctx is NULL");
        if (ctxs[i]->device != ctxs[0]->device) return fail(PWGS_EINVAL, "pwgs_mh_launch_multi: all contexts must live on one device");
        for (int j = 0; j < i; j++) if (ctxs[j] == ctxs[i]) return fail(PWGS_EINVAL, "pwgs_mh_launch_multi: a context appears twice");
    }
    pwgs_ctx *c0 = ctxs[0];
    if (n_chains > c0->n_sms) return fail(PWGS_ELIMIT, "more chains than SMs");
    int cpc = c0->n_sms / n_chains;
    if (c0->opt_ctas > 0) cpc = std::min(cpc, c0->opt_ctas);
    std::vector<MhParams> plist(n_chains);
    size_t smem = 0;
    int kind = 0;
    for (int i = 0; i < n_chains; i++) {
        MhPlan plan;
        int rc_ = mh_prepare(ctxs[i], iters, mh_std[i], seed[i], stream[i], cpc, plan);
        if (rc_ != PWGS_OK) return rc_;
        plist[i] = plan.p;
        smem = std::max(smem, plan.smem);
        kind = std::max(kind, plan.kind);                // one kernel for all chains: the most general variant any of them needs
    }
    if (kind == 2) for (int i = 0; i < n_chains; i++) if (!plist[i].gscratch)
        return fail(PWGS_ELIMIT, "pwgs_mh_launch_multi: chains with and without shared-memory node tables cannot share a launch");
    // everything the chains' own streams have queued (tree uploads, work lists) before the kernel; their later work after it
    CUDA_TRY(c0->plist.alloc(n_chains));
    for (int i = 1; i < n_chains; i++) {
        CUDA_TRY(cudaEventRecord(ctxs[i]->ev0, ctxs[i]->stream));
        CUDA_TRY(cudaStreamWaitEvent(c0->stream, ctxs[i]->ev0, 0));
    }
    CUDA_TRY(cudaMemcpyAsync(c0->plist.p, plist.data(), n_chains * sizeof(MhParams), cudaMemcpyHostToDevice, c0->stream));
    for (int i = 0; i < n_chains; i++) CUDA_TRY(cudaEventRecord(ctxs[i]->ev0, c0->stream));
    { int rc_ = mh_launch_grid(c0->device, kind, c0->plist.p, n_chains, cpc, smem, c0->stream); if (rc_ != PWGS_OK) return rc_; }
    for (int i = 0; i < n_chains; i++) {
        CUDA_TRY(cudaEventRecord(ctxs[i]->ev1, c0->stream));
        if (i) CUDA_TRY(cudaStreamWaitEvent(ctxs[i]->stream, ctxs[i]->ev1, 0));
        ctxs[i]->launched = true;
        ctxs[i]->launched_iters = iters;
    }
    c0->n_launches++;
    return PWGS_OK;
}

int pwgs_mh_wait(pwgs_ctx *c, double *phi_out, double *pi_out, double *accept_ratio, double *llh_out, float *kernel_ms)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    if (!c->launched) return fail(PWGS_ESTATE, "no MH launch pending");
    CUDA_TRY(cudaSetDevice(c->device));
    c->launched = false;
    size_t KT = (size_t)c->K * c->T;
    int acc = 0, aborted = 0;
    double llh = 0;
    if (phi_out) CUDA_TRY(cudaMemcpyAsync(phi_out, c->phi_out.p, KT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (pi_out) CUDA_TRY(cudaMemcpyAsync(pi_out, c->pi_out.p, KT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&acc, c->acc_out.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&aborted, c->abort_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&llh, c->llh_out.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (c->opt_prof && c->prof.p) CUDA_TRY(cudaMemcpyAsync(c->prof_host, c->prof.p, sizeof(c->prof_host), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (aborted) return fail(PWGS_ECUDA, "MH kernel watchdog: the chain barrier did not complete (kernel aborted itself)");
    if (accept_ratio) *accept_ratio = c->launched_iters > 0 ? (double)acc / c->launched_iters : 0.0;   // mh.cpp:104-107
    if (llh_out) *llh_out = llh;
    if (kernel_ms) CUDA_TRY(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
    return PWGS_OK;
}

int pwgs_mh_run(pwgs_ctx *c, int iters, double mh_std, uint64_t seed, uint64_t stream,
                double *phi_out, double *pi_out, double *accept_ratio, double *llh_out)
{
    int rc = pwgs_mh_launch(c, iters, mh_std, seed, stream);
    if (rc != PWGS_OK) return rc;
    return pwgs_mh_wait(c, phi_out, pi_out, accept_ratio, llh_out, nullptr);
}

// ------------------------------------------------------------------ likelihood grid / totals
static int llh_block_impl(pwgs_ctx *c, int row0, int n_rows, const int32_t *rows, int n_cols, const int32_t *cols, int semantics, double *out,
                          int64_t out_pitch = 0)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    if (!c->tree_set) return fail(PWGS_ESTATE, "pwgs_set_tree must be called first");
    if (n_rows < 0 || (n_rows > 0 && !out)) return fail(PWGS_EINVAL, "rows/out NULL");
    const int col_major = (semantics & PWGS_LAYOUT_COLS) ? 1 : 0;
    semantics &= ~PWGS_LAYOUT_COLS;
    if (semantics != PWGS_SEM_MH && semantics != PWGS_SEM_PY) return fail(PWGS_EINVAL, "unknown semantics");
    if (cols == nullptr) n_cols = c->K;
    if (n_cols < 0) return fail(PWGS_EINVAL, "n_cols must be >= 0");
    if (n_rows == 0 || n_cols == 0) return PWGS_OK;
    if (rows) { for (int r = 0; r < n_rows; r++) if (rows[r] < 0 || rows[r] >= c->D) return fail(PWGS_EINVAL, "row index out of range"); }
    else if (row0 < 0 || row0 > c->D - n_rows) return fail(PWGS_EINVAL, "row range out of range");
    if (cols) for (int k = 0; k < n_cols; k++) if (cols[k] < 0 || cols[k] >= c->K) return fail(PWGS_EINVAL, "node index out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    size_t total = (size_t)n_rows * n_cols;
    // row / column lists through the page-locked staging area (no implicit stream synchronisation)
    const size_t nr = rows ? (size_t)n_rows : 0, nc = cols ? (size_t)n_cols : 0;
    if (nr + nc) {
        CUDA_TRY(stage_acquire(c, (nr + nc) * sizeof(int)));
        int *h = (int *)c->stage.p;
        if (nr) memcpy(h, rows, nr * sizeof(int));
        if (nc) memcpy(h + nr, cols, nc * sizeof(int));
        CUDA_TRY(c->rows.alloc(nr + nc));
        CUDA_TRY(cudaMemcpyAsync(c->rows.p, h, (nr + nc) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(stage_release(c));
    }
    CUDA_TRY(c->scratch.alloc(total));
    CUDA_TRY(cudaEventRecord(c->ev2, c->stream));
    if (!rows && col_major && n_cols <= 65535 && !c->opt_grid_plain) {
        // a range of data, column-major (the sweep's calls): plain data and CNV-affected SSMs of the range in separate warps
        const std::vector<int> &pl = c->pidx_host, &cv = c->ssm_of_m;
        const int p0 = (int)(std::lower_bound(pl.begin(), pl.end(), row0) - pl.begin());
        const int p1 = (int)(std::lower_bound(pl.begin(), pl.end(), row0 + n_rows) - pl.begin());
        const int c0 = (int)(std::lower_bound(cv.begin(), cv.end(), row0) - cv.begin());
        const int c1 = (int)(std::lower_bound(cv.begin(), cv.end(), row0 + n_rows) - cv.begin());
        const int np = p1 - p0, nc = c1 - c0, np_pad = (np + 31) & ~31;
        dim3 grid((unsigned)((np_pad + nc + 127) / 128), (unsigned)n_cols);
        pair_llh_sorted_kernel<<<grid, 128, 0, c->stream>>>(data_view(c), tree_view(c), row0, n_rows, np, np_pad, c->pidx.p + p0, nc, c->cidx.p + c0,
                                                            cols ? c->rows.p + nr : nullptr, semantics, kLogTinyMh, kLogTinyPy, kLogQuarter, c->scratch.p);
    } else
    pair_llh_kernel<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(data_view(c), tree_view(c), 0, row0, n_rows, rows ? c->rows.p : nullptr, n_cols,
                                                                            cols ? c->rows.p + nr : nullptr, col_major, semantics, kLogTinyMh, kLogTinyPy,
                                                                            kLogQuarter, c->scratch.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev3, c->stream));
    c->n_launches++;
    if (out_pitch > 0) {                                   // column-major block into the columns of a wider host array
        if (!col_major || out_pitch < n_rows) return fail(PWGS_EINVAL, "a pitched output needs PWGS_LAYOUT_COLS and pitch >= n_rows");
        CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)out_pitch * sizeof(double), c->scratch.p, (size_t)n_rows * sizeof(double),
                                   (size_t)n_rows * sizeof(double), (size_t)n_cols, cudaMemcpyDeviceToHost, c->stream));
    } else
        CUDA_TRY(cudaMemcpyAsync(out, c->scratch.p, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->grid_timed = true;
    return PWGS_OK;
}

int pwgs_llh_block(pwgs_ctx *c, int n_rows, const int32_t *rows, int n_cols, const int32_t *cols, int semantics, double *out)
{
    if (n_rows > 0 && !rows) return fail(PWGS_EINVAL, "rows/out NULL");
    return llh_block_impl(c, 0, n_rows, rows, n_cols, cols, semantics, out);
}

int pwgs_llh_range(pwgs_ctx *c, int row0, int n_rows, int n_cols, const int32_t *cols, int semantics, double *out)
{
    return llh_block_impl(c, row0, n_rows, nullptr, n_cols, cols, semantics, out);
}

int pwgs_llh_range_pitched(pwgs_ctx *c, int row0, int n_rows, int n_cols, const int32_t *cols, int semantics, double *out, int64_t out_pitch)
{
    if (out_pitch <= 0) return fail(PWGS_EINVAL, "out_pitch must be positive");
    return llh_block_impl(c, row0, n_rows, nullptr, n_cols, cols, semantics, out, out_pitch);
}

int pwgs_llh_rows(pwgs_ctx *c, int n_rows, const int32_t *rows, int semantics, double *out)
{
    return pwgs_llh_block(c, n_rows, rows, 0, nullptr, semantics, out);
}

int pwgs_total_llh(pwgs_ctx *c, int semantics, double *out)
{
    if (!c || !out) return fail(PWGS_EINVAL, "NULL argument");
    if (!c->tree_set) return fail(PWGS_ESTATE, "pwgs_set_tree must be called first");
    if (semantics == PWGS_SEM_MH) {
        // exactly the likelihood path of the MH kernel: zero iterations leave the state alone and report its posterior
        return pwgs_mh_run(c, 0, 100.0, 0, 0, nullptr, nullptr, nullptr, out);
    }
    if (semantics != PWGS_SEM_PY) return fail(PWGS_EINVAL, "unknown semantics");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(c->scratch.alloc((size_t)c->D + 1));
    pair_llh_kernel<<<(c->D + 127) / 128, 128, 0, c->stream>>>(data_view(c), tree_view(c), 1, 0, c->D, nullptr, 0, nullptr, 0, semantics,
                                                               kLogTinyMh, kLogTinyPy, kLogQuarter, c->scratch.p);
    CUDA_TRY(cudaGetLastError());
    sum_kernel<<<1, 1024, 0, c->stream>>>(c->D, c->scratch.p, c->scratch.p + c->D);
    CUDA_TRY(cudaGetLastError());
    c->n_launches += 2;
    CUDA_TRY(cudaMemcpyAsync(out, c->scratch.p + c->D, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PWGS_OK;
}

int pwgs_get_lbc(pwgs_ctx *c, double *out)
{
    if (!c || !out) return fail(PWGS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(out, c->lbc.p, (size_t)c->D * c->T * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PWGS_OK;
}

int pwgs_get_data_states(pwgs_ctx *c, int32_t *M, int32_t *ssm_of_m, int32_t *coef)
{
    if (!c) return fail(PWGS_EINVAL, "ctx is NULL");
    if (M) *M = c->M;
    if (!ssm_of_m && !coef) return PWGS_OK;
    if (!c->tree_set) return fail(PWGS_ESTATE, "pwgs_set_tree must be called first");
    if (ssm_of_m) for (int m = 0; m < c->M; m++) ssm_of_m[m] = c->ssm_of_m[m];
    if (coef && c->M > 0) {
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(c->coef.alloc((size_t)c->M * c->K));
        data_states_kernel<<<(c->M + 127) / 128, 128, 0, c->stream>>>(data_view(c), tree_view(c), c->M, c->cidx.p, c->coef.p);
        CUDA_TRY(cudaGetLastError());
        c->n_launches++;
        std::vector<int4> h((size_t)c->M * c->K);
        CUDA_TRY(cudaMemcpyAsync(h.data(), c->coef.p, h.size() * sizeof(int4), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const int K = c->K;
        for (int m = 0; m < c->M; m++)
            for (int k = 0; k < K; k++) {
                int4 v = h[(size_t)k * c->M + m];
                int w[4] = {v.x, v.y, v.z, v.w};
                for (int q = 0; q < 4; q++) {          // unpack to the c_data_states layout [m][state][node][nr,nv]
                    coef[(((size_t)m * 4 + q) * K + k) * 2 + 0] = (short)(w[q] & 0xffff);
                    coef[(((size_t)m * 4 + q) * K + k) * 2 + 1] = w[q] >> 16;
                }
            }
    }
    return PWGS_OK;
}

int pwgs_set_option(pwgs_ctx *c, const char *name, int64_t value)
{
    if (!c || !name) return fail(PWGS_EINVAL, "NULL argument");
    std::string n(name);
    if (n == "mh_batch") {
        if (value < 1 || value > PWGS_MAX_BATCH || (value & (value - 1))) return fail(PWGS_EINVAL, "mh_batch must be a power of two in 1..256");
        c->opt_batch = (int)value;
    } else if (n == "mh_ctas") {
        if (value < 0) return fail(PWGS_EINVAL, "mh_ctas must be >= 0");
        c->opt_ctas = (int)value;
    } else if (n == "mh_dist") {
        if (value < -1 || value > 1) return fail(PWGS_EINVAL, "mh_dist must be -1 (auto), 0 or 1");
        c->opt_dist = (int)value;
    } else if (n == "mh_big") {
        if (value < -1 || value > 1) return fail(PWGS_EINVAL, "mh_big must be -1 (auto), 0 or 1");
        c->opt_big = (int)value;
    } else if (n == "grid_plain") {
        c->opt_grid_plain = value != 0;
    } else if (n == "mh_collapse") {
        c->opt_collapse = value != 0;
    } else if (n == "mh_stage") {
        c->opt_stage = value != 0;
    } else if (n == "mh_profile") {
        c->opt_prof = value != 0;
    } else if (n == "mh_w_plain" || n == "mh_w_cnv1" || n == "mh_w_cnv2" || n == "mh_w_cnv3") {
        if (value < 1 || value > 1000000) return fail(PWGS_EINVAL, "round weights must be in 1..1000000");
        c->opt_w[n == "mh_w_plain" ? 0 : n == "mh_w_cnv1" ? 1 : n == "mh_w_cnv2" ? 2 : 3] = (int)value;
        c->work_valid = false;
    } else return fail(PWGS_EINVAL, "unknown option " + n);
    return PWGS_OK;
}

// Work statistics of the sparse CNV list (for the roofline model): number of CNV-affected SSMs with 1 / 2 / 3 merged states
// and sum over SSMs of entries x states.
static int cnv_work_stats(pwgs_ctx *c, int64_t out[4])
{
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!c->tree_set) return fail(PWGS_ESTATE, "pwgs_set_tree must be called first");
    if (c->M == 0) return PWGS_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    int cpc = c->opt_ctas > 0 ? std::min(c->opt_ctas, c->n_sms) : c->n_sms;
    int rc = prepare_cnv_work(c, cpc, (c->opt_collapse && c->n_classes > 0 && c->Dp > 0) ? c->n_classes : 0, c->work_ncls > 0 ? c->work_ncls : mh_group_classes(c->opt_batch));
    if (rc != PWGS_OK) return rc;
    std::vector<int> code(c->M);
    CUDA_TRY(cudaMemcpyAsync(code.data(), c->code_tmp.p, sizeof(int) * c->M, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int m = 0; m < c->M; m++) {
        int ne = code[m] & 0xff, nu = (code[m] >> 8) & 3;
        if (nu >= 1 && nu <= 3) out[nu - 1]++;
        out[3] += (int64_t)ne * nu;
    }
    return PWGS_OK;
}

int pwgs_get_option(pwgs_ctx *c, const char *name, int64_t *value)
{
    if (!c || !name || !value) return fail(PWGS_EINVAL, "NULL argument");
    std::string n(name);
    if (n == "cnv_nu1" || n == "cnv_nu2" || n == "cnv_nu3" || n == "cnv_entry_states") {
        int64_t st[4];
        int rc = cnv_work_stats(c, st);
        if (rc != PWGS_OK) return rc;
        *value = st[n == "cnv_nu1" ? 0 : n == "cnv_nu2" ? 1 : n == "cnv_nu3" ? 2 : 3];
        return PWGS_OK;
    }
    if (n == "mh_batch") *value = c->opt_batch;
    else if (n == "mh_ctas") *value = c->opt_ctas;
    else if (n == "n_sms") *value = c->n_sms;
    else if (n == "grid_kernel_ns") {                     // device time of the last pwgs_llh_rows / _block / _range kernel
        float ms = 0.f;
        if (c->grid_timed) CUDA_TRY(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
        *value = (int64_t)(ms * 1e6);
    }
    else if (n == "n_plain") *value = c->Dp;
    else if (n == "n_cnv_ssm") *value = c->M;
    else if (n == "launches") *value = c->n_launches;
    else if (n == "n_classes") *value = c->n_classes;
    else if (n == "mh_collapse") *value = c->opt_collapse;
    else if (n.rfind("mh_lik_cta_", 0) == 0) {
        const int r = atoi(n.c_str() + 11);
        if (r < 0 || r >= c->prof_cpc) return fail(PWGS_EINVAL, "no such CTA");
        *value = (int64_t)c->prof_host[16 + r];
    }
    else if (n == "mh_plan_classes") *value = c->work_ncls;
    else if (n.rfind("mh_plan_", 0) == 0) {         // mh_plan_<class>_<cta>: rounds of class 0 (plain thirds) / 1..3 (CNV states) of a CTA
        int cls = 0, r = 0;
        if (sscanf(n.c_str() + 8, "%d_%d", &cls, &r) != 2 || cls < 0 || cls > 3 || r < 0 || (size_t)r * 4 + 3 >= c->plan_counts.size())
            return fail(PWGS_EINVAL, "no such plan entry");
        *value = c->plan_counts[(size_t)r * 4 + cls];
    }
    else if (n == "mh_lik_min" || n == "mh_lik_max" || n == "mh_lik_sum") {
        unsigned long long mn = ~0ull, mx = 0, sm = 0;
        for (int i = 0; i < c->prof_cpc; i++) { const unsigned long long v = c->prof_host[16 + i]; mn = std::min(mn, v); mx = std::max(mx, v); sm += v; }
        *value = (int64_t)(n == "mh_lik_min" ? (c->prof_cpc ? mn : 0) : n == "mh_lik_max" ? mx : sm);
    }
    else if (n.rfind("mh_prof_", 0) == 0 && n.size() == 9 && n[8] >= '0' && n[8] <= '9') *value = (int64_t)c->prof_host[n[8] - '0'];
    else if (n.rfind("mh_prof_1", 0) == 0 && n.size() == 10 && n[9] >= '0' && n[9] <= '5') *value = (int64_t)c->prof_host[10 + n[9] - '0'];
    else return fail(PWGS_EINVAL, "unknown option " + n);
    return PWGS_OK;
}

int pwgs_get_stream(pwgs_ctx *c, void **stream)
{
    if (!c || !stream) return fail(PWGS_EINVAL, "NULL argument");
    *stream = (void *)c->stream;
    return PWGS_OK;
}

int pwgs_math_selftest(int device, int n, const double *x, double *out)
{
    if (!x || !out || n <= 0) return fail(PWGS_EINVAL, "bad argument");
    CUDA_TRY(cudaSetDevice(device));
    double *dx = nullptr, *dout = nullptr;
    CUDA_TRY(cudaMalloc((void **)&dx, sizeof(double) * n));
    CUDA_TRY(cudaMalloc((void **)&dout, sizeof(double) * 3 * (size_t)n));
    CUDA_TRY(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    math_selftest_kernel<<<(n + 4 * 128 - 1) / (4 * 128), 128>>>(n, dx, dout);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dout);
    return PWGS_OK;
}

int pwgs_plan_partition(int cpc, int M, const int32_t *n_states, int Dp, int ncls, const int32_t *w,
                        int32_t *round_slot, int32_t *cta_nc, int32_t *pstart, int32_t *cchunk, int32_t *pchunk)
{
    if (cpc < 1 || M < 0 || Dp < 0 || ncls < 1 || ncls > MH_PCLASSES || !n_states || !w || !round_slot || !cta_nc || !pstart || !cchunk || !pchunk)
        return fail(PWGS_EINVAL, "pwgs_plan_partition: bad argument");
    if (n_states[0] + n_states[1] + n_states[2] != M) return fail(PWGS_EINVAL, "n_states must add up to M");
    const int ns[3] = {n_states[0], n_states[1], n_states[2]}, ww[4] = {w[0], w[1], w[2], w[3]};
    const Partition P = plan_partition(cpc, M, ns, Dp, ww, ncls);
    for (int u = 0; u < (M + 31) / 32; u++) round_slot[u] = P.round_slot[u];
    for (int i = 0; i < cpc; i++) cta_nc[i] = P.cta_nc[i];
    for (size_t i = 0; i < P.pstart.size(); i++) pstart[i] = P.pstart[i];
    *cchunk = P.cchunk; *pchunk = P.pchunk;
    return PWGS_OK;
}

int pwgs_selftest_llh_host(int n_ssm, int n_cnv, int ntps, const int32_t *a, const int32_t *d, const double *mu_r, const double *mu_v,
                           const int32_t *cnv_ptr, const int32_t *cnv_idx, const int32_t *cnv_c1, const int32_t *cnv_c2,
                           int K, const int32_t *parent, const int32_t *node_of_datum, const double *phi, const double *pi,
                           int n_rows, const int32_t *rows, int semantics, double *out)
{
    if (n_ssm < 0 || n_cnv < 0 || ntps < 1 || K < 1 || n_rows < 0 || !a || !d || !mu_r || !mu_v || !parent || !node_of_datum || !phi || !pi ||
        (n_rows > 0 && (!rows || !out)))
        return fail(PWGS_EINVAL, "pwgs_selftest_llh_host: bad argument");
    if (semantics != PWGS_SEM_MH && semantics != PWGS_SEM_PY) return fail(PWGS_EINVAL, "unknown semantics");
    const int D = n_ssm + n_cnv;
    HostTree ht;
    std::string err;
    if (!host_tree_prep(K, ntps, parent, pi, ht, err)) return fail(PWGS_EINVAL, err);
    std::vector<double> lbc((size_t)D * ntps);
    for (size_t i = 0; i < lbc.size(); i++) lbc[i] = lgamma((double)d[i] + 1.0) - lgamma((double)a[i] + 1.0) - lgamma((double)(d[i] - a[i]) + 1.0);
    std::vector<int32_t> zero_ptr(n_ssm + 1, 0);
    DataView dv;
    dv.n_ssm = n_ssm; dv.n_cnv = n_cnv; dv.T = ntps;
    dv.a = a; dv.d = d; dv.mu_r = mu_r; dv.mu_v = mu_v; dv.lbc = lbc.data();
    dv.cnv_ptr = cnv_ptr ? cnv_ptr : zero_ptr.data(); dv.cnv_idx = cnv_idx; dv.cnv_c1 = cnv_c1; dv.cnv_c2 = cnv_c2;
    TreeView tv;
    tv.K = K; tv.T = ntps;
    tv.parent = parent; tv.depth = ht.depth.data(); tv.tin = ht.tin.data(); tv.tout = ht.tout.data();
    tv.phi = phi; tv.pi = pi; tv.node_of_datum = node_of_datum; tv.phisub = ht.phisub.data(); tv.root = ht.root;
    for (int r = 0; r < n_rows; r++) {
        if (rows[r] < 0 || rows[r] >= D) return fail(PWGS_EINVAL, "row index out of range");
        for (int k = 0; k < K; k++)
            out[(size_t)r * K + k] = datum_ll_at(dv, tv, rows[r], k, semantics, kLogTinyMh, kLogTinyPy, kLogQuarter);
    }
    return PWGS_OK;
}

int pwgs_fp64_peak(int device, double *dfma_per_s)
{
    if (!dfma_per_s) return fail(PWGS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    double *out;
    CUDA_TRY(cudaMalloc((void **)&out, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 20000, block = 512, grid = prop.multiProcessorCount * 4;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        dfma_peak_kernel<<<grid, block>>>(iters, 1.0 + rep, out);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double)grid * block * 8.0 * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *dfma_per_s = best;
    return PWGS_OK;
}

} // extern "C"
