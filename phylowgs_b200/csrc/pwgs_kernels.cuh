// pwgs_kernels.cuh — the sm_100a kernels of libpwgs_b200.so.
//   K0  lbc_kernel            log_bin_coeff(d,a) per datum x sample            (util.cpp:12-14 @ mh.cpp:234,284)
//   K0b data_states_kernel    dense integer (nr,nv) state coefficients         (params.py:114-171; diagnostics only)
//   K0c cnv_sparse/sort/scatter  sparse, state-merged, sorted CNV work list    (what K1 reads instead of K0b's table)
//   K0d collapse_kernel       integer totals of the plain data per (node, class, sample)   (opt-in "mh_collapse")
//   K1  mh_kernel<STAGED>     persistent Metropolis-Hastings sampler           (mh.cpp:65-177, mh.hpp:80-163, util.cpp:24-33)
//   K2  pair_llh_kernel       datum x candidate-node log-likelihoods           (data.py:26-126 / mh.hpp:80-163)
// Citations are file:line relative to the reference checkout.
#pragma once
#include "pwgs_device.cuh"
#include <type_traits>

// ============================================================================ K0
__global__ void lbc_kernel(int n, const int *__restrict__ a, const int *__restrict__ d, double *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int dd = d[i], aa = a[i];
        out[i] = lgamma((double)dd + 1.0) - lgamma((double)aa + 1.0) - lgamma((double)(dd - aa) + 1.0);
    }
}

// fixed-order single-block sum (deterministic): out[0] = sum_{i<n} x[idx ? idx[i]*T+tp : i]
__global__ void sum_lbc_plain_kernel(int Dp, int T, const int *__restrict__ pidx, const double *__restrict__ lbc, double *out)
{
    __shared__ double red[32];
    double s = 0;
    for (int i = threadIdx.x; i < Dp; i += blockDim.x)
        for (int tp = 0; tp < T; tp++) s += lbc[pidx[i] * T + tp];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (blockDim.x >> 5); w++) t += red[w];
        out[0] = t;
    }
}

// gather kernels for the permuted (plain-first) work lists
__global__ void gather_node_kernel(int n, const int *__restrict__ idx, const int *__restrict__ node_of_datum, int *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = node_of_datum[idx[i]];
}

// assignments of a few data changed (pwgs_move_data)
__global__ void move_data_kernel(int n, const int *__restrict__ datum, const int *__restrict__ node, int *__restrict__ node_of_datum)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) node_of_datum[datum[i]] = node[i];
}

// ============================================================================ K0b
// One thread per CNV-affected SSM m.  Pass 1 accumulates the tp=0 genome counts exactly like
// compute_n_genomes(0) (params.py:120) to decide the substitutions of params.py:160-166; pass 2 writes the final
// coefficients, packed as 8 x int16 {nr1,nv1,nr2,nv2,nr3,nv3,nr4,nv4} at coef[k*M + m] (k-major: coalesced in m).
__global__ void data_states_kernel(DataView dv, TreeView tv, int M, const int *__restrict__ ssm_of_m, int4 *__restrict__ coef)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int j = ssm_of_m[m], s = tv.node_of_datum[j];
    double pv0 = 0, pv1 = 0;
    for (int k = 0; k < tv.K; k++) {
        int nr[4], nv[4];
        state_coefs(dv, tv, j, s, k, nr, nv);
        double p = tv.pi[k * tv.T];
        pv0 += p * nv[0];
        pv1 += p * nv[1];
    }
    bool four = (dv.cnv_ptr[j + 1] - dv.cnv_ptr[j] == 1) &&
                (s == tv.node_of_datum[dv.n_ssm + dv.cnv_idx[dv.cnv_ptr[j]]]);      // data.py:122-123
    for (int k = 0; k < tv.K; k++) {
        int nr[4], nv[4];
        state_coefs(dv, tv, j, s, k, nr, nv);
        if (pv0 == 0) { nr[0] = nr[1]; nv[0] = nv[1]; }
        else if (pv1 == 0) { nr[1] = nr[0]; nv[1] = nv[0]; }
        if (!four) { nr[2] = nr[0]; nv[2] = nv[0]; nr[3] = nr[1]; nv[3] = nv[1]; }
        int4 o;
        o.x = (nr[0] & 0xffff) | (nv[0] << 16);
        o.y = (nr[1] & 0xffff) | (nv[1] << 16);
        o.z = (nr[2] & 0xffff) | (nv[2] << 16);
        o.w = (nr[3] & 0xffff) | (nv[3] << 16);
        coef[(size_t)k * M + m] = o;
    }
}

// ============================================================================ K0c  sparse CNV work list for the MH kernel
// For a CNV-affected SSM the state coefficients c[q][k] of params.py:114-171 are piecewise constant over the tree: they depend
// on node k only through "is k under the SSM's node" and "which linked CNV is the most recent one above k".  Writing
// c[q][k] = sum over ancestors-or-self a of k of delta[q][a] gives  sum_k pi_k c[q][k] = sum_a delta[q][a] * phi_a  (phi = subtree
// sums, mh.cpp:134-141), and delta is non-zero only at the root, at the SSM's node and at nodes hosting one of its CNVs:
// <= 2 + (#links) entries instead of K.  States that coincide on every node (st3 == st4 always; st1 == st2 == st3 for most
// 2-state SSMs) are merged into one "unique state" with a multiplicity, so  logsumexp_q(ll_q) = logsumexp_u(ll_u + log w_u).
//
// code[m]: bits 0..7 number of entries, bits 8..9 number of unique states (1..3), bits 12..14 / 16..18 / 20..22 multiplicities.
#define CNV_CODE(ne, nu, w0, w1, w2) ((ne) | ((nu) << 8) | ((w0) << 12) | ((w1) << 16) | ((w2) << 20))

__device__ __forceinline__ void final_state_coefs(const DataView &dv, const TreeView &tv, int j, int s, int k, bool sub0, bool sub1,
                                                  bool four, int nr[4], int nv[4])
{
    state_coefs(dv, tv, j, s, k, nr, nv);
    if (sub0) { nr[0] = nr[1]; nv[0] = nv[1]; }                       // params.py:160-161
    else if (sub1) { nr[1] = nr[0]; nv[1] = nv[0]; }                  // params.py:162-163
    if (!four) { nr[2] = nr[0]; nv[2] = nv[0]; nr[3] = nr[1]; nv[3] = nv[1]; }   // params.py:164-166
}

// One thread per CNV-affected SSM m (input order).  Writes code[m] and up to Emax entries (enode, ecoef) at [e*M + m];
// ecoef packs (nr,nv) deltas of unique states 0..2 as int16 pairs in x,y,z.  *overflow is set if an SSM needs more entries.
__global__ void cnv_sparse_kernel(DataView dv, TreeView tv, int M, int Emax, const int *__restrict__ ssm_of_m,
                                  int *__restrict__ code, int *__restrict__ enode, int4 *__restrict__ ecoef, int *overflow)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int j = ssm_of_m[m], s = tv.node_of_datum[j], K = tv.K;
    double pv0 = 0, pv1 = 0;
    for (int k = 0; k < K; k++) {
        int nr[4], nv[4];
        state_coefs(dv, tv, j, s, k, nr, nv);
        double p = tv.pi[k * tv.T];
        pv0 += p * nv[0];
        pv1 += p * nv[1];
    }
    const bool sub0 = pv0 == 0, sub1 = !sub0 && pv1 == 0;
    const bool four = (dv.cnv_ptr[j + 1] - dv.cnv_ptr[j] == 1) && (s == tv.node_of_datum[dv.n_ssm + dv.cnv_idx[dv.cnv_ptr[j]]]);
    bool e01 = true, e02 = true, e12 = true;                           // state 3 always equals state 2 (or state 1 when !four)
    for (int k = 0; k < K; k++) {
        int nr[4], nv[4];
        final_state_coefs(dv, tv, j, s, k, sub0, sub1, four, nr, nv);
        e01 &= nr[0] == nr[1] && nv[0] == nv[1];
        e02 &= nr[0] == nr[2] && nv[0] == nv[2];
        e12 &= nr[1] == nr[2] && nv[1] == nv[2];
    }
    // unique states in first-occurrence order with multiplicities over the four states {0,1,2,3}; 3 mirrors 2 when four, else 1
    int uq[3] = {0, -1, -1}, w[3] = {1, 0, 0}, nu = 1;
    int rep1 = e01 ? 0 : 1;
    if (rep1 == 1) { uq[nu] = 1; w[nu] = 1; nu++; } else w[0]++;
    int rep2 = e02 ? 0 : (e12 ? 1 : 2);
    int twice = 2;                                                     // states 2 and 3 coincide (st3 == st4 / mirrored pair below)
    if (!four) {
        // 2-state SSM: state 2 := state 0, state 3 := state 1
        w[0] += 1;
        if (rep1 == 1) w[1] += 1; else w[0] += 1;
    } else {
        if (rep2 == 2) { uq[nu] = 2; w[nu] = twice; nu++; }
        else if (rep2 == 0) w[0] += twice;
        else w[1] += twice;                                            // rep2 == 1 implies rep1 == 1: unique index 1 is state 1
    }
    int ne = 0;
    for (int k = 0; k < K; k++) {
        int nr[4], nv[4], pr[4] = {0, 0, 0, 0}, pvv[4] = {0, 0, 0, 0};
        final_state_coefs(dv, tv, j, s, k, sub0, sub1, four, nr, nv);
        if (tv.parent[k] >= 0) final_state_coefs(dv, tv, j, s, tv.parent[k], sub0, sub1, four, pr, pvv);
        int d[3] = {0, 0, 0};
        bool any = false;
        for (int u = 0; u < nu; u++) {
            int q = uq[u];
            int dr = nr[q] - pr[q], dvv = nv[q] - pvv[q];
            d[u] = (dr & 0xffff) | (dvv << 16);
            any |= (dr != 0) || (dvv != 0);
        }
        if (any) {
            if (ne < Emax) {
                enode[(size_t)ne * M + m] = k;
                ecoef[(size_t)ne * M + m] = make_int4(d[0], d[1], d[2], 0);
            } else *overflow = 1;
            ne++;
        }
    }
    code[m] = CNV_CODE(min(ne, Emax), nu, w[0], w[1], w[2]);
}

// Stable 3-bucket counting sort of the CNV-affected SSMs by their number of unique states, most expensive first (single CTA,
// deterministic): pos[m] = rank of m in the sorted order.  Warps of the MH kernel then see (nearly) uniform work per round and
// the partially filled last warp of a CTA holds the cheapest SSMs.
__global__ void cnv_sort_kernel(int M, const int *__restrict__ code, int *__restrict__ pos, int *__restrict__ counts)
{
    __shared__ int cnt[1024][3];
    __shared__ int base[3];
    const int t = threadIdx.x, per = (M + blockDim.x - 1) / blockDim.x, lo = min(M, t * per), hi = min(M, lo + per);
    int c[3] = {0, 0, 0};
    for (int m = lo; m < hi; m++) c[3 - ((code[m] >> 8) & 3)]++;
    cnt[t][0] = c[0]; cnt[t][1] = c[1]; cnt[t][2] = c[2];
    __syncthreads();
    if (t == 0) {
        int run[3] = {0, 0, 0};
        for (int i = 0; i < (int)blockDim.x; i++)
            for (int b = 0; b < 3; b++) { int v = cnt[i][b]; cnt[i][b] = run[b]; run[b] += v; }
        base[0] = 0; base[1] = run[0]; base[2] = run[0] + run[1];
        counts[0] = run[0]; counts[1] = run[1]; counts[2] = run[2];     // SSMs with 3 / 2 / 1 merged states
    }
    __syncthreads();
    int o[3] = {base[0] + cnt[t][0], base[1] + cnt[t][1], base[2] + cnt[t][2]};
    for (int m = lo; m < hi; m++) pos[m] = o[3 - ((code[m] >> 8) & 3)]++;
}

// Scatter into the CTA-major layout the MH kernel streams.  The sorted list is cut into warp rounds of 32 consecutive SSMs
// (uniform state count inside a round but for the two rounds straddling a class boundary); the host deals whole rounds to the
// CTAs by cost (pwgs_api.cu: plan_partition) and passes round_slot[u] = first slot of round u: sorted rank r goes to
// round_slot[r / 32] + r % 32.  A CTA's rounds are contiguous, most expensive first; it holds cta_nc[rank] SSMs.
struct CnvWork {
    int *a, *d;            // [T][Mp]
    double *lbc;           // [T][Mp]
    double *mu_r;          // [Mp]
    int *code;             // [Mp]
    int *enode;            // [Emax][Mp]
    int4 *ecoef;           // [Emax][Mp]
    int Mp, cchunk, Emax;
    const int *cta_nc;     // [cpc] CNV-affected SSMs of each CTA
};
__global__ void cnv_scatter_kernel(int M, int T, const int *__restrict__ round_slot, CnvWork w, const int *__restrict__ pos, const int *__restrict__ ca,
                                   const int *__restrict__ cd, const double *__restrict__ clbc, const double *__restrict__ cmu_r,
                                   const int *__restrict__ code, const int *__restrict__ enode, const int4 *__restrict__ ecoef)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int r = pos[m], dst = round_slot[r >> 5] + (r & 31);
    for (int tp = 0; tp < T; tp++) {
        w.a[(size_t)tp * w.Mp + dst] = ca[(size_t)tp * M + m];
        w.d[(size_t)tp * w.Mp + dst] = cd[(size_t)tp * M + m];
        w.lbc[(size_t)tp * w.Mp + dst] = clbc[(size_t)tp * M + m];
    }
    w.mu_r[dst] = cmu_r[m];
    const int cd_ = code[m];
    w.code[dst] = cd_;
    for (int e = 0; e < (cd_ & 0xff); e++) {
        w.enode[(size_t)e * w.Mp + dst] = enode[(size_t)e * M + m];
        w.ecoef[(size_t)e * w.Mp + dst] = ecoef[(size_t)e * M + m];
    }
}

// ============================================================================ K2
// llh of datum j placed in node s under either convention (see oracle/pwgs_oracle.c datum_ll_cpp_at / orc_datum_ll_py).
// Sum over all nodes k of pi_k * (nr, nv)_k for the four states of SSM j sitting in node s, at sample tp, WITHOUT walking the
// nodes: the integer coefficients depend on k only through "is k under s" and "which linked CNV is the most recent above k"
// (state_coefs), so they change, going down the tree, only at s and at the nodes hosting j's CNVs.  With delta(a) =
// coef(a) - coef(parent(a)) (coef of the root's parent = 0) the sum is sum_a delta(a) * phisub(a) over those few nodes, phisub
// being the subtree sums of pi.  O(links^2) per (datum, node) pair instead of O(K * links): the grid of a 100-node tree in the
// time the dense walk needed for 10 nodes.  (K0c uses the same identity for the MH kernel's work list.)
__host__ __device__ inline void cnv_state_sums(const DataView &dv, const TreeView &tv, int j, int s, int tp, double nr[4], double nv[4])
{
    const int l0 = dv.cnv_ptr[j], L = dv.cnv_ptr[j + 1] - l0, T = dv.T;
#pragma unroll
    for (int q = 0; q < 4; q++) { nr[q] = 0.0; nv[q] = 0.0; }
    for (int i = 0; i < L + 2; i++) {                                       // candidates: root, s, the CNVs' nodes
        const int a = i == 0 ? tv.root : (i == 1 ? s : tv.node_of_datum[dv.n_ssm + dv.cnv_idx[l0 + i - 2]]);
        bool seen = false;
        for (int i2 = 0; i2 < i && !seen; i2++)
            seen = a == (i2 == 0 ? tv.root : (i2 == 1 ? s : tv.node_of_datum[dv.n_ssm + dv.cnv_idx[l0 + i2 - 2]]));
        if (seen) continue;
        int cr[4], cv[4], pr[4] = {0, 0, 0, 0}, pv[4] = {0, 0, 0, 0};
        state_coefs(dv, tv, j, s, a, cr, cv);
        if (tv.parent[a] >= 0) state_coefs(dv, tv, j, s, tv.parent[a], pr, pv);
        const double ph = tv.phisub[a * T + tp];
#pragma unroll
        for (int q = 0; q < 4; q++) { nr[q] += ph * (double)(cr[q] - pr[q]); nv[q] += ph * (double)(cv[q] - pv[q]); }
    }
    // The callers DECIDE on these sums (a state with nv == 0 is dropped or substituted, nr + nv == 0 counts log 1e-99), and the
    // signed deltas above can cancel to an exact 0 or a tiny negative where the reference's sum of non-negative terms is a tiny
    // positive (and the other way round).  So whenever a sum a decision looks at is within rounding of zero, take all of them
    // from the reference's own form, sum_k pi_k * coef_k: non-negative terms, zero exactly when every term is.  Rare (a node
    // with pi ~ 0 carrying all of a state's variant copies), O(K) then.
    bool near_zero = false;
#pragma unroll
    for (int q = 0; q < 4; q++) near_zero |= fabs(nv[q]) < 1e-7 || fabs(nr[q] + nv[q]) < 1e-7;
    if (near_zero) {
#pragma unroll
        for (int q = 0; q < 4; q++) { nr[q] = 0.0; nv[q] = 0.0; }
        for (int k = 0; k < tv.K; k++) {
            int cr[4], cv[4];
            state_coefs(dv, tv, j, s, k, cr, cv);
            const double p = tv.pi[k * T + tp];
#pragma unroll
            for (int q = 0; q < 4; q++) { nr[q] += p * (double)cr[q]; nv[q] += p * (double)cv[q]; }
        }
    }
}

__host__ __device__ inline double datum_ll_at(const DataView &dv, const TreeView &tv, int j, int s, int semantics,
                                              double log_tiny_mh, double log_tiny_py, double log_quarter)
{
    int T = dv.T;
    double mu_r = dv.mu_r[j], mu_v = dv.mu_v[j];
    bool cnv_ssm = j < dv.n_ssm && dv.cnv_ptr[j + 1] > dv.cnv_ptr[j];
    double total = 0;
    if (!cnv_ssm) {
        for (int tp = 0; tp < T; tp++) {
            double phi = tv.phi[s * T + tp];
            double mu = (1.0 - phi) * mu_r + phi * mu_v;
            total += log_binomial_likelihood(dv.a[j * T + tp], dv.d[j * T + tp], mu) + dv.lbc[j * T + tp];
        }
        return total;
    }
    bool four = (dv.cnv_ptr[j + 1] - dv.cnv_ptr[j] == 1) &&
                (s == tv.node_of_datum[dv.n_ssm + dv.cnv_idx[dv.cnv_ptr[j]]]);
    bool sub0 = false, sub1 = false;
    for (int tp = 0; tp < T; tp++) {
        double nr[4], nv[4];
        cnv_state_sums(dv, tv, j, s, tp, nr, nv);
        int a = dv.a[j * T + tp], d = dv.d[j * T + tp];
        double lbc = dv.lbc[j * T + tp], llh;
        if (semantics == 0) {                                   // mh.hpp:80-163 after params.py:160-166
            if (tp == 0) { sub0 = (nv[0] == 0); sub1 = !sub0 && (nv[1] == 0); }
            if (sub0) { nr[0] = nr[1]; nv[0] = nv[1]; }
            else if (sub1) { nr[1] = nr[0]; nv[1] = nv[0]; }
            if (!four) { nr[2] = nr[0]; nv[2] = nv[0]; nr[3] = nr[1]; nv[3] = nv[1]; }
            double ll[4];
#pragma unroll
            for (int q = 0; q < 4; q++) ll[q] = mh_state_ll(nr[q], nv[q], a, d, mu_r, lbc, log_quarter, log_tiny_mh);
            llh = logsumexp4(ll);
        } else {                                                // data.py:47-62
            int ns = four ? 4 : 2, kept = 0;
            for (int q = 0; q < ns; q++) kept += (nv[q] > 0);
            double ll[4], mx = -INFINITY;
            int n = 0;
            double lw = log(1.0 / (double)max(kept, 1));
            for (int q = 0; q < ns; q++) {
                if (!(nv[q] > 0)) continue;
                double mu = (nr[q] * mu_r + nv[q] * (1.0 - mu_r)) / (nr[q] + nv[q]);
                ll[n] = log_binomial_likelihood(a, d, mu) + lw + lbc;
                mx = fmax(mx, ll[n]);
                n++;
            }
            if (kept == 0) { ll[0] = log_tiny_py; mx = ll[0]; n = 1; }
            double sm = 0;
            for (int q = 0; q < n; q++) sm += exp(ll[q] - mx);
            llh = log(sm) + mx;
        }
        total += llh;
    }
    return total;
}

// mode 0: pairs (rows[r], cols[c]) -> out[r*n_cols + c]     (block of the datum x node grid; cols NULL = all K nodes)
// mode 1: pairs (j, node_of_datum[j]) for all j -> out[j]    (terms of the complete-data likelihood)
__global__ void pair_llh_kernel(DataView dv, TreeView tv, int mode, int row0, int n_rows, const int *__restrict__ rows, int n_cols,
                                const int *__restrict__ cols, int col_major, int semantics, double log_tiny_mh, double log_tiny_py, double log_quarter,
                                double *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = mode == 0 ? (size_t)n_rows * n_cols : (size_t)n_rows;
    if (i >= total) return;
    int j, s;
    if (mode == 0) {
        // row-major: out[r][c], neighbouring threads = neighbouring nodes; column-major: out[c][r], neighbouring threads = data
        const size_t r = col_major ? i % n_rows : i / n_cols;
        j = rows ? rows[r] : row0 + (int)r;                      // rows == NULL: the range row0 .. row0 + n_rows - 1
        s = (int)(col_major ? i / n_rows : i % n_cols);
        if (cols) s = cols[s];                                   // cols == NULL: all K nodes
    }
    else { j = (int)i; s = tv.node_of_datum[j]; }
    out[i] = datum_ll_at(dv, tv, j, s, semantics, log_tiny_mh, log_tiny_py, log_quarter);
}

// The grid in the form an assignment sweep asks for -- a range of data x all nodes (or a list of them), column-major -- with
// warp-uniform work: one block row per node, and within it the plain data of the range first, then (from the next multiple of
// 32) its CNV-affected SSMs, whose path through datum_ll_at is ~6 times longer.  With the data in file order, where 30 % of the
// SSMs of a WGS tumour sit in CNVs at random positions, every warp of pair_llh_kernel waited for a handful of such lanes (ncu:
// 10.9 of 32 threads active per instruction).  Same function per pair, so the same bits.  plain / cnvs: ascending datum indices
// of the two kinds (fixed for the chain), already offset to the first one inside the range.
__global__ void pair_llh_sorted_kernel(DataView dv, TreeView tv, int row0, int n_rows, int np, int np_pad, const int *__restrict__ plain,
                                       int nc, const int *__restrict__ cnvs, const int *__restrict__ cols, int semantics, double log_tiny_mh,
                                       double log_tiny_py, double log_quarter, double *__restrict__ out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int j;
    if (r < np) j = plain[r];
    else if (r >= np_pad && r - np_pad < nc) j = cnvs[r - np_pad];
    else return;
    const int s = cols ? cols[blockIdx.y] : (int)blockIdx.y;
    out[(size_t)blockIdx.y * n_rows + (j - row0)] = datum_ll_at(dv, tv, j, s, semantics, log_tiny_mh, log_tiny_py, log_quarter);
}

// deterministic single-block sum (thread-strided partial sums, fixed-order combine)
__global__ void sum_kernel(int n, const double *__restrict__ x, double *out)
{
    __shared__ double red[32];
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (blockDim.x >> 5); w++) t += red[w];
        out[0] = t;
    }
}

// ============================================================================ K0d  sufficient statistics of the plain data
// With the assignments frozen during mh_loop, every plain datum in node k with error rates (mu_r, mu_v) of class c contributes
// a*log(mu_kc) + (d-a)*log(1-mu_kc) with the SAME mu_kc: the sum over those data is A_kc*log(mu_kc) + B_kc*log(1-mu_kc) with the
// integer totals A = sum a, B = sum (d-a) (exact in 64-bit integers).  Opt-in ("mh_collapse"): the general path streams
// every datum, because the input format allows arbitrary per-SSM error rates (README.md:30-41); real inputs have <= 3 classes.
__global__ void collapse_kernel(int Dp, int T, int C, const int *__restrict__ pa, const int *__restrict__ pd,
                                const int *__restrict__ pnode, const int *__restrict__ pcls, unsigned long long *__restrict__ sums)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Dp) return;
    const int k = pnode[i], c = pcls[i];
    for (int tp = 0; tp < T; tp++) {
        const int a = pa[(size_t)tp * Dp + i], d = pd[(size_t)tp * Dp + i];
        unsigned long long *s = sums + 2 * (((size_t)k * C + c) * T + tp);
        atomicAdd(s, (unsigned long long)a);                       // integer sums: exact and order-independent
        atomicAdd(s + 1, (unsigned long long)(d - a));
    }
}

// ============================================================================ K1
// consumer warps (likelihood + bookkeeping); with the producer warp that is 12 warps = 3 per SM sub-partition at 168 registers.
// Tried in round 2 and measured slower: a 13th warp (12 consumers + producer at 128 registers, the most a sub-partition with
// four warps allows: 1.52 vs 1.43 ms, a warp's three groups take longer when four warps share a dispatch port), and full
// batches cut into thirds of the data rounds so that 12 warps divide them evenly (1.69 ms: shares cost ~20 % more per datum).
#define MH_CWARPS 11
#define MH_MAXNREG 168
#define MH_UNITS 32                      // work units (proposal group x share of the data rounds) a batch is cut into: ~3 per consumer warp
#define MH_RED_DOUBLES(maxb) ((size_t)(maxb) > MH_BU * MH_UNITS ? (size_t)(maxb) : (size_t)MH_BU * MH_UNITS)
#define MH_CONSUMERS (MH_CWARPS * 32)
static_assert(PWGS_MAX_BATCH / 32 <= MH_CWARPS, "the publish / read-totals steps use one consumer thread per proposal");
#define MH_THREADS (MH_CONSUMERS + 32)   // + one producer warp (proposal draws for the ring)
#define MH_BU 4                          // proposals evaluated per thread in one pass over its data
#define MH_RING 4                        // slots of the global proposal ring
#define MH_WATCHDOG_CYCLES 8000000000ll  // ~4 s of SM clock: a barrier that has not completed by then never will
#define MH_SUM_STRIDE 16                 // 64-bit words between the running totals of two proposals: one L2 line each
#define MH_PCLASSES 6                    // proposal-group classes of the plain partition: a warp's 1st .. 6th group of a full batch
#define MH_NARROW_BATCH 128              // batch cap while accepts are recent; the full cap (PWGS_MAX_BATCH) opens up after
#define MH_WIDE_AFTER 8                  // this many all-reject batches in a row: an accept wastes the rest of its batch, so the wide
                                         // batch (half as many chain barriers per iteration) only pays at very low acceptance
#define MH_DIST_MIN_CTAS 8               // chains with fewer CTAs draw every proposal locally in every CTA
#ifndef MH_MIN_EPOCH_BATCH
#define MH_MIN_EPOCH_BATCH 8             // smallest first batch of an epoch that needs the chain barrier anyway
#endif
#ifndef MH_EPOCH_LOCAL
#define MH_EPOCH_LOCAL 1                 // epochs whose first two batches fit one round of local draws start without a chain barrier
#endif

struct MhParams {
    int Dp, M, T, K;
    const int *pa, *pd;                 // [T][Dp]  plain data (SSMs outside CNVs + CNV data), sample-major
    const double *pmu_r, *pmu_v;        // [Dp]
    const int *pnode;                   // [Dp] node index of each plain datum
    const int *pstart;                  // [MH_PCLASSES][cpc+1] for proposal groups of class i, CTA r walks plain data
                                        // pstart[i][r] .. pstart[i][r+1]-1: whole warp rounds, dealt by cost in thirds of a round
    int pchunk;                         // largest plain slice of any CTA (union over the classes; stride of the staged copy)
    int ncls;                           // classes the partition was made for: group g of a batch is of class g * ncls / (max_batch / MH_BU)
    CnvWork cw;                         // CNV-affected SSMs: sorted, CTA-major sparse work list (K0c)
    int n_classes;                      // > 0: plain data enter through their per-(node, class, sample) totals (K0d)
    const unsigned long long *cl_sums;  // [K][n_classes][T][2] = {sum a, sum d-a}
    const double *cl_mu_r, *cl_mu_v;    // [n_classes]
    const int *tin, *tout;              // [K] Euler-tour interval of every node (subtree test)
    const double *phi0, *pi0;           // [K*T] state on entry
    double *phi_out, *pi_out;           // [K*T] state on exit
    int *acc_out;                       // accepted proposals
    double *llh_out;                    // multi_param_post of the final state
    int iters, max_batch;               // max_batch: power of two in 1..PWGS_MAX_BATCH
    int stage;                          // 1: each CTA copies its slice of the data into shared memory once (normal case)
    int distributed;                    // 1: proposals two batches ahead are drawn once per chain and shared through `ring`
    double mh_std;
    unsigned long long seed, stream;
    unsigned long long *sums;           // [2][PWGS_MAX_BATCH][MH_SUM_STRIDE] running fixed-point totals of the CTAs' partial sums: words 0, 1
                                        // of each proposal's own 128-byte line (atomics to one line serialise in L2); zeroed before launch
    double *ring;                       // [MH_RING][bufstride] proposal ring (distributed mode)
    unsigned long long *counter;        // arrival counter of the chain's grid barrier (zeroed before launch)
    int *abort_flag;                    // watchdog: set when a spin wait exceeds MH_WATCHDOG_CYCLES; every CTA then leaves the loop
    double lbc_plain_sum, log_tiny;
    double *gscratch;                   // mh_kernel<*, true> only: [cpc][gstride] the CTAs' state copies, proposal buffers and draw scratch in global
    unsigned long long gstride;         // memory (trees whose node tables do not fit in shared memory), doubles per CTA
    unsigned long long *prof;           // optional [16 + cpc]: cycle counters of CTA 0 / thread 0, then every CTA's likelihood cycles (NULL = off)
};

// ---- shared-memory carve-up of one CTA (doubles first, then ints, then the staged data slice)
struct MhSmem {
    // two copies of the chain state (an accept writes the other copy, so the producer warp never reads a half-written one)
    double *state0;                                 // per copy: pi[KT], phi[KT], log(pi)[KT], lgamma(std*pi)[KT], lgsum[T], slg[T]
    int statestride, KT, T;
    __device__ __forceinline__ double *pi(int e) const { return state0 + e * statestride; }
    __device__ __forceinline__ double *phi(int e) const { return state0 + e * statestride + KT; }
    __device__ __forceinline__ double *lp(int e) const { return state0 + e * statestride + 2 * KT; }
    __device__ __forceinline__ double *lg(int e) const { return state0 + e * statestride + 3 * KT; }
    __device__ __forceinline__ double *lgsum(int e) const { return state0 + e * statestride + 4 * KT; }
    __device__ __forceinline__ double *slg(int e) const { return state0 + e * statestride + 4 * KT + T; }
    // two proposal buffers (current batch / next batch), `bufstride` doubles apart; the ring slots use the same layout
    double *buf0;                                   // per buffer: pi1, phi1 (groups of four proposals interleaved), hastings, log r: see mh_bufstride
    int bufstride, o_phi, o_hast, o_logr;
    __device__ __forceinline__ double *prop_pi(int b) const { return buf0 + b * bufstride; }
    __device__ __forceinline__ double *prop_phi(int b) const { return buf0 + b * bufstride + o_phi; }
    __device__ __forceinline__ double *hast(int b) const { return buf0 + b * bufstride + o_hast; }
    __device__ __forceinline__ double *logr(int b) const { return buf0 + b * bufstride + o_logr; }
    double *red;                                    // [nb][shares per proposal group]: per-share sums of the current batch
    double *llh;                                    // [MAXB]
    unsigned long long *seen;                       // [2][MAXB][2]: p.sums as of the last time this CTA read each parity
    double *pscratch;                               // [MH_CWARPS+1][4K] per warp: pi1 and phi1 of a proposal drawn for the ring, Euler-tour scan
    volatile int *ctl;                              // producer mailbox, see MhCtl
    int *tin, *tout;                                // [K]
};
enum MhCtl { CTL_GEN = 0, CTL_DONE = 1, CTL_PROD_DONE = 2, CTL_ABORT = 3, CTL_JOB = 4 /* 2 x {state, it, nb} */, CTL_PRANGE = 12 /* MH_PCLASSES x {first, count} */,
             CTL_OK = 12 + 2 * MH_PCLASSES /* PWGS_MAX_BATCH / 32 ballots of the accept test */, CTL_WORDS = 12 + 2 * MH_PCLASSES + PWGS_MAX_BATCH / 32 };

// Proposal buffer (shared memory) = ring slot (global memory) layout, in doubles:
//   pi1   [maxb][KT]                       proposal b, node/sample i at b*KT + i
//   phi1  [ceil(maxb/4)][KT][4]            proposal b, node/sample i at mh_phi_at(b, i, KT): the four proposals a likelihood thread
//                                          evaluates together are adjacent, so it reads them with two 16-byte loads
//   hastings [maxb][T][3]: {Hastings correction, sum_k lgamma(mh_std*pi1_k), lgamma(sum_k mh_std*pi1_k)} -- the two lgamma terms
//                          are what the chain state needs if this proposal is accepted (no lgamma pass on the accept path);  log r [maxb]
// Region offsets are even so that the phi region is 16-byte aligned.
__host__ __device__ inline size_t mh_even(size_t x) { return (x + 1) & ~(size_t)1; }
__host__ __device__ inline size_t mh_phi_doubles(int KT, int nb) { return (size_t)((nb + 3) & ~3) * KT; }
__host__ __device__ inline size_t mh_o_phi(int KT, int maxb) { return mh_even((size_t)maxb * KT); }
__host__ __device__ inline size_t mh_o_hast(int KT, int maxb) { return mh_o_phi(KT, maxb) + mh_phi_doubles(KT, maxb); }
__host__ __device__ inline size_t mh_bufstride(int K, int T, int maxb) { return mh_even(mh_o_hast(K * T, maxb) + 3 * (size_t)maxb * T + maxb); }
__host__ __device__ __forceinline__ int mh_phi_at(int b, int i, int KT) { return (b >> 2) * 4 * KT + i * 4 + (b & 3); }
// the part that grows with the tree: two state copies, two proposal buffers, the draw scratch of every warp.  Normally in shared
// memory; mh_kernel<false, true> keeps it in global memory (p.gscratch), one region per CTA
__host__ __device__ inline size_t mh_tree_doubles(int K, int T, int maxb)
{
    size_t KT = (size_t)K * T;
    return 2 * (4 * KT + 2 * T) + 2 * mh_bufstride(K, T, maxb) + 4 * (size_t)K * (MH_CWARPS + 1);
}
// the rest: tables of the logarithm / exponential, per-batch sums and totals
__host__ __device__ inline size_t mh_fixed_doubles(int maxb)
{
    return PWGS_LOG_TAB_DOUBLES + PWGS_EXP_TAB_DOUBLES + MH_RED_DOUBLES(maxb) + maxb + 4 * (size_t)maxb;
}
__host__ __device__ inline size_t mh_smem_doubles(int K, int T, int maxb) { return mh_fixed_doubles(maxb) + mh_tree_doubles(K, T, maxb); }
__host__ __device__ inline size_t mh_smem_ints(int K) { return CTL_WORDS + 2 * (size_t)K; }

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MH_CONSUMERS) : "memory"); }
__device__ __forceinline__ int pow2_floor(int x) { return x <= 0 ? 0 : 1 << (31 - __clz(x)); }
// size of the batch that follows a batch of n rejected proposals, with `left` iterations still to do
__device__ __forceinline__ int next_batch(int n, int left, int maxb) { return pow2_floor(min(min(maxb, 2 * n), left)); }
// cap of the batch that runs after jn all-reject batches in a row
__device__ __forceinline__ int batch_cap(int jn, int maxb) { return jn >= MH_WIDE_AFTER ? maxb : min(maxb, MH_NARROW_BATCH); }

// single out-of-line copy of the (large) lgamma routine
__device__ __noinline__ double pwgs_lgamma(double x) { return lgamma(x); }

// One warp draws proposal `it` for sample tp from chain state copy e: pi1 ~ Dirichlet(mh_std*pi + 1) -> +1e-4 -> renormalise
// (mh.cpp:113-131, util.cpp:24-33), phi1 = subtree sums of pi1 (mh.cpp:134-141) and the Hastings correction of mh.cpp:80-93
// for this sample; for tp == 0 also the log of the iteration's acceptance uniform (mh.cpp:95-96).  Lanes own the nodes
// k = lane, lane+32, ...; every reduction is a fixed-shape shuffle tree, so all CTAs of a chain get bit-identical results.
// x / ph: pi1 / phi1 of node k at x[k * xs] / ph[k * phs].
// (Trees of more than 32 nodes take their subtree sums from a prefix scan over the Euler
// tour instead of one pass over all nodes per node: at the ~110 nodes of a burn-in tree the K^2 loop, over a scratch array that
// lives in global memory in that mode, was a third of a draw, and the draws pace the chain there.)
__device__ __noinline__ void mh_propose(const MhParams &p, const MhSmem &S, int e, double *x, double *ph, int xs, int phs, double *hast_out,
                                        double *logr_out, int tp, uint32_t it, int lane)
{
    double *esc = S.pscratch + (size_t)(threadIdx.x >> 5) * 4 * p.K + 2 * p.K;   // this warp's scan scratch (third and fourth K of its 4K)
    const int T = p.T, K = p.K;
    const double mstd = p.mh_std;
    const double *cpi = S.pi(e), *clp = S.lp(e);
    double s = 0.0;
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
        double g = philox_gamma(mstd * cpi[k * T + tp] + 1.0, p.seed, p.stream, it, (uint32_t)k, (uint32_t)tp);
        x[k * xs] = g;
        s += g;
    }
    const double norm = warp_sum(s);                                             // gsl_ran_dirichlet normalisation
    s = 0.0;
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
        double v = x[k * xs] / norm + 0.0001;                                    // util.cpp:26-27
        x[k * xs] = v;
        s += v;
    }
    const double sum = warp_sum(s);                                              // util.cpp:28-30
    double sa = 0.0;
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
        double v = x[k * xs] / sum;                                              // util.cpp:31-32
        x[k * xs] = v;
        sa += mstd * v;
    }
    const double sa1 = warp_sum(sa);
    __syncwarp();
    const bool wide = K > 32;
    if (wide) {
        // E[tin(k)] = pi1_k, E[tout(k)] = 0; after the inclusive scan phi_k = E[tout(k)] - E[tin(k)] + pi1_k.  Fixed-shape shuffle
        // scan with a carry between the chunks of 32: the same bits in every CTA; |error| ~ 1e-15 absolute (the sums are <= 1)
#pragma unroll 1
        for (int k = lane; k < K; k += 32) { esc[S.tin[k]] = x[k * xs]; esc[S.tout[k]] = 0.0; }
        __syncwarp();
        double carry = 0.0;
#pragma unroll 1
        for (int base = 0; base < 2 * K; base += 32) {
            const int t = base + lane;
            double v = t < 2 * K ? esc[t] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(PWGS_FULL_MASK, v, o);
                if (lane >= o) v += u;
            }
            v += carry;
            if (t < 2 * K) esc[t] = v;
            carry = __shfl_sync(PWGS_FULL_MASK, v, 31);
        }
        __syncwarp();
    }
    double t1 = 0.0, t2 = 0.0, slg = 0.0, lgs = 0.0;
#pragma unroll 1
    for (int k = lane; k <= K; k += 32) {                                        // virtual node K: lgamma(sum alpha) rides along
        const bool real = k < K;
        const int kk = real ? k : 0;
        const int lo = S.tin[kk], hi = S.tout[kk];
        double acc = 0.0;
        if (wide) acc = esc[hi] - esc[lo] + x[kk * xs];
        else {
#pragma unroll 1
            for (int j = 0; j < K; j++) {                                        // phi_k = sum of pi over the subtree of k
                double v = x[j * xs];
                if (lo <= S.tin[j] && S.tout[j] <= hi) acc += v;
            }
        }
        if (real) ph[k * phs] = acc;
        const double pn = real ? x[kk * xs] : 1.0, pv = cpi[kk * T + tp];
        const double lg = pwgs_lgamma(real ? mstd * pn : sa1), lpn = log(pn);
        if (real) {
            t1 += (mstd * pn - 1.0) * clp[k * T + tp];                           // lnDir(pi ; mh_std*pi1), mh.cpp:86-88
            slg += lg;
            t2 += (mstd * pv - 1.0) * lpn;                                       // lnDir(pi1 ; mh_std*pi), mh.cpp:90-92
        } else lgs = lg;
    }
    t1 = warp_sum(t1); t2 = warp_sum(t2); slg = warp_sum(slg); lgs = warp_sum(lgs);
    if (lane == 0) {
        double lp1 = t1 + lgs - slg;
        double lp2 = t2 + S.lgsum(e)[tp] - S.slg(e)[tp];
        hast_out[0] = lp1 - lp2;
        hast_out[1] = slg;                                                       // the state's slg / lgsum once this proposal is
        hast_out[2] = lgs;                                                       // accepted (same bits as mh_refresh_state computes)
        if (tp == 0) {
            double u0, u1;
            stream_draw(p.seed, p.stream, it, 0, 0, PWGS_PURPOSE_ACCEPT, 0, u0, u1);
            *logr_out = log(u0);
        }
    }
}

// the consumer warps draw a whole batch (nb proposals x T samples) into proposal buffer `buf`
__device__ __forceinline__ void mh_draw_local(const MhParams &p, const MhSmem &S, int e, int buf, int nb, int it_first, int warp, int lane)
{
    const int T = p.T, KT = p.K * T;
#pragma unroll 1
    for (int task = warp; task < nb * T; task += MH_CWARPS) {
        const int b = task / T, tp = task - b * T;
        mh_propose(p, S, e, S.prop_pi(buf) + b * KT + tp, S.prop_phi(buf) + mh_phi_at(b, tp, KT), T, 4 * T, S.hast(buf) + 3 * (b * T + tp),
                   S.logr(buf) + b, tp, (uint32_t)(it_first + b), lane);
    }
}

// Start of an epoch whose first two batches are small enough for ONE round of draws by this CTA's own warps ((n0 + n1) * T <=
// MH_CWARPS): batch n0 goes into buffer `buf`, batch n1 into the other one, where the main loop expects the batch after.  Every
// CTA draws the same proposals (they are functions of the state and the iteration alone), so no chain barrier and no copy out of
// the epoch slots is needed: at acceptance rates of 10 % and more, where an accept comes every few iterations, this is most
// epochs and saves ~4k of the ~26k cycles an accept costs.  Out of line: cold code.
__device__ __noinline__ void mh_epoch_local(const MhParams &p, const MhSmem &S, int e, int buf, int n0, int it0, int n1, int warp, int lane)
{
    const int T = p.T, KT = p.K * T;
#pragma unroll 1
    for (int task = warp; task < (n0 + n1) * T; task += MH_CWARPS) {
        const int bsel = task / T, tp = task - bsel * T;
        const int bb = bsel < n0 ? buf : buf ^ 1, b = bsel < n0 ? bsel : bsel - n0;
        mh_propose(p, S, e, S.prop_pi(bb) + b * KT + tp, S.prop_phi(bb) + mh_phi_at(b, tp, KT), T, 4 * T, S.hast(bb) + 3 * (b * T + tp),
                   S.logr(bb) + b, tp, (uint32_t)(it0 + bsel), lane);
    }
}

// log(pi), lgamma(mh_std*pi) and their per-sample sums for state copy e (consumer warps; caller syncs before and after)
__device__ __noinline__ void mh_refresh_state(const MhParams &p, const MhSmem &S, int e, int tid, int warp, int lane)
{
    const int T = p.T, K = p.K, KT = K * T;
#pragma unroll 1
    for (int i = tid; i < KT; i += MH_CONSUMERS) {
        S.lp(e)[i] = log(S.pi(e)[i]);
        S.lg(e)[i] = pwgs_lgamma(p.mh_std * S.pi(e)[i]);
    }
    consumer_sync();
#pragma unroll 1
    for (int tp = warp; tp < T; tp += MH_CWARPS) {
        double sa = 0.0, sl = 0.0;
#pragma unroll 1
        for (int k = lane; k < K; k += 32) { sa += p.mh_std * S.pi(e)[k * T + tp]; sl += S.lg(e)[k * T + tp]; }
        sa = warp_sum(sa); sl = warp_sum(sl);
        if (lane == 0) { S.lgsum(e)[tp] = pwgs_lgamma(sa); S.slg(e)[tp] = sl; }
    }
}

// State copy e becomes proposal `sel` of buffer `buf` (its pi / phi are already copied): log(pi), and the two lgamma terms that
// came with the proposal (mh_propose) -- bit for bit what mh_refresh_state would compute, without its lgamma pass and the second
// one for the sums (~5k cycles of the ~26k an accept costs).  Out of line: cold code, kept out of the likelihood loops' registers.
__device__ __noinline__ void mh_adopt_state(const MhSmem &S, int e, int buf, int sel, int tid)
{
    const int T = S.T, KT = S.KT;
#pragma unroll 1
    for (int i = tid; i < KT; i += MH_CONSUMERS) S.lp(e)[i] = log(S.pi(e)[i]);
    if (tid < T) {
        S.slg(e)[tid] = S.hast(buf)[3 * (sel * T + tid) + 1];
        S.lgsum(e)[tid] = S.hast(buf)[3 * (sel * T + tid) + 2];
    }
}

// a / b for b > 0 without the slow-path branch of the IEEE division (|error| <= 1 ulp)
__device__ __forceinline__ double pwgs_div_pos(double a, double b)
{
    double y = pwgs_rcp_approx(b);
    double t = fma(-b, y, 1.0);
    y = fma(y, t, y);
    t = fma(-b, y, 1.0);
    y = fma(y, t, y);
    double q = a * y;
    return fma(fma(-b, q, a), y, q);
}

// ---- the CTA's slice of the two work lists as the likelihood loop reads it: element l of the slice lives at [tp * stride + l].
// Normally the slice is staged in shared memory once per launch (SliceShared: offsets into the dynamic shared array, so the
// loads are LDS); slices too large for shared memory are read in place from global memory (SliceGlobal).
extern __shared__ __align__(16) double mh_sm[];
struct SliceShared {
    int np, pstride, nc, cstride;
    int o_pint, o_pdbl, o_cint, o_cdbl, o_ecoef;       // offsets in elements of int / double / int / double / int4
    int T;
    __device__ __forceinline__ double pmu_r(int i) const { return mh_sm[o_pdbl + i]; }
    __device__ __forceinline__ double pmu_v(int i) const { return mh_sm[o_pdbl + pstride + i]; }
    __device__ __forceinline__ int pnode(int i) const { return ((const int *)mh_sm)[o_pint + i]; }
    __device__ __forceinline__ int pa(int tp, int i) const { return ((const int *)mh_sm)[o_pint + (1 + tp) * pstride + i]; }
    __device__ __forceinline__ int pd(int tp, int i) const { return ((const int *)mh_sm)[o_pint + (1 + T + tp) * pstride + i]; }
    __device__ __forceinline__ double cmu_r(int i) const { return mh_sm[o_cdbl + i]; }
    __device__ __forceinline__ double clbc(int tp, int i) const { return mh_sm[o_cdbl + (1 + tp) * cstride + i]; }
    __device__ __forceinline__ int ccode(int i) const { return ((const int *)mh_sm)[o_cint + i]; }
    __device__ __forceinline__ int ca(int tp, int i) const { return ((const int *)mh_sm)[o_cint + (1 + tp) * cstride + i]; }
    __device__ __forceinline__ int cd(int tp, int i) const { return ((const int *)mh_sm)[o_cint + (1 + T + tp) * cstride + i]; }
    __device__ __forceinline__ int cenode(int e, int i) const { return ((const int *)mh_sm)[o_cint + (1 + 2 * T + e) * cstride + i]; }
    __device__ __forceinline__ int cecoef(int e, int i, int u) const { return ((const int *)mh_sm)[4 * (o_ecoef + e * cstride + i) + u]; }
    __device__ __forceinline__ int4 cecoef4(int e, int i) const { return ((const int4 *)mh_sm)[o_ecoef + e * cstride + i]; }
};
struct SliceGlobal {
    int np, pstride, nc, cstride;
    const int *g_pa, *g_pd, *g_pnode, *g_ca, *g_cd, *g_ccode, *g_cenode;
    const double *g_pmu_r, *g_pmu_v, *g_clbc, *g_cmu_r;
    const int4 *g_cecoef;
    __device__ __forceinline__ double pmu_r(int i) const { return __ldg(g_pmu_r + i); }
    __device__ __forceinline__ double pmu_v(int i) const { return __ldg(g_pmu_v + i); }
    __device__ __forceinline__ int pnode(int i) const { return __ldg(g_pnode + i); }
    __device__ __forceinline__ int pa(int tp, int i) const { return __ldg(g_pa + (size_t)tp * pstride + i); }
    __device__ __forceinline__ int pd(int tp, int i) const { return __ldg(g_pd + (size_t)tp * pstride + i); }
    __device__ __forceinline__ double cmu_r(int i) const { return __ldg(g_cmu_r + i); }
    __device__ __forceinline__ double clbc(int tp, int i) const { return __ldg(g_clbc + (size_t)tp * cstride + i); }
    __device__ __forceinline__ int ccode(int i) const { return __ldg(g_ccode + i); }
    __device__ __forceinline__ int ca(int tp, int i) const { return __ldg(g_ca + (size_t)tp * cstride + i); }
    __device__ __forceinline__ int cd(int tp, int i) const { return __ldg(g_cd + (size_t)tp * cstride + i); }
    __device__ __forceinline__ int cenode(int e, int i) const { return __ldg(g_cenode + (size_t)e * cstride + i); }
    __device__ __forceinline__ int cecoef(int e, int i, int u) const { return __ldg((const int *)(g_cecoef + (size_t)e * cstride + i) + u); }
    __device__ __forceinline__ int4 cecoef4(int e, int i) const { return __ldg(g_cecoef + (size_t)e * cstride + i); }
};

__host__ __device__ inline size_t mh_align16(size_t x) { return (x + 15) & ~(size_t)15; }
// bytes of shared memory for a staged slice of pchunk plain data and cchunk CNV-affected SSMs
__host__ __device__ inline size_t mh_stage_bytes(int pchunk, int cchunk, int T, int Emax)
{
    size_t b = 0;
    b += mh_align16((size_t)cchunk * Emax * sizeof(int4));
    b += mh_align16((size_t)pchunk * 2 * sizeof(double)) + mh_align16((size_t)cchunk * (1 + (size_t)T) * sizeof(double));
    b += mh_align16((size_t)pchunk * (1 + 2 * (size_t)T) * sizeof(int)) + mh_align16((size_t)cchunk * (1 + 2 * (size_t)T + Emax) * sizeof(int));
    return b;
}

// phi1 of the BU proposals of a thread's group at node/sample index n: they are adjacent in the proposal buffer (mh_phi_at)
template <int BU>
__device__ __forceinline__ void mh_load_phi(const double *bphi, int n, double (&ph)[BU])
{
    if constexpr (BU == 4) {
        const double2 lo = *(const double2 *)(bphi + 4 * n), hi = *(const double2 *)(bphi + 4 * n + 2);
        ph[0] = lo.x; ph[1] = lo.y; ph[2] = hi.x; ph[3] = hi.y;
    } else {
#pragma unroll
        for (int bb = 0; bb < BU; bb++) ph[bb] = bphi[4 * n + bb];
    }
}

// One merged genome state u of a CNV-affected SSM for the BU proposals (mh.hpp:95-101): nr, nv = sum over the sparse entries of
// delta * phi1, mixture mean, the two logarithms.  nd[e]: node/sample index of entry e (0 beyond the SSM's entries), cu[e]: its
// packed (dnr, dnv) coefficients for this state (0 beyond).  A state whose nr + nv is not positive counts log(1e-99) (mh.hpp:100):
// the arithmetic runs unconditionally (garbage in, garbage out, nothing traps) and the warp patches those values afterwards,
// which keeps the selects out of the path every other SSM takes.
template <int BU, class Slice>
__device__ __forceinline__ void mh_cnv_state(const Slice &sl, const double *bphi, const double2 *ltab, int idx, int ne, int T, int tp,
                                             const int (&nd)[4], const int (&cu)[4], int u, double mu_r, double omr, double fa, double fb,
                                             double lw, double tiny, double (&val)[BU])
{
    double nr[BU], nv[BU];
#pragma unroll
    for (int bb = 0; bb < BU; bb++) nr[bb] = nv[bb] = 0.0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const double dr = (double)(short)(cu[e] & 0xffff), dv = (double)(cu[e] >> 16);
        double ph[BU];
        mh_load_phi<BU>(bphi, nd[e], ph);
#pragma unroll
        for (int bb = 0; bb < BU; bb++) { nr[bb] = fma(ph[bb], dr, nr[bb]); nv[bb] = fma(ph[bb], dv, nv[bb]); }
    }
#pragma unroll 1
    for (int e = 4; e < ne; e++) {                                              // SSMs in three or more CNVs
        const int c = sl.cecoef(e, idx, u);
        const double dr = (double)(short)(c & 0xffff), dv = (double)(c >> 16);
        double ph[BU];
        mh_load_phi<BU>(bphi, sl.cenode(e, idx) * T + tp, ph);
#pragma unroll
        for (int bb = 0; bb < BU; bb++) { nr[bb] = fma(ph[bb], dr, nr[bb]); nv[bb] = fma(ph[bb], dv, nv[bb]); }
    }
    double num[BU], den[BU], mu[BU];
    bool bad = false;
#pragma unroll
    for (int bb = 0; bb < BU; bb++) {
        den[bb] = nr[bb] + nv[bb];
        num[bb] = fma(nv[bb], omr, nr[bb] * mu_r);
        bad |= !(den[bb] > 0.0);
    }
    pwgs_div_pos_n<BU>(num, den, mu);                                            // mh.hpp:96
    double xx[2 * BU], ll[2 * BU];
#pragma unroll
    for (int bb = 0; bb < BU; bb++) { xx[bb] = mu[bb]; xx[BU + bb] = 1.0 - mu[bb]; }
    pwgs_log_pos_n<2 * BU>(xx, ll, ltab);
#pragma unroll
    for (int bb = 0; bb < BU; bb++) val[bb] = fma(fa, ll[bb], fma(fb, ll[BU + bb], lw));
    if (__any_sync(PWGS_FULL_MASK, bad)) {
#pragma unroll
        for (int bb = 0; bb < BU; bb++) val[bb] = den[bb] > 0.0 ? val[bb] : tiny;
    }
}

// Likelihood of this CTA's slice of the data for the nb proposals of the batch (param_post, mh.cpp:153-166).  The consumer
// warps are split into nb/BU groups; group g evaluates proposals g*BU .. g*BU+BU-1, each thread walking the CTA's data with
// the group's stride and accumulating BU sums, so that every load / int->fp64 conversion is shared by BU proposals and the BU
// independent chains of logarithms overlap in the FP64 pipe.  Leaves red[bb][warp] = the warp's sum for proposal g*BU+bb.
// What bounds these loops is the warp scheduler's dispatch port: an FP64 instruction holds it for two cycles, any other
// instruction for one (scripts/micro/plain_round.cu: time per round = 2 F + O at every occupancy), so beside the FP64 count the
// number of address / select / move instructions around it decides the speed -- hence the 16-byte phi loads, the states peeled
// out of their loop (no selects on the first one, no register shuffling between them) and the warp-level patch for nr + nv <= 0.
template <int BU, class Slice>
__device__ __forceinline__ void mh_likelihood(const MhParams &p, const MhSmem &S, const Slice &sl, int cur, int wpg, int role, int g, int pcls,
                                              int lane, int rank, int cpc)
{
    // proposal group g (proposals g*BU .. g*BU+BU-1) is evaluated by wpg warps; this warp is number `role` of them
    const int T = p.T, KT = p.K * T;
    const int tpg = wpg * 32, lt = role * 32 + lane;
    const double2 *ltab = (const double2 *)mh_sm + (lane & (PWGS_LOG_TAB_COPIES - 1));   // tables of the logarithm (this lane's copy) and the
    const double *etab = mh_sm + PWGS_LOG_TAB_DOUBLES;                  // exponential: first thing in the dynamic shared array
    const double *bphi = S.prop_phi(cur) + mh_phi_at(g * BU, 0, KT);    // BU == 1: proposal g of its group of four, stride 4
    double acc[BU];
#pragma unroll
    for (int bb = 0; bb < BU; bb++) acc[bb] = 0.0;

    // ---- CNV-affected SSMs (mh.hpp:90-161) through the sparse, state-merged work list
    const int cnt = sl.nc;
#pragma unroll 1
    for (int l = lt; l - lane < cnt; l += tpg) {
        const bool act = l < cnt;
        const int idx = act ? l : 0;
        const int code = act ? sl.ccode(idx) : 0;                                // lanes past the end: no entries, no states (the
        const int ne = code & 0xff, nu = (code >> 8) & 3;                        // round costs what its real SSMs cost)
        const int numax = __reduce_max_sync(PWGS_FULL_MASK, nu);
        const double actd = act ? 1.0 : 0.0;                                     // lanes past the end of the list add nothing
        const double mu_r = sl.cmu_r(idx), omr = 1.0 - mu_r;
        int nd0[4];
        int4 cq[4];                                                              // entry e: packed coefficients of its (up to) 4 states
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const bool v = e < ne;
            nd0[e] = v ? sl.cenode(e, idx) * T : 0;
            cq[e] = v ? sl.cecoef4(e, idx) : make_int4(0, 0, 0, 0);
        }
#pragma unroll 1
        for (int tp = 0; tp < T; tp++) {
            const int a = sl.ca(tp, idx), d = sl.cd(tp, idx);
            const double lbc = sl.clbc(tp, idx), fa = (double)a, fb = (double)(d - a);
            int nd[4];
#pragma unroll
            for (int e = 0; e < 4; e++) nd[e] = nd0[e] + tp;
            // states one after the other: util.cpp:35-43 over the merged states, weights folded into the terms
            double v0[BU];
            {
                const int w = (code >> 12) & 7;
                const int cu[4] = {cq[0].x, cq[1].x, cq[2].x, cq[3].x};
                mh_cnv_state<BU>(sl, bphi, ltab, idx, ne, T, tp, nd, cu, 0, mu_r, omr, fa, fb, lbc + c_logw4[w], p.log_tiny + c_logw[w], v0);
            }
            if (numax <= 1) {
#pragma unroll
                for (int bb = 0; bb < BU; bb++) acc[bb] = fma(actd, v0[bb], acc[bb]);
                continue;
            }
            double mx[BU], sm[BU];
            {
                double v1[BU], xe[BU], ee[BU];
                const int w = (code >> 16) & 7;
                const int cu[4] = {cq[0].y, cq[1].y, cq[2].y, cq[3].y};
                mh_cnv_state<BU>(sl, bphi, ltab, idx, ne, T, tp, nd, cu, 1, mu_r, omr, fa, fb, lbc + c_logw4[w], p.log_tiny + c_logw[w], v1);
                // logsumexp of two: max + log(1 + exp(-|difference|)); an SSM with a single state contributes exp(-inf) = 0
#pragma unroll
                for (int bb = 0; bb < BU; bb++) {
                    v1[bb] = 1 < nu ? v1[bb] : -INFINITY;
                    mx[bb] = fmax(v0[bb], v1[bb]);
                    xe[bb] = -fabs(v0[bb] - v1[bb]);
                }
                pwgs_exp_nonpos_n<BU>(xe, ee, etab);
#pragma unroll
                for (int bb = 0; bb < BU; bb++) sm[bb] = 1.0 + ee[bb];
            }
            if (numax > 2) {
                double v2[BU], xe[BU], ee[BU];
                const int w = (code >> 20) & 7;
                const int cu[4] = {cq[0].z, cq[1].z, cq[2].z, cq[3].z};
                mh_cnv_state<BU>(sl, bphi, ltab, idx, ne, T, tp, nd, cu, 2, mu_r, omr, fa, fb, lbc + c_logw4[w], p.log_tiny + c_logw[w], v2);
                // running (max, sum): of exp(mx - nm) and exp(v2 - nm) one is exp(0) = 1, so one exponential does
#pragma unroll
                for (int bb = 0; bb < BU; bb++) { v2[bb] = 2 < nu ? v2[bb] : -INFINITY; xe[bb] = -fabs(mx[bb] - v2[bb]); }
                pwgs_exp_nonpos_n<BU>(xe, ee, etab);
#pragma unroll
                for (int bb = 0; bb < BU; bb++) {
                    const bool up = v2[bb] > mx[bb];
                    sm[bb] = fma(sm[bb], up ? ee[bb] : 1.0, up ? 1.0 : ee[bb]);
                    mx[bb] = up ? v2[bb] : mx[bb];
                }
            }
            double ls[BU];
            pwgs_log_pos_n<BU>(sm, ls, ltab);
#pragma unroll
            for (int bb = 0; bb < BU; bb++) acc[bb] = fma(actd, mx[bb] + ls[bb], acc[bb]);
        }
    }

    // ---- plain data (mh.hpp:86-89): SSMs outside CNVs and the CNV data themselves; the slice depends on the group's class
    const int pbeg = S.ctl[CTL_PRANGE + 2 * pcls], pend = pbeg + S.ctl[CTL_PRANGE + 2 * pcls + 1];
#pragma unroll 1
    for (int j = pbeg + lt; j - lane < pend; j += tpg) {
        const bool act = j < pend;
        const int jj = act ? j : 0;
        const double mu_r = sl.pmu_r(jj), dmu = sl.pmu_v(jj) - mu_r;
        const int nodeT = sl.pnode(jj) * T;
#pragma unroll 1
        for (int tp = 0; tp < T; tp++) {
            const int a = sl.pa(tp, jj), d = sl.pd(tp, jj);
            const double fa = act ? (double)a : 0.0, fb = act ? (double)(d - a) : 0.0;
            double xx[2 * BU], ll[2 * BU], ph[BU];
            mh_load_phi<BU>(bphi, nodeT + tp, ph);
#pragma unroll
            for (int bb = 0; bb < BU; bb++) {
                xx[bb] = fma(ph[bb], dmu, mu_r);                    // (1-phi)*mu_r + phi*mu_v (mh.hpp:87) in one rounding
                xx[BU + bb] = 1.0 - xx[bb];
            }
            pwgs_log_pos_n<2 * BU>(xx, ll, ltab);
#pragma unroll
            for (int bb = 0; bb < BU; bb++) acc[bb] = fma(fa, ll[bb], fma(fb, ll[BU + bb], acc[bb]));
        }
    }
    // ---- plain data through their sufficient statistics (K0d): entry (node, class, sample) e belongs to CTA e mod cpc
    if (p.n_classes > 0) {
        const int C = p.n_classes, n_ent = p.K * C * T;
#pragma unroll 1
        for (int e = rank + cpc * lt; e - cpc * lane < n_ent; e += cpc * tpg) {
            const bool act = e < n_ent;
            const int ee = act ? e : 0, k = ee / (C * T), c = (ee / T) % C, tp = ee % T;
            const double A = act ? (double)p.cl_sums[2 * ee] : 0.0, B = act ? (double)p.cl_sums[2 * ee + 1] : 0.0;
            const double mu_r = p.cl_mu_r[c], dmu = p.cl_mu_v[c] - mu_r;
            double xx[2 * BU], ll[2 * BU], ph[BU];
            mh_load_phi<BU>(bphi, k * T + tp, ph);
#pragma unroll
            for (int bb = 0; bb < BU; bb++) {
                xx[bb] = fma(ph[bb], dmu, mu_r);
                xx[BU + bb] = 1.0 - xx[bb];
            }
            pwgs_log_pos_n<2 * BU>(xx, ll, ltab);
#pragma unroll
            for (int bb = 0; bb < BU; bb++) acc[bb] = fma(A, ll[bb], fma(B, ll[BU + bb], acc[bb]));
        }
    }
#pragma unroll
    for (int bb = 0; bb < BU; bb++) {
        const double t = warp_sum(acc[bb]);
        if (lane == 0) S.red[(g * BU + bb) * wpg + role] = t;
    }
}

// one warp draws (proposal b, sample tp) of a batch into a slot of the global ring, through its own shared-memory scratch
__device__ __forceinline__ void mh_draw_to_slot(const MhParams &p, const MhSmem &S, int e, double *slot, int b, int tp, uint32_t it,
                                                int scratch, int lane)
{
    const int T = p.T, K = p.K, KT = K * T;
    double *x = S.pscratch + (size_t)scratch * 4 * K;
    double hast[3] = {0.0, 0.0, 0.0}, logr = 0.0;
    mh_propose(p, S, e, x, x + K, 1, 1, hast, &logr, tp, it, lane);
    __syncwarp();
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
        slot[(size_t)b * KT + k * T + tp] = x[k];
        slot[(size_t)S.o_phi + mh_phi_at(b, k * T + tp, KT)] = x[K + k];
    }
    if (lane == 0) {
        slot[(size_t)S.o_hast + 3 * (b * T + tp)] = hast[0];
        slot[(size_t)S.o_hast + 3 * (b * T + tp) + 1] = hast[1];
        slot[(size_t)S.o_hast + 3 * (b * T + tp) + 2] = hast[2];
        if (tp == 0) slot[(size_t)S.o_logr + b] = logr;
    }
    __syncwarp();
}

// ---- ring jobs.  Job g = "draw this CTA's share of batch g+2 from state copy e into ring slot (g+2) % MH_RING"; a share is the
// (proposal, sample) pairs t with (t + 37 g) % cpc == rank, enumerated directly (first + i * cpc): walking all nb*T pairs with a
// modulo each cost the producer warp more than the draws themselves once T > 1 and made it the pace-setter of the chain.
// Jobs are posted in the mailbox S.ctl[CTL_JOB..] by thread 0 at the top of batch g and must be complete before the CTA arrives
// at the barrier of batch g+1.
__device__ __forceinline__ int mh_job_first(int job, int rank, int cpc) { return (rank - (job * 37) % cpc + cpc) % cpc; }
__device__ __noinline__ void mh_ring_tasks(const MhParams &p, const MhSmem &S, int job, int rank, int cpc, int scratch, int lane)
{
    const int T = p.T;
    const volatile int *jb = S.ctl + CTL_JOB + 3 * (job & 1);
    const int e = jb[0], it_first = jb[1], nb = jb[2];
    double *slot = p.ring + (size_t)((job + 2) % MH_RING) * mh_bufstride(p.K, T, p.max_batch);
#pragma unroll 1
    for (int task = mh_job_first(job, rank, cpc); task < nb * T; task += cpc) {
        const int b = task / T, tp = task - b * T;
        mh_draw_to_slot(p, S, e, slot, b, tp, (uint32_t)(it_first + b), scratch, lane);
    }
    __threadfence();                                                             // ring stores ordered before the "done" flag
    __syncwarp();
    if (lane == 0) S.ctl[CTL_PROD_DONE] = job;
}

__device__ __noinline__ void mh_producer(const MhParams &p, const MhSmem &S, int rank, int cpc, int lane)
{
    int job = 1;                                                                // jobs are numbered by global batch, from 1
#pragma unroll 1
    for (;;) {
        while (S.ctl[CTL_GEN] < job && !S.ctl[CTL_DONE]) __nanosleep(256);    // idle most of the time: stay out of the issue slots
        if (S.ctl[CTL_GEN] < job) return;                                        // done and nothing new posted
        __threadfence_block();
        mh_ring_tasks(p, S, job, rank, cpc, MH_CWARPS, lane);
        job++;
    }
}

// dst[0 .. n) = src[0 .. n) by the consumer threads, both 16-byte aligned: 16-byte L2 loads, eight in flight per thread
__device__ __forceinline__ void mh_copy_wide(double *dst, const double *src, int n, int tid)
{
    if (n <= 0) return;
    const int n2 = n >> 1;
    const double2 *s2 = (const double2 *)src;
    double2 *d2 = (double2 *)dst;
    int i = tid;
#pragma unroll 1
    for (; i + 7 * MH_CONSUMERS < n2; i += 8 * MH_CONSUMERS) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcg(s2 + i + u * MH_CONSUMERS);
#pragma unroll
        for (int u = 0; u < 8; u++) d2[i + u * MH_CONSUMERS] = v[u];
    }
#pragma unroll 1
    for (; i < n2; i += MH_CONSUMERS) d2[i] = __ldcg(s2 + i);
    if ((n & 1) && tid == 0) dst[n - 1] = __ldcg(src + n - 1);
}

// Start of an epoch (right after an accept) in distributed mode: the first two batches of the new epoch are drawn once per
// chain as well.  Every (proposal, sample) has one owner warp somewhere in the chain, the draws land in the two epoch slots
// of the ring, and one extra chain barrier makes them visible; the first batch is then fetched into proposal buffer `buf`.
// (Drawing them redundantly in every CTA cost ~10k cycles per 11 proposals and dominated the price of an accept.)
// Out of line on purpose: it is cold code and must not disturb the register allocation of the likelihood loops.
// Returns true if the watchdog fired.
__device__ __noinline__ bool mh_epoch_start(const MhParams &p, const MhSmem &S, int e, int buf, int n0, int it0, int n1, int rank, int cpc,
                                            unsigned long long bar_no, int tid, int warp, int lane)
{
    const int T = p.T, KT = p.K * T, it1 = it0 + n0, total = (n0 + n1) * T;
    double *es0 = p.ring + (size_t)MH_RING * (size_t)S.bufstride, *es1 = es0 + S.bufstride;
#pragma unroll 1
    for (int li = warp; li * cpc + rank < total; li += MH_CWARPS) {
        const int t = li * cpc + rank, bsel = t / T, tp = t - bsel * T;
        if (bsel < n0) mh_draw_to_slot(p, S, e, es0, bsel, tp, (uint32_t)(it0 + bsel), warp, lane);
        else mh_draw_to_slot(p, S, e, es1, bsel - n0, tp, (uint32_t)(it1 + bsel - n0), warp, lane);
    }
    __threadfence();
    consumer_sync();
    if (tid == 0) {
        red_release_add(p.counter, 1ull);
        const unsigned long long target = (bar_no + 1) * (unsigned long long)cpc;
        const long long t0 = clock64();
        unsigned spins = 0;
        while (ld_acquire(p.counter) < target)
            if ((++spins & 0xff) == 0 && (*(volatile int *)p.abort_flag || clock64() - t0 > MH_WATCHDOG_CYCLES)) {
                atomicExch(p.abort_flag, 1);
                S.ctl[CTL_ABORT] = 1;
                break;
            }
    }
    consumer_sync();
    if (S.ctl[CTL_ABORT]) return true;
    double *dst = S.prop_pi(buf);
    for (int i = tid; i < n0 * KT; i += MH_CONSUMERS) dst[i] = __ldcg(es0 + i);
    for (int i = tid; i < (int)mh_phi_doubles(KT, n0); i += MH_CONSUMERS) dst[S.o_phi + i] = __ldcg(es0 + S.o_phi + i);
    for (int i = tid; i < 3 * n0 * T; i += MH_CONSUMERS) dst[S.o_hast + i] = __ldcg(es0 + S.o_hast + i);
    for (int i = tid; i < n0; i += MH_CONSUMERS) dst[S.o_logr + i] = __ldcg(es0 + S.o_logr + i);
    return false;
}

// One chain = cpc co-resident CTAs (cooperative launch).  Every CTA keeps the node table in shared memory, evaluates its slice
// of the data (whole warp rounds, dealt by cost: plan_partition in pwgs_api.cu) for a batch of nb speculative proposals, adds
// its partial sums to the chain's running integer totals and meets the other CTAs in a per-chain barrier; every CTA then reads
// the same totals and takes the same accept / reject decisions.
// "Prefetching" MH: proposals for iterations it0..it0+nb-1 are all drawn from the current state; the first accepted one ends
// the batch and the later ones are discarded, which is exactly the sequential chain because a rejected iteration leaves the
// state untouched and each iteration owns its random substream.  The batch size adapts to the acceptance rate (doubling after
// all-reject batches up to 128, to 256 after MH_WIDE_AFTER of them in a row; shrinking after accepts).  Proposals are drawn
// ahead of time under the assumption that the batches in flight reject everything: each (proposal, sample) is drawn ONCE per
// chain -- the first two batches after an accept by all warps of the chain (mh_epoch_start, one extra barrier), every later
// batch by the producer warp of its owner CTA two batches ahead (mh_ring_tasks) -- and travels through a ring in global memory
// under the chain barrier's release / acquire.  Chains with few CTAs draw locally in every CTA.
template <bool STAGED, bool BIG>
__global__ void __maxnreg__(MH_MAXNREG) mh_kernel(const MhParams *__restrict__ plist, int cpc)
{
    double *sm = mh_sm;
    __shared__ MhParams p;                              // this chain's descriptor (grid = n_chains * cpc CTAs)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = blockIdx.x % cpc;
    if (tid == 0) p = plist[blockIdx.x / cpc];
    __syncthreads();
    const int T = p.T, K = p.K, KT = K * T, MAXB = p.max_batch;
    const bool dist = p.distributed != 0;

    MhSmem S;
    {
        // BIG: the tree-sized part lives in this CTA's region of p.gscratch (generic loads instead of LDS in the likelihood
        // loops: several times slower, but any tree runs); a separate instantiation so that the normal kernel keeps its
        // shared-memory addressing
        double *q = sm + PWGS_LOG_TAB_DOUBLES + PWGS_EXP_TAB_DOUBLES;
        double *t = BIG ? p.gscratch + (size_t)rank * p.gstride : q;
        S.KT = KT; S.T = T; S.statestride = 4 * KT + 2 * T;
        S.state0 = t; t += 2 * S.statestride;
        S.bufstride = (int)mh_bufstride(K, T, MAXB); S.o_phi = (int)mh_o_phi(KT, MAXB); S.o_hast = (int)mh_o_hast(KT, MAXB); S.o_logr = S.o_hast + 3 * MAXB * T;
        S.buf0 = t; t += 2 * S.bufstride;
        S.pscratch = t; t += 4 * K * (MH_CWARPS + 1);
        if (!BIG) q = t;
        S.red = q; q += MH_RED_DOUBLES(MAXB); S.llh = q; q += MAXB;
        S.seen = (unsigned long long *)q; q += 4 * MAXB;
        S.ctl = (volatile int *)q; S.tin = (int *)q + CTL_WORDS; S.tout = S.tin + K;
    }
    pwgs_log_tab_fill(sm, tid, MH_THREADS);
    for (int i = tid; i < PWGS_EXP_TAB_DOUBLES; i += MH_THREADS) sm[PWGS_LOG_TAB_DOUBLES + i] = g_pwgs_exp_tab[i];
    for (int i = tid; i < KT; i += MH_THREADS) {
        S.pi(0)[i] = p.pi0[i];
        S.phi(0)[i] = p.phi0[i];
        S.prop_pi(0)[i] = p.pi0[i];                    // slot 0 of buffer 0 carries the entry state through the first pass
        S.prop_phi(0)[mh_phi_at(0, i, KT)] = p.phi0[i];
    }
    for (int k = tid; k < K; k += MH_THREADS) { S.tin[k] = p.tin[k]; S.tout[k] = p.tout[k]; }
    if (tid < CTL_PRANGE) S.ctl[tid] = 0;             // (the CTL_PRANGE words are written below)
    for (int i = tid; i < 4 * MAXB; i += MH_THREADS) S.seen[i] = 0ull;

    // this CTA's slices of the plain work list (contiguous) and of the CNV work list (already CTA-major); both are whole
    // warp rounds, dealt so that the CTAs' costs (not their counts) are level
    int p0 = p.pstart[rank], p1 = p.pstart[rank + 1];
    for (int i = 1; i < MH_PCLASSES; i++) { p0 = min(p0, p.pstart[i * (cpc + 1) + rank]); p1 = max(p1, p.pstart[i * (cpc + 1) + rank + 1]); }
    const int pchunk = p.pchunk;
    if (tid < MH_PCLASSES) {                           // per class: first datum (relative to the staged slice) and count
        const int b = p.pstart[tid * (cpc + 1) + rank], e = p.pstart[tid * (cpc + 1) + rank + 1];
        S.ctl[CTL_PRANGE + 2 * tid] = b - p0;
        S.ctl[CTL_PRANGE + 2 * tid + 1] = p.n_classes > 0 ? 0 : e - b;
    }
    const CnvWork &cw = p.cw;
    const int cbase = rank * cw.cchunk, np = p.n_classes > 0 ? 0 : p1 - p0, nc = cw.cta_nc[rank];
    typename std::conditional<STAGED, SliceShared, SliceGlobal>::type sl;
    if constexpr (STAGED) {
        // carve the staging area behind the node tables (16-byte aligned) and copy the slice once
        char *q = (char *)(S.tout + K);
        q = (char *)mh_align16((size_t)q);
        int4 *s_ecoef = (int4 *)q; q += mh_align16((size_t)cw.cchunk * cw.Emax * sizeof(int4));
        double *s_pmu = (double *)q; q += mh_align16((size_t)pchunk * 2 * sizeof(double));
        double *s_cd = (double *)q; q += mh_align16((size_t)cw.cchunk * (1 + (size_t)T) * sizeof(double));
        int *s_pi = (int *)q; q += mh_align16((size_t)pchunk * (1 + 2 * (size_t)T) * sizeof(int));
        int *s_ci = (int *)q;
        for (int i = tid; i < np; i += MH_THREADS) {
            s_pmu[i] = p.pmu_r[p0 + i]; s_pmu[pchunk + i] = p.pmu_v[p0 + i]; s_pi[i] = p.pnode[p0 + i];
            for (int tp = 0; tp < T; tp++) {
                s_pi[(1 + tp) * pchunk + i] = p.pa[(size_t)tp * p.Dp + p0 + i];
                s_pi[(1 + T + tp) * pchunk + i] = p.pd[(size_t)tp * p.Dp + p0 + i];
            }
        }
        const int cc = cw.cchunk;
        for (int i = tid; i < nc; i += MH_THREADS) {
            s_cd[i] = cw.mu_r[cbase + i];
            const int code = cw.code[cbase + i];
            s_ci[i] = code;
            for (int tp = 0; tp < T; tp++) {
                s_cd[(1 + tp) * cc + i] = cw.lbc[(size_t)tp * cw.Mp + cbase + i];
                s_ci[(1 + tp) * cc + i] = cw.a[(size_t)tp * cw.Mp + cbase + i];
                s_ci[(1 + T + tp) * cc + i] = cw.d[(size_t)tp * cw.Mp + cbase + i];
            }
            for (int e = 0; e < (code & 0xff); e++) {
                s_ci[(1 + 2 * T + e) * cc + i] = cw.enode[(size_t)e * cw.Mp + cbase + i];
                s_ecoef[(size_t)e * cc + i] = cw.ecoef[(size_t)e * cw.Mp + cbase + i];
            }
        }
        sl.np = np; sl.pstride = pchunk; sl.nc = nc; sl.cstride = cc; sl.T = T;
        sl.o_ecoef = (int)(((char *)s_ecoef - (char *)sm) / 16); sl.o_pdbl = (int)(s_pmu - sm); sl.o_cdbl = (int)(s_cd - sm);
        sl.o_pint = (int)(s_pi - (int *)sm); sl.o_cint = (int)(s_ci - (int *)sm);
    } else {
        sl.np = np; sl.pstride = p.Dp; sl.nc = nc; sl.cstride = cw.Mp;
        sl.g_pmu_r = p.pmu_r + p0; sl.g_pmu_v = p.pmu_v + p0; sl.g_pnode = p.pnode + p0; sl.g_pa = p.pa + p0; sl.g_pd = p.pd + p0;
        sl.g_cmu_r = cw.mu_r + cbase; sl.g_clbc = cw.lbc + cbase; sl.g_ccode = cw.code + cbase; sl.g_ca = cw.a + cbase;
        sl.g_cd = cw.d + cbase; sl.g_cenode = cw.enode + cbase; sl.g_cecoef = cw.ecoef + cbase;
    }
    __syncthreads();

    if (warp == MH_CWARPS) {                           // ---- producer warp
        if (dist) mh_producer(p, S, rank, cpc, lane);
        return;
    }

    // =================================================================== consumer warps
    mh_refresh_state(p, S, 0, tid, warp, lane);
    consumer_sync();

    const bool prof = p.prof != nullptr && tid == 0;
    unsigned long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#define MH_PROF(slot) do { if (prof) { long long t_ = clock64(); pc[slot] += (unsigned long long)(t_ - t_prev); t_prev = t_; } } while (0)

    int e = 0;                                         // live copy of the chain state
    int cur = 0;                                       // proposal buffer of the current batch
    int b_it = 0, b_nb = 1;                            // current batch: first iteration, size
    int j = -1;                                        // batches since the last accept; -1: the initial evaluation of the entry state
    int accepted = 0;
    int g = 0;                                         // global batch number (proposal ring slots, partial-sum parity)
    unsigned long long bar_no = 0;                     // chain barriers passed so far (one per batch + one per accept)
    int es_ready = 0;                                  // the second batch of the current epoch waits in epoch slot 1 (1), or is
                                                       // already in the other proposal buffer (2: mh_epoch_local)
    double cur_llh = 0.0;

#pragma unroll 1
    while (b_nb > 0) {
        // the batches that follow if everything is rejected
        const int n1_it = j < 0 ? 0 : b_it + b_nb;
        const int n1_nb = j < 0 ? pow2_floor(min(1, p.iters)) : next_batch(b_nb, p.iters - n1_it, batch_cap(j + 1, MAXB));
        const int n2_it = n1_it + n1_nb, n2_nb = next_batch(n1_nb, p.iters - n2_it, batch_cap(j + 2, MAXB));
        if (dist && tid == 0 && g >= 1) {              // post job g: draw batch g+2 from state copy e
            volatile int *jb = S.ctl + CTL_JOB + 3 * (g & 1);
            jb[0] = e; jb[1] = n2_it; jb[2] = n2_nb;
            __threadfence_block();
            S.ctl[CTL_GEN] = g;
        }

        // ---------------- likelihood (param_post) of the CTA's slice for every proposal of the batch
        const int bu = b_nb >= MH_BU ? MH_BU : 1;      // thread groups of MH_BU proposals, or one proposal per group
        const int ngroups = b_nb / bu;                  // proposal groups of the batch
        const int wpg = ngroups < MH_UNITS ? MH_UNITS / ngroups : 1;
        if (ngroups < MH_UNITS) {
            // smaller batches: every group is cut into wpg shares of the data rounds (share r walks rounds r, r+wpg, ...) so that
            // there are MH_UNITS units of work again, dealt to the warps in turn like the groups of a full batch.  A group's
            // class (which partition of the plain data it walks) depends on the group alone, so its shares tile the list.
            // (Shares cost ~20 % more per datum than whole groups -- per-unit set-up and reductions -- so full batches do not use them.)
            const int gfull = max(1, MAXB / MH_BU);
#pragma unroll 1
            for (int v = warp; v < MH_UNITS; v += MH_CWARPS) {
                const int gq = v & (ngroups - 1), role = v / ngroups;
                if (bu == MH_BU) mh_likelihood<MH_BU>(p, S, sl, cur, wpg, role, gq, (gq * p.ncls) / gfull, lane, rank, cpc);
                else mh_likelihood<1>(p, S, sl, cur, wpg, role, gq, 0, lane, rank, cpc);
            }
        } else {
            // large batches: several groups per warp, in turn; a warp's i-th group is of class i (at most MH_PCLASSES: 128
            // proposals = 32 groups over 11 warps), and each class has its own partition of the plain data over the CTAs
            int pcls = 0;
#pragma unroll 1
            for (int gq = warp; gq < ngroups; gq += MH_CWARPS, pcls++)
                mh_likelihood<MH_BU>(p, S, sl, cur, 1, 0, gq, pcls, lane, rank, cpc);
        }
        consumer_sync();
        MH_PROF(0);

        // ---------------- publish the CTA's partial sums and arrive at the chain barrier (release)
        const int parity = g & 1;
        if (warp < PWGS_MAX_BATCH / 32) {
            // The CTA's partial sum t goes into the chain's running totals as two integers, rint(t) and the remainder in
            // units of 2^-52: integer adds commute, so every CTA later reads the same bits whatever order the adds landed in,
            // and the total is exact (no rounding between CTAs).  The totals are never reset: readers take differences.
            // Thread b publishes proposal b; the four warps meet in their own barrier before thread 0 arrives (release).
            unsigned long long *acc = p.sums + (size_t)parity * MH_SUM_STRIDE * PWGS_MAX_BATCH;
            if (tid < b_nb) {
                double t = 0.0;
                for (int r = 0; r < wpg; r++) t += S.red[tid * wpg + r];
                const double th = rint(t);
                atomicAdd(acc + MH_SUM_STRIDE * tid, (unsigned long long)(long long)th);
                atomicAdd(acc + MH_SUM_STRIDE * tid + 1, (unsigned long long)__double2ll_rn((t - th) * 4503599627370496.0));
            }
            asm volatile("bar.sync 2, %0;" ::"n"(PWGS_MAX_BATCH) : "memory");
            if (tid == 0) {
                if (dist) {                                                      // our share of batch g+1 must be in the ring
                    const long long t0 = clock64();
                    unsigned spins = 0;
                    while (S.ctl[CTL_PROD_DONE] < g - 1)
                        if ((++spins & 0xfff) == 0 && clock64() - t0 > MH_WATCHDOG_CYCLES) { atomicExch(p.abort_flag, 1); break; }
                }
                __threadfence_block();
                red_release_add(p.counter, 1ull);
            }
        }
        MH_PROF(1);

        // ---------------- while the barrier is in flight: draw the next batch locally when the ring cannot provide it
        const bool have_next = j == 0 && es_ready == 2;
        const bool next_local = !dist || j < 0 || (j == 0 && es_ready == 0);
        if (next_local) mh_draw_local(p, S, e, cur ^ 1, n1_nb, n1_it, warp, lane);
        MH_PROF(2);
        if (tid == 0) {
            const unsigned long long target = (bar_no + 1) * (unsigned long long)cpc;
            const long long t0 = clock64();
            unsigned spins = 0;
            while (ld_acquire(p.counter) < target)
                if ((++spins & 0xff) == 0 && (*(volatile int *)p.abort_flag || clock64() - t0 > MH_WATCHDOG_CYCLES)) {
                    atomicExch(p.abort_flag, 1);
                    S.ctl[CTL_ABORT] = 1;
                    break;
                }
        }
        bar_no++;
        consumer_sync();
        if (S.ctl[CTL_ABORT]) break;
        MH_PROF(3);

        // ---------------- every CTA reads the same integer totals -> bit-identical decisions everywhere;
        // the loads of the next batch's ring slot (complete: every CTA posted its share before arriving) go out first so that
        // all L2 round trips of this step overlap
        const bool fetch = !next_local && !have_next && n1_nb > 0;
        const double *slot = p.ring + (size_t)(j == 0 ? MH_RING + 1 : (g + 1) % MH_RING) * (size_t)S.bufstride;   // epoch slot 1 / ring
        double f_pi = 0.0, f_phi = 0.0, f_h = 0.0, f_h2 = 0.0, f_h3 = 0.0, f_r = 0.0;
        const int n1_phi = (int)mh_phi_doubles(KT, n1_nb);
        if (fetch) {
            if (tid < n1_nb * KT) f_pi = __ldcg(slot + tid);
            if (tid < n1_phi) f_phi = __ldcg(slot + S.o_phi + tid);
            if (tid < 3 * n1_nb * T) f_h = __ldcg(slot + S.o_hast + tid);                  // the three-word Hastings records: all of a
            if (tid + MH_CONSUMERS < 3 * n1_nb * T) f_h2 = __ldcg(slot + S.o_hast + tid + MH_CONSUMERS);           // full batch at T = 1 in this
            if (tid + 2 * MH_CONSUMERS < 3 * n1_nb * T) f_h3 = __ldcg(slot + S.o_hast + tid + 2 * MH_CONSUMERS);   // one round of loads
            if (tid < n1_nb) f_r = __ldcg(slot + S.o_logr + tid);
        }
        if (warp < PWGS_MAX_BATCH / 32) {
            // thread b: the batch's total for proposal b = what was added to this parity since we last looked, and right away
            // the accept test of proposal b (mh.cpp:73-100); one ballot per warp, combined after the sync below
            bool ok = false;
            if (tid < MAXB) {
                const ulonglong2 now = __ldcg((const ulonglong2 *)(p.sums + ((size_t)parity * PWGS_MAX_BATCH + tid) * MH_SUM_STRIDE));
                unsigned long long *sn = S.seen + (size_t)parity * 2 * MAXB + 2 * tid;
                const unsigned long long was_x = sn[0], was_y = sn[1];
                sn[0] = now.x; sn[1] = now.y;
                const double val = (double)(long long)(now.x - was_x) + (double)(long long)(now.y - was_y) * 2.220446049250313e-16;
                S.llh[tid] = val;
                if (j >= 0 && tid < b_nb) {
                    double a = val - cur_llh;
                    for (int tp = 0; tp < T; tp++) a += S.hast(cur)[3 * (tid * T + tp)];
                    ok = S.logr(cur)[tid] < a;
                }
            }
            const unsigned m = __ballot_sync(PWGS_FULL_MASK, ok);
            if (lane == 0) S.ctl[CTL_OK + warp] = (int)m;
        }
        if (fetch) {
            double *dst = S.prop_pi(cur ^ 1);
            if (tid < n1_nb * KT) dst[tid] = f_pi;
            if (tid < n1_phi) dst[S.o_phi + tid] = f_phi;
            if (tid < 3 * n1_nb * T) dst[S.o_hast + tid] = f_h;
            if (tid + MH_CONSUMERS < 3 * n1_nb * T) dst[S.o_hast + tid + MH_CONSUMERS] = f_h2;
            if (tid + 2 * MH_CONSUMERS < 3 * n1_nb * T) dst[S.o_hast + tid + 2 * MH_CONSUMERS] = f_h3;
            if (tid < n1_nb) dst[S.o_logr + tid] = f_r;
            if constexpr (BIG) {
                // node tables in global memory = wide trees: 256 proposals x K nodes x (pi, phi) are tens of thousands of doubles
                // per batch; with one 8-byte L2 load per thread and trip this copy was half of the kernel at 100 nodes
                mh_copy_wide(dst + MH_CONSUMERS, slot + MH_CONSUMERS, n1_nb * KT - MH_CONSUMERS, tid);
                mh_copy_wide(dst + S.o_phi + MH_CONSUMERS, slot + S.o_phi + MH_CONSUMERS, n1_phi - MH_CONSUMERS, tid);
            } else {
                for (int i = tid + MH_CONSUMERS; i < n1_nb * KT; i += MH_CONSUMERS) dst[i] = __ldcg(slot + i);
                for (int i = tid + MH_CONSUMERS; i < n1_phi; i += MH_CONSUMERS) dst[S.o_phi + i] = __ldcg(slot + S.o_phi + i);
            }
            for (int i = tid + 3 * MH_CONSUMERS; i < 3 * n1_nb * T; i += MH_CONSUMERS) dst[S.o_hast + i] = __ldcg(slot + S.o_hast + i);
        }
        consumer_sync();
        MH_PROF(4);

        // ---------------- accept / reject: the tests were made while the totals were read; pick the first accepted proposal
        int sel = -1;
        if (j < 0) {
            cur_llh = S.llh[0];
        } else {
            for (int w = 0; w < PWGS_MAX_BATCH / 32 && sel < 0; w++) {          // the first accepted proposal wins
                const unsigned m = (unsigned)S.ctl[CTL_OK + w];
                if (m) sel = 32 * w + __ffs(m) - 1;
            }
        }
        pc[7]++;
        g++;
        MH_PROF(5);
        if (sel < 0) {                                  // the batch drawn ahead is the real one
            b_it = n1_it; b_nb = n1_nb; j++;
            cur ^= 1;
            continue;
        }
        // ---------------- update_params (mh.cpp:170-177) into the other state copy and redraw from the new state
        cur_llh = S.llh[sel];
        for (int i = tid; i < KT; i += MH_CONSUMERS) {
            S.pi(e ^ 1)[i] = S.prop_pi(cur)[sel * KT + i];
            S.phi(e ^ 1)[i] = S.prop_phi(cur)[mh_phi_at(sel, i, KT)];
        }
        accepted++;
        e ^= 1;
        b_it += sel + 1;
        // after an accept at position sel, expect the next one about as far away: batch = pow2 >= sel+1
        // ... but no fewer than MH_MIN_EPOCH_BATCH unless the epoch is small enough to start without a chain barrier
        // (mh_epoch_local): a batch of 8 costs a chain what a batch of 2 costs (latency, not arithmetic, at these sizes) and
        // wastes less of the fixed price of the next accept
        int nb0 = 2 * pow2_floor(sel + 1);
        if (!(MH_EPOCH_LOCAL && dist && 3 * nb0 * T <= MH_CWARPS)) nb0 = max(MH_MIN_EPOCH_BATCH, nb0);
        b_nb = pow2_floor(min(min(batch_cap(0, MAXB), nb0), p.iters - b_it));
        j = 0;
        mh_adopt_state(S, e, cur, sel, tid);
        consumer_sync();
        es_ready = 0;
        if (!dist) {
            mh_draw_local(p, S, e, cur ^ 1, b_nb, b_it, warp, lane);
        } else {
            const int n1b = next_batch(b_nb, p.iters - (b_it + b_nb), batch_cap(1, MAXB));
            if (MH_EPOCH_LOCAL && (b_nb + n1b) * T <= MH_CWARPS) {
                mh_epoch_local(p, S, e, cur ^ 1, b_nb, b_it, n1b, warp, lane);
                es_ready = n1b > 0 ? 2 : 0;
            } else {
                if (mh_epoch_start(p, S, e, cur ^ 1, b_nb, b_it, n1b, rank, cpc, bar_no, tid, warp, lane)) break;
                bar_no++;
                es_ready = n1b > 0 ? 1 : 0;
            }
        }
        cur ^= 1;
        consumer_sync();
        MH_PROF(6);
    }
    if (tid == 0) { __threadfence_block(); S.ctl[CTL_DONE] = 1; }
    if (prof) {
        if (blockIdx.x == 0) for (int i = 0; i < 8; i++) p.prof[i] = pc[i];
        p.prof[16 + rank] = pc[0];                      // every CTA's likelihood cycles: the balance of the work partition
    }
#undef MH_PROF

    if (rank == 0) {
        for (int i = tid; i < KT; i += MH_CONSUMERS) { p.pi_out[i] = S.pi(e)[i]; p.phi_out[i] = S.phi(e)[i]; }
        if (tid == 0) { *p.acc_out = accepted; *p.llh_out = cur_llh + p.lbc_plain_sum; }
    }
}

// ============================================================================ math self-test (diagnostics)
// out[0][i] = pwgs_log_pos_n(x[i]) (the MH kernel's logarithm), out[1][i] = pwgs_exp_nonpos_n(-x[i]), out[2][i] = x[i] / (1 + x[i])
// through pwgs_div_pos_n; four values per thread, like the kernel uses them.
__global__ void math_selftest_kernel(int n, const double *__restrict__ x, double *__restrict__ out)
{
    __shared__ __align__(16) double tabs[PWGS_LOG_TAB_DOUBLES + PWGS_EXP_TAB_DOUBLES];
    pwgs_log_tab_fill(tabs, threadIdx.x, blockDim.x);
    for (int t = threadIdx.x; t < PWGS_EXP_TAB_DOUBLES; t += blockDim.x) tabs[PWGS_LOG_TAB_DOUBLES + t] = g_pwgs_exp_tab[t];
    __syncthreads();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    double v[4], l[4], e[4], nx[4], den[4], q[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = x[min(i + k, n - 1)]; nx[k] = -v[k]; den[k] = 1.0 + v[k]; }
    pwgs_log_pos_n<4>(v, l, (const double2 *)tabs + (threadIdx.x & (PWGS_LOG_TAB_COPIES - 1)));
    pwgs_exp_nonpos_n<4>(nx, e, tabs + PWGS_LOG_TAB_DOUBLES);
    pwgs_div_pos_n<4>(v, den, q);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (i + k < n) { out[i + k] = l[k]; out[n + i + k] = e[k]; out[2 * (size_t)n + i + k] = q[k]; }
}

// ============================================================================ FP64 peak probe
__global__ void dfma_peak_kernel(int iters, double seed, double *out)
{
    double x0 = seed + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    const double a = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678) out[0] = x0;
}
