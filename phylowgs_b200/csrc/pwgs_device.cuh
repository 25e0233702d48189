// pwgs_device.cuh — device-side building blocks shared by the kernels of libpwgs_b200.so.
// sm_100a only.  Citations are file:line relative to the reference checkout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PWGS_MAX_BATCH 16
#define PWGS_FULL_MASK 0xffffffffu

// ------------------------------------------------------------------ tree tables
// Ancestor tests use Euler-tour intervals computed on the host in pwgs_set_tree:
// u is an ancestor-or-self of v  <=>  tin[u] <= tin[v] && tout[v] <= tout[u]   (node.py:86-92 get_ancestors includes self)
struct TreeView {
    int K, T;
    const int *parent;   // [K]
    const int *depth;    // [K]
    const int *tin;      // [K]
    const int *tout;     // [K]
    const double *phi;   // [K*T]
    const double *pi;    // [K*T]
    const int *node_of_datum;  // [D]
};

struct DataView {
    int n_ssm, n_cnv, T;
    const int *a, *d;            // [D*T] row-major
    const double *mu_r, *mu_v;   // [D]
    const double *lbc;           // [D*T]
    const int *cnv_ptr, *cnv_idx, *cnv_c1, *cnv_c2;
};

__device__ __forceinline__ bool anc_or_self(const TreeView &t, int u, int v)
{
    return t.tin[u] <= t.tin[v] && t.tout[v] <= t.tout[u];
}

// data.py:128-134 == params.py:174-180: deepest ancestor-or-self of node k hosting one of SSM j's CNVs;
// two CNVs in the same node -> the first link in file order (strict '>' keeps it).
__device__ __forceinline__ int most_recent_cnv(const DataView &d, const TreeView &t, int j, int k)
{
    int best = -1, best_depth = -1;
    for (int l = d.cnv_ptr[j]; l < d.cnv_ptr[j + 1]; l++) {
        int n = t.node_of_datum[d.n_ssm + d.cnv_idx[l]];
        if (anc_or_self(t, n, k) && t.depth[n] > best_depth) { best = l; best_depth = t.depth[n]; }
    }
    return best;
}

// The four-way case split of data.py:71-109 / params.py:131-158: integer (nr,nv) that node k contributes to the
// four genome states of SSM j sitting in node s.
__device__ __forceinline__ void state_coefs(const DataView &d, const TreeView &t, int j, int s, int k, int nr[4], int nv[4])
{
    bool in_sub = anc_or_self(t, s, k);
    int l = most_recent_cnv(d, t, j, k);
    if (l < 0) {
        int r = in_sub ? 1 : 2, v = in_sub ? 1 : 0;
#pragma unroll
        for (int q = 0; q < 4; q++) { nr[q] = r; nv[q] = v; }
    } else if (!in_sub) {
        int tt = d.cnv_c1[l] + d.cnv_c2[l];
#pragma unroll
        for (int q = 0; q < 4; q++) { nr[q] = tt; nv[q] = 0; }
    } else {
        int c1 = d.cnv_c1[l], c2 = d.cnv_c2[l], tt = c1 + c2;
        nr[2] = nr[3] = max(0, tt - 1);
        nv[2] = nv[3] = min(1, tt);
        int cnv_node = t.node_of_datum[d.n_ssm + d.cnv_idx[l]];
        if (anc_or_self(t, s, cnv_node)) { nr[0] = c1; nv[0] = c2; nr[1] = c2; nv[1] = c1; }
        else { nr[0] = nr[1] = nr[2]; nv[0] = nv[1] = nv[2]; }
    }
}

// ------------------------------------------------------------------ likelihood terms
// util.cpp:16-18 / util2.py:30
__device__ __forceinline__ double log_binomial_likelihood(int x, int n, double mu)
{
    return (double)x * log(mu) + (double)(n - x) * log(1.0 - mu);
}

// util.cpp:35-43 over exactly four values
__device__ __forceinline__ double logsumexp4(const double x[4])
{
    double m = x[0];
#pragma unroll
    for (int i = 1; i < 4; i++) if (x[i] > m) m = x[i];
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += exp(x[i] - m);
    return log(s) + m;
}

// mh.hpp:95-101 for one state
__device__ __forceinline__ double mh_state_ll(double nr, double nv, int a, int d, double mu_r, double lbc,
                                              double log_quarter, double log_tiny)
{
    if (nr + nv > 0) {
        double mu = (nv * (1.0 - mu_r) + nr * mu_r) / (nr + nv);
        return log_binomial_likelihood(a, d, mu) + log_quarter + lbc;
    }
    return log_tiny;
}

// ------------------------------------------------------------------ Philox4x32-10 counter RNG
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u01_from_u64(uint64_t x)
{
    return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

// Stream layout (DESIGN.md "Random streams"):
//   key = {seed_lo, seed_hi ^ stream_hi}; ctr = {iteration, (tp<<16)|node, (purpose<<24)|index, stream_lo}
#define PWGS_PURPOSE_GAMMA 0u
#define PWGS_PURPOSE_ACCEPT 1u
__device__ __forceinline__ void stream_draw(uint64_t seed, uint64_t stream, uint32_t it, uint32_t node, uint32_t tp,
                                            uint32_t purpose, uint32_t idx, double &u0, double &u1)
{
    uint32_t o[4];
    philox4x32_10(it, (tp << 16) | node, (purpose << 24) | idx, (uint32_t)stream,
                  (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(stream >> 32), o);
    u0 = u01_from_u64(((uint64_t)o[1] << 32) | o[0]);
    u1 = u01_from_u64(((uint64_t)o[3] << 32) | o[2]);
}

// Marsaglia & Tsang (2000) Gamma(alpha >= 1, 1).  Attempt j consumes draws 2j (Box-Muller normal) and 2j+1
// (acceptance uniform) of the (iteration, node, tp) substream.  Replaces gsl_ran_gamma under gsl_ran_dirichlet (util.cpp:25).
__device__ __forceinline__ double philox_gamma(double alpha, uint64_t seed, uint64_t stream, uint32_t it, uint32_t node, uint32_t tp)
{
    double d = alpha - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d);
    for (uint32_t j = 0;; j++) {
        double u0, u1, ua, ub;
        stream_draw(seed, stream, it, node, tp, PWGS_PURPOSE_GAMMA, 2 * j, u0, u1);
        double x = sqrt(-2.0 * log(u0)) * cospi(2.0 * u1);
        double v = 1.0 + c * x;
        if (v <= 0) continue;
        v = v * v * v;
        stream_draw(seed, stream, it, node, tp, PWGS_PURPOSE_GAMMA, 2 * j + 1, ua, ub);
        if (ua < 1.0 - 0.0331 * x * x * x * x) return d * v;
        if (log(ua) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
    }
}

// ------------------------------------------------------------------ reductions / sync
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PWGS_FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ void red_release_add(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
