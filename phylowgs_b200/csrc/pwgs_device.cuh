// pwgs_device.cuh — device-side building blocks shared by the kernels of libpwgs_b200.so.
// sm_100a only.  Citations are file:line relative to the reference checkout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "pwgs_tables.cuh"

#define PWGS_MAX_BATCH 256
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
    const double *phisub; // [K*T] subtree sums of pi (children-first on the host, pwgs_set_tree): what the sparse grid sums multiply
    int root;
};

struct DataView {
    int n_ssm, n_cnv, T;
    const int *a, *d;            // [D*T] row-major
    const double *mu_r, *mu_v;   // [D]
    const double *lbc;           // [D*T]
    const int *cnv_ptr, *cnv_idx, *cnv_c1, *cnv_c2;
};

__host__ __device__ __forceinline__ bool anc_or_self(const TreeView &t, int u, int v)
{
    return t.tin[u] <= t.tin[v] && t.tout[v] <= t.tout[u];
}

// data.py:128-134 == params.py:174-180: deepest ancestor-or-self of node k hosting one of SSM j's CNVs;
// two CNVs in the same node -> the first link in file order (strict '>' keeps it).
__host__ __device__ __forceinline__ int most_recent_cnv(const DataView &d, const TreeView &t, int j, int k)
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
__host__ __device__ __forceinline__ void state_coefs(const DataView &d, const TreeView &t, int j, int s, int k, int nr[4], int nv[4])
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
        nr[2] = nr[3] = tt > 1 ? tt - 1 : 0;          // max(0, t-1), min(1, t)
        nv[2] = nv[3] = tt < 1 ? tt : 1;
        int cnv_node = t.node_of_datum[d.n_ssm + d.cnv_idx[l]];
        if (anc_or_self(t, s, cnv_node)) { nr[0] = c1; nv[0] = c2; nr[1] = c2; nv[1] = c1; }
        else { nr[0] = nr[1] = nr[2]; nv[0] = nv[1] = nv[2]; }
    }
}

// ------------------------------------------------------------------ reciprocal seed
__device__ __forceinline__ double pwgs_rcp_approx(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}

// ------------------------------------------------------------------ the MH kernel's logarithm and exponential
// Table-driven (the scheme of the ARM optimized-routines / glibc 2.28+ `log` and `exp`, published algorithms; tables generated
// by scripts/gen_tables.py into pwgs_tables.cuh and copied to shared memory by the kernel):
//   x = 2^k z, z in [sqrt(2)/2, sqrt(2));  the top 7 bits of z's mantissa field pick (1/c, log c);  r = z/c - 1, |r| <= 2^-8;
//   log x = (k ln2 + log c) + (r + r^2 (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6)).
// 9 FP64 instructions per logarithm, no reciprocal and no division (the fdlibm form it replaces: 16 + one MUFU.RCP64H: the XU pipe
// under the reciprocal seeds and int->fp64 conversions was a second limiter of the likelihood loop beside the FP64 pipe, see
// scripts/micro/plain_round.cu), a dependency chain of 7 instead of 16.  Measured 2e-14 relative (pwgs_math_selftest), the
// interval around z = 1 is exact in the table so arguments next to 1 keep their relative accuracy without a branch.
// N independent evaluations are written step by step across the N values so that ptxas emits the chains interleaved.
// NORMAL POSITIVE FINITE arguments only: the binomial means mu, 1 - mu (pwgs_create / pwgs_set_tree keep them inside (0, 1))
// and the logsumexp sums in [1, 4].
#define PWGS_LOG_TAB_COPIES 8            // interleaved copies of the logarithm table in shared memory, see pwgs_log_pos_n
#define PWGS_LOG_TAB_DOUBLES (256 * PWGS_LOG_TAB_COPIES)

#define PWGS_EXP_TAB_DOUBLES 128
// shared-memory image of the table: entry j of copy c at double2 index 8 j + c
__device__ __forceinline__ void pwgs_log_tab_fill(double *dst, int tid, int nthreads)
{
    for (int i = tid; i < 256 * PWGS_LOG_TAB_COPIES; i += nthreads) {
        const int j = i / (2 * PWGS_LOG_TAB_COPIES), rest = i % (2 * PWGS_LOG_TAB_COPIES);      // rest = 2 c + {0: 1/c, 1: log c}
        dst[i] = g_pwgs_log_tab[2 * j + (rest & 1)];
    }
}
// `tab`: the calling lane's copy, i.e. (const double2 *)table + (lane & 7)
template <int N>
__device__ __forceinline__ void pwgs_log_pos_n(const double (&x)[N], double (&res)[N], const double2 *__restrict__ tab)
{
    double r[N], w[N], r2[N], P[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        // hs: the high word moved down so that its exponent field IS k and changes exactly at sqrt(2) (0x3fe6a09e = high word
        // of sqrt(2)/2); few integer instructions on purpose: the loop is close to issue-bound
        const int hi = __double2hiint(x[i]);
        const int hs = hi - 0x3fe6a09e;
        const int k20 = hs & 0xfff00000;                                                             // k * 2^20
        const double z = __hiloint2double(hi - k20, __double2loint(x[i]));
        // interval j = (hs >> 13) & 127 lives at tab[8 j + (lane & 7)] (`tab` already points at the lane's copy): the 8 lanes of a
        // quarter-warp then read 8 different 16-byte slots of the 128-byte rows, i.e. an LDS.128 is 4 conflict-free wavefronts
        // whatever the 32 intervals are.  (One copy: 9.5 wavefronts per lookup with the 11 distinct means of a round, and the
        // shared-memory pipe at 99 % was what bounded the loop: scripts/micro/plain_round.cu under ncu.)
        const double2 t = *(const double2 *)((const char *)tab + ((hs & 0x000fe000) >> 6));
        r[i] = fma(z, t.x, -1.0);
        // k ln2 + log c in one rounding (the conversion takes k * 2^20 as it is, the constant is ln2 * 2^-20); |k| <= 1074 keeps
        // the representation error of the single ln2 constant below 2.5e-14 absolute on |log| >= 0.69 |k|
        w[i] = fma((double)k20, 6.931471805599453094e-01 / 1048576.0, t.y);
    }
#pragma unroll
    for (int i = 0; i < N; i++) r2[i] = r[i] * r[i];
#pragma unroll
    for (int i = 0; i < N; i++) P[i] = fma(r[i], -1.0 / 6.0, 0.2);
#pragma unroll
    for (int i = 0; i < N; i++) P[i] = fma(r[i], P[i], -0.25);
#pragma unroll
    for (int i = 0; i < N; i++) P[i] = fma(r[i], P[i], 1.0 / 3.0);
#pragma unroll
    for (int i = 0; i < N; i++) P[i] = fma(r[i], P[i], -0.5);
#pragma unroll
    for (int i = 0; i < N; i++) res[i] = w[i] + fma(r2[i], P[i], r[i]);
}
template <int N>
__device__ __forceinline__ void pwgs_div_pos_n(const double (&a)[N], const double (&b)[N], double (&r)[N])
{
    // a / b to 2^-40 relative (hardware reciprocal, 2^-20, and one Newton step); see pwgs_log_pos_n for why that is enough here
    double y[N], t[N];
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = pwgs_rcp_approx(b[i]);
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = fma(-b[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = fma(y[i], t[i], y[i]);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = a[i] * y[i];
}
// exp(x) for x <= 0 (arguments below -700 count as -700):  x = (128 k + i) ln2/128 + r, |r| <= ln2/256;
// exp x = 2^k 2^(i/128) (1 + r + r^2 (1/2 + r/6 + r^2/24)): 10 FP64 instructions (the 13-term Horner form it replaces: 18);
// 3e-15 relative (pwgs_math_selftest).
template <int N>
__device__ __forceinline__ void pwgs_exp_nonpos_n(const double (&xin)[N], double (&out)[N], const double *__restrict__ tab)
{
    double x[N], kd[N], r[N], r2[N], p[N], sc[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = fmax(xin[i], -700.0);
#pragma unroll
    for (int i = 0; i < N; i++) kd[i] = fma(x[i], 1.8466496523378731e+02, 6755399441055744.0);      // 128/ln2; round to nearest integer
#pragma unroll
    for (int i = 0; i < N; i++) {
        const int n = __double2loint(kd[i]);                                                        // 128 k + i  (<= 0)
        kd[i] -= 6755399441055744.0;
        const double t = tab[n & 127];
        sc[i] = __hiloint2double(__double2hiint(t) + ((n >> 7) << 20), __double2loint(t));         // 2^k 2^(i/128): normal for x >= -700
    }
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(kd[i], -5.41521234663378e-03, x[i]);                     // ln2/128, high 32 bits: kd * hi is exact
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(kd[i], -1.4907929134926466e-12, r[i]);                   // low part
#pragma unroll
    for (int i = 0; i < N; i++) r2[i] = r[i] * r[i];
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = fma(r[i], 1.0 / 24.0, 1.0 / 6.0);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = fma(r[i], p[i], 0.5);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = fma(r2[i], p[i], r[i]);
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = fma(sc[i], p[i], sc[i]);
}

// log(w/4) and log(w) for state multiplicities w = 0..4 (index 0 unused)
__constant__ double c_logw4[5] = {0.0, -1.3862943611198906, -0.6931471805599453, -0.2876820724517809, 0.0};
__constant__ double c_logw[5] = {0.0, 0.0, 0.6931471805599453, 1.0986122886681098, 1.3862943611198906};

// ------------------------------------------------------------------ likelihood terms
// util.cpp:16-18 / util2.py:30
__host__ __device__ __forceinline__ double log_binomial_likelihood(int x, int n, double mu)
{
    return (double)x * log(mu) + (double)(n - x) * log(1.0 - mu);
}

// util.cpp:35-43 over exactly four values
__host__ __device__ __forceinline__ double logsumexp4(const double x[4])
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
__host__ __device__ __forceinline__ double mh_state_ll(double nr, double nv, int a, int d, double mu_r, double lbc,
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
    // the attempt cap only guards against a non-finite alpha (never reached otherwise: each attempt succeeds with p > 0.95)
    for (uint32_t j = 0; j < 4096u; j++) {
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
    return d;
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
