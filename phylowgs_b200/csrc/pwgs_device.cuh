// pwgs_device.cuh — device-side building blocks shared by the kernels of libpwgs_b200.so.
// sm_100a only.  Citations are file:line relative to the reference checkout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

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

// ------------------------------------------------------------------ straight-line fp64 log / exp
// The MH kernel evaluates up to eight independent logarithms per datum.  CUDA's log()/exp() carry slow-path branches that
// keep the compiler from interleaving neighbouring calls; these versions are branch-free (special cases are selects at the
// end), so independent calls overlap in the FP64 pipe.  Accuracy < 1 ulp (same argument reduction and minimax coefficients
// as Sun's fdlibm e_log.c, a published algorithm; not GSL, not the reference).
__device__ __forceinline__ double pwgs_rcp_approx(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}

__device__ __forceinline__ double pwgs_log(double x)
{
    int hi = __double2hiint(x), lo = __double2loint(x);
    const bool sub = hi < 0x00100000;                        // subnormal (or <= 0: fixed up below)
    const double xs = x * 18014398509481984.0;               // 2^54
    hi = sub ? __double2hiint(xs) : hi;
    lo = sub ? __double2loint(xs) : lo;
    int e = (sub ? -54 : 0) + (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;                 // mantissa >= sqrt(2): halve it, bump the exponent
    hi |= i ^ 0x3ff00000;
    e += i >> 20;
    const double f = __hiloint2double(hi, lo) - 1.0;         // m - 1, m in [sqrt(2)/2, sqrt(2))
    const double d = 2.0 + f;
    double y = pwgs_rcp_approx(d);                           // ~2^-20; two Newton steps -> full precision
    double t = fma(-d, y, 1.0);
    y = fma(y, t, y);
    t = fma(-d, y, 1.0);
    y = fma(y, t, y);
    double q = f * y;
    q = fma(fma(-d, q, f), y, q);                            // s = f / (2 + f), correctly rounded to < 1 ulp
    const double z = q * q, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)e;
    double r = fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(q, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
    r = (x == 0.0) ? -INFINITY : r;
    r = (x < 0.0 || x != x) ? NAN : r;
    r = (x == INFINITY) ? INFINITY : r;
    return r;
}

// log(x) for NORMAL POSITIVE FINITE x only (no zero / negative / subnormal / inf / nan handling): the binomial means mu and
// 1 - mu of the likelihood, which pwgs_create / pwgs_set_tree keep strictly inside (0, 1).  Same arithmetic as pwgs_log.
__device__ __forceinline__ double pwgs_log_pos(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;
    hi |= i ^ 0x3ff00000;
    e += i >> 20;
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double d = 2.0 + f;
    double y = pwgs_rcp_approx(d);
    double t = fma(-d, y, 1.0);
    y = fma(y, t, y);
    t = fma(-d, y, 1.0);
    y = fma(y, t, y);
    double q = f * y;
    q = fma(fma(-d, q, f), y, q);
    const double z = q * q, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)e;
    return fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(q, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
}

// N independent pwgs_log_pos evaluations written step by step across the N values, so that the instruction stream the
// compiler emits alternates between the N dependency chains (it does not interleave N separate inlined calls on its own, and a
// single chain leaves the FP64 pipe idle for most of its latency).
template <int N>
__device__ __forceinline__ void pwgs_log_pos_n(const double (&x)[N], double (&r)[N])
{
    // Few values stay live per chain (f, q, z, the Horner accumulator and the integer exponent) so that ptxas can keep all N
    // chains interleaved inside the register budget: the polynomial is ONE Horner chain in z = s^2 (fdlibm splits it into even
    // and odd halves for scalar ILP, which costs two more live doubles per chain; here the N chains are the ILP), the exponent
    // is converted to fp64 only where it is used, and the tail needs nothing but f, s and the polynomial.
    double f[N], d[N], y[N], t[N], q[N], z[N], R[N];
    int k[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        // x = 2^e * m with m in [sqrt(2)/2, sqrt(2)): shift the high word so that the exponent field changes exactly at sqrt(2)
        const int hs = __double2hiint(x[i]) + (0x3ff00000 - 0x3fe6a09e);
        k[i] = (hs >> 20) - 1023;
        f[i] = __hiloint2double((hs & 0x000fffff) + 0x3fe6a09e, __double2loint(x[i])) - 1.0;
    }
#pragma unroll
    for (int i = 0; i < N; i++) d[i] = 2.0 + f[i];
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = pwgs_rcp_approx(d[i]);
    // y ~ 1/d to 2^-20 from the hardware, 2^-40 after one Newton step, and s = f*y is used as it is: s only multiplies the
    // O(f^2) part of log(1+f) = f - hf + s*(hf + R), so the logarithm comes out within 2e-13 relative (measured by
    // pwgs_math_selftest) -- four orders of magnitude inside the 1e-9 the likelihood is specified to, at 20 FP64 instructions
    // instead of the 24 a correctly rounded quotient costs.  K2 and pwgs_total_llh use the full-precision log().
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = fma(-d[i], y[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = fma(y[i], t[i], y[i]);
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = f[i] * y[i];
#pragma unroll
    for (int i = 0; i < N; i++) z[i] = q[i] * q[i];
    // R = z*(Lg1 + z*(Lg2 + ... + z*Lg7))   (fdlibm e_log.c coefficients)
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = fma(z[i], 1.479819860511658591e-01, 1.531383769920937332e-01);
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = fma(z[i], R[i], 1.818357216161805012e-01);
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = fma(z[i], R[i], 2.222219843214978396e-01);
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = fma(z[i], R[i], 2.857142874366239149e-01);
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = fma(z[i], R[i], 3.999999999940941908e-01);
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = fma(z[i], R[i], 6.666666666666735130e-01);
#pragma unroll
    for (int i = 0; i < N; i++) R[i] = z[i] * R[i];
#pragma unroll
    for (int i = 0; i < N; i++) {
        // log(1+f) = 2s + s*R and 2s = f - s*f  =>  log(1+f) = f - s*(f - R) (fdlibm's form for small exponents; its 0.5 f^2
        // variant only buys the last ulp), then + k*ln2 in one fma: the hi/lo split of ln2 matters for huge |k| only; here
        // |k| <= 1074 makes the representation error of the single constant at most 2.3e-14 absolute on |log| >= 0.69 |k|
        const double u = fma(-q[i], f[i] - R[i], f[i]);
        r[i] = fma((double)k[i], 6.931471805599453094e-01, u);
    }
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
template <int N>
__device__ __forceinline__ void pwgs_exp_nonpos_n(const double (&xin)[N], double (&out)[N])
{
    double x[N], kd[N], r[N], p[N];
    int n[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = fmax(xin[i], -700.0);
#pragma unroll
    for (int i = 0; i < N; i++) kd[i] = fma(x[i], 1.4426950408889634074, 6755399441055744.0);
#pragma unroll
    for (int i = 0; i < N; i++) { n[i] = __double2loint(kd[i]); kd[i] -= 6755399441055744.0; }
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(kd[i], -6.93147180369123816490e-01, x[i]);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(kd[i], -1.90821492927058770002e-10, r[i]);
    const double c[13] = {2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
                          2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
                          4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = 1.6059043836821613e-10;
#pragma unroll
    for (int j = 0; j < 13; j++) {
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], c[j]);
    }
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = __hiloint2double(__double2hiint(p[i]) + (n[i] << 20), __double2loint(p[i]));
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
