/* gsl_shim.c — TEST INFRASTRUCTURE.  A small stand-in for the six GSL entry points
 * the reference MH sampler links against (mh.cpp:66,88,92,95,102; util.cpp:25).
 * libgsl is absent from this image, so the unmodified reference sources are linked
 * against this file to obtain a runnable `mh.o` (oracle/_ref/).  Written from the
 * published algorithms, not from GSL source:
 *   - MT19937 (Matsumoto & Nishimura 1998, 2002 initialisation), default seed 4357
 *     (GSL maps seed 0 -> 4357);
 *   - uniform_pos: 32-bit draw / 2^32, redrawn while 0;
 *   - Gaussian: Marsaglia polar method;
 *   - Gamma(a>=1): Marsaglia & Tsang (2000); a<1 via the U^(1/a) boost;
 *   - Dirichlet: normalised independent Gammas;
 *   - Dirichlet log-pdf: lgamma(sum a) - sum lgamma(a_i) + sum (a_i-1) log(theta_i).
 * "Parity unpinned" at this boundary: the real GSL uses a ziggurat Gaussian, so the
 * random stream differs from a true GSL build; only distributions agree.
 * The same generator is exported as pwgs_shim_* for oracle/pwgs_oracle.c so the C
 * restatement of mh_loop can be replayed draw-for-draw against the compiled reference. */
#include <stdlib.h>
#include <math.h>
#include "gsl/gsl_rng.h"
#include "gsl/gsl_randist.h"

#define MT_N 624
#define MT_M 397
struct gsl_rng_s { unsigned long mt[MT_N]; int mti; };

static const gsl_rng_type mt_type = { "mt19937" };
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;

void gsl_rng_set(gsl_rng *r, unsigned long s)
{
    if (s == 0) s = 4357;
    r->mt[0] = s & 0xffffffffUL;
    for (int i = 1; i < MT_N; i++)
        r->mt[i] = (1812433253UL * (r->mt[i-1] ^ (r->mt[i-1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
    r->mti = MT_N;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    (void)T;
    gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
    gsl_rng_set(r, 0);
    return r;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

static unsigned long mt_get(gsl_rng *r)
{
    unsigned long *mt = r->mt;
    if (r->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk+1] & 0x7fffffffUL);
            mt[kk] = mt[kk+MT_M] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        for (; kk < MT_N - 1; kk++) {
            unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk+1] & 0x7fffffffUL);
            mt[kk] = mt[kk+(MT_M-MT_N)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        unsigned long y = (mt[MT_N-1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
        mt[MT_N-1] = mt[MT_M-1] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        r->mti = 0;
    }
    unsigned long k = mt[r->mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680UL;
    k ^= (k << 15) & 0xefc60000UL;
    k ^= (k >> 18);
    return k & 0xffffffffUL;
}

static double mt_uniform(gsl_rng *r) { return mt_get(r) / 4294967296.0; }

double gsl_rng_uniform_pos(gsl_rng *r)
{
    double x;
    do { x = mt_uniform(r); } while (x == 0);
    return x;
}

double gsl_ran_gaussian(gsl_rng *r, double sigma)
{
    double x, y, r2;
    do {
        x = -1 + 2 * gsl_rng_uniform_pos(r);
        y = -1 + 2 * gsl_rng_uniform_pos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

double gsl_ran_gamma(gsl_rng *r, double a, double b)
{
    if (a < 1) {
        double u = gsl_rng_uniform_pos(r);
        return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
    }
    double x, v, u;
    double d = a - 1.0 / 3.0;
    double c = (1.0 / 3.0) / sqrt(d);
    for (;;) {
        do {
            x = gsl_ran_gaussian(r, 1.0);
            v = 1.0 + c * x;
        } while (v <= 0);
        v = v * v * v;
        u = gsl_rng_uniform_pos(r);
        if (u < 1 - 0.0331 * x * x * x * x) break;
        if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
    }
    return b * d * v;
}

void gsl_ran_dirichlet(gsl_rng *r, size_t K, const double alpha[], double theta[])
{
    double norm = 0.0;
    for (size_t i = 0; i < K; i++) theta[i] = gsl_ran_gamma(r, alpha[i], 1.0);
    for (size_t i = 0; i < K; i++) norm += theta[i];
    for (size_t i = 0; i < K; i++) theta[i] /= norm;
}

double gsl_ran_dirichlet_lnpdf(size_t K, const double alpha[], const double theta[])
{
    double log_p = 0.0, sum_alpha = 0.0;
    for (size_t i = 0; i < K; i++) log_p += (alpha[i] - 1.0) * log(theta[i]);
    for (size_t i = 0; i < K; i++) sum_alpha += alpha[i];
    log_p += lgamma(sum_alpha);
    for (size_t i = 0; i < K; i++) log_p -= lgamma(alpha[i]);
    return log_p;
}
