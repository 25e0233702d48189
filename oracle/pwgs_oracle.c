/* pwgs_oracle.c — TEST INFRASTRUCTURE ONLY (see pwgs_oracle.h for the contract).
 * Plain-C CPU restatement of the PhyloWGS MH / likelihood hot path.  Citations are
 * file:line relative to /root/reference.  Deliberately scalar and sequential: it is
 * the checker, never the product, and never shipped behind libpwgs_b200.so. */
#include "pwgs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gsl/gsl_rng.h"
#include "gsl/gsl_randist.h"

/* ---------------------------------------------------------------- util.cpp:8-43 */
static double log_factorial(int n) { return lgamma(n + 1); }                 /* util.cpp:8-10 */
double orc_log_bin_coeff(int n, int k)                                        /* util.cpp:12-14 */
{ return log_factorial(n) - log_factorial(k) - log_factorial(n - k); }
static double log_binomial_likelihood(int x, int n, double mu)                /* util.cpp:16-18, util2.py:30 */
{ return x * log(mu) + (n - x) * log(1 - mu); }
double orc_logsumexp(const double *x, int n)                                  /* util.cpp:35-43, util2.py:36-38 */
{
    double maxes = x[0], sum = 0.0;
    for (int i = 1; i < n; i++) if (x[i] > maxes) maxes = x[i];
    for (int i = 0; i < n; i++) sum += exp(x[i] - maxes);
    return log(sum) + maxes;
}
void orc_lbc_table(const orc_data *D, double *out)                            /* mh.cpp:234,284; data.py:14 */
{
    int n = (D->n_ssm + D->n_cnv) * D->ntps;
    for (int i = 0; i < n; i++) out[i] = orc_log_bin_coeff(D->d[i], D->a[i]);   /* log_bin_coeff(d,a) */
}

/* ---------------------------------------------------------------- tree helpers */
static int anc_or_self(const int *parent, int u, int v)   /* u in v.get_ancestors()  (node.py:86-92: includes v itself) */
{
    while (v >= 0) { if (v == u) return 1; v = parent[v]; }
    return 0;
}
/* data.py:128-134 == params.py:174-180: walk nd's ancestors from nd upwards; at the first node that hosts
 * one of this SSM's CNVs return the FIRST such link in dat.cnv order.  -1 when none. */
static int most_recent_cnv(const orc_data *D, const orc_tree *T, int j, int nd)
{
    for (int n = nd; n >= 0; n = T->parent[n])
        for (int l = D->cnv_ptr[j]; l < D->cnv_ptr[j + 1]; l++)
            if (T->node_of_datum[D->n_ssm + D->cnv_idx[l]] == n) return l;
    return -1;
}
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* The four-way case split shared by data.py:71-109 (fp64 accumulation) and params.py:131-158
 * (text coefficients): integer (nr,nv) of states 1..4 that node k contributes for SSM j sitting in node s. */
static void state_coefs(const orc_data *D, const orc_tree *T, int j, int s, int k, int nr[4], int nv[4])
{
    int in_sub = anc_or_self(T->parent, s, k);        /* ssm_node in node.get_ancestors() */
    int l = most_recent_cnv(D, T, j, k);
    if (!in_sub && l < 0) {
        for (int q = 0; q < 4; q++) { nr[q] = 2; nv[q] = 0; }
    } else if (in_sub && l < 0) {
        for (int q = 0; q < 4; q++) { nr[q] = 1; nv[q] = 1; }
    } else if (!in_sub && l >= 0) {
        int t = D->cnv_c1[l] + D->cnv_c2[l];
        for (int q = 0; q < 4; q++) { nr[q] = t; nv[q] = 0; }
    } else {
        int c1 = D->cnv_c1[l], c2 = D->cnv_c2[l], t = c1 + c2;
        nr[2] = nr[3] = imax(0, t - 1);
        nv[2] = nv[3] = imin(1, t);
        int cnv_node = T->node_of_datum[D->n_ssm + D->cnv_idx[l]];
        if (anc_or_self(T->parent, s, cnv_node)) {    /* ssm_node in mr_cnv[0].node.get_ancestors() */
            nr[0] = c1; nv[0] = c2;                    /* maternal */
            nr[1] = c2; nv[1] = c1;                    /* paternal */
        } else {
            nr[0] = nr[1] = nr[2];
            nv[0] = nv[1] = nv[2];
        }
    }
}

/* data.py:65-126 */
int orc_compute_n_genomes(const orc_data *D, const orc_tree *T, int j, int s, int tp,
                          const double *pi, double nr[4], double nv[4])
{
    int Tn = D->ntps;
    for (int q = 0; q < 4; q++) nr[q] = nv[q] = 0;
    for (int k = 0; k < T->K; k++) {
        int cr[4], cv[4];
        double p = pi[k * Tn + tp];
        state_coefs(D, T, j, s, k, cr, cv);
        for (int q = 0; q < 4; q++) { nr[q] += p * cr[q]; nv[q] += p * cv[q]; }
    }
    int nl = D->cnv_ptr[j + 1] - D->cnv_ptr[j];
    if (nl == 1 && s == T->node_of_datum[D->n_ssm + D->cnv_idx[D->cnv_ptr[j]]]) return 4;   /* data.py:122-123 */
    return 2;
}

/* params.py:114-171 */
int orc_data_states(const orc_data *D, const orc_tree *T, int *ssm_of_m, int *coef)
{
    int K = T->K, m = 0;
    for (int j = 0; j < D->n_ssm; j++) {
        if (D->cnv_ptr[j + 1] == D->cnv_ptr[j]) continue;            /* "if not dat.cnv: continue" */
        if (coef) {
            double pr[4], pv[4];
            int s = T->node_of_datum[j];
            int ns = orc_compute_n_genomes(D, T, j, s, 0, T->pi, pr, pv);   /* poss_n_genomes = compute_n_genomes(0) */
            int *c = coef + (size_t)m * 4 * K * 2;
            for (int k = 0; k < K; k++) {
                int nr[4], nv[4];
                state_coefs(D, T, j, s, k, nr, nv);
                if (pv[0] == 0) { nr[0] = nr[1]; nv[0] = nv[1]; }              /* params.py:160-161 */
                else if (pv[1] == 0) { nr[1] = nr[0]; nv[1] = nv[0]; }         /* params.py:162-163 */
                if (ns == 2) { nr[2] = nr[0]; nv[2] = nv[0]; nr[3] = nr[1]; nv[3] = nv[1]; }   /* 164-166 */
                for (int q = 0; q < 4; q++) { c[(q * K + k) * 2] = nr[q]; c[(q * K + k) * 2 + 1] = nv[q]; }
            }
        }
        if (ssm_of_m) ssm_of_m[m] = j;
        m++;
    }
    return m;
}

/* mh.hpp:80-163 */
double orc_datum_ll_cpp(const orc_data *D, int K, int Tn, int j, int tp, double phi,
                        const double *pi, const int *coef_m, const double *lbc)
{
    int a = D->a[j * Tn + tp], d = D->d[j * Tn + tp];
    double mu_r = D->mu_r[j], mu_v = D->mu_v[j];
    if (!coef_m) {                                                   /* cnv==NULL: CNV datum or SSM outside CNVs */
        double mu = (1 - phi) * mu_r + phi * mu_v;
        return log_binomial_likelihood(a, d, mu) + lbc[j * Tn + tp];
    }
    double ll[4];
    for (int q = 0; q < 4; q++) {
        double nr = 0, nv = 0;
        for (int k = 0; k < K; k++) {
            nr += pi[k * Tn + tp] * coef_m[(q * K + k) * 2];
            nv += pi[k * Tn + tp] * coef_m[(q * K + k) * 2 + 1];
        }
        if (nr + nv > 0) {
            double mu = (nv * (1 - mu_r) + nr * mu_r) / (nr + nv);
            ll[q] = log_binomial_likelihood(a, d, mu) + log(0.25) + lbc[j * Tn + tp];
        } else {
            ll[q] = log(pow(10, -99));
        }
    }
    return orc_logsumexp(ll, 4);
}

/* data.py:26-62 (update_tree bookkeeping of 33-38 is implicit: the caller's tree arrays are the tree) */
double orc_datum_ll_py(const orc_data *D, const orc_tree *T, int j, int s, const double *phi_s, const double *lbc)
{
    int Tn = D->ntps;
    double total = 0;
    int is_cnv_ssm = j < D->n_ssm && D->cnv_ptr[j + 1] > D->cnv_ptr[j];      /* "if self.cnv" */
    for (int tp = 0; tp < Tn; tp++) {
        int a = D->a[j * Tn + tp], d = D->d[j * Tn + tp];
        double mu_r = D->mu_r[j], mu_v = D->mu_v[j], llh;
        if (is_cnv_ssm) {
            double nr[4], nv[4], ll[4];
            int ns = orc_compute_n_genomes(D, T, j, s, tp, T->pi, nr, nv), n = 0, kept = 0;
            for (int q = 0; q < ns; q++) if (nv[q] > 0) kept++;             /* data.py:52 */
            for (int q = 0; q < ns; q++) {
                if (!(nv[q] > 0)) continue;
                double mu = (nr[q] * mu_r + nv[q] * (1 - mu_r)) / (nr[q] + nv[q]);
                ll[n++] = log_binomial_likelihood(a, d, mu) + log(1.0 / kept) + lbc[j * Tn + tp];
            }
            if (kept == 0) ll[n++] = log(1e-99);                            /* data.py:56-57 */
            llh = orc_logsumexp(ll, n);
        } else {
            double mu = (1 - phi_s[tp]) * mu_r + phi_s[tp] * mu_v;
            llh = log_binomial_likelihood(a, d, mu) + lbc[j * Tn + tp];
        }
        total += llh;
    }
    return total;
}

/* mh.hpp semantics evaluated for SSM j as if it sat in node s: states rebuilt as params.py:114-171 would. */
static double datum_ll_cpp_at(const orc_data *D, const orc_tree *T, int j, int s, const double *phi_s, const double *lbc)
{
    int Tn = D->ntps, K = T->K;
    int is_cnv_ssm = j < D->n_ssm && D->cnv_ptr[j + 1] > D->cnv_ptr[j];
    double total = 0;
    if (!is_cnv_ssm) {
        for (int tp = 0; tp < Tn; tp++) total += orc_datum_ll_cpp(D, K, Tn, j, tp, phi_s[tp], T->pi, NULL, lbc);
        return total;
    }
    int *coef = (int *)malloc(sizeof(int) * 4 * K * 2);
    double pr[4], pv[4];
    int ns = orc_compute_n_genomes(D, T, j, s, 0, T->pi, pr, pv);
    for (int k = 0; k < K; k++) {
        int nr[4], nv[4];
        state_coefs(D, T, j, s, k, nr, nv);
        if (pv[0] == 0) { nr[0] = nr[1]; nv[0] = nv[1]; }
        else if (pv[1] == 0) { nr[1] = nr[0]; nv[1] = nv[0]; }
        if (ns == 2) { nr[2] = nr[0]; nv[2] = nv[0]; nr[3] = nr[1]; nv[3] = nv[1]; }
        for (int q = 0; q < 4; q++) { coef[(q * K + k) * 2] = nr[q]; coef[(q * K + k) * 2 + 1] = nv[q]; }
    }
    for (int tp = 0; tp < Tn; tp++) total += orc_datum_ll_cpp(D, K, Tn, j, tp, phi_s[tp], T->pi, coef, lbc);
    free(coef);
    return total;
}

/* mh.cpp:153-166: nodes in index order, each node's data in increasing datum index, sequential fp64 sum */
double orc_param_post(const orc_data *D, const orc_tree *T, const double *phi, const double *pi,
                      const int *ssm_of_m, const int *coef, const double *lbc, int tp)
{
    int Dn = D->n_ssm + D->n_cnv, K = T->K, Tn = D->ntps;
    int *m_of_ssm = (int *)malloc(sizeof(int) * (D->n_ssm + 1));
    for (int j = 0; j < D->n_ssm; j++) m_of_ssm[j] = -1;
    int M = 0;
    for (int j = 0; j < D->n_ssm; j++) if (D->cnv_ptr[j + 1] > D->cnv_ptr[j]) M++;
    for (int m = 0; m < M; m++) m_of_ssm[ssm_of_m[m]] = m;
    double llh = 0.0;
    for (int i = 0; i < K; i++) {
        double p = phi[i * Tn + tp];
        for (int j = 0; j < Dn; j++) {
            if (T->node_of_datum[j] != i) continue;
            const int *cm = (j < D->n_ssm && m_of_ssm[j] >= 0) ? coef + (size_t)m_of_ssm[j] * 4 * K * 2 : NULL;
            llh += orc_datum_ll_cpp(D, K, Tn, j, tp, p, pi, cm, lbc);
        }
    }
    free(m_of_ssm);
    return llh;
}

typedef struct { int M; int *ssm_of_m; int *coef; double *lbc; } prep_t;
static void prep_make(const orc_data *D, const orc_tree *T, prep_t *P)
{
    int Dn = D->n_ssm + D->n_cnv;
    P->M = orc_data_states(D, T, NULL, NULL);
    P->ssm_of_m = (int *)malloc(sizeof(int) * (P->M + 1));
    P->coef = (int *)malloc(sizeof(int) * ((size_t)P->M * 4 * T->K * 2 + 1));
    P->lbc = (double *)malloc(sizeof(double) * Dn * D->ntps);
    orc_data_states(D, T, P->ssm_of_m, P->coef);
    orc_lbc_table(D, P->lbc);
}
static void prep_free(prep_t *P) { free(P->ssm_of_m); free(P->coef); free(P->lbc); }

static double multi_post(const orc_data *D, const orc_tree *T, const prep_t *P, const double *phi, const double *pi)
{                                                                    /* mh.cpp:147-152 */
    double post = 0.0;
    for (int tp = 0; tp < D->ntps; tp++) post += orc_param_post(D, T, phi, pi, P->ssm_of_m, P->coef, P->lbc, tp);
    return post;
}
double orc_multi_param_post(const orc_data *D, const orc_tree *T, const double *phi, const double *pi)
{
    prep_t P; prep_make(D, T, &P);
    double r = multi_post(D, T, &P, phi, pi);
    prep_free(&P);
    return r;
}

void orc_llh_rows(const orc_data *D, const orc_tree *T, int n_rows, const int *rows, int semantics, double *out)
{
    int Dn = D->n_ssm + D->n_cnv, Tn = D->ntps;
    double *lbc = (double *)malloc(sizeof(double) * Dn * Tn);
    orc_lbc_table(D, lbc);
    for (int r = 0; r < n_rows; r++)
        for (int k = 0; k < T->K; k++)
            out[(size_t)r * T->K + k] = semantics == ORC_SEM_PY
                ? orc_datum_ll_py(D, T, rows[r], k, T->phi + k * Tn, lbc)
                : datum_ll_cpp_at(D, T, rows[r], k, T->phi + k * Tn, lbc);
    free(lbc);
}

void orc_llh_block(const orc_data *D, const orc_tree *T, int n_rows, const int *rows, int n_cols, const int *cols,
                   int semantics, double *out)
{
    int Dn = D->n_ssm + D->n_cnv, Tn = D->ntps;
    double *lbc = (double *)malloc(sizeof(double) * Dn * Tn);
    orc_lbc_table(D, lbc);
    for (int r = 0; r < n_rows; r++)
        for (int c = 0; c < n_cols; c++) {
            int k = cols[c];
            out[(size_t)r * n_cols + c] = semantics == ORC_SEM_PY
                ? orc_datum_ll_py(D, T, rows[r], k, T->phi + k * Tn, lbc)
                : datum_ll_cpp_at(D, T, rows[r], k, T->phi + k * Tn, lbc);
        }
    free(lbc);
}

double orc_total_llh(const orc_data *D, const orc_tree *T, int semantics)
{
    if (semantics == ORC_SEM_CPP) return orc_multi_param_post(D, T, T->phi, T->pi);
    /* tssb.py:404-410 data term: sum over nodes (index order) of complete_logprob (alleles.py:55-56) */
    int Dn = D->n_ssm + D->n_cnv, Tn = D->ntps;
    double *lbc = (double *)malloc(sizeof(double) * Dn * Tn);
    orc_lbc_table(D, lbc);
    double total = 0;
    for (int i = 0; i < T->K; i++) {
        double node_sum = 0;
        for (int j = 0; j < Dn; j++)
            if (T->node_of_datum[j] == i) node_sum += orc_datum_ll_py(D, T, j, i, T->phi + i * Tn, lbc);
        total += node_sum;
    }
    free(lbc);
    return total;
}

/* ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
double orc_u01_from_u64(uint64_t x) { return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0); }

/* Stream layout (DESIGN.md "Random streams"): key = {seed_lo, seed_hi ^ stream_hi};
 * ctr = {iteration, (tp<<16)|node, (purpose<<24)|index, stream_lo}. */
enum { PURPOSE_GAMMA = 0, PURPOSE_ACCEPT = 1 };
static void stream_draw(uint64_t seed, uint64_t stream, uint32_t it, uint32_t node, uint32_t tp,
                        uint32_t purpose, uint32_t idx, double u[2])
{
    uint32_t ctr[4] = { it, (tp << 16) | node, (purpose << 24) | idx, (uint32_t)stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(stream >> 32) };
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    u[0] = orc_u01_from_u64(((uint64_t)o[1] << 32) | o[0]);
    u[1] = orc_u01_from_u64(((uint64_t)o[3] << 32) | o[2]);
}
/* Marsaglia-Tsang Gamma(alpha>=1, 1) on the counter stream: attempt j consumes draws idx 2j (Box-Muller
 * normal from two uniforms) and 2j+1 (acceptance uniform). */
static double philox_gamma(double alpha, uint64_t seed, uint64_t stream, uint32_t it, uint32_t node, uint32_t tp)
{
    double d = alpha - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d);
    for (uint32_t j = 0;; j++) {
        double un[2], ua[2];
        stream_draw(seed, stream, it, node, tp, PURPOSE_GAMMA, 2 * j, un);
        double x = sqrt(-2.0 * log(un[0])) * cos(2.0 * M_PI * un[1]);
        double v = 1.0 + c * x;
        if (v <= 0) continue;
        v = v * v * v;
        stream_draw(seed, stream, it, node, tp, PURPOSE_GAMMA, 2 * j + 1, ua);
        double u = ua[0];
        if (u < 1 - 0.0331 * x * x * x * x) return d * v;
        if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return d * v;
    }
}
double orc_philox_accept_uniform(uint64_t seed, uint64_t stream, uint32_t iter)
{
    double u[2];
    stream_draw(seed, stream, iter, 0, 0, PURPOSE_ACCEPT, 0, u);
    return u[0];
}

/* children-before-parent accumulation of mh.cpp:134-141: param1_i = pi1_i + sum_{c in cids} param1_c,
 * children taken in increasing index (= cids order for the post-order file of params.py:78-99). */
static void accumulate_phi(int K, int Tn, const int *parent, const double *pi1, double *phi1, int tp)
{
    /* process nodes deepest-first; a node's children are summed left-to-right in index order */
    int *depth = (int *)malloc(sizeof(int) * K), maxd = 0;
    for (int i = 0; i < K; i++) { int dd = 0; for (int v = parent[i]; v >= 0; v = parent[v]) dd++; depth[i] = dd; if (dd > maxd) maxd = dd; }
    for (int dd = maxd; dd >= 0; dd--)
        for (int i = 0; i < K; i++) {
            if (depth[i] != dd) continue;
            double p = pi1[i * Tn + tp];
            for (int c = 0; c < K; c++) if (parent[c] == i) p += phi1[c * Tn + tp];
            phi1[i * Tn + tp] = p;
        }
    free(depth);
}

void orc_philox_proposal(int K, int Tn, const int *parent, const double *pi, double mh_std,
                         uint64_t seed, uint64_t stream, uint32_t iter, double *pi1, double *phi1)
{
    double std = (double)(float)mh_std;                                        /* mh.hpp:28: MH_STD is float */
    for (int tp = 0; tp < Tn; tp++) {
        double sum = 0;
        for (int i = 0; i < K; i++) {                                          /* mh.cpp:125-128, util.cpp:25 */
            double alpha = std * pi[i * Tn + tp] + 1;
            pi1[i * Tn + tp] = philox_gamma(alpha, seed, stream, iter, (uint32_t)i, (uint32_t)tp);
            sum += pi1[i * Tn + tp];
        }
        for (int i = 0; i < K; i++) pi1[i * Tn + tp] /= sum;
        for (int i = 0; i < K; i++) pi1[i * Tn + tp] = pi1[i * Tn + tp] + 0.0001;   /* util.cpp:26-27 */
        sum = 0.0;
        for (int i = 0; i < K; i++) sum += pi1[i * Tn + tp];                   /* util.cpp:28-30 */
        for (int i = 0; i < K; i++) pi1[i * Tn + tp] /= sum;                   /* util.cpp:31-32 */
        accumulate_phi(K, Tn, parent, pi1, phi1, tp);
    }
}

/* mh.cpp:65-109 */
int orc_mh_loop(const orc_data *D, const orc_tree *T, int iters, double mh_std, int rng_kind,
                uint64_t seed, uint64_t stream, int recompute_old,
                double *phi_out, double *pi_out, double *llh_out, double *trace)
{
    int K = T->K, Tn = D->ntps, KT = K * Tn;
    double std = (double)(float)mh_std;
    prep_t P; prep_make(D, T, &P);
    double *phi = phi_out, *pi = pi_out;
    double *phi1 = (double *)calloc(KT, sizeof(double)), *pi1 = (double *)calloc(KT, sizeof(double));
    double *alpha = (double *)malloc(sizeof(double) * K), *x = (double *)malloc(sizeof(double) * K);
    double *theta = (double *)malloc(sizeof(double) * K), *pv = (double *)malloc(sizeof(double) * K), *pn = (double *)malloc(sizeof(double) * K);
    memcpy(phi, T->phi, sizeof(double) * KT);
    memcpy(pi, T->pi, sizeof(double) * KT);
    gsl_rng *rng = rng_kind == ORC_RNG_MT19937_SHIM ? gsl_rng_alloc(gsl_rng_mt19937) : NULL;   /* mh.cpp:66: default seed */
    int accepted = 0;
    double post_old = multi_post(D, T, &P, phi, pi);
    for (int itr = 0; itr < iters; itr++) {
        if (rng_kind == ORC_RNG_MT19937_SHIM) {
            for (int tp = 0; tp < Tn; tp++) {                                  /* sample_cons_params, mh.cpp:113-142 */
                for (int i = 0; i < K; i++) alpha[i] = std * pi[i * Tn + tp] + 1;
                gsl_ran_dirichlet(rng, K, alpha, x);                           /* dirichlet_sample, util.cpp:24-33 */
                for (int i = 0; i < K; i++) x[i] = x[i] + 0.0001;
                double sum = 0.0;
                for (int i = 0; i < K; i++) sum += x[i];
                for (int i = 0; i < K; i++) x[i] /= sum;
                for (int i = 0; i < K; i++) pi1[i * Tn + tp] = x[i];
                accumulate_phi(K, Tn, T->parent, pi1, phi1, tp);
            }
        } else {
            orc_philox_proposal(K, Tn, T->parent, pi, mh_std, seed, stream, (uint32_t)itr, pi1, phi1);
        }
        if (recompute_old) post_old = multi_post(D, T, &P, phi, pi);           /* mh.cpp:73 recomputes both */
        double post_new = multi_post(D, T, &P, phi1, pi1);
        double a = post_new - post_old;
        for (int tp = 0; tp < Tn; tp++) {                                      /* mh.cpp:80-93 */
            for (int i = 0; i < K; i++) { pn[i] = pi1[i * Tn + tp]; pv[i] = pi[i * Tn + tp]; }
            for (int i = 0; i < K; i++) theta[i] = std * pn[i];
            a += gsl_ran_dirichlet_lnpdf(K, theta, pv);
            for (int i = 0; i < K; i++) theta[i] = std * pv[i];
            a -= gsl_ran_dirichlet_lnpdf(K, theta, pn);
        }
        double r = rng_kind == ORC_RNG_MT19937_SHIM ? gsl_rng_uniform_pos(rng)
                                                    : orc_philox_accept_uniform(seed, stream, (uint32_t)itr);
        int acc = log(r) < a;                                                  /* mh.cpp:95-100 */
        if (acc) {
            accepted++;
            memcpy(phi, phi1, sizeof(double) * KT);                            /* update_params, mh.cpp:170-177 */
            memcpy(pi, pi1, sizeof(double) * KT);
            post_old = post_new;
        }
        if (trace) { trace[2 * itr] = a; trace[2 * itr + 1] = acc; }
    }
    if (rng) gsl_rng_free(rng);
    if (llh_out) *llh_out = multi_post(D, T, &P, phi, pi);
    free(phi1); free(pi1); free(alpha); free(x); free(theta); free(pv); free(pn);
    prep_free(&P);
    return accepted;
}
