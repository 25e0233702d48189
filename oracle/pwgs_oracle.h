/* pwgs_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the PhyloWGS hot path: the fp64 CNV-aware log-binomial
 * likelihood in its two conventions (C++ mh.hpp and Python data.py), the
 * data-state coefficient construction of params.py, and the Metropolis-Hastings loop
 * of mh.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (phylowgs_b200/, libpwgs_b200.so) never does.
 *
 * Pinned against: the UNMODIFIED reference mh.cpp/util.cpp compiled into
 * oracle/_ref/libpwgs_ref.so (tests/test_oracle_vs_ref.py when /root/reference or the
 * prebuilt _ref exists) and the committed golden vectors tests/golden/ (JSON files) generated
 * from it (tests/golden/make_golden.py).  The Python-semantics half (data.py:26-126,
 * params.py:68-171) is pinned against tests/golden/py_vectors.json: outputs of the
 * reference's own data.py / params.py / tssb.py executed under Python 3 by
 * oracle/ref_py3.py (declared mechanical patches; tests/golden/make_py_golden.py,
 * tests/test_py_golden.py).
 *
 * All citations are file:line relative to /root/reference.
 */
#ifndef PWGS_ORACLE_H
#define PWGS_ORACLE_H
#include <stdint.h>

typedef struct {
    int n_ssm, n_cnv, ntps;      /* D = n_ssm + n_cnv data; datum index: SSM k -> k, CNV k -> n_ssm+k (mh.cpp:256, params.py:71-76) */
    const int *a, *d;            /* [D*T] row-major [datum][tp]; a = reference reads, d = total (README.md:21-41) */
    const double *mu_r, *mu_v;   /* [D] */
    const int *cnv_ptr;          /* [n_ssm+1] CSR: links of SSM j are cnv_ptr[j]..cnv_ptr[j+1]-1, in cnv-file order (util2.py:86-90) */
    const int *cnv_idx;          /* link -> CNV index 0..n_cnv-1 */
    const int *cnv_c1, *cnv_c2;  /* link -> tok[1], tok[2] of "sK,tok1,tok2" (util2.py:90) = mr_cnv[1], mr_cnv[2] */
} orc_data;

typedef struct {
    int K;                       /* number of tree nodes incl. the (empty) root */
    const int *parent;           /* [K] parent index, -1 for the root; any node order */
    const int *node_of_datum;    /* [D] node index holding each datum */
    const double *phi, *pi;      /* [K*T] row-major [node][tp]: node.params, node.pi */
} orc_tree;

enum { ORC_SEM_CPP = 0, ORC_SEM_PY = 1 };
enum { ORC_RNG_MT19937_SHIM = 0, ORC_RNG_PHILOX = 1 };

double orc_log_bin_coeff(int n, int k);
void   orc_lbc_table(const orc_data *D, double *out /*[D*T]*/);
double orc_logsumexp(const double *x, int n);

/* Philox4x32-10 block function (known-answer tested) and the derived uniforms. */
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_u01_from_u64(uint64_t x);

/* params.py:114-171 — integer (nr,nv) coefficients; coef layout [m][state 0..3][node k][0=nr,1=nv],
 * m enumerates CNV-affected SSMs in increasing SSM index; ssm_of_m (may be NULL) receives the SSM index.
 * Returns M (number of CNV-affected SSMs); call with coef==NULL to count. */
int    orc_data_states(const orc_data *D, const orc_tree *T, int *ssm_of_m, int *coef);

/* data.py:65-126 for SSM j placed at node s; returns 4 or 2 (number of states). */
int    orc_compute_n_genomes(const orc_data *D, const orc_tree *T, int j, int s, int tp,
                             const double *pi, double nr[4], double nv[4]);

/* mh.hpp:80-163 for one datum at one sample.  coef_m = this SSM's [4][K][2] block or NULL for plain data. */
double orc_datum_ll_cpp(const orc_data *D, int K, int ntps, int j, int tp, double phi,
                        const double *pi /*[K*T]*/, const int *coef_m, const double *lbc);
/* data.py:26-62 for datum j placed at node s with node parameters phi_s[T]. */
double orc_datum_ll_py(const orc_data *D, const orc_tree *T, int j, int s, const double *phi_s, const double *lbc);

/* mh.cpp:147-166 (semantics CPP: needs coef from orc_data_states) */
double orc_param_post(const orc_data *D, const orc_tree *T, const double *phi, const double *pi,
                      const int *ssm_of_m, const int *coef, const double *lbc, int tp);
double orc_multi_param_post(const orc_data *D, const orc_tree *T, const double *phi, const double *pi);

/* Rows of the datum x node grid: out[r*K + k] = llh of datum rows[r] if it sat in node k. */
void   orc_llh_rows(const orc_data *D, const orc_tree *T, int n_rows, const int *rows, int semantics, double *out);
/* The same for a subset of the nodes: out[r*n_cols + c] = llh of datum rows[r] if it sat in node cols[c]. */
void   orc_llh_block(const orc_data *D, const orc_tree *T, int n_rows, const int *rows, int n_cols, const int *cols,
                     int semantics, double *out);
/* Sum over data of llh(j, node_of_datum[j]) (tssb.py:404-410 data term / mh.cpp:147). */
double orc_total_llh(const orc_data *D, const orc_tree *T, int semantics);

/* mh.cpp:65-142.  Returns accepted count; phi_out/pi_out [K*T]; llh_out = multi_param_post of final state.
 * trace (may be NULL) receives per-iteration {a, accepted} pairs [iters*2]. */
int    orc_mh_loop(const orc_data *D, const orc_tree *T, int iters, double mh_std, int rng_kind,
                   uint64_t seed, uint64_t stream, int recompute_old,
                   double *phi_out, double *pi_out, double *llh_out, double *trace);

/* One proposal exactly as mh.cpp:113-142 + util.cpp:24-33 under the Philox stream (for kernel unit tests). */
void   orc_philox_proposal(int K, int ntps, const int *parent, const double *pi, double mh_std,
                           uint64_t seed, uint64_t stream, uint32_t iter, double *pi1, double *phi1);
double orc_philox_accept_uniform(uint64_t seed, uint64_t stream, uint32_t iter);
#endif
