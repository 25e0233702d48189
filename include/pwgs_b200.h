/* pwgs_b200.h — C ABI of libpwgs_b200.so, the B200 (sm_100a) replacement for the
 * PhyloWGS Metropolis-Hastings / SSM-likelihood hot path.
 *
 * The reference has no FFI for this path: it crosses a PROCESS boundary
 * (params.py:61 `subprocess.check_call([mh.o, MH_ITR, MH_STD, N_SSM, N_CNV, NNODES,
 * TREE_HEIGHT, ssm_file, cnv_file, c_tree, c_data_states, c_params, c_mh_ar, NTPS])`,
 * parsed at mh.cpp:22-61) plus three Python call sites (tssb.py:111,128; evolve.py:188,213).
 * Each entry point below names the reference interface it replaces.  All citations are
 * file:line relative to the reference checkout.
 *
 * Conventions
 *  - every function returns 0 on success or a negative PWGS_E* code; the message is
 *    available from pwgs_last_error() (thread-local).  Nothing throws or exits.
 *  - all pointers are HOST pointers to C-contiguous arrays owned by the caller; the
 *    library copies what it needs during the call and never retains host pointers.
 *  - int32 / float64 only.  D = n_ssm + n_cnv data; datum index = SSM k -> k,
 *    CNV k -> n_ssm + k (mh.cpp:256, params.py:71-76).  T = ntps samples, K = tree nodes.
 *  - a context is bound to one CUDA device and one chain; it is not thread-safe, but
 *    any number of contexts may coexist in a process (one per chain / per GPU).
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with
 *    PWGS_ECUDA.
 */
#ifndef PWGS_B200_H
#define PWGS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pwgs_ctx pwgs_ctx;

enum {
    PWGS_OK = 0,
    PWGS_EINVAL = -1,   /* bad argument (message says which) */
    PWGS_ECUDA = -2,    /* CUDA runtime error / no device */
    PWGS_ESTATE = -3,   /* call order (e.g. mh_run before set_tree) */
    PWGS_ELIMIT = -4    /* problem exceeds a documented limit (copy numbers > 32767, more than ~20 000 tree nodes) */
};

/* Likelihood conventions. */
enum {
    PWGS_SEM_MH = 0,    /* C++: mh.hpp:80-163 — 4 states, 1/4 weights, states substituted as params.py:160-166 */
    PWGS_SEM_PY = 1,    /* Python: data.py:47-126 — states with nv==0 dropped, rest weighted 1/len */
    PWGS_LAYOUT_COLS = 0x100  /* OR-ed into `semantics` of pwgs_llh_rows / pwgs_llh_block: out[k*n_rows + r] (one contiguous column per
                                 node, what the assignment sweep reads) instead of out[r*n_cols + k] */
};

const char *pwgs_last_error(void);
int pwgs_version(void);
/* Number of visible CUDA devices (0 and PWGS_ECUDA when there is none). */
int pwgs_device_count(int *n);

/* Replaces load_ssm_data / load_cnv_data (mh.cpp:208-305) and util2.load_data (util2.py:53-94):
 * uploads the read counts ONCE per chain instead of re-parsing ssm_data.txt / cnv_data.txt on every
 * mh.o invocation, and evaluates log_bin_coeff(d,a) (util.cpp:12-14) for every datum x sample on the
 * device.
 *   a, d      [D*T] row-major [datum][sample]   (columns a / d of the input files)
 *   mu_r,mu_v [D]   (CNV rows carry 0.999 / 0.5: mh.cpp:285-286, util2.py:84)
 *   cnv_ptr   [n_ssm+1] CSR offsets; links of SSM j in cnv-file order (util2.py:86-90)
 *   cnv_idx   [L] CNV index 0..n_cnv-1 ; cnv_c1/cnv_c2 [L] = tok[1], tok[2] of "sK,c1,c2"
 * cnv_* may be NULL when n_cnv == 0. */
int pwgs_create(int device, int n_ssm, int n_cnv, int ntps,
                const int32_t *a, const int32_t *d, const double *mu_r, const double *mu_v,
                const int32_t *cnv_ptr, const int32_t *cnv_idx, const int32_t *cnv_c1, const int32_t *cnv_c2,
                pwgs_ctx **out);
int pwgs_destroy(pwgs_ctx *ctx);

/* Replaces params.write_tree + params.write_data_state (params.py:68-171) and load_tree /
 * load_data_states (mh.cpp:309-463): hands the current tree to the device and builds the
 * integer (nr,nv) state coefficients there.
 *   parent        [K]  parent index, -1 for the single root; any node order
 *   node_of_datum [D]  node holding each datum
 *   phi, pi       [K*T] row-major [node][sample]  (node.params, node.pi) */
int pwgs_set_tree(pwgs_ctx *ctx, int K, const int32_t *parent, const int32_t *node_of_datum /* or NULL, see pwgs_move_data */,
                  const double *phi, const double *pi);

/* Tree edits in the middle of an assignment sweep (tssb.py:122-124 moves a datum, find_node tssb.py:353-361 spawns nodes) without
 * re-sending what did not change:
 *  - pwgs_set_tree with node_of_datum == NULL keeps the assignments the device already has (allowed when the new tree has at
 *    least as many nodes as the last one: spawned nodes are appended and hold no data yet);
 *  - pwgs_move_data re-assigns n data: datum[i] now sits in node[i].
 * Both only queue work on the context's stream (no host synchronisation). */
int pwgs_move_data(pwgs_ctx *ctx, int n, const int32_t *datum, const int32_t *node);

/* A page-locked host buffer of at least `bytes`, owned by the context (valid until the next call with a larger size or
 * pwgs_destroy).  Output arrays that live in it (the datum x node grid of a sweep, 45 MB at 50k SSMs x 110 nodes) are written by
 * DMA at PCIe speed; results written to ordinary pageable memory go through the driver's staging copy (4x slower). */
int pwgs_host_buffer(pwgs_ctx *ctx, uint64_t bytes, void **ptr);

/* Replaces `mh.o` (mh.cpp:22-109) as invoked by params.metropolis (params.py:30-65): `iters`
 * Metropolis-Hastings iterations over (pi, phi) for the tree of the last pwgs_set_tree, in ONE
 * persistent kernel.  mh_std is rounded to float like mh.hpp:28.  Random stream: Philox4x32-10 keyed
 * by (seed, stream) — the reference re-seeds MT19937 with GSL's default at every call (mh.cpp:66).
 *   phi_out, pi_out [K*T]  final node.params / node.pi (update_tree_params, params.py:183-194; full fp64,
 *                          not the 6 significant digits of write_params mh.cpp:191-204)
 *   accept_ratio           accepted / iters (c_mh_ar.txt, mh.cpp:104-107)
 *   llh_out                multi_param_post of the final state (may be NULL) */
int pwgs_mh_run(pwgs_ctx *ctx, int iters, double mh_std, uint64_t seed, uint64_t stream,
                double *phi_out, double *pi_out, double *accept_ratio, double *llh_out);

/* Same kernel with the state kept on the device (inputs already resident): launches asynchronously and
 * records CUDA events around it; pwgs_mh_wait blocks and returns what pwgs_mh_run returns. */
int pwgs_mh_launch(pwgs_ctx *ctx, int iters, double mh_std, uint64_t seed, uint64_t stream);
int pwgs_mh_wait(pwgs_ctx *ctx, double *phi_out, double *pi_out, double *accept_ratio, double *llh_out,
                 float *kernel_ms);

/* Several chains of ONE device in one cooperative launch (multievolve.py:85-105 runs its chains as independent processes, whose
 * kernels time-slice the GPU; small inputs leave most of it idle): chain i gets n_sms / n_chains CTAs and runs exactly what
 * pwgs_mh_launch would with "mh_ctas" set to that number -- same random streams, same result bits.  Every context then takes its
 * own pwgs_mh_wait.  mh_std, seed, stream: [n_chains]. */
int pwgs_mh_launch_multi(int n_chains, pwgs_ctx *const *ctxs, int iters, const double *mh_std, const uint64_t *seed,
                         const uint64_t *stream);

/* Replaces Datum._log_likelihood as called per candidate node by TSSB.resample_assignments
 * (tssb.py:111,128 -> alleles.py:52 -> data.py:26-62): out[r*K + k] = log-likelihood of datum rows[r]
 * if it sat in node k of the current tree (all K candidates at once). */
int pwgs_llh_rows(pwgs_ctx *ctx, int n_rows, const int32_t *rows, int semantics, double *out);
/* The same for a subset of candidate nodes: out[r*n_cols + c] for node cols[c] (cols == NULL: all K nodes).  After find_node
 * spawns nodes (tssb.py:353-361) only the new nodes' columns are missing: a spawn splits its parent's pi but leaves every
 * existing node's phi, and therefore every existing grid entry, unchanged. */
int pwgs_llh_block(pwgs_ctx *ctx, int n_rows, const int32_t *rows, int n_cols, const int32_t *cols, int semantics, double *out);
/* The same for the data row0 .. row0 + n_rows - 1 (the data an assignment sweep has still to visit): no row list to send. */
int pwgs_llh_range(pwgs_ctx *ctx, int row0, int n_rows, int n_cols, const int32_t *cols, int semantics, double *out);
/* The same, column-major (PWGS_LAYOUT_COLS required), written in place into a wider array: node cols[c] of datum row0 + r lands at
 * out[c * out_pitch + r] (out_pitch >= n_rows, in doubles) -- the columns of newly spawned nodes go straight into the sweep's
 * [node][datum] grid with one strided DMA. */
int pwgs_llh_range_pitched(pwgs_ctx *ctx, int row0, int n_rows, int n_cols, const int32_t *cols, int semantics, double *out,
                           int64_t out_pitch);

/* semantics MH: multi_param_post of the current state (mh.cpp:147-166).
 * semantics PY: the data term of TSSB.complete_data_log_likelihood (tssb.py:404-410, alleles.py:55-56). */
int pwgs_total_llh(pwgs_ctx *ctx, int semantics, double *out);

/* Diagnostics mirroring intermediate files of the reference, used by the parity tests. */
int pwgs_get_lbc(pwgs_ctx *ctx, double *out /* [D*T] log_bin_coeff(d,a), mh.cpp:234,284 */);
/* c_data_states.txt (params.py:167-168): *M = number of CNV-affected SSMs; ssm_of_m [M]; coef [M][4][K][2] int32.
 * Pass NULL arrays to query M only. */
int pwgs_get_data_states(pwgs_ctx *ctx, int32_t *M, int32_t *ssm_of_m, int32_t *coef);

/* Tuning knobs: "mh_batch" (largest speculative batch of proposals per grid barrier: a power of two in 1..256 = default; the kernel
 * adapts the actual batch to the running acceptance rate), "mh_ctas" (CTAs per chain, 0 = all SMs) and "mh_collapse" (1: plain
 * data enter the MH likelihood through exact integer totals per (node, error-rate class, sample) instead of one term per datum;
 * off by default, available when the data have at most 16 distinct (mu_r, mu_v) pairs), "mh_big" (-1 = automatic: trees whose
 * node tables (about 18 K T doubles) do not fit in shared memory keep them in global memory; 1 / 0 force / forbid that), "grid_plain"
 * (1: pwgs_llh_range's column-major form walks the data in file order instead of plain data and CNV-affected SSMs in separate warps;
 * for A/B timing, the results are the same bits).  "mh_w_plain", "mh_w_cnv1", "mh_w_cnv2",
 * "mh_w_cnv3": relative cost of a warp round (32 data) of plain data / of CNV-affected SSMs with 1, 2, 3 merged states, used to
 * deal whole rounds to the chain's CTAs (defaults 100 / 183 / 429 / 675, fitted to measured per-CTA cycles).  Diagnostics with
 * "mh_profile" = 1: "mh_prof_0".."mh_prof_7" (CTA 0's cycles per phase), "mh_lik_min/max/sum" and "mh_lik_cta_<r>" (likelihood
 * cycles of the chain's CTAs), "mh_plan_<class>_<r>" (rounds of CTA r: class 0 = plain, in thirds; 1..3 = CNV by states). */
int pwgs_set_option(pwgs_ctx *ctx, const char *name, int64_t value);
int pwgs_get_option(pwgs_ctx *ctx, const char *name, int64_t *value);

/* The context's cudaStream_t (as void*), so a caller can order its own device work / CUDA events against the
 * library's launches (bench.py times with events on this stream).  Read-only "launches" option = number of this
 * library's kernels launched for the context so far. */
int pwgs_get_stream(pwgs_ctx *ctx, void **stream);

/* ---- host side of TSSB.resample_assignments (tssb.py:82-150 with find_node 340-377): the serial slice sampler over the unit
 * interval, native, reading the datum x node log-likelihood grid from host memory (one column per node, filled from
 * pwgs_llh_rows / pwgs_llh_block).  No device work happens inside; the callbacks let the caller mirror tree edits and fetch
 * the grid entries an edit makes necessary.
 *   nodes: preorder at entry (node 0 = root, parent[k] < k); spawned nodes get the next indices.  Children of node k and their
 *   sticks: child_idx / child_stick[child_ptr[k] .. child_ptr[k+1]) in stick order.
 *   uniforms: consumed strictly in this order - per datum one for the slice level (log u + llh), then per candidate one for
 *   the point of the interval; inside find_node, per spawned child: stick (only below depth 0; Beta(1,dp_gamma) by inversion
 *   1-u^(1/b)), the fraction of the parent's pi it takes (alleles.py:35), main (Beta(1, alpha_decay^(depth+1)*dp_alpha)).
 *   refill() is called when the buffer runs out; it refills `uniforms` in place and returns how many it wrote.
 *   on_spawn(user, n): called after find_node spawned nodes while datum n is being placed, before any column of a new node is
 *   read: the callee mirrors spawn_*[already mirrored .. n_spawned) and stores the new nodes' columns in cols[].
 *   on_cnv_move(user, n): CNV datum n moved (assignments[n] is already updated): refresh the rows of its SSMs that are still
 *   to be visited.
 *   With ctx != NULL and on_spawn == on_cnv_move == NULL the library refreshes the grid itself, with no trip through the
 *   caller's language per tree edit (a sweep at 50k SSMs has ~100 of them): a spawn event appends the new nodes to the tables
 *   below (child pi = frac * parent pi, parent pi -= child pi, child phi = child pi: alleles.py:35-37), sends the grown tree
 *   (pwgs_set_tree with NULL assignments) and fills the new nodes' columns for the data still to come (rows n.. of
 *   cols[k] = grid + k * n_data for k < cap_cols, library-owned memory beyond); a CNV move is one pwgs_move_data plus the rows of
 *   the CNV's SSMs still to come (ssm_ptr / ssm_idx, ascending) in all nodes.  The context must hold the sweep's tree, in the
 *   sweep's node order, and its assignments.  n_refreshes counts the device round trips.
 * Returns 0, or a negative code (PWGS_EINVAL; -3 uniforms exhausted; -4 node / spawn capacity; -6/-7 callback failure;
 * -8 device refresh failed, see pwgs_last_error). */
typedef struct pwgs_sweep {
    int32_t n_data, n_ssm;
    int32_t n_nodes, cap_nodes;
    const int32_t *parent;            /* [n_nodes] */
    const double *main;               /* [n_nodes] */
    const int32_t *child_ptr;         /* [n_nodes+1] */
    const int32_t *child_idx;         /* [n_nodes-1] */
    const double *child_stick;        /* [n_nodes-1] */
    int32_t *assignments;             /* [n_data] node index per datum, in/out */
    double **cols;                    /* [cap_nodes] cols[k][n] = llh of datum n in node k */
    double *uniforms;
    int32_t n_uniforms;
    double dp_alpha, dp_gamma, alpha_decay;
    int32_t min_depth, max_depth;
    int32_t cap_spawn, n_spawned;     /* log of spawned nodes, in spawn order: node index = n_nodes + position */
    int32_t *spawn_parent;
    double *spawn_stick, *spawn_main, *spawn_frac;
    void *user;
    int (*on_spawn)(void *user, int32_t n);
    int (*on_cnv_move)(void *user, int32_t n);
    int (*refill)(void *user);
    int32_t n_moved, n_shrunk, uniforms_used;   /* out */
    /* library-side refreshes (see above); all unused when the callbacks are given */
    pwgs_ctx *ctx;
    int32_t ntps, cap_cols;
    int32_t *tree_parent;             /* [cap_nodes] in/out: parent of every node incl. the spawned ones */
    double *phi, *pi;                 /* [cap_nodes][ntps] in/out */
    double *grid;                     /* [cap_cols][n_data], ideally inside pwgs_host_buffer (page-locked) */
    const int32_t *ssm_ptr, *ssm_idx; /* CSR over the CNV data: SSMs linked to CNV datum n_ssm + j, ascending */
    double *scratch;                  /* landing area for the row blocks of a CNV move (page-locked if possible), or NULL */
    int64_t scratch_len;              /* doubles */
    int32_t n_refreshes;              /* out */
} pwgs_sweep;
int pwgs_resample_assignments(pwgs_sweep *sweep);

/* Host code, no device: byte assembly of the data list of a tree pickle (evolve.py:233 pickles the whole TSSB with all its
 * Datum objects for trees.zip, util2.py:250-262; host/fastpickle.py keeps one constant byte fragment per datum).  Writes, for
 * i = 0..n-1:  frags[frag_off[i] .. frag_off[i+1])  |  reference ref_of[i]  |  tail.   refs is a table of 8-byte records
 * {length <= 7, opcode bytes}; ref_of[i] < 0 writes the fragment alone.  Returns the number of bytes written (<= cap), or a
 * negative PWGS_E* code (PWGS_ELIMIT: cap too small). */
int64_t pwgs_pickle_join(int32_t n, const uint8_t *frags, const int64_t *frag_off, const int32_t *ref_of, int32_t n_refs,
                         const uint8_t *refs, const uint8_t *tail, int32_t tail_len, uint8_t *out, int64_t cap);

/* Diagnostics: the MH kernel's branch-free fp64 routines on caller-supplied arguments x[n] (positive, normal):
 * out[0..n) = log(x), out[n..2n) = exp(-x), out[2n..3n) = x / (1 + x).  Used by the accuracy test. */
int pwgs_math_selftest(int device, int n, const double *x, double *out);

/* Measured FP64-pipe issue peak of the device (DFMA chains, thread-instructions/s) — the roofline
 * denominator for the MH kernel, which is FP64-pipe-bound, not HBM-bound. */
int pwgs_fp64_peak(int device, double *dfma_per_s);

/* Diagnostics, host code only (no device needed): the MH kernel's work partition for a chain of `cpc` CTAs.  M CNV-affected SSMs
 * sorted by merged states (n_states = how many have 3, 2, 1), Dp plain data, ncls proposal-group classes (1..6), w = relative cost of
 * a warp round {plain, cnv1, cnv2, cnv3}.  Out: round_slot [ceil(M/32)] first slot of each CNV round in the CTA-major list,
 * cta_nc [cpc] CNV-affected SSMs per CTA, pstart [6][cpc+1] plain range boundaries per class, the two slice strides.  Used by the
 * CPU test-suite to check that every datum is walked exactly once per proposal whatever the shape. */
/* Diagnostics, host code only (no device needed): the grid likelihood of pwgs_llh_rows -- the very function the kernel runs,
 * compiled for the host -- on caller-owned arrays (layouts as in pwgs_create / pwgs_set_tree): out[r*K + k].  The CPU test-suite
 * holds it against the oracle, so the device code's logic is checked where there is no GPU. */
int pwgs_selftest_llh_host(int n_ssm, int n_cnv, int ntps, const int32_t *a, const int32_t *d, const double *mu_r, const double *mu_v,
                           const int32_t *cnv_ptr, const int32_t *cnv_idx, const int32_t *cnv_c1, const int32_t *cnv_c2,
                           int K, const int32_t *parent, const int32_t *node_of_datum, const double *phi, const double *pi,
                           int n_rows, const int32_t *rows, int semantics, double *out);

int pwgs_plan_partition(int cpc, int M, const int32_t *n_states, int Dp, int ncls, const int32_t *w,
                        int32_t *round_slot, int32_t *cta_nc, int32_t *pstart, int32_t *cchunk, int32_t *pchunk);

#ifdef __cplusplus
}
#endif
#endif
