// ref_harness.cpp — TEST INFRASTRUCTURE (oracle/_ref build only).
// A thin extern "C" shell around the UNMODIFIED reference sources
// /root/reference/mh.cpp + util.cpp (compiled where they lie, with -Dmain=ref_main),
// so tests and the fixture generator can call the reference's own loaders and
// likelihood from Python/ctypes.  Nothing here restates the algorithm: every
// number comes from the reference's functions (mh.cpp:65-463, mh.hpp:55-164).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "mh.hpp"   // from -I/root/reference (never copied into this repo)

struct ref_handle {
    struct config conf;
    struct datum *data;
    struct node *nodes;
};

extern "C" {

ref_handle *ref_load(const char *ssm_file, const char *cnv_file, const char *tree_file,
                     const char *states_file, int mh_itr, double mh_std,
                     int n_ssm, int n_cnv, int nnodes, int ntps)
{
    // Mirrors the call order of the reference main() (mh.cpp:22-61).
    ref_handle *h = new ref_handle;
    h->conf.MH_ITR = mh_itr;
    h->conf.MH_STD = (float)mh_std;
    h->conf.N_SSM_DATA = n_ssm;
    h->conf.N_CNV_DATA = n_cnv;
    h->conf.NNODES = nnodes;
    h->conf.TREE_HEIGHT = 0;
    h->conf.NTPS = ntps;
    h->data = new datum[n_ssm + n_cnv];
    load_ssm_data((char *)ssm_file, h->data, h->conf);
    if (n_cnv > 0) load_cnv_data((char *)cnv_file, h->data, h->conf);
    h->nodes = new node[nnodes];
    load_tree((char *)tree_file, h->nodes, h->conf);
    load_data_states((char *)states_file, h->data, h->nodes, h->conf);
    return h;
}

void ref_free(ref_handle *h) { delete[] h->data; delete[] h->nodes; delete h; }

double ref_multi_param_post(ref_handle *h, int old) { return multi_param_post(h->nodes, h->data, old, h->conf); }
double ref_param_post(ref_handle *h, int old, int tp) { return param_post(h->nodes, h->data, old, h->conf, tp); }
double ref_datum_ll(ref_handle *h, int datum_idx, double phi, int old, int tp) { return h->data[datum_idx].log_ll(phi, old, tp); }
double ref_log_bin_coeff(int n, int k) { return log_bin_coeff(n, k); }
double ref_logsumexp(double *x, int n) { return logsumexp(x, n); }
int ref_datum_is_cnv_affected(ref_handle *h, int datum_idx) { return h->data[datum_idx].cnv != NULL; }

// Runs the reference mh_loop (mh.cpp:65-109) on the loaded state; the acceptance
// ratio is written by the reference to ar_file.  Final phi/pi come back row-major [node][tp]
// in the file (post-)order the tree was loaded in; node ids in ids_out.
void ref_mh_loop(ref_handle *h, const char *ar_file) { mh_loop(h->nodes, h->data, (char *)ar_file, h->conf); }
void ref_get_params(ref_handle *h, int *ids_out, double *phi_out, double *pi_out)
{
    for (int i = 0; i < h->conf.NNODES; i++) {
        ids_out[i] = h->nodes[i].id;
        for (int tp = 0; tp < h->conf.NTPS; tp++) {
            phi_out[i * h->conf.NTPS + tp] = h->nodes[i].param[tp];
            pi_out[i * h->conf.NTPS + tp] = h->nodes[i].pi[tp];
        }
    }
}
// One proposal draw exactly as the reference does it (mh.cpp:113-142), for RNG-replay tests.
void ref_write_params(ref_handle *h, const char *fname) { write_params((char *)fname, h->nodes, h->conf); }

} // extern "C"
