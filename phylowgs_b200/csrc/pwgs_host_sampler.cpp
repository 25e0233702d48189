// pwgs_host_sampler.cpp — the sequential part of TSSB.resample_assignments (tssb.py:82-150, find_node 340-377) as native host
// code.  The slice sampler over the unit interval is inherently serial and branchy (every move changes node occupancy, and
// find_node may break new sticks), so it stays on the host, as north_star prescribes; what it needs from the device is the
// datum x node log-likelihood grid, which it reads from host memory (columns filled by pwgs_llh_rows / pwgs_llh_block).
// Python keeps ownership of the tree: spawned nodes and moves are reported back and mirrored there.
//
// Random numbers: ONE stream of uniforms, consumed in a fixed order (documented at pwgs_resample_assignments in the header);
// Beta(1, b) draws are taken by inversion, 1 - u^(1/b), so that the Python implementation of the same sweep
// (phylowgs_b200/host/sampler.py) consumes the identical stream and the two are bit-for-bit interchangeable.
#include "../../include/pwgs_b200.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <memory>
#include <vector>

namespace {

struct Tree {
    std::vector<int> parent, index_in_parent, depth;
    std::vector<double> main;
    std::vector<std::vector<int>> kids;
    std::vector<std::vector<double>> sticks;
    // right edges of the children's intervals, 1 - prod(1 - stick), kept as the running product tssb.py:349-352 recomputes at
    // every visit (same operations in the same order, so the same bits)
    std::vector<std::vector<double>> edges;
    std::vector<double> rem;
};

struct Sampler {
    pwgs_sweep *s;
    Tree t;
    int u_pos = 0;
    int err = 0;

    double uniform()
    {
        if (u_pos >= s->n_uniforms) {
            int n = s->refill ? s->refill(s->user) : 0;
            if (n <= 0) { err = -3; return 0.5; }
            s->n_uniforms = n;
            u_pos = 0;
        }
        return s->uniforms[u_pos++];
    }
    // util.py:29-30 with the Beta(1, b) variate taken by inversion
    double boundbeta1(double b) { return (1.0 - DBL_EPSILON) * ((1.0 - std::pow(uniform(), 1.0 / b)) - 0.5) + 0.5; }

    void spawn_child(int node)
    {
        const int depth = t.depth[node];
        if ((int)t.parent.size() >= s->cap_nodes || s->n_spawned >= s->cap_spawn) { err = -4; return; }
        const double stick = depth != 0 ? boundbeta1(s->dp_gamma) : 0.999;                        // tssb.py:355
        const double frac = uniform();                                                             // alleles.py:35
        const double mn = s->min_depth <= depth + 1 ? boundbeta1(std::pow(s->alpha_decay, depth + 1) * s->dp_alpha) : 0.0;
        const int id = (int)t.parent.size();
        t.parent.push_back(node);
        t.index_in_parent.push_back((int)t.kids[node].size());
        t.depth.push_back(depth + 1);
        t.main.push_back(mn);
        t.kids.emplace_back();
        t.sticks.emplace_back();
        t.edges.emplace_back();
        t.rem.push_back(1.0);
        t.kids[node].push_back(id);
        t.sticks[node].push_back(stick);
        t.rem[node] *= 1.0 - stick;
        t.edges[node].push_back(1.0 - t.rem[node]);
        const int k = s->n_spawned++;
        s->spawn_parent[k] = node; s->spawn_stick[k] = stick; s->spawn_main[k] = mn; s->spawn_frac[k] = frac;
    }

    // tssb.py:340-377; returns the node, fills `path` with the child indices from the root
    int find_node(double u, std::vector<int> &path)
    {
        int node = 0, depth = 0;
        path.clear();
        for (;;) {
            const double mn = t.main[node];
            if (depth >= s->max_depth || u < mn) return node;
            u = (u - mn) / (1.0 - mn);
            int index = 0;
            if (depth > 0) {
                while (t.edges[node].empty() || t.edges[node].back() < u) {
                    spawn_child(node);
                    if (err) return node;
                }
                const std::vector<double> &edges = t.edges[node];
                while (edges[index] < u) index++;
                const double left = index ? edges[index - 1] : 0.0;
                u = (u - left) / (edges[index] - left);
            }
            if (t.kids[node].empty()) { err = -5; return node; }      // depth 0 always has its first child (evolve.py:84-98)
            path.push_back(index);
            node = t.kids[node][index];
            depth++;
        }
    }

    void path_of(int node, std::vector<int> &path)
    {
        path.clear();
        for (int v = node; t.parent[v] >= 0; v = t.parent[v]) path.push_back(t.index_in_parent[v]);
        for (size_t i = 0, j = path.size(); i + 1 < j; i++, j--) std::swap(path[i], path[j - 1]);
    }

    // tssb.py:84-94 (lexicographic order of the index tuples; see sampler.py::_path_cmp)
    static int path_cmp(const std::vector<int> &cur, const std::vector<int> &cand)
    {
        if (cur.empty() && cand.empty()) return 0;
        if (cur.empty()) return 1;
        if (cand.empty()) return -1;
        const size_t n = cur.size() < cand.size() ? cur.size() : cand.size();
        for (size_t i = 0; i < n; i++)
            if (cand[i] != cur[i]) return cand[i] > cur[i] ? 1 : -1;
        return cand.size() > cur.size() ? 1 : (cand.size() < cur.size() ? -1 : 0);
    }

    // ---- library-side refreshes of the grid (s->ctx given, no callbacks)
    std::vector<std::unique_ptr<double[]>> own_cols;                  // columns beyond the caller's grid buffer
    std::vector<double> own_scratch;
    std::vector<int32_t> idx_buf;
    int n_known = 0;                                                  // nodes whose columns exist

    // find_node spawned nodes n_known .. size-1 while datum n was being placed
    int spawn_refresh(int n)
    {
        const int K = (int)t.parent.size(), T = s->ntps, first = n_known, m = K - first;
        for (int k = first; k < K; k++) {
            const int pk = t.parent[k];
            const double frac = s->spawn_frac[k - s->n_nodes];
            s->tree_parent[k] = pk;
            for (int tp = 0; tp < T; tp++) {                          // alleles.py:35-37
                const double child = frac * s->pi[(size_t)pk * T + tp];
                s->pi[(size_t)k * T + tp] = child;
                s->pi[(size_t)pk * T + tp] = s->pi[(size_t)pk * T + tp] - child;
                s->phi[(size_t)k * T + tp] = child;
            }
        }
        if (pwgs_set_tree(s->ctx, K, s->tree_parent, nullptr, s->phi, s->pi) != PWGS_OK) return -8;
        double *dst;
        if (K <= s->cap_cols) dst = s->grid + (size_t)first * s->n_data;
        else {                                                        // one block for the event: its columns stay adjacent
            own_cols.emplace_back(new double[(size_t)m * s->n_data]);
            dst = own_cols.back().get();
        }
        for (int i = 0; i < m; i++) s->cols[first + i] = dst + (size_t)i * s->n_data;
        idx_buf.resize(m);
        for (int i = 0; i < m; i++) idx_buf[i] = first + i;
        const int rows = s->n_data - n;
        if (pwgs_llh_range_pitched(s->ctx, n, rows, m, idx_buf.data(), PWGS_SEM_PY | PWGS_LAYOUT_COLS, dst + n, s->n_data) != PWGS_OK) return -8;
        n_known = K;
        s->n_refreshes++;
        return 0;
    }

    // CNV datum n moved to node k: the rows of its SSMs that are still to be visited change in every node
    int cnv_refresh(int n, int k)
    {
        const int32_t dn = n, dk = k;
        if (pwgs_move_data(s->ctx, 1, &dn, &dk) != PWGS_OK) return -8;
        const int j = n - s->n_ssm;
        const int32_t *lo = s->ssm_idx + s->ssm_ptr[j], *hi = s->ssm_idx + s->ssm_ptr[j + 1];
        while (lo < hi && *lo <= n) lo++;                             // ascending: skip the rows already visited
        const int nr = (int)(hi - lo), K = (int)t.parent.size();
        if (nr == 0) return 0;
        double *block = s->scratch;
        if ((int64_t)nr * K > s->scratch_len || !block) {
            own_scratch.resize((size_t)nr * K);
            block = own_scratch.data();
        }
        if (pwgs_llh_block(s->ctx, nr, lo, 0, nullptr, PWGS_SEM_PY | PWGS_LAYOUT_COLS, block) != PWGS_OK) return -8;
        for (int c = 0; c < K; c++) {
            double *col = s->cols[c];
            const double *src = block + (size_t)c * nr;
            for (int i = 0; i < nr; i++) col[lo[i]] = src[i];
        }
        s->n_refreshes++;
        return 0;
    }

    int run()
    {
        std::vector<int> cur_path, cand_path;
        int mirrored = s->n_spawned;
        const bool native = s->on_spawn == nullptr;
        n_known = (int)t.parent.size();
        for (int n = 0; n < s->n_data; n++) {
            const int cur = s->assignments[n];
            path_of(cur, cur_path);
            double lo = 0.0, hi = 1.0;
            const double llh_s = std::log(uniform()) + s->cols[cur][n];
            for (;;) {
                const double u = (hi - lo) * uniform() + lo;
                int cand = find_node(u, cand_path);
                if (err) return err;
                if (t.parent[cand] < 0) { cand = t.kids[cand][0]; cand_path.assign(1, 0); }       // tssb.py:118-120
                if (s->n_spawned != mirrored) {
                    if (native) { if (int rc = spawn_refresh(n)) return rc; }
                    else if (s->on_spawn(s->user, n) != 0) return -6;                             // new columns are now in s->cols
                    mirrored = s->n_spawned;
                }
                if (s->cols[cand][n] > llh_s) {
                    if (cand != cur) {
                        s->assignments[n] = cand;
                        s->n_moved++;
                        if (n >= s->n_ssm) {                                                      // rows of its SSMs still to come
                            if (native) { if (int rc = cnv_refresh(n, cand)) return rc; }
                            else if (s->on_cnv_move(s->user, n) != 0) return -7;
                        }
                    }
                    break;
                }
                if (std::fabs(hi - lo) < DBL_EPSILON) { s->n_shrunk++; break; }                   // tssb.py:132-137
                if (path_cmp(cur_path, cand_path) < 0) lo = u; else hi = u;
            }
            if (err) return err;
        }
        s->uniforms_used = u_pos;
        return 0;
    }
};

}  // namespace

extern "C" int pwgs_resample_assignments(pwgs_sweep *s)
{
    if (!s || !s->assignments || !s->cols || !s->uniforms) return PWGS_EINVAL;
    if ((s->on_spawn == nullptr) != (s->on_cnv_move == nullptr)) return PWGS_EINVAL;
    if (!s->on_spawn) {                                                                           // the library refreshes the grid itself
        if (!s->ctx || !s->tree_parent || !s->phi || !s->pi || s->ntps <= 0 || (s->cap_cols > 0 && !s->grid)) return PWGS_EINVAL;
        if (s->n_data > s->n_ssm && (!s->ssm_ptr || !s->ssm_idx)) return PWGS_EINVAL;
        s->n_refreshes = 0;
    }
    if (s->n_nodes <= 0 || s->n_nodes > s->cap_nodes) return PWGS_EINVAL;
    Sampler S;
    S.s = s;
    Tree &t = S.t;
    const int K = s->n_nodes;
    t.parent.assign(s->parent, s->parent + K);
    t.main.assign(s->main, s->main + K);
    t.index_in_parent.assign(K, 0);
    t.depth.assign(K, 0);
    t.kids.assign(K, {});
    t.sticks.assign(K, {});
    t.edges.assign(K, {});
    t.rem.assign(K, 1.0);
    for (int k = 0; k < K; k++) {
        for (int c = s->child_ptr[k]; c < s->child_ptr[k + 1]; c++) {
            const int ch = s->child_idx[c];
            if (ch <= 0 || ch >= K) return PWGS_EINVAL;
            t.index_in_parent[ch] = (int)t.kids[k].size();
            t.kids[k].push_back(ch);
            t.sticks[k].push_back(s->child_stick[c]);
            t.rem[k] *= 1.0 - s->child_stick[c];
            t.edges[k].push_back(1.0 - t.rem[k]);
        }
    }
    if (t.parent[0] >= 0) return PWGS_EINVAL;                                                     // node 0 is the root
    for (int k = 1; k < K; k++) {
        if (t.parent[k] < 0 || t.parent[k] >= k) return PWGS_EINVAL;                              // preorder: parents come first
        t.depth[k] = t.depth[t.parent[k]] + 1;
    }
    for (int n = 0; n < s->n_data; n++) if (s->assignments[n] <= 0 || s->assignments[n] >= K) return PWGS_EINVAL;
    s->n_moved = s->n_shrunk = 0;
    return S.run();
}
