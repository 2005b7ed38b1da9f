// build_model.cpp — host-side model preparation for libcsmc:
//   * per-basis interaction "perspectives" of the reference's Lattice constructor
//     (src/lattice.jl:176-283) in closed form, O(N * terms) instead of O(N^2 * terms);
//   * colouring of the interaction hypergraph (periodic pattern search, greedy fallback);
//   * colour-major / class-major storage order, explicit neighbour tables, and the descriptors
//     the kernels receive as by-value parameters.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <tuple>

#include "csmc_internal.h"

namespace csmc {

namespace {

inline int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
inline int posmod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

inline int64_t site_index(const HostModel &hm, int b, const int *i) {
    int64_t p = b;
    for (int d = 0; d < hm.D; ++d) p = p * hm.L[d] + i[d];
    return p;
}
inline void site_coords(const HostModel &hm, int64_t p, int &b, int *i) {
    for (int d = hm.D - 1; d >= 0; --d) { i[d] = (int)(p % hm.L[d]); p /= hm.L[d]; }
    for (int d = hm.D; d < MAXD; ++d) i[d] = 0;
    b = (int)p;
}
// neighbour cell under the boundary condition (src/lattice.jl:101-109); false == missing (open)
inline bool neighbour_cell(const HostModel &hm, const int *i, const int *off, int *out) {
    for (int d = 0; d < hm.D; ++d) {
        int v = i[d] + off[d];
        if (hm.periodic) v = posmod(v, hm.L[d]);
        else if (v < 0 || v >= hm.L[d]) return false;
        out[d] = v;
    }
    for (int d = hm.D; d < MAXD; ++d) out[d] = 0;
    return true;
}

// ---- perspectives -----------------------------------------------------------------------------
void build_basis_terms(const csmc_model *m, HostModel &hm) {
    const int D = hm.D;
    hm.basis_terms.assign(hm.n_basis, {});
    hm.coefs.clear();
    std::map<std::tuple<int, int, int>, int> coef_of;  // (kind, term, perspective) -> offset

    auto bil_coef = [&](int t, int persp) {
        auto key = std::make_tuple(2, t, persp);
        auto it = coef_of.find(key);
        if (it != coef_of.end()) return it->second;
        int off = (int)hm.coefs.size();
        const double *J = m->bil_matrix + 9 * t;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) hm.coefs.push_back(persp == 0 ? J[3 * r + c] : J[3 * c + r]);
        coef_of[key] = off;
        return off;
    };
    auto cub_coef = [&](int t, int persp) {
        auto key = std::make_tuple(3, t, persp);
        auto it = coef_of.find(key);
        if (it != coef_of.end()) return it->second;
        int off = (int)hm.coefs.size();
        const double *C = m->cub_tensor + 27 * t;  // column-major: [a,b,c] at a + 3b + 9c
        auto in = [&](int a, int b, int c) { return C[a + 3 * b + 9 * c]; };
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    hm.coefs.push_back(persp == 0 ? in(a, b, c) : persp == 1 ? in(b, a, c) : in(c, b, a));
        coef_of[key] = off;
        return off;
    };
    auto quar_coef = [&](int t, int persp) {
        auto key = std::make_tuple(4, t, persp);
        auto it = coef_of.find(key);
        if (it != coef_of.end()) return it->second;
        int off = (int)hm.coefs.size();
        const double *R = m->quar_tensor + 81 * t;
        auto in = [&](int a, int b, int c, int d) { return R[a + 3 * b + 9 * c + 27 * d]; };
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    for (int d = 0; d < 3; ++d)
                        hm.coefs.push_back(persp == 0   ? in(a, b, c, d)
                                           : persp == 1 ? in(b, a, c, d)
                                           : persp == 2 ? in(c, b, a, d)
                                                        : in(d, b, c, a));
        coef_of[key] = off;
        return off;
    };

    for (int b0 = 0; b0 < hm.n_basis; ++b0) {
        const int b = b0 + 1;
        auto &terms = hm.basis_terms[b0];
        for (int t = 0; t < hm.N2; ++t) {  // src/lattice.jl:176-204
            const int b1 = m->bil_basis[2 * t], b2 = m->bil_basis[2 * t + 1];
            const int *off = m->bil_offset + D * t;
            if (b != b1 && b != b2) continue;
            HostTerm h{};
            h.kind = 2; h.row = t;
            const bool fwd = (b1 == b2) || (b1 == b);
            h.coef = bil_coef(t, fwd ? 0 : 1);
            h.nb_basis[0] = (fwd ? b2 : b1) - 1;
            for (int d = 0; d < D; ++d) h.off[0][d] = fwd ? off[d] : -off[d];
            terms.push_back(h);
        }
        for (int t = 0; t < hm.N3; ++t) {  // src/lattice.jl:209-237
            const int *bb = m->cub_basis + 3 * t;
            const int *oj = m->cub_offset + 2 * D * t, *ok = oj + D;
            if (b != bb[0] && b != bb[1] && b != bb[2]) continue;
            HostTerm h{};
            h.kind = 3; h.row = hm.N2 + 2 * t;
            if (bb[0] == b) {
                h.coef = cub_coef(t, 0); h.nb_basis[0] = bb[1] - 1; h.nb_basis[1] = bb[2] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = oj[d]; h.off[1][d] = ok[d]; }
            } else if (bb[1] == b) {
                h.coef = cub_coef(t, 1); h.nb_basis[0] = bb[0] - 1; h.nb_basis[1] = bb[2] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = -oj[d]; h.off[1][d] = ok[d] - oj[d]; }
            } else {
                h.coef = cub_coef(t, 2); h.nb_basis[0] = bb[1] - 1; h.nb_basis[1] = bb[0] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = oj[d] - ok[d]; h.off[1][d] = -ok[d]; }
            }
            terms.push_back(h);
        }
        for (int t = 0; t < hm.N4; ++t) {  // src/lattice.jl:243-283
            const int *bb = m->quar_basis + 4 * t;
            const int *oj = m->quar_offset + 3 * D * t, *ok = oj + D, *ol = ok + D;
            if (b != bb[0] && b != bb[1] && b != bb[2] && b != bb[3]) continue;
            HostTerm h{};
            h.kind = 4; h.row = hm.N2 + 2 * hm.N3 + 3 * t;
            if (bb[0] == b) {
                h.coef = quar_coef(t, 0);
                h.nb_basis[0] = bb[1] - 1; h.nb_basis[1] = bb[2] - 1; h.nb_basis[2] = bb[3] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = oj[d]; h.off[1][d] = ok[d]; h.off[2][d] = ol[d]; }
            } else if (bb[1] == b) {
                h.coef = quar_coef(t, 1);
                h.nb_basis[0] = bb[0] - 1; h.nb_basis[1] = bb[2] - 1; h.nb_basis[2] = bb[3] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = -oj[d]; h.off[1][d] = ok[d] - oj[d]; h.off[2][d] = ol[d] - oj[d]; }
            } else if (bb[2] == b) {
                h.coef = quar_coef(t, 2);
                h.nb_basis[0] = bb[1] - 1; h.nb_basis[1] = bb[0] - 1; h.nb_basis[2] = bb[3] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = oj[d] - ok[d]; h.off[1][d] = -ok[d]; h.off[2][d] = ol[d] - ok[d]; }
            } else {
                h.coef = quar_coef(t, 3);
                h.nb_basis[0] = bb[1] - 1; h.nb_basis[1] = bb[2] - 1; h.nb_basis[2] = bb[0] - 1;
                for (int d = 0; d < D; ++d) { h.off[0][d] = oj[d] - ol[d]; h.off[1][d] = ok[d] - ol[d]; h.off[2][d] = -ol[d]; }
            }
            terms.push_back(h);
        }
    }
    hm.onsite_coef.assign(hm.n_basis, -1);
    for (int b0 = 0; b0 < hm.n_basis; ++b0) {
        bool nz = false;
        for (int k = 0; k < 9; ++k) nz |= (hm.onsite[9 * b0 + k] != 0.0);
        if (!nz) continue;
        hm.onsite_coef[b0] = (int)hm.coefs.size();
        for (int k = 0; k < 9; ++k) hm.coefs.push_back(hm.onsite[9 * b0 + k]);
    }
}

// ---- conflicts ----------------------------------------------------------------------------------
struct Conflict { int ba, bc; int delta[MAXD]; };
inline bool operator<(const Conflict &x, const Conflict &y) {
    return std::tie(x.ba, x.bc, x.delta[0], x.delta[1], x.delta[2]) <
           std::tie(y.ba, y.bc, y.delta[0], y.delta[1], y.delta[2]);
}

std::vector<Conflict> build_conflicts(const HostModel &hm) {
    std::set<Conflict> s;
    auto add = [&](int ba, const int *oa, int bc, const int *oc) {
        Conflict c{}; c.ba = ba; c.bc = bc;
        for (int d = 0; d < hm.D; ++d) c.delta[d] = oc[d] - oa[d];
        s.insert(c);
        Conflict r{}; r.ba = bc; r.bc = ba;
        for (int d = 0; d < hm.D; ++d) r.delta[d] = -c.delta[d];
        s.insert(r);
    };
    const int zero[MAXD] = {0, 0, 0};
    for (int b = 0; b < hm.n_basis; ++b)
        for (const auto &t : hm.basis_terms[b]) {
            const int nn = t.kind - 1;
            for (int j = 0; j < nn; ++j) {
                add(b, zero, t.nb_basis[j], t.off[j]);
                for (int k = j + 1; k < nn; ++k) add(t.nb_basis[j], t.off[j], t.nb_basis[k], t.off[k]);
            }
        }
    return std::vector<Conflict>(s.begin(), s.end());
}

// is (b, i) and (b, i + delta) the same site on this lattice?
inline bool same_site_delta(const HostModel &hm, const Conflict &c) {
    if (c.ba != c.bc) return false;
    for (int d = 0; d < hm.D; ++d) {
        if (hm.periodic) { if (posmod(c.delta[d], hm.L[d]) != 0) return false; }
        else if (c.delta[d] != 0) return false;
    }
    return true;
}

// ---- small-graph colouring (DSATUR greedy + bounded exact improvement) ----------------------------
struct Graph { int n; std::vector<std::vector<int>> adj; };

int greedy_dsatur(const Graph &g, std::vector<int> &col) {
    col.assign(g.n, -1);
    int ncol = 0;
    for (int it = 0; it < g.n; ++it) {
        int best = -1, best_sat = -1, best_deg = -1;
        for (int v = 0; v < g.n; ++v) {
            if (col[v] >= 0) continue;
            std::set<int> seen;
            for (int u : g.adj[v]) if (col[u] >= 0) seen.insert(col[u]);
            int sat = (int)seen.size(), deg = (int)g.adj[v].size();
            if (sat > best_sat || (sat == best_sat && deg > best_deg)) { best = v; best_sat = sat; best_deg = deg; }
        }
        std::vector<char> used(ncol + 1, 0);
        for (int u : g.adj[best]) if (col[u] >= 0) used[col[u]] = 1;
        int c = 0;
        while (used[c]) ++c;
        col[best] = c;
        ncol = std::max(ncol, c + 1);
    }
    return ncol;
}

bool try_k(const Graph &g, int k, std::vector<int> &col, const std::vector<int> &order, int idx, long &budget) {
    if (idx == g.n) return true;
    if (--budget < 0) return false;
    const int v = order[idx];
    int maxc = -1;
    for (int i = 0; i < idx; ++i) maxc = std::max(maxc, col[order[i]]);
    for (int c = 0; c <= std::min(maxc + 1, k - 1); ++c) {  // symmetry breaking
        bool ok = true;
        for (int u : g.adj[v]) if (col[u] == c) { ok = false; break; }
        if (!ok) continue;
        col[v] = c;
        if (try_k(g, k, col, order, idx + 1, budget)) return true;
        col[v] = -1;
    }
    return false;
}

int colour_graph(const Graph &g, std::vector<int> &col) {
    int ub = greedy_dsatur(g, col);
    std::vector<int> order(g.n);
    for (int i = 0; i < g.n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return g.adj[a].size() > g.adj[b].size(); });
    while (ub > 1) {
        std::vector<int> trial(g.n, -1);
        long budget = 200000;
        if (!try_k(g, ub - 1, trial, order, 0, budget)) break;
        col = trial;
        int mx = 0;
        for (int c : col) mx = std::max(mx, c + 1);
        ub = mx;
    }
    return ub;
}

struct Pattern { bool ok = false; int P[MAXD] = {1, 1, 1}; int ncol = 0; int Q = 0; std::vector<int> class_colour; };

inline int class_index(const HostModel &hm, const int *P, int b, const int *r) {
    int q = b;
    for (int d = 0; d < hm.D; ++d) q = q * P[d] + r[d];
    return q;
}

Pattern search_pattern(const HostModel &hm, const std::vector<Conflict> &conf) {
    Pattern best;
    const int PMAX = 8, QMAX = 128;
    std::vector<int> cand[MAXD];
    for (int d = 0; d < MAXD; ++d) {
        if (d >= hm.D) { cand[d] = {1}; continue; }
        for (int p = 1; p <= std::min(PMAX, hm.L[d]); ++p)
            if (!hm.periodic || hm.L[d] % p == 0) cand[d].push_back(p);
    }
    for (int p0 : cand[0]) for (int p1 : cand[1]) for (int p2 : cand[2]) {
        const int P[MAXD] = {p0, p1, p2};
        int cells = 1;
        for (int d = 0; d < hm.D; ++d) cells *= P[d];
        const int Q = hm.n_basis * cells;
        if (Q > QMAX) continue;
        if (best.ok && best.ncol == 1) continue;
        Graph g; g.n = Q; g.adj.assign(Q, {});
        bool valid = true;
        std::vector<std::set<int>> adjs(Q);
        for (int b = 0; b < hm.n_basis && valid; ++b) {
            int r[MAXD] = {0, 0, 0};
            for (int c = 0; c < cells && valid; ++c) {
                int t = c;
                for (int d = hm.D - 1; d >= 0; --d) { r[d] = t % P[d]; t /= P[d]; }
                const int q = class_index(hm, P, b, r);
                for (const auto &cf : conf) {
                    if (cf.ba != b) continue;
                    if (same_site_delta(hm, cf)) continue;  // self-interaction, not a race
                    int r2[MAXD] = {0, 0, 0};
                    for (int d = 0; d < hm.D; ++d) r2[d] = posmod(r[d] + cf.delta[d], P[d]);
                    const int q2 = class_index(hm, P, cf.bc, r2);
                    if (q2 == q) { valid = false; break; }
                    adjs[q].insert(q2); adjs[q2].insert(q);
                }
            }
        }
        if (!valid) continue;
        for (int q = 0; q < Q; ++q) g.adj[q].assign(adjs[q].begin(), adjs[q].end());
        std::vector<int> col;
        const int ncol = colour_graph(g, col);
        // prefer: fewer colours, then fewer classes, then a longer fastest dimension
        auto better = [&]() {
            if (!best.ok) return true;
            if (ncol != best.ncol) return ncol < best.ncol;
            if (Q != best.Q) return Q < best.Q;
            return P[hm.D - 1] < best.P[hm.D - 1];
        };
        if (better()) {
            best.ok = true; best.ncol = ncol; best.Q = Q; best.class_colour = col;
            for (int d = 0; d < MAXD; ++d) best.P[d] = P[d];
        }
    }
    return best;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
std::string build_host_model(const csmc_model *m, int flags, HostModel &hm) {
    if (!m) return "model is NULL";
    if (m->dim < 1 || m->dim > MAXD) return "dim must be 1..3";
    if (m->n_basis < 1) return "n_basis must be >= 1";
    if (m->n_bilinear < 0 || m->n_cubic < 0 || m->n_quartic < 0) return "negative term count";
    hm.D = m->dim; hm.n_basis = m->n_basis; hm.periodic = m->periodic ? 1 : 0; hm.S = m->S;
    int64_t cells = 1;
    for (int d = 0; d < MAXD; ++d) {
        hm.L[d] = d < hm.D ? m->shape[d] : 1;
        if (hm.L[d] < 1) return "shape entries must be >= 1";
        cells *= hm.L[d];
    }
    hm.N = cells * hm.n_basis;
    if (hm.N > (int64_t)1 << 30) return "lattice too large (N must be < 2^30 sites)";
    hm.N2 = m->n_bilinear; hm.N3 = m->n_cubic; hm.N4 = m->n_quartic;
    hm.n_rows = hm.N2 + 2 * hm.N3 + 3 * hm.N4;
    if (hm.n_rows > 30000) return "too many interaction terms";
    if (!m->field || !m->onsite) return "field/onsite pointers are NULL";
    hm.field.assign(m->field, m->field + 3 * hm.n_basis);
    hm.onsite.assign(m->onsite, m->onsite + 9 * hm.n_basis);
    auto chk_basis = [&](const int32_t *b, int n) {
        for (int i = 0; i < n; ++i) if (b[i] < 1 || b[i] > hm.n_basis) return false;
        return true;
    };
    if (hm.N2 && !chk_basis(m->bil_basis, 2 * hm.N2)) return "bilinear basis index out of range";
    if (hm.N3 && !chk_basis(m->cub_basis, 3 * hm.N3)) return "cubic basis index out of range";
    if (hm.N4 && !chk_basis(m->quar_basis, 4 * hm.N4)) return "quartic basis index out of range";

    build_basis_terms(m, hm);
    const auto conf = build_conflicts(hm);
    for (const auto &c : conf) if (same_site_delta(hm, c)) hm.self_loop = true;

    const int D = hm.D;
    const int ALIGN = 16;
    Pattern pat = search_pattern(hm, conf);
    hm.segs.clear();
    hm.colour_of_site.assign(hm.N, 0);
    hm.pos_of_ref.assign(hm.N, -1);

    if (pat.ok) {
        hm.pattern = true;
        hm.n_colours = pat.ncol;
        for (int d = 0; d < MAXD; ++d) hm.P[d] = pat.P[d];
        int pcells = 1;
        for (int d = 0; d < D; ++d) pcells *= hm.P[d];
        std::vector<HostSeg> tmp;
        for (int b = 0; b < hm.n_basis; ++b)
            for (int c = 0; c < pcells; ++c) {
                HostSeg s{};
                int t = c;
                for (int d = 0; d < MAXD; ++d) { s.r[d] = 0; s.P[d] = hm.P[d]; s.M[d] = 1; }
                for (int d = D - 1; d >= 0; --d) { s.r[d] = t % hm.P[d]; t /= hm.P[d]; }
                s.basis = b;
                s.colour = pat.class_colour[class_index(hm, hm.P, b, s.r)];
                s.count = 1;
                for (int d = 0; d < D; ++d) {
                    s.M[d] = s.r[d] < hm.L[d] ? (hm.L[d] - s.r[d] + hm.P[d] - 1) / hm.P[d] : 0;
                    s.count *= s.M[d];
                }
                if (s.count > 0) tmp.push_back(s);
            }
        std::stable_sort(tmp.begin(), tmp.end(), [](const HostSeg &a, const HostSeg &b) { return a.colour < b.colour; });
        hm.segs = tmp;
    } else {
        // greedy per-site colouring on the explicit conflict graph
        hm.pattern = false;
        std::vector<int> col(hm.N, -1);
        int ncol = 0;
        std::vector<char> used;
        for (int64_t p = 0; p < hm.N; ++p) {
            int b, i[MAXD], j[MAXD];
            site_coords(hm, p, b, i);
            used.assign(ncol + 1, 0);
            for (const auto &c : conf) {
                if (c.ba != b || same_site_delta(hm, c)) continue;
                if (!neighbour_cell(hm, i, c.delta, j)) continue;
                int64_t q = site_index(hm, c.bc, j);
                if (q != p && col[q] >= 0) used[col[q]] = 1;
            }
            int cc = 0;
            while (used[cc]) ++cc;
            col[p] = cc;
            ncol = std::max(ncol, cc + 1);
        }
        hm.n_colours = ncol;
        std::map<std::pair<int, int>, int> seg_of;
        for (int c = 0; c < ncol; ++c)
            for (int b = 0; b < hm.n_basis; ++b) {
                HostSeg s{};
                s.colour = c; s.basis = b; s.count = 0;
                for (int d = 0; d < MAXD; ++d) { s.P[d] = 1; s.r[d] = 0; s.M[d] = 1; }
                seg_of[{c, b}] = (int)hm.segs.size();
                hm.segs.push_back(s);
            }
        for (int64_t p = 0; p < hm.N; ++p) {
            int b, i[MAXD];
            site_coords(hm, p, b, i);
            hm.segs[seg_of[{col[p], b}]].sites.push_back((int)p);
        }
        std::vector<HostSeg> keep;
        for (auto &s : hm.segs) if (!s.sites.empty()) { s.count = (int)s.sites.size(); s.M[0] = s.count; keep.push_back(s); }
        hm.segs = keep;
    }

    // storage positions
    int pos = 0;
    for (auto &s : hm.segs) {
        s.start = pos;
        pos += (s.count + ALIGN - 1) / ALIGN * ALIGN;
    }
    hm.npad = std::max(pos, ALIGN);
    hm.ref_of_pos.assign(hm.npad, -1);
    for (auto &s : hm.segs) {
        if (hm.pattern) {
            int mm[MAXD];
            for (int lin = 0; lin < s.count; ++lin) {
                int t = lin, i[MAXD] = {0, 0, 0};
                for (int d = D - 1; d >= 0; --d) { mm[d] = t % s.M[d]; t /= s.M[d]; }
                for (int d = 0; d < D; ++d) i[d] = mm[d] * s.P[d] + s.r[d];
                const int64_t p = site_index(hm, s.basis, i);
                hm.ref_of_pos[s.start + lin] = (int32_t)p;
                hm.pos_of_ref[p] = s.start + lin;
                hm.colour_of_site[p] = s.colour;
            }
        } else {
            for (int lin = 0; lin < s.count; ++lin) {
                const int p = s.sites[lin];
                hm.ref_of_pos[s.start + lin] = p;
                hm.pos_of_ref[p] = s.start + lin;
                hm.colour_of_site[p] = s.colour;
            }
        }
    }
    hm.colour_seg_begin.assign(hm.n_colours + 1, 0);
    for (const auto &s : hm.segs) hm.colour_seg_begin[s.colour + 1]++;
    for (int c = 0; c < hm.n_colours; ++c) hm.colour_seg_begin[c + 1] += hm.colour_seg_begin[c];

    // can the arithmetic-neighbour kernels be used?
    hm.structured = hm.pattern && !(flags & CSMC_FLAG_FORCE_GENERIC);
    if (hm.structured) {
        for (const auto &s : hm.segs)
            for (const auto &t : hm.basis_terms[s.basis])
                for (int k = 0; k < t.kind - 1; ++k)
                    for (int d = 0; d < D; ++d) {
                        const int delta = floordiv(s.r[d] + t.off[k][d], hm.P[d]);
                        if (std::abs(delta) > 120) hm.structured = false;
                        if (hm.periodic && std::abs(delta) > hm.L[d] / hm.P[d]) hm.structured = false;
                    }
    }

    // explicit neighbour table, storage order: only the explicit-table kernels read it (1 GB at L=8192)
    if (!hm.structured) {
    hm.nbr.assign((size_t)std::max(hm.n_rows, 1) * hm.npad, -1);
    for (int64_t p = 0; p < hm.N; ++p) {
        int b, i[MAXD], j[MAXD];
        site_coords(hm, p, b, i);
        const int pp = hm.pos_of_ref[p];
        for (const auto &t : hm.basis_terms[b]) {
            const int nn = t.kind - 1;
            int64_t q[3];
            bool all = true;
            for (int k = 0; k < nn; ++k) {
                if (!neighbour_cell(hm, i, t.off[k], j)) { all = false; break; }
                q[k] = site_index(hm, t.nb_basis[k], j);
                if (q[k] == p) hm.self_loop = true;
            }
            if (!all) continue;
            for (int k = 0; k < nn; ++k) hm.nbr[(size_t)(t.row + k) * hm.npad + pp] = hm.pos_of_ref[q[k]];
        }
    }
    }

    // parameter-struct size class
    hm.need_large = false;
    if ((int)hm.segs.size() > PassSmall::NG || (int)hm.coefs.size() > PassSmall::NC) hm.need_large = true;
    for (int c = 0; c < hm.n_colours; ++c) {
        int ns = hm.colour_seg_begin[c + 1] - hm.colour_seg_begin[c], nt = 0;
        for (int s = hm.colour_seg_begin[c]; s < hm.colour_seg_begin[c + 1]; ++s) nt += (int)hm.basis_terms[hm.segs[s].basis].size();
        if (ns > PassSmall::NS || nt > PassSmall::NT) hm.need_large = true;
        if (ns > PassLarge::NS || nt > PassLarge::NT)
            return "model too large for the kernel parameter block (segments/terms per colour)";
    }
    if ((int)hm.segs.size() > PassLarge::NG || (int)hm.coefs.size() > PassLarge::NC)
        return "model too large for the kernel parameter block (segments/coefficients)";
    return "";
}

template <class P>
std::string fill_pass_params(const HostModel &hm, int colour, P &p) {
    std::memset((void *)&p, 0, sizeof(P));
    if ((int)hm.segs.size() > P::NG || (int)hm.coefs.size() > P::NC) return "parameter block overflow";
    p.npad = hm.npad;
    p.rep_stride = 3LL * hm.npad;
    p.periodic = hm.periodic;
    p.S = hm.S;
    for (int d = 0; d < MAXD; ++d) p.L[d] = hm.L[d];
    for (size_t s = 0; s < hm.segs.size(); ++s) {
        p.geom[s].start = hm.segs[s].start;
        for (int d = 0; d < MAXD; ++d) p.geom[s].M[d] = hm.segs[s].M[d];
    }
    for (size_t k = 0; k < hm.coefs.size(); ++k) p.coefs[k] = hm.coefs[k];
    // class lookup for structured neighbours
    std::map<std::tuple<int, int, int, int>, int> seg_of_class;
    if (hm.pattern)
        for (size_t s = 0; s < hm.segs.size(); ++s)
            seg_of_class[std::make_tuple(hm.segs[s].basis, hm.segs[s].r[0], hm.segs[s].r[1], hm.segs[s].r[2])] = (int)s;
    const int s0 = hm.colour_seg_begin[colour], s1 = hm.colour_seg_begin[colour + 1];
    if (s1 - s0 > P::NS) return "parameter block overflow (segments)";
    p.n_segs = s1 - s0;
    int nt = 0;
    for (int s = s0; s < s1; ++s) {
        const HostSeg &hs = hm.segs[s];
        DevSeg &ds = p.segs[s - s0];
        ds.start = hs.start; ds.count = hs.count;
        for (int d = 0; d < MAXD; ++d) { ds.M[d] = hs.M[d]; ds.P[d] = (int8_t)hs.P[d]; ds.r[d] = (int8_t)hs.r[d]; }
        ds.term_begin = nt;
        ds.basis = (int16_t)hs.basis;
        ds.geom = (int16_t)s;
        ds.onsite = hm.onsite_coef[hs.basis];
        for (int k = 0; k < 3; ++k) ds.h[k] = hm.field[3 * hs.basis + k];
        int n2 = 0, n3 = 0, n4 = 0;
        for (const auto &t : hm.basis_terms[hs.basis]) {
            if (nt >= P::NT) return "parameter block overflow (terms)";
            DevTerm &dt = p.terms[nt++];
            dt.coef = t.coef; dt.row = (int16_t)t.row;
            for (int k = 0; k < 3; ++k) { dt.nseg[k] = -1; for (int d = 0; d < MAXD; ++d) dt.d[k][d] = 0; }
            if (hm.pattern)
                for (int k = 0; k < t.kind - 1; ++k) {
                    int r2[MAXD] = {0, 0, 0};
                    for (int d = 0; d < hm.D; ++d) {
                        r2[d] = posmod(hs.r[d] + t.off[k][d], hm.P[d]);
                        int delta = floordiv(hs.r[d] + t.off[k][d], hm.P[d]);
                        dt.d[k][d] = (int8_t)std::max(-127, std::min(127, delta));
                    }
                    auto it = seg_of_class.find(std::make_tuple(t.nb_basis[k], r2[0], r2[1], r2[2]));
                    dt.nseg[k] = it == seg_of_class.end() ? -1 : (int16_t)it->second;
                }
            if (t.kind == 2) ++n2; else if (t.kind == 3) ++n3; else ++n4;
        }
        ds.n2 = (int16_t)n2; ds.n3 = (int16_t)n3; ds.n4 = (int16_t)n4;
    }
    return "";
}

template std::string fill_pass_params<PassSmall>(const HostModel &, int, PassSmall &);
template std::string fill_pass_params<PassLarge>(const HostModel &, int, PassLarge &);

// reference-layout tables (lat.bilinear_sites / cubic_sites / quartic_sites): 1-based, 0 == null
void reference_tables(const csmc_model *m, int64_t *bil, int64_t *cub, int64_t *quar) {
    HostModel hm;
    hm.D = m->dim; hm.n_basis = m->n_basis; hm.periodic = m->periodic ? 1 : 0;
    int64_t cells = 1;
    for (int d = 0; d < MAXD; ++d) { hm.L[d] = d < hm.D ? m->shape[d] : 1; cells *= hm.L[d]; }
    hm.N = cells * hm.n_basis;
    hm.N2 = m->n_bilinear; hm.N3 = m->n_cubic; hm.N4 = m->n_quartic;
    hm.onsite.assign(9 * hm.n_basis, 0.0);
    build_basis_terms(m, hm);
    if (bil) std::fill(bil, bil + hm.N * hm.N2, 0);
    if (cub) std::fill(cub, cub + hm.N * hm.N3 * 2, 0);
    if (quar) std::fill(quar, quar + hm.N * hm.N4 * 3, 0);
    for (int64_t p = 0; p < hm.N; ++p) {
        int b, i[MAXD], j[MAXD];
        site_coords(hm, p, b, i);
        for (const auto &t : hm.basis_terms[b]) {
            const int nn = t.kind - 1;
            int64_t q[3];
            bool all = true;
            for (int k = 0; k < nn; ++k) {
                if (!neighbour_cell(hm, i, t.off[k], j)) { all = false; break; }
                q[k] = site_index(hm, t.nb_basis[k], j) + 1;
            }
            if (!all) continue;
            if (t.kind == 2 && bil) bil[p * hm.N2 + t.row] = q[0];
            if (t.kind == 3 && cub) { int tt = (t.row - hm.N2) / 2; cub[(p * hm.N3 + tt) * 2] = q[0]; cub[(p * hm.N3 + tt) * 2 + 1] = q[1]; }
            if (t.kind == 4 && quar) { int tt = (t.row - hm.N2 - 2 * hm.N3) / 3; for (int k = 0; k < 3; ++k) quar[(p * hm.N4 + tt) * 3 + k] = q[k]; }
        }
    }
}

}  // namespace csmc
