/*
 * csmc_oracle.c — CPU restatement of ClassicalSpinMC.jl's sweep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline.  libcsmc.so never links it.
 *
 * The reference is pure Julia and cannot run here (no julia / MPI / HDF5 in the image), so this
 * file restates its algorithm site for site; every function cites the reference file:line it
 * follows (paths relative to the reference repo).  Parity status:
 *   - pinned by the reference's own golden values (tests/test_oracle_golden.py):
 *       test/latticetests.jl:17 (E == -1.0), :18 (field == (-1,-0,-0)), :30 (E/N == -2.0),
 *       :6 (|s| == S), test/mctests.jl:49,57 (annealed E/N rounds to -0.6444);
 *   - cubic / quartic contractions (Einsum.jl 0.4.1, not vendored; src/hamiltonian.jl:46-48,
 *     62-64,114,127,180,193): no reference test touches them -> "parity unpinned"; restated
 *     from the formulae as written, checked for self-consistency and against an independent
 *     numpy restatement of the same reference lines (tests/test_oracle_golden.py).
 *   - RNG: the reference uses Julia's task-local Xoshiro256++, unseeded in its tests, so no
 *     stream is pinned.  Reference-order drivers here use xoshiro256++; the colour-order
 *     Metropolis uses the same counter-based Philox4x32-10 stream as the CUDA kernels so the two
 *     can be compared proposal by proposal.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/csmc.h"

#define MAXD CSMC_MAX_DIM

typedef struct orc_lattice {
    int D, n_basis, periodic;
    int shape[MAXD];
    int64_t N;
    int N2, N3, N4;
    double S;
    /* per-site tables, reference layout (src/lattice.jl:14-22) */
    double *field;     /* N x 3 */
    double *onsite;    /* N x 9 */
    int64_t *bil_site; /* N x N2, 1-based, 0 null */
    int32_t *bil_mat;  /* N x N2 -> row in mats (0 == zero matrix) */
    int64_t *cub_site; /* N x N3 x 2 */
    int32_t *cub_ten;  /* N x N3 -> row in tens3 (0 == zeros) */
    int64_t *quar_site; /* N x N4 x 3 */
    int32_t *quar_ten;  /* N x N4 -> row in tens4 (0 == zeros) */
    /* unique coupling tables: the reference stores one copy per site (src/lattice.jl:17,19,21);
       identical values, shared storage */
    double *mats;  /* (1 + 2*N2) x 9   : 0 zero, 1+2t = J_t, 2+2t = transpose */
    double *tens3; /* (1 + 3*N3) x 27  : perspectives b1,b2,b3 ; index [a][b][c] row-major here */
    double *tens4; /* (1 + 4*N4) x 81 */
} orc_lattice;

/* ------------------------------------------------------------------------------------------ */
/* site indexing: src/lattice.jl:29-33 — sorted (basis, i1..iD) tuples, last index fastest      */
static int64_t site_index0(const orc_lattice *L, int b0, const int *idx0) {
    int64_t p = b0;
    for (int d = 0; d < L->D; ++d) p = p * L->shape[d] + idx0[d];
    return p;
}

static void site_coords(const orc_lattice *L, int64_t p0, int *b0, int *idx0) {
    for (int d = L->D - 1; d >= 0; --d) {
        idx0[d] = (int)(p0 % L->shape[d]);
        p0 /= L->shape[d];
    }
    *b0 = (int)p0;
}

/* BC(index, offset): src/lattice.jl:101-109.  Returns 0 when (open bc) the neighbour is missing,
   which is the `isnothing(findfirst(...))` branch of :197-203. */
static int apply_bc(const orc_lattice *L, const int *idx0, const int *off, int sign, int *out) {
    for (int d = 0; d < L->D; ++d) {
        int v = idx0[d] + sign * off[d];
        if (L->periodic) {
            v %= L->shape[d];
            if (v < 0) v += L->shape[d];
        } else if (v < 0 || v >= L->shape[d]) {
            return 0;
        }
        out[d] = v;
    }
    return 1;
}

/* literal restatement of `findfirst(x->x == (bj, new_ind...), indices)` over the sorted tuple
   list (src/lattice.jl:196): O(N) scan, used only to validate the closed form on small N. */
static int64_t find_literal(const orc_lattice *L, int b0, const int *idx0) {
    int bb, ii[MAXD];
    for (int64_t p = 0; p < L->N; ++p) {
        site_coords(L, p, &bb, ii);
        int same = (bb == b0);
        for (int d = 0; d < L->D && same; ++d) same = (ii[d] == idx0[d]);
        if (same) return p + 1;
    }
    return 0;
}

static int64_t find_site(const orc_lattice *L, int b0, const int *idx0, int literal) {
    if (literal) return find_literal(L, b0, idx0);
    return site_index0(L, b0, idx0) + 1;
}

/* tensor storage inside the oracle: row-major [a][b][c]; the ABI hands Julia column-major. */
static void load_tensor3(const double *colmajor, double *T) {
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c) T[a * 9 + b * 3 + c] = colmajor[a + 3 * b + 9 * c];
}
static void load_tensor4(const double *colmajor, double *T) {
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d)
                    T[a * 27 + b * 9 + c * 3 + d] = colmajor[a + 3 * b + 9 * c + 27 * d];
}

void orc_free(orc_lattice *L) {
    if (!L) return;
    free(L->field); free(L->onsite); free(L->bil_site); free(L->bil_mat);
    free(L->cub_site); free(L->cub_ten); free(L->quar_site); free(L->quar_ten);
    free(L->mats); free(L->tens3); free(L->tens4);
    free(L);
}

/* Lattice(...) table construction: src/lattice.jl:65-291 */
orc_lattice *orc_build(const csmc_model *m, int literal) {
    orc_lattice *L = (orc_lattice *)calloc(1, sizeof(orc_lattice));
    L->D = m->dim; L->n_basis = m->n_basis; L->periodic = m->periodic; L->S = m->S;
    int64_t cells = 1;
    for (int d = 0; d < L->D; ++d) { L->shape[d] = m->shape[d]; cells *= m->shape[d]; }
    L->N = cells * m->n_basis;                                   /* :72 */
    L->N2 = m->n_bilinear; L->N3 = m->n_cubic; L->N4 = m->n_quartic; /* :88-90 */
    const int D = L->D; const int64_t N = L->N;

    L->field = (double *)calloc((size_t)N * 3, sizeof(double));
    L->onsite = (double *)calloc((size_t)N * 9, sizeof(double));
    L->bil_site = (int64_t *)calloc((size_t)N * (L->N2 ? L->N2 : 1), sizeof(int64_t));
    L->bil_mat = (int32_t *)calloc((size_t)N * (L->N2 ? L->N2 : 1), sizeof(int32_t));
    L->cub_site = (int64_t *)calloc((size_t)N * (L->N3 ? L->N3 : 1) * 2, sizeof(int64_t));
    L->cub_ten = (int32_t *)calloc((size_t)N * (L->N3 ? L->N3 : 1), sizeof(int32_t));
    L->quar_site = (int64_t *)calloc((size_t)N * (L->N4 ? L->N4 : 1) * 3, sizeof(int64_t));
    L->quar_ten = (int32_t *)calloc((size_t)N * (L->N4 ? L->N4 : 1), sizeof(int32_t));
    L->mats = (double *)calloc((size_t)(1 + 2 * L->N2) * 9, sizeof(double));
    L->tens3 = (double *)calloc((size_t)(1 + 3 * L->N3) * 27, sizeof(double));
    L->tens4 = (double *)calloc((size_t)(1 + 4 * L->N4) * 81, sizeof(double));

    /* coupling perspectives.  transposeJ: src/interaction_matrix.jl:27-31 */
    for (int t = 0; t < L->N2; ++t) {
        const double *J = m->bil_matrix + 9 * t;
        double *A = L->mats + 9 * (1 + 2 * t), *B = L->mats + 9 * (2 + 2 * t);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) { A[3 * r + c] = J[3 * r + c]; B[3 * r + c] = J[3 * c + r]; }
    }
    /* permutedims(J,[2,1,3]) / [3,2,1]: src/lattice.jl:221,226 (transpositions: B[a,b,c] = J[b,a,c]
       resp. J[c,b,a]) */
    for (int t = 0; t < L->N3; ++t) {
        double J[27];
        load_tensor3(m->cub_tensor + 27 * t, J);
        double *P1 = L->tens3 + 27 * (1 + 3 * t), *P2 = P1 + 27, *P3 = P2 + 27;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c) {
                    P1[a * 9 + b * 3 + c] = J[a * 9 + b * 3 + c];
                    P2[a * 9 + b * 3 + c] = J[b * 9 + a * 3 + c];
                    P3[a * 9 + b * 3 + c] = J[c * 9 + b * 3 + a];
                }
    }
    /* permutedims(J,[2,1,3,4]) / [3,2,1,4] / [4,2,3,1]: src/lattice.jl:259,265,271 */
    for (int t = 0; t < L->N4; ++t) {
        double J[81];
        load_tensor4(m->quar_tensor + 81 * t, J);
        double *P1 = L->tens4 + 81 * (1 + 4 * t), *P2 = P1 + 81, *P3 = P2 + 81, *P4 = P3 + 81;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    for (int d = 0; d < 3; ++d) {
                        int o = a * 27 + b * 9 + c * 3 + d;
                        P1[o] = J[o];
                        P2[o] = J[b * 27 + a * 9 + c * 3 + d];
                        P3[o] = J[c * 27 + b * 9 + a * 3 + d];
                        P4[o] = J[d * 27 + b * 9 + c * 3 + a];
                    }
    }

    for (int64_t i = 0; i < N; ++i) {                             /* :168 */
        int b0, idx[MAXD], nj[MAXD], nk[MAXD], nl[MAXD];
        site_coords(L, i, &b0, idx);
        const int b = b0 + 1;
        memcpy(L->field + 3 * i, m->field + 3 * b0, 3 * sizeof(double));   /* :172 */
        memcpy(L->onsite + 9 * i, m->onsite + 9 * b0, 9 * sizeof(double)); /* :173 */

        for (int t = 0; t < L->N2; ++t) {                         /* :176-204 */
            const int b1 = m->bil_basis[2 * t], b2 = m->bil_basis[2 * t + 1];
            const int *off = m->bil_offset + D * t;
            int64_t *slot = L->bil_site + i * L->N2 + t;
            int32_t *mat = L->bil_mat + i * L->N2 + t;
            if (b != b1 && b != b2) { *slot = 0; *mat = 0; continue; } /* :178-182 */
            int bj, sign, which;
            if (b1 == b2) { bj = b; sign = 1; which = 1 + 2 * t; }     /* :184-186 */
            else if (b1 == b) { bj = b2; sign = 1; which = 1 + 2 * t; } /* :187-189 */
            else { bj = b1; sign = -1; which = 2 + 2 * t; }            /* :190-193 */
            int64_t j = 0;
            if (apply_bc(L, idx, off, sign, nj)) j = find_site(L, bj - 1, nj, literal); /* :195-196 */
            if (j) { *slot = j; *mat = which; } else { *slot = 0; *mat = 0; }           /* :197-203 */
        }

        for (int t = 0; t < L->N3; ++t) {                         /* :209-237 */
            const int b1 = m->cub_basis[3 * t], b2 = m->cub_basis[3 * t + 1], b3 = m->cub_basis[3 * t + 2];
            const int *oj = m->cub_offset + 2 * D * t, *ok = oj + D;
            int jo[MAXD], ko[MAXD], bj, bk, which;
            int64_t *slot = L->cub_site + (i * L->N3 + t) * 2;
            int32_t *ten = L->cub_ten + i * L->N3 + t;
            if (b != b1 && b != b2 && b != b3) { slot[0] = slot[1] = 0; *ten = 0; continue; }
            if (b1 == b) {                                        /* :215-216 */
                bj = b2; bk = b3; which = 1 + 3 * t;
                for (int d = 0; d < D; ++d) { jo[d] = oj[d]; ko[d] = ok[d]; }
            } else if (b2 == b) {                                 /* :217-221 */
                bj = b1; bk = b3; which = 2 + 3 * t;
                for (int d = 0; d < D; ++d) { ko[d] = ok[d] - oj[d]; jo[d] = -oj[d]; }
            } else {                                              /* :222-226 */
                bj = b2; bk = b1; which = 3 + 3 * t;
                for (int d = 0; d < D; ++d) { jo[d] = oj[d] - ok[d]; ko[d] = -ok[d]; }
            }
            int64_t j = 0, k = 0;
            if (apply_bc(L, idx, jo, 1, nj)) j = find_site(L, bj - 1, nj, literal); /* :228 */
            if (apply_bc(L, idx, ko, 1, nk)) k = find_site(L, bk - 1, nk, literal); /* :229 */
            if (!j || !k) { slot[0] = slot[1] = 0; *ten = 0; }    /* :230-232 */
            else { slot[0] = j; slot[1] = k; *ten = which; }      /* :233-235 */
        }

        for (int t = 0; t < L->N4; ++t) {                         /* :243-283 */
            const int *bb = m->quar_basis + 4 * t;
            const int *oj = m->quar_offset + 3 * D * t, *ok = oj + D, *ol = ok + D;
            int jo[MAXD], ko[MAXD], lo[MAXD], bj, bk, bl, which;
            int64_t *slot = L->quar_site + (i * L->N4 + t) * 3;
            int32_t *ten = L->quar_ten + i * L->N4 + t;
            if (b != bb[0] && b != bb[1] && b != bb[2] && b != bb[3]) {
                slot[0] = slot[1] = slot[2] = 0; *ten = 0; continue;
            }
            if (bb[0] == b) {                                     /* :251-252 */
                bj = bb[1]; bk = bb[2]; bl = bb[3]; which = 1 + 4 * t;
                for (int d = 0; d < D; ++d) { jo[d] = oj[d]; ko[d] = ok[d]; lo[d] = ol[d]; }
            } else if (bb[1] == b) {                              /* :253-259 */
                bj = bb[0]; bk = bb[2]; bl = bb[3]; which = 2 + 4 * t;
                for (int d = 0; d < D; ++d) { jo[d] = -oj[d]; ko[d] = ok[d] + jo[d]; lo[d] = ol[d] + jo[d]; }
            } else if (bb[2] == b) {                              /* :260-265 */
                bj = bb[1]; bk = bb[0]; bl = bb[3]; which = 3 + 4 * t;
                for (int d = 0; d < D; ++d) { ko[d] = -ok[d]; lo[d] = ol[d] + ko[d]; jo[d] = oj[d] + ko[d]; }
            } else {                                              /* :266-271 */
                bj = bb[1]; bk = bb[2]; bl = bb[0]; which = 4 + 4 * t;
                for (int d = 0; d < D; ++d) { lo[d] = -ol[d]; jo[d] = oj[d] + lo[d]; ko[d] = ok[d] + lo[d]; }
            }
            int64_t j = 0, k = 0, l = 0;
            if (apply_bc(L, idx, jo, 1, nj)) j = find_site(L, bj - 1, nj, literal); /* :273 */
            if (apply_bc(L, idx, ko, 1, nk)) k = find_site(L, bk - 1, nk, literal); /* :274 */
            if (apply_bc(L, idx, lo, 1, nl)) l = find_site(L, bl - 1, nl, literal); /* :275 */
            if (!j || !k || !l) { slot[0] = slot[1] = slot[2] = 0; *ten = 0; }     /* :276-278 */
            else { slot[0] = j; slot[1] = k; slot[2] = l; *ten = which; }         /* :279-281 */
        }
    }
    return L;
}

int64_t orc_n_sites(const orc_lattice *L) { return L->N; }

void orc_get_tables(const orc_lattice *L, int64_t *bil, int64_t *cub, int64_t *quar) {
    if (bil) memcpy(bil, L->bil_site, sizeof(int64_t) * (size_t)L->N * L->N2);
    if (cub) memcpy(cub, L->cub_site, sizeof(int64_t) * (size_t)L->N * L->N3 * 2);
    if (quar) memcpy(quar, L->quar_site, sizeof(int64_t) * (size_t)L->N * L->N4 * 3);
}

/* per-site bilinear matrices as the reference stores them (lat.bilinear_matrices): N x N2 x 9 */
void orc_get_bilinear_matrices(const orc_lattice *L, double *out) {
    for (int64_t i = 0; i < L->N * L->N2; ++i) memcpy(out + 9 * i, L->mats + 9 * L->bil_mat[i], 72);
}

/* ------------------------------------------------------------------------------------------ */
/* get_local_field(lattice, point): src/hamiltonian.jl:3-67.  p is 1-based. */
void orc_local_field(const orc_lattice *L, const double *spins, int64_t p, double *out) {
    const int64_t i = p - 1;
    const double *s = spins + 3 * i, *o = L->onsite + 9 * i, *h = L->field + 3 * i;
    double Hx = 0.0, Hy = 0.0, Hz = 0.0;
    Hx += 2 * (o[0] * s[0] + o[1] * s[1] + o[2] * s[2]);         /* :20-22 */
    Hy += 2 * (o[3] * s[0] + o[4] * s[1] + o[5] * s[2]);
    Hz += 2 * (o[6] * s[0] + o[7] * s[1] + o[8] * s[2]);
    for (int n = 0; n < L->N2; ++n) {                            /* :25-34 */
        int64_t j = L->bil_site[i * L->N2 + n];
        if (j == 0) continue;
        const double *J = L->mats + 9 * L->bil_mat[i * L->N2 + n], *sj = spins + 3 * (j - 1);
        Hx += J[0] * sj[0] + J[1] * sj[1] + J[2] * sj[2];
        Hy += J[3] * sj[0] + J[4] * sj[1] + J[5] * sj[2];
        Hz += J[6] * sj[0] + J[7] * sj[1] + J[8] * sj[2];
    }
    for (int n = 0; n < L->N3; ++n) {                            /* :37-49 */
        const int64_t *c = L->cub_site + (i * L->N3 + n) * 2;
        if (c[0] == 0 && c[1] == 0) continue;
        const double *C = L->tens3 + 27 * L->cub_ten[i * L->N3 + n];
        const double *sj = spins + 3 * (c[0] - 1), *sk = spins + 3 * (c[1] - 1);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                Hx += C[0 * 9 + a * 3 + b] * sj[a] * sk[b];
                Hy += C[1 * 9 + a * 3 + b] * sj[a] * sk[b];
                Hz += C[2 * 9 + a * 3 + b] * sj[a] * sk[b];
            }
    }
    for (int n = 0; n < L->N4; ++n) {                            /* :52-65 */
        const int64_t *r = L->quar_site + (i * L->N4 + n) * 3;
        if (r[0] == 0 && r[1] == 0 && r[2] == 0) continue;
        const double *R = L->tens4 + 81 * L->quar_ten[i * L->N4 + n];
        const double *sj = spins + 3 * (r[0] - 1), *sk = spins + 3 * (r[1] - 1), *sl = spins + 3 * (r[2] - 1);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c) {
                    Hx += R[0 * 27 + a * 9 + b * 3 + c] * sj[a] * sk[b] * sl[c];
                    Hy += R[1 * 27 + a * 9 + b * 3 + c] * sj[a] * sk[b] * sl[c];
                    Hz += R[2 * 27 + a * 9 + b * 3 + c] * sj[a] * sk[b] * sl[c];
                }
    }
    out[0] = Hx - h[0]; out[1] = Hy - h[1]; out[2] = Hz - h[2];  /* :66 */
}

/* energy(lattice, point): src/hamiltonian.jl:139-196 */
double orc_site_energy(const orc_lattice *L, const double *spins, int64_t p) {
    const int64_t i = p - 1;
    const double *s = spins + 3 * i, *o = L->onsite + 9 * i, *h = L->field + 3 * i;
    double E = 0.0;
    E += s[0] * (o[0] * s[0] + o[1] * s[1] + o[2] * s[2]) +      /* :155-157 */
         s[1] * (o[3] * s[0] + o[4] * s[1] + o[5] * s[2]) +
         s[2] * (o[6] * s[0] + o[7] * s[1] + o[8] * s[2]);
    for (int n = 0; n < L->N2; ++n) {                            /* :160-169 */
        int64_t j = L->bil_site[i * L->N2 + n];
        if (j == 0) continue;
        const double *J = L->mats + 9 * L->bil_mat[i * L->N2 + n], *sj = spins + 3 * (j - 1);
        E += s[0] * (J[0] * sj[0] + J[1] * sj[1] + J[2] * sj[2]) +
             s[1] * (J[3] * sj[0] + J[4] * sj[1] + J[5] * sj[2]) +
             s[2] * (J[6] * sj[0] + J[7] * sj[1] + J[8] * sj[2]);
    }
    for (int n = 0; n < L->N3; ++n) {                            /* :172-181 */
        const int64_t *c = L->cub_site + (i * L->N3 + n) * 2;
        if (c[0] == 0 && c[1] == 0) continue;
        const double *C = L->tens3 + 27 * L->cub_ten[i * L->N3 + n];
        const double *sj = spins + 3 * (c[0] - 1), *sk = spins + 3 * (c[1] - 1);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int cc = 0; cc < 3; ++cc) E += C[a * 9 + b * 3 + cc] * s[a] * sj[b] * sk[cc];
    }
    for (int n = 0; n < L->N4; ++n) {                            /* :184-194 */
        const int64_t *r = L->quar_site + (i * L->N4 + n) * 3;
        if (r[0] == 0 && r[1] == 0 && r[2] == 0) continue;
        const double *R = L->tens4 + 81 * L->quar_ten[i * L->N4 + n];
        const double *sj = spins + 3 * (r[0] - 1), *sk = spins + 3 * (r[1] - 1), *sl = spins + 3 * (r[2] - 1);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    for (int d = 0; d < 3; ++d)
                        E += R[a * 27 + b * 9 + c * 3 + d] * s[a] * sj[b] * sk[c] * sl[d];
    }
    return E - (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);        /* :195 */
}

/* total_energy(lattice): src/hamiltonian.jl:70-132.  Also returns sum |e_i| contributions in
   *abs_sum (not in the reference) so tests can state "1e-12 relative to sum |e_i|". */
double orc_total_energy(const orc_lattice *L, const double *spins, double *abs_sum) {
    double E2 = 0.0, E3 = 0.0, E4 = 0.0, Ez = 0.0, Eo = 0.0, A = 0.0;
    for (int64_t i = 0; i < L->N; ++i) {
        const double *s = spins + 3 * i, *o = L->onsite + 9 * i, *h = L->field + 3 * i;
        double t = (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);    /* :89 */
        Ez -= t; A += fabs(t);
        t = s[0] * (o[0] * s[0] + o[1] * s[1] + o[2] * s[2]) +   /* :90-92 */
            s[1] * (o[3] * s[0] + o[4] * s[1] + o[5] * s[2]) +
            s[2] * (o[6] * s[0] + o[7] * s[1] + o[8] * s[2]);
        Eo += t; A += fabs(t);
        for (int n = 0; n < L->N2; ++n) {                        /* :94-103 */
            int64_t j = L->bil_site[i * L->N2 + n];
            if (j == 0) continue;
            const double *J = L->mats + 9 * L->bil_mat[i * L->N2 + n], *sj = spins + 3 * (j - 1);
            t = s[0] * (J[0] * sj[0] + J[1] * sj[1] + J[2] * sj[2]) +
                s[1] * (J[3] * sj[0] + J[4] * sj[1] + J[5] * sj[2]) +
                s[2] * (J[6] * sj[0] + J[7] * sj[1] + J[8] * sj[2]);
            E2 += t; A += fabs(t) / 2;
        }
        for (int n = 0; n < L->N3; ++n) {                        /* :106-115 */
            const int64_t *c = L->cub_site + (i * L->N3 + n) * 2;
            if (c[0] == 0 && c[1] == 0) continue;
            const double *C = L->tens3 + 27 * L->cub_ten[i * L->N3 + n];
            const double *sj = spins + 3 * (c[0] - 1), *sk = spins + 3 * (c[1] - 1);
            t = 0.0;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    for (int cc = 0; cc < 3; ++cc) t += C[a * 9 + b * 3 + cc] * s[a] * sj[b] * sk[cc];
            E3 += t; A += fabs(t) / 3;
        }
        for (int n = 0; n < L->N4; ++n) {                        /* :118-128 */
            const int64_t *r = L->quar_site + (i * L->N4 + n) * 3;
            if (r[0] == 0 && r[1] == 0 && r[2] == 0) continue;
            const double *R = L->tens4 + 81 * L->quar_ten[i * L->N4 + n];
            const double *sj = spins + 3 * (r[0] - 1), *sk = spins + 3 * (r[1] - 1), *sl = spins + 3 * (r[2] - 1);
            t = 0.0;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    for (int c = 0; c < 3; ++c)
                        for (int d = 0; d < 3; ++d)
                            t += R[a * 27 + b * 9 + c * 3 + d] * s[a] * sj[b] * sk[c] * sl[d];
            E4 += t; A += fabs(t) / 4;
        }
    }
    if (abs_sum) *abs_sum = A;
    return E2 / 2 + E3 / 3 + E4 / 4 + Ez + Eo;                   /* :131 */
}

/* get_magnetization: src/observables.jl:12-18; m3 (optional) receives the vector sum */
double orc_magnetization(const orc_lattice *L, const double *spins, double *m3) {
    double mx = 0, my = 0, mz = 0;
    for (int64_t i = 0; i < L->N; ++i) { mx += spins[3 * i]; my += spins[3 * i + 1]; mz += spins[3 * i + 2]; }
    if (m3) { m3[0] = mx; m3[1] = my; m3[2] = mz; }
    return sqrt(mx * mx + my * my + mz * mz);
}

/* ------------------------------------------------------------------------------------------ */
/* one overrelaxation site update: src/monte_carlo.jl:128-137 */
static inline void or_update(const orc_lattice *L, double *spins, int64_t p) {
    double H[3];
    double *s = spins + 3 * (p - 1);
    orc_local_field(L, spins, p, H);
    if (H[0] == 0.0 && H[1] == 0.0 && H[2] == 0.0) return;        /* :131-134 */
    double proj = 2.0 * (s[0] * H[0] + s[1] * H[1] + s[2] * H[2]) / (H[0] * H[0] + H[1] * H[1] + H[2] * H[2]);
    double n0 = -s[0] + proj * H[0], n1 = -s[1] + proj * H[1], n2 = -s[2] + proj * H[2];
    s[0] = n0; s[1] = n1; s[2] = n2;
}

/* deterministic single-site update: src/monte_carlo.jl:205-210 */
static inline void det_update(const orc_lattice *L, double *spins, int64_t p) {
    double H[3];
    double *s = spins + 3 * (p - 1);
    orc_local_field(L, spins, p, H);
    if (H[0] == 0.0 && H[1] == 0.0 && H[2] == 0.0) return;
    double nrm = sqrt(H[0] * H[0] + H[1] * H[1] + H[2] * H[2]);
    s[0] = -H[0] / nrm * L->S; s[1] = -H[1] / nrm * L->S; s[2] = -H[2] / nrm * L->S;
}

/* overrelaxation!(lattice): src/monte_carlo.jl:126-139.  order == NULL: reference order 1..N
   (Gauss-Seidel, in place); otherwise the explicit visiting order (1-based), e.g. colour order. */
void orc_overrelax(const orc_lattice *L, double *spins, const int64_t *order, int64_t n, int n_sweeps) {
    for (int sw = 0; sw < n_sweeps; ++sw) {
        if (!order) for (int64_t p = 1; p <= L->N; ++p) or_update(L, spins, p);
        else for (int64_t q = 0; q < n; ++q) or_update(L, spins, order[q]);
    }
}

void orc_deterministic_order(const orc_lattice *L, double *spins, const int64_t *order, int64_t n, int n_sweeps) {
    for (int sw = 0; sw < n_sweeps; ++sw) {
        if (!order) for (int64_t p = 1; p <= L->N; ++p) det_update(L, spins, p);
        else for (int64_t q = 0; q < n; ++q) det_update(L, spins, order[q]);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* RNG 1: xoshiro256++ (the generator family Julia >= 1.7 uses for rand()); stream unpinned. */
typedef struct { uint64_t s[4]; } orc_rng;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(orc_rng *r) {
    uint64_t *s = r->s, result = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
void orc_rng_seed(orc_rng *r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) { /* splitmix64 */
        uint64_t z = (seed += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        r->s[i] = z ^ (z >> 31);
    }
}
static inline double xo_uniform(orc_rng *r) { return (double)(xo_next(r) >> 11) * 0x1.0p-53; }

/* random_spin_orientation(S): src/lattice.jl:306-311 */
static inline void rso(double S, double u1, double u2, double *out) {
    double phi = 2.0 * M_PI * u1;
    double z = 2.0 * u2 - 1.0;
    double r = sqrt(1.0 - z * z);
    out[0] = S * (r * cos(phi)); out[1] = S * (r * sin(phi)); out[2] = S * z;
}

/* gaussian_move: src/metropolis.jl:84-87 */
static inline void cone_move(double S, const double *s, double sigma, double u1, double u2, double *out) {
    double r[3]; rso(S, u1, u2, r);
    double n0 = s[0] + sigma * r[0], n1 = s[1] + sigma * r[1], n2 = s[2] + sigma * r[2];
    double nrm = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    out[0] = n0 / nrm * S; out[1] = n1 / nrm * S; out[2] = n2 / nrm * S;
}

/* metropolis!(mc, T): src/metropolis.jl:65-82 with calculate_energy_diff! :94-101 (sigma < 0) or
   the cone version :103-110 (sigma >= 0).  N proposals at uniformly random sites, with
   replacement; two energy() evaluations per proposal; third uniform only drawn when dE >= 0. */
double orc_metropolis_ref(const orc_lattice *L, double *spins, double T, double sigma, orc_rng *rng) {
    double accepted = 0.0;
    for (int64_t sweep = 0; sweep < L->N; ++sweep) {
        int64_t point = 1 + (int64_t)(xo_uniform(rng) * (double)L->N);   /* :70 */
        if (point > L->N) point = L->N;
        double *s = spins + 3 * (point - 1);
        double old[3] = {s[0], s[1], s[2]}, prop[3];                      /* :71 */
        double E_old = orc_site_energy(L, spins, point);                  /* :95 */
        double u1 = xo_uniform(rng), u2 = xo_uniform(rng);
        if (sigma < 0) rso(L->S, u1, u2, prop); else cone_move(L->S, old, sigma, u1, u2, prop);
        s[0] = prop[0]; s[1] = prop[1]; s[2] = prop[2];                   /* :97 */
        double dE = orc_site_energy(L, spins, point) - E_old;             /* :98-99 */
        int accept = dE < 0 ? 1 : (xo_uniform(rng) < exp(-dE / T));       /* :73 */
        if (!accept) { s[0] = old[0]; s[1] = old[1]; s[2] = old[2]; }     /* :74-75 */
        else accepted += 1;
    }
    return accepted;
}

/* RNG 2: Philox4x32-10 (Salmon et al., SC'11), the counter-based stream shared with the CUDA
   kernels.  key = (seed_lo, seed_hi); counter = (site0, global replica, ctr_lo, ctr_hi<<8 | tag). */
static inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }

enum { TAG_PROPOSE = 0, TAG_ACCEPT = 1, TAG_INIT = 2, TAG_EXCHANGE = 3 };

void orc_philox_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint64_t ctr, uint32_t tag, uint32_t *out) {
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, (uint32_t)ctr, (uint32_t)((ctr >> 32) << 8) | tag, out);
}

/* Lattice(...; initialCondition=:random): src/lattice.jl:76-79, Philox stream (seed, replica, site) */
void orc_randomize_spins(const orc_lattice *L, double *spins, uint64_t seed, uint32_t replica) {
    for (int64_t i = 0; i < L->N; ++i) {
        uint32_t r[4];
        orc_philox_raw(seed, (uint32_t)i, replica, 0, TAG_INIT, r);
        rso(L->S, u53(r[0], r[1]), u53(r[2], r[3]), spins + 3 * i);
    }
}

/* One Metropolis sweep visiting `order` once (colour order), same arithmetic as metropolis!
   (two energy() evaluations, src/metropolis.jl:94-101) but the Philox stream of the CUDA kernels.
   sigma < 0: uniform proposal; sigma >= 0: cone move. */
double orc_metropolis_philox(const orc_lattice *L, double *spins, const int64_t *order, int64_t n,
                             double T, double sigma, uint64_t seed, uint32_t replica, uint64_t sweep_ctr) {
    double accepted = 0.0;
    for (int64_t q = 0; q < n; ++q) {
        int64_t point = order ? order[q] : q + 1;
        double *s = spins + 3 * (point - 1);
        double old[3] = {s[0], s[1], s[2]}, prop[3];
        uint32_t r[4];
        orc_philox_raw(seed, (uint32_t)(point - 1), replica, sweep_ctr, TAG_PROPOSE, r);
        /* one Philox call per proposal, 128 bits -> u1 (43 bits), u2 (43 bits), u3 (42 bits) */
        double u1 = (double)(((uint64_t)r[0] << 11) | (r[1] >> 21)) * 0x1.0p-43;
        double u2 = (double)(((uint64_t)(r[1] & 0x1FFFFFu) << 22) | (r[2] >> 10)) * 0x1.0p-43;
        double u3 = (double)(((uint64_t)(r[2] & 0x3FFu) << 32) | r[3]) * 0x1.0p-42;
        double E_old = orc_site_energy(L, spins, point);
        if (sigma < 0) rso(L->S, u1, u2, prop); else cone_move(L->S, old, sigma, u1, u2, prop);
        s[0] = prop[0]; s[1] = prop[1]; s[2] = prop[2];
        double dE = orc_site_energy(L, spins, point) - E_old;
        int accept = dE < 0 ? 1 : (u3 < exp(-dE / T));
        if (!accept) { s[0] = old[0]; s[1] = old[1]; s[2] = old[2]; } else accepted += 1;
    }
    return accepted;
}

/* ------------------------------------------------------------------------------------------ */
/* Test helpers that are not in the reference: all-sites field, and colour-order sweeps that carry a
   per-site forward error bound along (the "conditioning-aware" tolerance of the full-size parity
   tests).  Two implementations of the same update that differ only in rounding (summation order,
   fused multiply-adds) differ after the update by at most TOL * S * kappa[i], where kappa follows
   the first-order recursion below with a rounding budget of TOL * ORC_ROUND per field term. */
void orc_local_field_all(const orc_lattice *L, const double *spins, double *out) {
    for (int64_t p = 1; p <= L->N; ++p) orc_local_field(L, spins, p, out + 3 * (p - 1));
}

#define ORC_ROUND (1.0 / 512.0)   /* in units of TOL = 1e-12: 1.95e-15 = 17.6 u per term of the field sum */

static double frob(const double *T, int n) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += T[k] * T[k];
    return sqrt(s);
}

/* spectral norm of a 3x3 matrix: sqrt of the largest eigenvalue of J^T J (trigonometric closed form for a
   symmetric 3x3 matrix), times 1 + 1e-9 so that it stays an upper bound under its own rounding */
static double spec3(const double *J) {
    double M[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) M[a][b] = J[0 + a] * J[0 + b] + J[3 + a] * J[3 + b] + J[6 + a] * J[6 + b];
    const double p1 = M[0][1] * M[0][1] + M[0][2] * M[0][2] + M[1][2] * M[1][2];
    const double q = (M[0][0] + M[1][1] + M[2][2]) / 3.0;
    const double p2 = (M[0][0] - q) * (M[0][0] - q) + (M[1][1] - q) * (M[1][1] - q) + (M[2][2] - q) * (M[2][2] - q) + 2.0 * p1;
    if (p2 <= 1e-300) return sqrt(fmax(q, 0.0)) * (1.0 + 1e-9);
    const double p = sqrt(p2 / 6.0);
    double B[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) B[a][b] = (M[a][b] - (a == b ? q : 0.0)) / p;
    double r = (B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0]) +
                B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0])) / 2.0;
    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    const double lmax = q + 2.0 * p * cos(acos(r) / 3.0);
    return sqrt(fmax(lmax, 0.0)) * (1.0 + 1e-9);
}

/* kF = S * sum_slots |T| S^(order-2) sum_nbrs kappa_nbr + ORC_ROUND * sum |terms of get_local_field|:
   bound (in units of TOL) on the error of the local field of site i given the neighbours' bounds */
static double field_error_bound(const orc_lattice *L, const double *spins, const double *kappa, int64_t i) {
    const double *s = spins + 3 * i, *o = L->onsite + 9 * i, *h = L->field + 3 * i;
    const double S = L->S;
    double A = fabs(h[0]) + fabs(h[1]) + fabs(h[2]), prop = 0.0;
    for (int k = 0; k < 9; ++k) A += 2 * fabs(o[k] * s[k % 3]);
    prop += 2 * spec3(o) * kappa[i];
    for (int n = 0; n < L->N2; ++n) {
        const int64_t j = L->bil_site[i * L->N2 + n];
        if (j == 0) continue;
        const double *J = L->mats + 9 * L->bil_mat[i * L->N2 + n], *sj = spins + 3 * (j - 1);
        for (int k = 0; k < 9; ++k) A += fabs(J[k] * sj[k % 3]);
        prop += spec3(J) * kappa[j - 1];
    }
    for (int n = 0; n < L->N3; ++n) {
        const int64_t *c = L->cub_site + (i * L->N3 + n) * 2;
        if (c[0] == 0 && c[1] == 0) continue;
        const double *C = L->tens3 + 27 * L->cub_ten[i * L->N3 + n];
        const double *sj = spins + 3 * (c[0] - 1), *sk = spins + 3 * (c[1] - 1);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int cc = 0; cc < 3; ++cc) A += fabs(C[a * 9 + b * 3 + cc] * sj[b] * sk[cc]);
        prop += frob(C, 27) * S * (kappa[c[0] - 1] + kappa[c[1] - 1]);
    }
    for (int n = 0; n < L->N4; ++n) {
        const int64_t *r = L->quar_site + (i * L->N4 + n) * 3;
        if (r[0] == 0 && r[1] == 0 && r[2] == 0) continue;
        const double *R = L->tens4 + 81 * L->quar_ten[i * L->N4 + n];
        const double *sj = spins + 3 * (r[0] - 1), *sk = spins + 3 * (r[1] - 1), *sl = spins + 3 * (r[2] - 1);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    for (int d = 0; d < 3; ++d) A += fabs(R[a * 27 + b * 9 + c * 3 + d] * sj[b] * sk[c] * sl[d]);
        prop += frob(R, 81) * S * S * (kappa[r[0] - 1] + kappa[r[1] - 1] + kappa[r[2] - 1]);
    }
    return S * prop + ORC_ROUND * A;
}

/* kind 0: overrelaxation (s' = R(F) s: |ds'| <= |ds| + 4 S |dF| / |F|), 1: deterministic
   (s' = -S F/|F|: |ds'| <= 2 S |dF| / |F|), 2: Metropolis with the shared Philox stream (an accepted
   proposal does not depend on the neighbours; a rejected one keeps the old spin).  Performs the same
   updates as orc_overrelax / orc_deterministic_order / orc_metropolis_philox and advances kappa[N]
   (start it at zeros for bit-identical inputs).  Returns the accepted count (kind 2). */
double orc_sweep_tracked(const orc_lattice *L, double *spins, const int64_t *order, int64_t n, int kind,
                         double T, uint64_t seed, uint32_t replica, uint64_t sweep_ctr, double *kappa) {
    double accepted = 0.0;
    for (int64_t q = 0; q < n; ++q) {
        const int64_t p = order ? order[q] : q + 1, i = p - 1;
        if (kind == 2) {
            const int64_t one = p;
            double *s = spins + 3 * i;
            const double o0 = s[0], o1 = s[1], o2 = s[2];
            const double a = orc_metropolis_philox(L, spins, &one, 1, T, -1.0, seed, replica, sweep_ctr);
            if (a != 0.0 || s[0] != o0 || s[1] != o1 || s[2] != o2) kappa[i] = ORC_ROUND;
            accepted += a;
            continue;
        }
        double H[3];
        orc_local_field(L, spins, p, H);
        const double nrm = sqrt(H[0] * H[0] + H[1] * H[1] + H[2] * H[2]);
        if (nrm == 0.0) continue;
        const double kF = field_error_bound(L, spins, kappa, i);
        if (kind == 0) { kappa[i] = kappa[i] + 4.0 * kF / nrm + ORC_ROUND; or_update(L, spins, p); }
        else { kappa[i] = 2.0 * kF / nrm + ORC_ROUND; det_update(L, spins, p); }
    }
    return accepted;
}

/* adaptive sigma rule: src/metropolis.jl:129-131 */
double orc_adapt_sigma(double sigma, double accepted, double n) {
    double a = accepted / n, f = 0.5 / fmax(1 - a, 0.05);
    double v = sigma * f;
    return v < 0.0 ? 0.0 : (v > 100.0 ? 100.0 : v);
}

/* ------------------------------------------------------------------------------------------ */
/* simulated_annealing!: src/monte_carlo.jl:157-190.  temps[] is the sequence T0, schedule(1), ...
   while T > mc.T (computed by the caller exactly as :168,184).  alg: 0 Metropolis, 1
   MetropolisAdaptive, 2 MetropolisFixedCone.  accept_out[n_temps] gets R per temperature. */
void orc_simulated_annealing(const orc_lattice *L, double *spins, const double *temps, int n_temps,
                             int64_t t_thermalization, int rate, int alg, double sigma0,
                             uint64_t seed, double *accept_out) {
    orc_rng rng; orc_rng_seed(&rng, seed);
    for (int it = 0; it < n_temps; ++it) {
        double T = temps[it], R = 0.0, sigma = sigma0;                 /* :169-171 */
        for (int64_t t = 1; t < t_thermalization; ++t) {               /* :172 */
            int do_metro = 1;
            if (rate != 0) { orc_overrelax(L, spins, NULL, 0, 1); do_metro = (t % rate == 0); } /* :173-177 */
            if (do_metro) {
                double acc = orc_metropolis_ref(L, spins, T, alg == 0 ? -1.0 : sigma, &rng);
                if (alg == 1) sigma = orc_adapt_sigma(sigma, acc, (double)L->N);
                R += acc;
            }
        }
        if (accept_out) accept_out[it] = R;
    }
}

/* deterministic_updates!: src/monte_carlo.jl:201-213 — t_deterministic-1 random single-site updates */
void orc_deterministic_updates(const orc_lattice *L, double *spins, int64_t t_deterministic, uint64_t seed) {
    orc_rng rng; orc_rng_seed(&rng, seed);
    for (int64_t sweeps = 1; sweeps < t_deterministic; ++sweeps) {
        int64_t point = 1 + (int64_t)(xo_uniform(&rng) * (double)L->N);
        if (point > L->N) point = L->N;
        det_update(L, spins, point);
    }
}

/* exchange decision: src/monte_carlo.jl:327-330.  (T,E) of the even-indexed member, (Tp,Ep) partner */
int orc_exchange_accept(double T, double E, double Tp, double Ep, double u) {
    double delta_beta = (1 / Tp - 1 / T), delta_E = (Ep - E);
    double w = exp(delta_beta * delta_E);
    return u < (w < 1.0 ? w : 1.0);
}

/* parallel_tempering!: src/monte_carlo.jl:235-398 with R "ranks" == replicas, one per OpenMP
   thread ("one temperature per CPU", examples/parallel_tempering/README.txt:17).  Configurations
   are swapped as in :336-347.  spins: R x N x 3.  Outputs (may be NULL): E_series/M_series
   [n_probe x R] in probe order, accepted_local[R], exchanges[R] (:269-274).  Returns probes taken.
   exchange_philox != 0 draws the exchange uniform from the shared Philox stream
   (pair's lower slot, exchange counter) so the device path can be compared decision by decision. */
int64_t orc_parallel_tempering(const orc_lattice *L, double *spins, const double *T, int R,
                               int64_t t_thermalization, int64_t t_measurement, int probe_rate,
                               int swap_rate, int rate, uint64_t seed, int n_threads,
                               double *E_series, double *M_series, double *accepted_local,
                               double *exchanges) {
    const int64_t total = t_thermalization + t_measurement, N3 = 3 * L->N;   /* :276 */
    const int dosweep = rate == 0 ? 1 : rate;                                 /* :289-293 */
    orc_rng *rng = (orc_rng *)malloc(sizeof(orc_rng) * R);
    double *E = (double *)malloc(sizeof(double) * R);
    double *tmp = (double *)malloc(sizeof(double) * N3);
    for (int r = 0; r < R; ++r) { orc_rng_seed(&rng[r], seed + 1000003ULL * (uint64_t)r); E[r] = orc_total_energy(L, spins + r * N3, NULL); } /* :265 */
    if (accepted_local) memset(accepted_local, 0, sizeof(double) * R);
    if (exchanges) memset(exchanges, 0, sizeof(double) * R);
    orc_rng xrng; orc_rng_seed(&xrng, seed ^ 0xabcdef12345ULL);
    int64_t n_probe = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    for (int64_t sweep = 0; sweep < total; ++sweep) {                         /* :295 */
        const int metro = (sweep % dosweep == 0);
#pragma omp parallel for schedule(static)
        for (int r = 0; r < R; ++r) {
            double *sp = spins + r * N3;
            if (rate != 0) orc_overrelax(L, sp, NULL, 0, 1);                  /* :298-300 */
            if (metro) {                                                      /* :302-305 */
                double a = orc_metropolis_ref(L, sp, T[r], -1.0, &rng[r]);
                if (accepted_local) accepted_local[r] += a;
                E[r] = orc_total_energy(L, sp, NULL);
            }
        }
        if (metro && R > 1 && sweep % swap_rate == 0) {                       /* :308 */
            const int first = ((sweep / swap_rate) % 2 == 0) ? 0 : 1;         /* :311-315 */
            for (int a = first; a + 1 < R; a += 2) {                          /* :317 */
                const int b = a + 1;
                /* the even-rank member decides (:327); pairs are (even, even+1) when first == 0
                   and (odd, odd+1) when first == 1, where the even member is b */
                const int ev = (a % 2 == 0) ? a : b, od = (ev == a) ? b : a;
                double u = xo_uniform(&xrng);
                if (orc_exchange_accept(T[ev], E[ev], T[od], E[od], u)) {     /* :328-330 */
                    memcpy(tmp, spins + a * N3, sizeof(double) * N3);         /* :336-345 */
                    memcpy(spins + a * N3, spins + b * N3, sizeof(double) * N3);
                    memcpy(spins + b * N3, tmp, sizeof(double) * N3);
                    double e = E[a]; E[a] = E[b]; E[b] = e;                   /* :346 */
                    if (exchanges) { exchanges[a] += 1; exchanges[b] += 1; }  /* :337 */
                }
            }
        }
        if (sweep >= t_thermalization && sweep % probe_rate == 0) {           /* :353,368-370 */
            for (int r = 0; r < R; ++r) {
                if (E_series) E_series[n_probe * R + r] = E[r];
                if (M_series) M_series[n_probe * R + r] = orc_magnetization(L, spins + r * N3, NULL);
            }
            ++n_probe;
        }
    }
    free(rng); free(E); free(tmp);
    return n_probe;
}

/* compute_equal_time_correlations(lat, ks): src/spin_correlations.jl:6-43.  pos: D x N column-major
   (site_positions), ks: D x N_k column-major, out: 9 x N_k column-major (Suv[3u+v, n]). */
void orc_structure_factor(const orc_lattice *L, const double *spins, const double *pos, const double *ks, int64_t n_k, double *out) {
    const int D = L->D;
    for (int64_t n = 0; n < n_k; ++n) {
        double re[3] = {0, 0, 0}, im[3] = {0, 0, 0};
        for (int64_t i = 0; i < L->N; ++i) {
            double kr = 0.0;
            for (int d = 0; d < D; ++d) kr += ks[n * D + d] * pos[i * D + d];      /* transpose(ks[:, n]) * pos, :20 */
            const double c = cos(kr), sn = -sin(kr);                                /* exp(-im * kr) */
            for (int u = 0; u < 3; ++u) { re[u] += c * spins[3 * i + u]; im[u] += sn * spins[3 * i + u]; }   /* :21-23 */
        }
        for (int u = 0; u < 3; ++u)
            for (int v = 0; v < 3; ++v)
                out[n * 9 + 3 * u + v] = (re[u] * re[v] + im[u] * im[v]) / (double)L->N;   /* real(s_u conj(s_v)) / N, :31-42 */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* timed CPU baseline for bench.py: n_threads independent replicas of the lattice, each running
   n_cycles x (or_per_cycle overrelaxation! sweeps + metro_per_cycle metropolis! sweeps) with the
   reference algorithm unchanged.  spins: n_threads x N x 3.  Returns total single-spin updates. */
double orc_cycles(const orc_lattice *L, double *spins, int n_threads, double T, int64_t n_cycles,
                  int or_per_cycle, int metro_per_cycle, uint64_t seed) {
    const int64_t N3 = 3 * L->N;
#ifdef _OPENMP
    omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n_threads; ++r) {
        orc_rng rng; orc_rng_seed(&rng, seed + 7919ULL * (uint64_t)r);
        double *sp = spins + r * N3;
        for (int64_t c = 0; c < n_cycles; ++c) {
            orc_overrelax(L, sp, NULL, 0, or_per_cycle);
            for (int mm = 0; mm < metro_per_cycle; ++mm) orc_metropolis_ref(L, sp, T, -1.0, &rng);
        }
    }
    return (double)n_threads * (double)n_cycles * (double)(or_per_cycle + metro_per_cycle) * (double)L->N;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
