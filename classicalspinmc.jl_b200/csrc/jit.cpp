// jit.cpp — runtime specialisation of the pass kernels: generates CUDA C++ for one lattice model
// (unrolled interaction terms, literal coefficients, constant geometry) and compiles it to an
// sm_100a cubin with NVRTC (loaded with dlopen so libcsmc.so has no link-time dependency on it).
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <stdexcept>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <tuple>

#include "csmc_internal.h"
#include "jit_prelude.h"

namespace csmc {

namespace {

inline int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
inline int posmod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

std::string lit(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "(%a)", v);   // hexadecimal floating literal: exact
    return buf;
}

int pow2_floor(int v) { int p = 1; while (2 * p <= v) p *= 2; return p; }
int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

struct LiveSlot { const HostTerm *t; int slot; int bit; int part; };

struct Gen {
    const HostModel &hm;
    std::ostringstream o;
    std::vector<std::vector<LiveSlot>> seg_live;   // per segment: live interaction slots (filled by segment())
    std::vector<int> seg_preload;             // per segment: neighbours preloaded into registers
    std::vector<int> seg_split;               // per segment: warps a site's slots are split over (1, 2, 4)
    std::vector<long> seg_flops;              // per segment: fp64 flops of one neighbour-field evaluation (fma = 2; a literal +-1 coefficient = 1)
    long flops_cur = 0;
    std::vector<double> ktab;                 // coefficients placed in the constant bank
    std::map<uint64_t, int> kslot;
    // interaction coefficient as an operand: simple values stay literals (the compiler folds +-1 into
    // adds), everything else is read from a __constant__ table with a compile-time index, so it is a
    // direct c[bank][offset] operand of the DFMA instead of an immediate built with two UMOVs
    std::string coef(double v) {
        static const bool use_table = std::getenv("CSMC_JIT_KTAB") != nullptr;   // A/B switch, default: literals
        if (!use_table || v == 1.0 || v == -1.0 || v == 2.0 || v == -2.0 || v == 0.5 || v == -0.5) return lit(v);
        uint64_t bits;
        std::memcpy(&bits, &v, 8);
        auto it = kslot.find(bits);
        int idx;
        if (it == kslot.end()) { idx = (int)ktab.size(); ktab.push_back(v); kslot[bits] = idx; }
        else idx = it->second;
        return "CSMC_K[" + std::to_string(idx) + "]";
    }
    int max_delta0 = 0;                       // largest neighbour shift along dimension 0, in supercells
    bool skew_ok = false;                     // see JitPlan::skew
    int skew_T0 = 0, skew_NT0 = 0, skew_row = 0;
    std::map<std::tuple<int, int, int, int>, int> seg_of_class;
    explicit Gen(const HostModel &h) : hm(h) {
        for (size_t s = 0; s < hm.segs.size(); ++s)
            seg_of_class[std::make_tuple(hm.segs[s].basis, hm.segs[s].r[0], hm.segs[s].r[1], hm.segs[s].r[2])] = (int)s;
    }

    // emits code computing `int j` (storage position of neighbour k of term t) and, for open
    // boundaries, clears `okt`.  Returns false when the neighbour class does not exist at all.
    bool neighbour(const HostSeg &hs, const HostTerm &t, int k) {
        int r2[MAXD] = {0, 0, 0}, delta[MAXD] = {0, 0, 0};
        for (int d = 0; d < hm.D; ++d) {
            r2[d] = posmod(hs.r[d] + t.off[k][d], hm.P[d]);
            delta[d] = floordiv(hs.r[d] + t.off[k][d], hm.P[d]);
        }
        auto it = seg_of_class.find(std::make_tuple(t.nb_basis[k], r2[0], r2[1], r2[2]));
        if (it == seg_of_class.end()) return false;
        max_delta0 = std::max(max_delta0, std::abs(delta[0]));
        const HostSeg &ns = hm.segs[it->second];
        o << "          int j" << k << ";\n          {\n";
        for (int d = 0; d < MAXD; ++d) {
            if (d >= hm.D) { o << "            const int n" << d << " = 0;\n"; continue; }
            o << "            int n" << d << " = m" << d << " + (" << delta[d] << ");\n";
            if (hm.periodic) {
                if (delta[d] > 0) o << "            n" << d << " = (n" << d << " >= " << ns.M[d] << ") ? n" << d << " - " << ns.M[d] << " : n" << d << ";\n";
                if (delta[d] < 0) o << "            n" << d << " = (n" << d << " < 0) ? n" << d << " + " << ns.M[d] << " : n" << d << ";\n";
            } else {
                if (delta[d] < 0) o << "            okt = okt && (n" << d << " >= 0);\n";
                if (hs.M[d] - 1 + delta[d] >= ns.M[d]) o << "            okt = okt && (n" << d << " < " << ns.M[d] << ");\n";
            }
        }
        o << "            j" << k << " = " << ns.start << " + (n0 * " << ns.M[1] << " + n1) * " << ns.M[2] << " + n2;\n          }\n";
        return true;
    }

    void segment(int s) {
        const HostSeg &hs = hm.segs[s];
        const int b = hs.basis;
        // which terms are live (some non-zero coefficient, every neighbour class exists)
        std::vector<LiveSlot> live;
        int nnb = 0;
        for (const auto &t : hm.basis_terms[b]) {
            const double *C = hm.coefs.data() + t.coef;
            const int ncoef = t.kind == 2 ? 9 : t.kind == 3 ? 27 : 81;
            bool any = false;
            for (int k = 0; k < ncoef; ++k) any |= (C[k] != 0.0);
            bool exists = true;
            for (int k = 0; k < t.kind - 1; ++k) {
                int r2[MAXD] = {0, 0, 0};
                for (int d = 0; d < hm.D; ++d) r2[d] = posmod(hs.r[d] + t.off[k][d], hm.P[d]);
                exists &= seg_of_class.count(std::make_tuple(t.nb_basis[k], r2[0], r2[1], r2[2])) > 0;
            }
            if (!any || !exists) continue;
            live.push_back({&t, nnb, (int)live.size(), 0});
            nnb += t.kind - 1;
        }
        if (live.size() > 32) throw std::runtime_error("more than 32 live interaction slots per site");
        // Heavy sites (cubic / quartic slots) are split over SP warps of the CTA: each warp evaluates a
        // subset of the slots for the same 32 sites, partial fields are summed through shared memory in a
        // fixed order.  4x the threads per site and a quarter of the live registers per thread.
        auto slot_cost = [](const HostTerm &t) { return t.kind == 2 ? 9 : t.kind == 3 ? 36 : 117; };
        int total_cost = 0;
        for (const auto &lv : live) total_cost += slot_cost(*lv.t);
        static const int sp_env = std::getenv("CSMC_JIT_SPLIT") ? std::atoi(std::getenv("CSMC_JIT_SPLIT")) : 0;
        // Measured on B200 (C5, L=512 / 1024): splitting is slower than one thread per site (8.7 / 7.2 us per
        // pass at SP = 4 / 2 against 5.7 us), the replicated index arithmetic and the barrier outweigh the
        // occupancy gain -- so it stays an experiment behind CSMC_JIT_SPLIT=2|4.
        (void)total_cost;
        int SP = sp_env > 0 ? sp_env : 1;
        if (SP != 1 && SP != 2 && SP != 4) SP = 1;
        if (nnb > 64) SP = 1;   // streaming variant is not split
        {
            std::vector<int> order(live.size()), load_of(SP, 0);
            for (size_t i = 0; i < live.size(); ++i) order[i] = (int)i;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return slot_cost(*live[a].t) > slot_cost(*live[b].t); });
            for (int i : order) {
                int best = 0;
                for (int p = 1; p < SP; ++p) if (load_of[p] < load_of[best]) best = p;
                live[i].part = best;
                load_of[best] += slot_cost(*live[i].t);
            }
        }
        seg_split.resize(hm.segs.size(), 1);
        seg_split[s] = SP;
        seg_live.resize(hm.segs.size());
        seg_live[s] = live;
        seg_preload.resize(hm.segs.size(), 1);

        o << "struct Seg" << s << " {\n";
        o << "    static constexpr int START = " << hs.start << ", COUNT = " << hs.count << ", NNB = " << nnb << ", SP = " << SP << ";\n";
        o << "    static constexpr double H0 = " << lit(hm.field[3 * b]) << ", H1 = " << lit(hm.field[3 * b + 1]) << ", H2 = " << lit(hm.field[3 * b + 2]) << ";\n";
        const bool ons = hm.onsite_coef[b] >= 0;
        o << "    static constexpr bool ONSITE = " << (ons ? "true" : "false") << ";\n";
        for (int k = 0; k < 9; ++k) o << "    static constexpr double O" << k << " = " << lit(hm.onsite[9 * b + k]) << ";\n";
        o << "    static __device__ __forceinline__ bool valid(int m0, int m1, int m2) { return m0 < " << hs.M[0] << " && m1 < " << hs.M[1] << " && m2 < " << hs.M[2] << "; }\n";
        o << "    static __device__ __forceinline__ int pos(int m0, int m1, int m2) { return " << hs.start << " + (m0 * " << hs.M[1] << " + m1) * " << hs.M[2] << " + m2; }\n";
        // linear index -> supercell coordinates (energy kernel)
        o << "    static __device__ __forceinline__ void locate(int idx, int &m0, int &m1, int &m2) {\n";
        o << "        m2 = idx % " << hs.M[2] << "; const int t = idx / " << hs.M[2] << "; m1 = t % " << hs.M[1] << "; m0 = t / " << hs.M[1] << ";\n    }\n";
        // reference site index (Philox counter)
        o << "    static __device__ __forceinline__ unsigned site(int m0, int m1, int m2) {\n";
        o << "        return (unsigned)(((" << b << " * " << hm.L[0] << " + (m0 * " << hs.P[0] << " + " << hs.r[0] << ")) * " << hm.L[1]
          << " + (m1 * " << hs.P[1] << " + " << hs.r[1] << ")) * " << hm.L[2] << " + (m2 * " << hs.P[2] << " + " << hs.r[2] << "));\n    }\n";
        static const int preload_max = std::getenv("CSMC_JIT_PRELOAD_MAX") ? std::atoi(std::getenv("CSMC_JIT_PRELOAD_MAX")) : 64;
        const bool preload = nnb <= preload_max;
        seg_preload[s] = preload ? 1 : 0;
        o << "    static constexpr bool PRELOAD = " << (preload ? "true" : "false") << ";\n";
        static const bool staged = !(std::getenv("CSMC_JIT_CONTRACT") && std::string(std::getenv("CSMC_JIT_CONTRACT")) == "outer");
        // sum_k x_k * y_k (+ init) as an explicit fma chain: the rounding order is fixed by the generator, so
        // the same term rounds identically in every kernel it is inlined into
        auto chain = [this](const std::vector<std::pair<std::string, std::string>> &xy, const std::string &init) {
            std::string e = init;
            for (const auto &p : xy) {
                const bool unit = p.first == lit(1.0) || p.first == lit(-1.0);   // folded into an add / a negation by the compiler
                flops_cur += e.empty() ? (unit ? 0 : 1) : (unit ? 1 : 2);
                e = e.empty() ? p.first + " * " + p.second : "fma(" + p.first + ", " + p.second + ", " + e + ")";
            }
            return e;
        };
        flops_cur = 0;
        using Terms = std::vector<std::pair<std::string, std::string>>;
        auto emit_accumulate = [&](const HostTerm &t) {
            const double *C = hm.coefs.data() + t.coef;
            if (t.kind == 2) {
                for (int a = 0; a < 3; ++a) {
                    Terms e;
                    for (int c = 0; c < 3; ++c)
                        if (C[3 * a + c] != 0.0) e.push_back({coef(C[3 * a + c]), "p" + std::to_string(c)});
                    if (!e.empty()) o << "        a" << a << " = " << chain(e, "a" + std::to_string(a)) << ";\n";
                }
            } else if (staged && t.kind == 3) {
                // b_a += sum_b (sum_c C[a,b,c] q_c) p_b : innermost index first, few live temporaries
                for (int a = 0; a < 3; ++a) {
                    Terms outer;
                    for (int bb = 0; bb < 3; ++bb) {
                        Terms inner;
                        for (int c = 0; c < 3; ++c)
                            if (C[a * 9 + bb * 3 + c] != 0.0) inner.push_back({coef(C[a * 9 + bb * 3 + c]), "q" + std::to_string(c)});
                        if (!inner.empty()) outer.push_back({"(" + chain(inner, "") + ")", "p" + std::to_string(bb)});
                    }
                    if (!outer.empty()) o << "        b" << a << " = " << chain(outer, "b" + std::to_string(a)) << ";\n";
                }
            } else if (staged && t.kind == 4) {
                // c_a += sum_b (sum_c (sum_d R[a,b,c,d] w_d) q_c) p_b
                for (int a = 0; a < 3; ++a) {
                    Terms sum_b;
                    for (int bb = 0; bb < 3; ++bb) {
                        Terms sum_c;
                        for (int c = 0; c < 3; ++c) {
                            Terms sum_d;
                            for (int d = 0; d < 3; ++d)
                                if (C[a * 27 + bb * 9 + c * 3 + d] != 0.0)
                                    sum_d.push_back({coef(C[a * 27 + bb * 9 + c * 3 + d]), "w" + std::to_string(d)});
                            if (!sum_d.empty()) sum_c.push_back({"(" + chain(sum_d, "") + ")", "q" + std::to_string(c)});
                        }
                        if (!sum_c.empty()) {
                            o << "        { const double t" << bb << " = " << chain(sum_c, "") << ";";
                            sum_b.push_back({"t" + std::to_string(bb), "p" + std::to_string(bb)});
                        } else {
                            o << "        {";
                        }
                        o << "\n";
                    }
                    if (!sum_b.empty()) o << "        c" << a << " = " << chain(sum_b, "c" + std::to_string(a)) << ";\n";
                    o << "        }}}\n";
                }
            } else if (t.kind == 3) {
                for (int bb = 0; bb < 3; ++bb)
                    for (int c = 0; c < 3; ++c) {
                        bool used = false;
                        for (int a = 0; a < 3; ++a) used |= (C[a * 9 + bb * 3 + c] != 0.0);
                        if (!used) continue;
                        o << "        { const double v = p" << bb << " * q" << c << ";";
                        for (int a = 0; a < 3; ++a)
                            if (C[a * 9 + bb * 3 + c] != 0.0) o << " b" << a << " += " << coef(C[a * 9 + bb * 3 + c]) << " * v;";
                        o << " }\n";
                    }
            } else {
                for (int bb = 0; bb < 3; ++bb)
                    for (int c = 0; c < 3; ++c) {
                        bool used_bc = false;
                        for (int d = 0; d < 3; ++d) for (int a = 0; a < 3; ++a) used_bc |= (C[a * 27 + bb * 9 + c * 3 + d] != 0.0);
                        if (!used_bc) continue;
                        o << "        { const double vbc = p" << bb << " * q" << c << ";\n";
                        for (int d = 0; d < 3; ++d) {
                            bool used = false;
                            for (int a = 0; a < 3; ++a) used |= (C[a * 27 + bb * 9 + c * 3 + d] != 0.0);
                            if (!used) continue;
                            o << "          { const double v = vbc * w" << d << ";";
                            for (int a = 0; a < 3; ++a)
                                if (C[a * 27 + bb * 9 + c * 3 + d] != 0.0) o << " c" << a << " += " << coef(C[a * 27 + bb * 9 + c * 3 + d]) << " * v;";
                            o << " }\n";
                        }
                        o << "        }\n";
                    }
            }
        };
        const char *nm[3] = {"p", "q", "w"};
        // phase 1 (PRELOAD): all neighbour loads up front (read-only during the pass: other colours) via ld.global.nc
        o << "    template <bool NC, int PART> static __device__ __forceinline__ void load(const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ sz,\n"
             "            int m0, int m1, int m2, double (&nb)[PRELOAD ? 3 * NNB + 1 : 1], unsigned &ok) {\n";
        if (preload)
            for (const auto &lv : live) {
                const HostTerm &t = *lv.t;
                o << "        if (PART < 0 || PART == " << lv.part << ") { // slot " << lv.bit << " kind " << t.kind << "\n";
                if (!hm.periodic) o << "          bool okt = true;\n";
                for (int k = 0; k < t.kind - 1; ++k) neighbour(hs, t, k);
                if (!hm.periodic) o << "          if (okt) {\n";
                for (int k = 0; k < t.kind - 1; ++k) {
                    const int base = 3 * (lv.slot + k);
                    o << "          nb[" << base << "] = ld<NC>(sx + j" << k << "); nb[" << base + 1 << "] = ld<NC>(sy + j" << k << "); nb[" << base + 2 << "] = ld<NC>(sz + j" << k << ");\n";
                }
                if (!hm.periodic) {
                    o << "          } else {\n            ok &= ~(1u << " << lv.bit << ");\n";
                    for (int k = 0; k < t.kind - 1; ++k) {
                        const int base = 3 * (lv.slot + k);
                        o << "            nb[" << base << "] = 0.0; nb[" << base + 1 << "] = 0.0; nb[" << base + 2 << "] = 0.0;\n";
                    }
                    o << "          }\n";
                }
                o << "        }\n";
            }
        o << "    }\n";
        // the same addresses as load(), as L1 prefetches (no destination registers): a CTA that handles several tiles pulls
        // the next tile's operands into L1 while it works on the current one
        o << "    static __device__ __forceinline__ void prefetch(const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ sz,\n"
             "            int m0, int m1, int m2) {\n";
        if (preload)
            for (const auto &lv : live) {
                const HostTerm &t = *lv.t;
                o << "        { // slot " << lv.bit << "\n";
                if (!hm.periodic) o << "          bool okt = true;\n";
                for (int k = 0; k < t.kind - 1; ++k) neighbour(hs, t, k);
                if (!hm.periodic) o << "          if (okt) {\n";
                for (int k = 0; k < t.kind - 1; ++k) o << "          pf_l1(sx + j" << k << "); pf_l1(sy + j" << k << "); pf_l1(sz + j" << k << ");\n";
                if (!hm.periodic) o << "          }\n";
                o << "        }\n";
            }
        o << "    }\n";
        // phase 2 (PRELOAD): unrolled neighbour field from registers: a* bilinear, b* cubic, c* quartic accumulators
        o << "    template <int PART> static __device__ __forceinline__ void field(const double (&nb)[PRELOAD ? 3 * NNB + 1 : 1], unsigned ok,\n"
             "            double &a0, double &a1, double &a2, double &b0, double &b1, double &b2, double &c0, double &c1, double &c2) {\n";
        if (preload)
            for (const auto &lv : live) {
                const HostTerm &t = *lv.t;
                o << "      if ((PART < 0 || PART == " << lv.part << ")" << (hm.periodic ? "" : " && (ok & (1u << " + std::to_string(lv.bit) + "))") << ") { // slot " << lv.bit << "\n";
                for (int k = 0; k < t.kind - 1; ++k) {
                    const int base = 3 * (lv.slot + k);
                    o << "        const double " << nm[k] << "0 = nb[" << base << "], " << nm[k] << "1 = nb[" << base + 1 << "], " << nm[k] << "2 = nb[" << base + 2 << "];\n";
                }
                emit_accumulate(t);
                o << "      }\n";
            }
        o << "    }\n";
        // streaming variant (many neighbours: keeping them all live would cost occupancy): gather and
        // accumulate slot by slot
        o << "    template <bool NC> static __device__ __forceinline__ void field_stream(const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ sz,\n"
             "            int m0, int m1, int m2, double &a0, double &a1, double &a2, double &b0, double &b1, double &b2, double &c0, double &c1, double &c2) {\n";
        if (!preload)
            for (const auto &lv : live) {
                const HostTerm &t = *lv.t;
                o << "      { // slot " << lv.bit << " kind " << t.kind << "\n";
                if (!hm.periodic) o << "          bool okt = true;\n";
                for (int k = 0; k < t.kind - 1; ++k) neighbour(hs, t, k);
                if (!hm.periodic) o << "        if (okt) {\n";
                for (int k = 0; k < t.kind - 1; ++k)
                    o << "        const double " << nm[k] << "0 = ld<NC>(sx + j" << k << "), " << nm[k] << "1 = ld<NC>(sy + j" << k << "), " << nm[k] << "2 = ld<NC>(sz + j" << k << ");\n";
                emit_accumulate(t);
                if (!hm.periodic) o << "        }\n";
                o << "      }\n";
            }
        o << "    }\n};\n\n";
        seg_flops.resize(hm.segs.size(), 0);
        seg_flops[s] = flops_cur;
    }

    // neighbour class and supercell shift of neighbour k of term t seen from class hs
    int nbr_class(const HostSeg &hs, const HostTerm &t, int k, int (&delta)[MAXD]) {
        int r2[MAXD] = {0, 0, 0};
        for (int d = 0; d < MAXD; ++d) delta[d] = 0;
        for (int d = 0; d < hm.D; ++d) {
            r2[d] = posmod(hs.r[d] + t.off[k][d], hm.P[d]);
            delta[d] = floordiv(hs.r[d] + t.off[k][d], hm.P[d]);
        }
        auto it = seg_of_class.find(std::make_tuple(t.nb_basis[k], r2[0], r2[1], r2[2]));
        return it == seg_of_class.end() ? -1 : it->second;
    }

    // ---- fused full-sweep kernel (two-colour periodic models) ------------------------------------------
    // One launch = one full sweep.  A CTA loads a tile of supercells plus the halo it needs into shared
    // memory, updates colour 0 on the tile extended by the ring colour 1 will read (recomputed by the
    // neighbouring CTAs with identical arithmetic, so identical values), then colour 1 on the tile, and
    // writes the tile to the *other* spin buffer (ping-pong: CTAs never read what another CTA writes in
    // the same launch).  Traffic per sweep: ~1.2 x 24 B read + 24 B written per site instead of
    // 2 x 72 B, and half the launches.
    bool emit_fused(JitPlan &plan) {
        plan.fused = false;
        if (!plan.want_fused || !hm.periodic || hm.n_colours != 2 || hm.D > 2) return false;
        const int nq = (int)hm.segs.size();
        for (int q = 0; q < nq; ++q) if (!seg_preload[q] || seg_split[q] != 1) return false;
        int W[MAXD] = {1, 1, 1}, NT[MAXD] = {1, 1, 1};
        int target[MAXD] = {hm.D == 1 ? 1024 : 16, 32, 1};
        int tpb = 256;
        if (const char *e = std::getenv("CSMC_JIT_FUSED_TILE")) std::sscanf(e, "%dx%d", &target[0], &target[1]);
        if (const char *e = std::getenv("CSMC_JIT_FUSED_TPB")) tpb = std::atoi(e);
        plan.fused_tpb = tpb;
        for (int d = 0; d < hm.D; ++d) {
            const int M = hm.segs[0].M[d];
            for (int q = 0; q < nq; ++q) if (hm.segs[q].M[d] != M) return false;
            int w = std::min(target[d], M);
            while (w > 1 && M % w != 0) --w;
            if (w < 4) return false;
            W[d] = w; NT[d] = M / w;
        }
        // extents: ext = update region of colour-0 classes beyond the tile, halo = data region
        std::vector<std::array<int, 2 * MAXD>> ext(nq), halo(nq);
        for (auto &e : ext) e.fill(0);
        for (auto &e : halo) e.fill(0);
        for (int q1 = 0; q1 < nq; ++q1) {
            if (hm.segs[q1].colour != 1) continue;
            for (const auto &lv : seg_live[q1])
                for (int k = 0; k < lv.t->kind - 1; ++k) {
                    int dl[MAXD];
                    const int q = nbr_class(hm.segs[q1], *lv.t, k, dl);
                    if (q < 0 || hm.segs[q].colour != 0) return false;
                    for (int d = 0; d < hm.D; ++d) {
                        if (std::abs(dl[d]) > 2) return false;
                        if (dl[d] < 0) ext[q][2 * d] = std::max(ext[q][2 * d], -dl[d]);
                        if (dl[d] > 0) ext[q][2 * d + 1] = std::max(ext[q][2 * d + 1], dl[d]);
                    }
                }
        }
        for (int q = 0; q < nq; ++q) halo[q] = ext[q];
        for (int q0 = 0; q0 < nq; ++q0) {
            if (hm.segs[q0].colour != 0) continue;
            for (const auto &lv : seg_live[q0])
                for (int k = 0; k < lv.t->kind - 1; ++k) {
                    int dl[MAXD];
                    const int q = nbr_class(hm.segs[q0], *lv.t, k, dl);
                    if (q < 0 || hm.segs[q].colour != 1) return false;
                    for (int d = 0; d < hm.D; ++d) {
                        if (std::abs(dl[d]) > 2) return false;
                        halo[q][2 * d] = std::max(halo[q][2 * d], ext[q0][2 * d] - dl[d]);
                        halo[q][2 * d + 1] = std::max(halo[q][2 * d + 1], ext[q0][2 * d + 1] + dl[d]);
                    }
                }
        }
        for (int q = 0; q < nq; ++q)
            for (int d = 0; d < hm.D; ++d)
                if (std::max(halo[q][2 * d], halo[q][2 * d + 1]) > hm.segs[q].M[d]) return false;   // single wrap-around only
        // shared-memory layout
        std::vector<int> shb(nq, 0);
        std::vector<std::array<int, MAXD>> S(nq);
        int total = 0;
        for (int q = 0; q < nq; ++q) {
            int n = 1;
            for (int d = 0; d < MAXD; ++d) { S[q][d] = d < hm.D ? W[d] + halo[q][2 * d] + halo[q][2 * d + 1] : 1; n *= S[q][d]; }
            shb[q] = total;
            total += (n + 1) / 2 * 2;
        }
        const size_t smem = (size_t)3 * total * sizeof(double);
        if (smem > 100 * 1024) return false;
        plan.fused = true;
        plan.fused_tiles = NT[0] * NT[1] * NT[2];
        plan.fused_smem = (int)smem;

        auto shpos = [&](int q, const std::string (&l)[MAXD]) {
            std::ostringstream e;
            e << shb[q] << " + ((" << l[0] << " + " << halo[q][0] << ") * " << S[q][1] << " + (" << l[1] << " + " << halo[q][2] << ")) * " << S[q][2]
              << " + (" << l[2] << " + " << halo[q][4] << ")";
            return e.str();
        };
        o << "#define SH_TOTAL " << total << "\n";
        // per class: tile-local neighbour loads + thin wrappers
        for (int q = 0; q < nq; ++q) {
            const HostSeg &hs = hm.segs[q];
            o << "struct Tile" << q << " {\n";
            o << "    static __device__ __forceinline__ int pos(int l0, int l1, int l2) { return " << shpos(q, {"l0", "l1", "l2"}) << "; }\n";
            o << "    static __device__ __forceinline__ void load(const double *__restrict__ shx, const double *__restrict__ shy, const double *__restrict__ shz,\n"
                 "            int l0, int l1, int l2, double (&nb)[3 * Seg" << q << "::NNB + 1]) {\n";
            for (const auto &lv : seg_live[q])
                for (int k = 0; k < lv.t->kind - 1; ++k) {
                    int dl[MAXD];
                    const int qn = nbr_class(hs, *lv.t, k, dl);
                    const std::string l[MAXD] = {"l0 + (" + std::to_string(dl[0]) + ")", "l1 + (" + std::to_string(dl[1]) + ")", "l2 + (" + std::to_string(dl[2]) + ")"};
                    const int base = 3 * (lv.slot + k);
                    o << "        { const int j = " << shpos(qn, l) << "; nb[" << base << "] = shx[j]; nb[" << base + 1 << "] = shy[j]; nb[" << base + 2 << "] = shz[j]; }\n";
                }
            o << "    }\n};\n";
        }
        for (int u = 0; u < 4; ++u) {
            o << "extern \"C\" __global__ void __launch_bounds__(" << tpb << ") csmc_fused_u" << u << "(const double *__restrict__ in, double *__restrict__ out, const SweepArgs a) {\n";
            o << "    extern __shared__ double sh[];\n    double *shx = sh, *shy = sh + SH_TOTAL, *shz = sh + 2 * SH_TOTAL;\n";
            o << "    const int rep = blockIdx.z + a.rep0;\n    int t = blockIdx.x;\n";
            o << "    const int t2 = t % " << NT[2] << "; t /= " << NT[2] << "; const int t1 = t % " << NT[1] << "; const int t0 = t / " << NT[1] << ";\n";
            o << "    const int o0 = t0 * " << W[0] << ", o1 = t1 * " << W[1] << ", o2 = t2 * " << W[2] << ";\n";
            o << "    const double *gx = in + (size_t)rep * (3ull * NPAD), *gy = gx + NPAD, *gz = gy + NPAD;\n";
            // load phase
            for (int q = 0; q < nq; ++q) {
                const HostSeg &hs = hm.segs[q];
                const int n = S[q][0] * S[q][1] * S[q][2];
                o << "    for (int e = threadIdx.x; e < " << n << "; e += " << tpb << ") {\n";
                o << "        const int e2 = e % " << S[q][2] << ", e1 = (e / " << S[q][2] << ") % " << S[q][1] << ", e0 = e / " << (S[q][2] * S[q][1]) << ";\n";
                o << "        int m0 = o0 + e0 - " << halo[q][0] << ", m1 = o1 + e1 - " << halo[q][2] << ", m2 = o2 + e2 - " << halo[q][4] << ";\n";
                for (int d = 0; d < hm.D; ++d)
                    o << "        m" << d << " = m" << d << " < 0 ? m" << d << " + " << hs.M[d] << " : (m" << d << " >= " << hs.M[d] << " ? m" << d << " - " << hs.M[d] << " : m" << d << ");\n";
                o << "        const int g = Seg" << q << "::pos(m0, m1, m2), j = " << shb[q] << " + e;\n";
                o << "        shx[j] = __ldg(gx + g); shy[j] = __ldg(gy + g); shz[j] = __ldg(gz + g);\n    }\n";
            }
            o << "    __syncthreads();\n    int n_acc = 0;\n";
            // update phases
            for (int c = 0; c < 2; ++c) {
                for (int q = 0; q < nq; ++q) {
                    if (hm.segs[q].colour != c) continue;
                    const HostSeg &hs = hm.segs[q];
                    int U[MAXD], lo[MAXD];
                    for (int d = 0; d < MAXD; ++d) { lo[d] = (c == 0 && d < hm.D) ? ext[q][2 * d] : 0; U[d] = d < hm.D ? W[d] + lo[d] + ((c == 0) ? ext[q][2 * d + 1] : 0) : 1; }
                    const int n = U[0] * U[1] * U[2];
                    o << "    for (int e = threadIdx.x; e < " << n << "; e += " << tpb << ") {\n";
                    o << "        const int l2 = e % " << U[2] << " - " << lo[2] << ", l1 = (e / " << U[2] << ") % " << U[1] << " - " << lo[1] << ", l0 = e / " << (U[2] * U[1]) << " - " << lo[0] << ";\n";
                    o << "        Site<Seg" << q << "> d;\n        d.valid = true; d.ok = 0xffffffffu;\n";
                    o << "        d.m0 = o0 + l0; d.m1 = o1 + l1; d.m2 = o2 + l2;\n";
                    for (int dd = 0; dd < hm.D; ++dd)
                        o << "        d.m" << dd << " = d.m" << dd << " < 0 ? d.m" << dd << " + " << hs.M[dd] << " : (d.m" << dd << " >= " << hs.M[dd] << " ? d.m" << dd << " - " << hs.M[dd] << " : d.m" << dd << ");\n";
                    o << "        d.pos = Tile" << q << "::pos(l0, l1, l2);\n";
                    o << "        d.s0 = shx[d.pos]; d.s1 = shy[d.pos]; d.s2 = shz[d.pos];\n";
                    o << "        Tile" << q << "::load(shx, shy, shz, l0, l1, l2, d.nb);\n";
                    o << "        const bool acc = site_finish_ptr<" << u << ", Seg" << q << ", false, -1>(d, shx, shy, shz, rep, a, 0ULL, 0.0, 0.0, 0.0);\n";
                    if (c == 0) {
                        o << "        if (acc && l0 >= 0 && l0 < " << W[0] << " && l1 >= 0 && l1 < " << W[1] << " && l2 >= 0 && l2 < " << W[2] << ") ++n_acc;\n";
                    } else {
                        o << "        if (acc) ++n_acc;\n";
                    }
                    o << "    }\n";
                }
                o << "    __syncthreads();\n";
            }
            // store phase
            o << "    double *hx = out + (size_t)rep * (3ull * NPAD), *hy = hx + NPAD, *hz = hy + NPAD;\n";
            for (int q = 0; q < nq; ++q) {
                const int n = W[0] * W[1] * W[2];
                o << "    for (int e = threadIdx.x; e < " << n << "; e += " << tpb << ") {\n";
                o << "        const int l2 = e % " << W[2] << ", l1 = (e / " << W[2] << ") % " << W[1] << ", l0 = e / " << (W[2] * W[1]) << ";\n";
                o << "        const int g = Seg" << q << "::pos(o0 + l0, o1 + l1, o2 + l2), j = Tile" << q << "::pos(l0, l1, l2);\n";
                o << "        hx[g] = shx[j]; hy[g] = shy[j]; hz[g] = shz[j];\n    }\n";
            }
            if (u >= 2) o << "    count_accepted(n_acc, rep, a);\n";
            o << "}\n";
        }
        return true;
    }

    // ---- tile-resident persistent kernel ---------------------------------------------------------------------
    // A per-colour pass over an L2-resident lattice is mostly fixed latency (launch, ramp, load latency, tail:
    // issue slots 25 % busy), so a sequence of sweeps is bound by the number of launches, not by bytes.  This kernel
    // removes the launches: the lattice of each replica is cut into CTA tiles over supercell coordinates (dimensions
    // 0 and 1; dimension 2 stays whole), one CTA per SM loads its tile plus the halo its sites read into shared
    // memory ONCE and then runs every colour pass of the whole sequence (n sweeps x C colours) on it.  After a pass
    // a CTA stores the sites other tiles read ("export region", within halo reach of the tile edge) to their home
    // in the global spin array and raises its progress counter (st.release.gpu); before a pass it waits for the
    // counters of its <= 8 neighbour tiles (ld.acquire.gpu, no grid-wide barrier) and reloads the halo cells of
    // the colour updated in the previous pass.  Same per-site arithmetic and Philox counters as the pass kernels
    // (site_finish_ptr), so results are bit-identical.  Launched cooperatively (co-residency guaranteed).
    bool emit_persist(JitPlan &plan) {
        plan.persist = false;
        if (plan.persist_replicas < 1 || !hm.periodic || hm.n_colours < 2) return false;
        const int nq = (int)hm.segs.size();
        int M[MAXD] = {1, 1, 1};
        for (int d = 0; d < MAXD; ++d) M[d] = hm.segs[0].M[d];
        int nnb_max = 0, nnb_colour_max = 0;
        for (int q = 0; q < nq; ++q) {
            if (!seg_preload[q] || seg_split[q] != 1) return false;
            for (int d = 0; d < MAXD; ++d) if (hm.segs[q].M[d] != M[d]) return false;
            int nnb = 0;
            for (const auto &lv : seg_live[q]) nnb += lv.t->kind - 1;
            nnb_max = std::max(nnb_max, nnb);
        }
        for (int c = 0; c < hm.n_colours; ++c) {
            int nnb = 0;
            for (int q = hm.colour_seg_begin[c]; q < hm.colour_seg_begin[c + 1]; ++q)
                for (const auto &lv : seg_live[q]) nnb += lv.t->kind - 1;
            nnb_colour_max = std::max(nnb_colour_max, nnb);
        }
        // halo of class q: how far beyond a tile the sites of the tile read class q (per tiled dimension and side)
        std::vector<std::array<int, 4>> halo(nq);
        for (auto &e : halo) e.fill(0);
        int reach2 = 0;
        for (int q = 0; q < nq; ++q)
            for (const auto &lv : seg_live[q])
                for (int k = 0; k < lv.t->kind - 1; ++k) {
                    int dl[MAXD];
                    const int qn = nbr_class(hm.segs[q], *lv.t, k, dl);
                    if (qn < 0) return false;
                    if (hm.segs[qn].colour == hm.segs[q].colour) return false;   // not a proper colouring for this scheme
                    for (int d = 0; d < 2; ++d) {
                        if (dl[d] < 0) halo[qn][2 * d] = std::max(halo[qn][2 * d], -dl[d]);
                        if (dl[d] > 0) halo[qn][2 * d + 1] = std::max(halo[qn][2 * d + 1], dl[d]);
                    }
                    reach2 = std::max(reach2, std::abs(dl[2]));
                }
        if (reach2 >= M[2] && M[2] > 1) return false;        // single wrap along the untiled dimension
        // every class is stored with the same padded tile geometry (halo = the largest any class needs), so that a
        // neighbour's shared-memory address is the site's own cell index plus a compile-time constant: lanes that walk
        // consecutive cells read consecutive addresses for every operand (no bank conflicts)
        int H[4] = {0, 0, 0, 0};
        for (int q = 0; q < nq; ++q) for (int k = 0; k < 4; ++k) H[k] = std::max(H[k], halo[q][k]);
        const int hmax[2] = {std::max(H[0], H[1]), std::max(H[2], H[3])};
        // classes of a colour fused per thread (shared neighbour loads are read once) while the registers allow it
        int fuse = nnb_colour_max <= 12 ? 2 : 1;
        if (const char *e = std::getenv("CSMC_PERSIST_FUSE")) fuse = std::max(1, std::atoi(e));
        int tpb = nnb_max * fuse <= 16 ? 512 : 256;
        if (const char *e = std::getenv("CSMC_PERSIST_TPB")) { const int v = std::atoi(e); if (v == 128 || v == 256 || v == 384 || v == 512 || v == 768 || v == 1024) tpb = v; }
        // tiling: minimise launches x loop iterations per pass (whole warps of cells, padded columns included)
        const int R = plan.persist_replicas, NSM = std::max(1, plan.persist_sms);
        double best_cost = 1e300;
        int bg[2] = {0, 0}, bw[2] = {0, 0};
        int force_g[2] = {0, 0};
        if (const char *e = std::getenv("CSMC_PERSIST_GRID")) std::sscanf(e, "%dx%d", &force_g[0], &force_g[1]);
        for (int g0 = 1; g0 <= std::min(M[0], NSM); ++g0)
            for (int g1 = 1; g1 <= std::min(M[1], NSM / g0); ++g1) {
                if (force_g[0] > 0 && (g0 != force_g[0] || g1 != force_g[1])) continue;
                const int g[2] = {g0, g1};
                int w[2];
                bool ok = true;
                for (int d = 0; d < 2; ++d) {
                    w[d] = (M[d] + g[d] - 1) / g[d];
                    if ((M[d] + w[d] - 1) / w[d] != g[d]) ok = false;                 // g tiles of extent w cover M exactly
                    const int last = M[d] - (g[d] - 1) * w[d];
                    if (last < std::max(1, hmax[d])) ok = false;                      // halo reaches one tile only
                    if (g[d] == 1 && hmax[d] > 0 && M[d] < 2 * hmax[d]) ok = false;   // tile wraps onto itself
                }
                if (!ok) continue;
                const long s0 = w[0] + H[0] + H[1], s1 = w[1] + H[2] + H[3];
                const long cells = (s0 * s1 * M[2] + 1) / 2 * 2;
                const size_t smem = (size_t)3 * cells * nq * sizeof(double);
                if (smem + 1024 > (size_t)plan.persist_smem_max) continue;
                const int T = g0 * g1, nrep = std::min(R, NSM / T);
                if (nrep < 1) continue;
                const int launches = (R + nrep - 1) / nrep;
                const long iters = ((long)w[0] * s1 * M[2] + tpb - 1) / tpb;          // per class group and pass
                const double cost = (double)launches * ((double)iters * tpb + 0.05 * (double)(s0 * s1 - (long)w[0] * w[1]) * M[2]);
                if (cost < best_cost) { best_cost = cost; bg[0] = g0; bg[1] = g1; bw[0] = w[0]; bw[1] = w[1]; }
            }
        if (bg[0] == 0) return false;
        const int G0 = bg[0], G1 = bg[1], W0 = bw[0], W1 = bw[1], T = G0 * G1;
        const int S0 = W0 + H[0] + H[1], S1 = W1 + H[2] + H[3], S2 = M[2];
        const int CELLS = (S0 * S1 * S2 + 1) / 2 * 2;
        plan.persist = true;
        plan.persist_tiles = T;
        plan.persist_nrep = std::min(R, NSM / T);
        plan.persist_smem = 3 * CELLS * nq * (int)sizeof(double);
        plan.persist_tpb = tpb;
        plan.persist_g[0] = G0; plan.persist_g[1] = G1; plan.persist_w[0] = W0; plan.persist_w[1] = W1;

        o << "\n// ---- tile-resident persistent kernel: " << G0 << " x " << G1 << " tiles of " << W0 << " x " << W1 << " x " << S2
          << " supercells per replica (padded " << S0 << " x " << S1 << "), " << plan.persist_smem << " B of shared memory, " << tpb << " threads, "
          << fuse << " class(es) per thread\n";
        o << "#define PT_CELLS " << CELLS << "\n#define PT_TOTAL " << (CELLS * nq) << "\n#define PT_TPB " << tpb << "\n";
        o << "#define PT_S1 " << S1 << "\n#define PT_S2 " << S2 << "\n#define PT_H0 " << H[0] << "\n#define PT_H1 " << H[2] << "\n";
        // neighbour loads of class q for the site in padded cell `cell` (row r = l0 + PT_H0, column col = l1 + PT_H1, l2)
        for (int q = 0; q < nq; ++q) {
            const HostSeg &hs = hm.segs[q];
            o << "__device__ __forceinline__ void pt_load" << q << "(const double *__restrict__ shx, const double *__restrict__ shy, const double *__restrict__ shz,\n"
                 "        int cell, int l2, double (&nb)[3 * Seg" << q << "::NNB + 1]) {\n";
            for (const auto &lv : seg_live[q])
                for (int k = 0; k < lv.t->kind - 1; ++k) {
                    int dl[MAXD];
                    const int qn = nbr_class(hs, *lv.t, k, dl);
                    const int b3 = 3 * (lv.slot + k);
                    const long off = (long)qn * CELLS + ((long)dl[0] * S1 + dl[1]) * S2 + dl[2];
                    std::string wrap;
                    if (dl[2] > 0) wrap = " + (l2 + " + std::to_string(dl[2]) + " >= " + std::to_string(S2) + " ? " + std::to_string(-S2) + " : 0)";
                    if (dl[2] < 0) wrap = " + (l2 - " + std::to_string(-dl[2]) + " < 0 ? " + std::to_string(S2) + " : 0)";
                    o << "    { const int j = cell + (" << off << ")" << wrap << "; nb[" << b3 << "] = shx[j]; nb[" << b3 + 1 << "] = shy[j]; nb[" << b3 + 2 << "] = shz[j]; }\n";
                }
            o << "}\n";
        }
        // one colour pass on the tile: every thread walks padded cells of the core rows; the classes of a colour are taken
        // `fuse` at a time: all loads of the group first (shared neighbours are read once), then the updates
        o << "template <int UPD> __device__ __forceinline__ int pt_pass(int colour, double *shx, double *shy, double *shz, double *gx, double *gy, double *gz,\n"
             "        int o0, int o1, int w0, int w1, int rep, const SweepArgs &a, unsigned long long ctr_extra) {\n    int n_acc = 0;\n    switch (colour) {\n";
        for (int c = 0; c < hm.n_colours; ++c) {
            o << "    case " << c << ": {\n";
            const int q0 = hm.colour_seg_begin[c], q1 = hm.colour_seg_begin[c + 1];
            for (int qa = q0; qa < q1; qa += fuse) {
                const int qb = std::min(q1, qa + fuse);
                o << "        for (int e = threadIdx.x; e < w0 * " << (S1 * S2) << "; e += PT_TPB) {\n";
                o << "            const int l2 = e % " << S2 << ", col = (e / " << S2 << ") % " << S1 << ", l0 = e / " << (S1 * S2) << ", l1 = col - " << H[2] << ";\n";
                o << "            if (l1 < 0 || l1 >= w1) continue;\n";
                o << "            const int cell = e + " << (H[0] * S1 * S2) << ";\n";
                for (int q = qa; q < qb; ++q) {
                    o << "            Site<Seg" << q << "> d" << q << ";\n            d" << q << ".valid = true; d" << q << ".ok = 0xffffffffu; d" << q << ".m0 = o0 + l0; d" << q
                      << ".m1 = o1 + l1; d" << q << ".m2 = l2;\n";
                    o << "            d" << q << ".pos = cell + " << ((long)q * CELLS) << ";\n            d" << q << ".s0 = shx[d" << q << ".pos]; d" << q << ".s1 = shy[d" << q
                      << ".pos]; d" << q << ".s2 = shz[d" << q << ".pos];\n";
                    o << "            pt_load" << q << "(shx, shy, shz, cell, l2, d" << q << ".nb);\n";
                }
                for (int q = qa; q < qb; ++q) {
                    o << "            n_acc += site_finish_ptr<UPD, Seg" << q << ", false, -1>(d" << q << ", shx, shy, shz, rep, a, ctr_extra, 0.0, 0.0, 0.0) ? 1 : 0;\n";
                    const bool exported = halo[q][0] || halo[q][1] || halo[q][2] || halo[q][3];
                    if (exported) {
                        // sites other tiles keep in their halo: upper halo of the tile below = my first rows, ...
                        o << "            if (l0 < " << halo[q][1] << " || l0 >= w0 - " << halo[q][0] << " || l1 < " << halo[q][3] << " || l1 >= w1 - " << halo[q][2] << ") {\n";
                        o << "                const int g = Seg" << q << "::pos(o0 + l0, o1 + l1, l2);\n";
                        o << "                gx[g] = shx[d" << q << ".pos]; gy[g] = shy[d" << q << ".pos]; gz[g] = shz[d" << q << ".pos];\n            }\n";
                    }
                }
                o << "        }\n";
            }
            o << "    } break;\n";
        }
        o << "    default: break;\n    }\n    return n_acc;\n}\n";
        // halo cells of the classes of one colour, re-read from their home in global memory (L2: ld.global.cg).  One flat
        // loop over every band of every class, so that the loads of a thread are independent and the pass pays one round trip.
        o << "__device__ __forceinline__ void pt_reload(int colour, double *shx, double *shy, double *shz, const double *gx, const double *gy, const double *gz,\n"
             "        int o0, int o1, int w0, int w1) {\n    switch (colour) {\n";
        for (int c = 0; c < hm.n_colours; ++c) {
            o << "    case " << c << ": {\n";
            // bands: (class, first row r0 [runtime expr], rows [expr], first col c0 [expr], cols [expr])
            struct Band { int q; std::string r0, nr, c0, nc; };
            std::vector<Band> bands;
            for (int q = hm.colour_seg_begin[c]; q < hm.colour_seg_begin[c + 1]; ++q) {
                const int HL0 = halo[q][0], HH0 = halo[q][1], HL1 = halo[q][2], HH1 = halo[q][3];
                const std::string cfull0 = std::to_string(H[2] - HL1), cfulln = "(w1 + " + std::to_string(HL1 + HH1) + ")";
                if (HL0) bands.push_back({q, std::to_string(H[0] - HL0), std::to_string(HL0), cfull0, cfulln});
                if (HH0) bands.push_back({q, "(" + std::to_string(H[0]) + " + w0)", std::to_string(HH0), cfull0, cfulln});
                if (HL1) bands.push_back({q, std::to_string(H[0]), "w0", std::to_string(H[2] - HL1), std::to_string(HL1)});
                if (HH1) bands.push_back({q, std::to_string(H[0]), "w0", "(" + std::to_string(H[2]) + " + w1)", std::to_string(HH1)});
            }
            if (!bands.empty()) {
                o << "        int n_[" << bands.size() + 1 << "];\n        n_[0] = 0;\n";
                for (size_t b = 0; b < bands.size(); ++b)
                    o << "        n_[" << b + 1 << "] = n_[" << b << "] + " << bands[b].nr << " * " << bands[b].nc << " * " << S2 << ";\n";
                o << "        for (int i = threadIdx.x; i < n_[" << bands.size() << "]; i += PT_TPB) {\n";
                o << "            int g, j;\n";
                for (size_t b = 0; b < bands.size(); ++b) {
                    const Band &B = bands[b];
                    o << "            " << (b ? "else " : "") << (b + 1 < bands.size() ? "if (i < n_[" + std::to_string(b + 1) + "]) " : "") << "{\n";
                    o << "                const int k = i - n_[" << b << "], e2 = k % " << S2 << ", t = k / " << S2 << ", cc = t % " << B.nc << ", rr = t / " << B.nc << ";\n";
                    o << "                const int r = " << B.r0 << " + rr, col = " << B.c0 << " + cc;\n";
                    o << "                int m0 = o0 + r - " << H[0] << ", m1 = o1 + col - " << H[2] << ";\n";
                    o << "                m0 = m0 < 0 ? m0 + " << M[0] << " : (m0 >= " << M[0] << " ? m0 - " << M[0] << " : m0); m1 = m1 < 0 ? m1 + " << M[1] << " : (m1 >= " << M[1] << " ? m1 - " << M[1] << " : m1);\n";
                    o << "                g = Seg" << B.q << "::pos(m0, m1, e2); j = " << ((long)B.q * CELLS) << " + (r * " << S1 << " + col) * " << S2 << " + e2;\n";
                    o << "            }\n";
                }
                o << "            shx[j] = __ldcg(gx + g); shy[j] = __ldcg(gy + g); shz[j] = __ldcg(gz + g);\n        }\n";
            }
            o << "    } break;\n";
        }
        o << "    default: break;\n    }\n}\n";
        o << "extern \"C\" __global__ void __launch_bounds__(PT_TPB, 1) csmc_persist(double *spins, const SweepArgs a, const PersistArgs pa) {\n";
        o << "    extern __shared__ double sh[];\n    double *shx = sh, *shy = sh + PT_TOTAL, *shz = sh + 2 * PT_TOTAL;\n";
        o << "    __shared__ int sh_acc;\n";
        o << "    const int tile = blockIdx.x, rep = blockIdx.y + a.rep0;\n";
        o << "    const int t1 = tile % " << G1 << ", t0 = tile / " << G1 << ";\n";
        o << "    const int o0 = t0 * " << W0 << ", o1 = t1 * " << W1 << ";\n";
        o << "    const int w0 = min(" << W0 << ", " << M[0] << " - o0), w1 = min(" << W1 << ", " << M[1] << " - o1);\n";
        o << "    double *gx = spins + (size_t)rep * (3ull * NPAD), *gy = gx + NPAD, *gz = gy + NPAD;\n";
        o << "    unsigned long long *flags = pa.flags + (size_t)rep * " << (T) << " * PERSIST_FLAG_STRIDE;\n";
        o << "    const unsigned long long base = flags[tile * PERSIST_FLAG_STRIDE];\n";
        // the (up to 8) neighbour tiles, one polling thread each
        o << "    const unsigned long long *nbr_flag = nullptr;\n";
        o << "    if (threadIdx.x < 8) {\n        const int k = threadIdx.x < 4 ? threadIdx.x : threadIdx.x + 1;   // 3 x 3 neighbourhood without the centre\n";
        o << "        const int n0 = (t0 + k / 3 - 1 + " << G0 << ") % " << G0 << ", n1 = (t1 + k % 3 - 1 + " << G1 << ") % " << G1 << ";\n";
        o << "        nbr_flag = flags + (n0 * " << G1 << " + n1) * PERSIST_FLAG_STRIDE;\n    }\n";
        o << "    if (threadIdx.x == 0) sh_acc = 0;\n";
        // load the tile and the halo each class needs (asynchronous copies: all of a thread's loads in flight at once)
        for (int q = 0; q < nq; ++q) {
            const int r_lo = H[0] - halo[q][0], c_lo = H[2] - halo[q][2];
            o << "    for (int e = threadIdx.x; e < (w0 + " << (halo[q][0] + halo[q][1]) << ") * " << (S1 * S2) << "; e += PT_TPB) {\n";
            o << "        const int e2 = e % " << S2 << ", col = (e / " << S2 << ") % " << S1 << ", r = " << r_lo << " + e / " << (S1 * S2) << ";\n";
            o << "        if (col < " << c_lo << " || col >= " << H[2] << " + w1 + " << halo[q][3] << ") continue;\n";
            o << "        int m0 = o0 + r - " << H[0] << ", m1 = o1 + col - " << H[2] << ";\n";
            o << "        m0 = m0 < 0 ? m0 + " << M[0] << " : (m0 >= " << M[0] << " ? m0 - " << M[0] << " : m0); m1 = m1 < 0 ? m1 + " << M[1] << " : (m1 >= " << M[1] << " ? m1 - " << M[1] << " : m1);\n";
            o << "        const int g = Seg" << q << "::pos(m0, m1, e2), j = " << ((long)q * CELLS) << " + (r * " << S1 << " + col) * " << S2 << " + e2;\n";
            o << "        cp_async8(shx + j, gx + g); cp_async8(shy + j, gy + g); cp_async8(shz + j, gz + g);\n    }\n";
        }
        o << "    cp_async_wait_all();\n    __syncthreads();\n";
        o << "    if (threadIdx.x == 0) st_release_gpu(flags + tile * PERSIST_FLAG_STRIDE, base + 1ULL);\n";
        o << "    int n_acc = 0;\n    unsigned long long step = 0;\n";
        o << "    long long pf_t = clock64(), pf[4] = {0, 0, 0, 0};\n";
        o << "#define PT_PROF(k) do { if (pa.prof) { const long long now_ = clock64(); pf[k] += now_ - pf_t; pf_t = now_; } } while (0)\n";
        o << "    for (int op = 0; op < pa.n_ops; ++op) {\n        const int upd = pa.upd[op];\n        const unsigned long long ce = pa.ctr_rel[op];\n";
        o << "        for (int c = 0; c < " << hm.n_colours << "; ++c, ++step) {\n";
        o << "            if (nbr_flag) persist_wait(nbr_flag, base + 1ULL + step, pa.err, pa.timeout_cycles);\n";
        o << "            __syncthreads();   // the polling threads' ld.acquire + this barrier order the halo loads below after the neighbours' stores\n";
        o << "            PT_PROF(0);\n";
        o << "            if (step) pt_reload(c == 0 ? " << (hm.n_colours - 1) << " : c - 1, shx, shy, shz, gx, gy, gz, o0, o1, w0, w1);\n";
        o << "            __syncthreads();\n";
        o << "            PT_PROF(1);\n";
        o << "            switch (upd) {\n";
        o << "            case UPD_OR: pt_pass<UPD_OR>(c, shx, shy, shz, gx, gy, gz, o0, o1, w0, w1, rep, a, ce); break;\n";
        o << "            case UPD_DET: pt_pass<UPD_DET>(c, shx, shy, shz, gx, gy, gz, o0, o1, w0, w1, rep, a, ce); break;\n";
        o << "            case UPD_METRO: n_acc += pt_pass<UPD_METRO>(c, shx, shy, shz, gx, gy, gz, o0, o1, w0, w1, rep, a, ce); break;\n";
        o << "            default: n_acc += pt_pass<UPD_CONE>(c, shx, shy, shz, gx, gy, gz, o0, o1, w0, w1, rep, a, ce); break;\n";
        o << "            }\n";
        o << "            __syncthreads();\n";
        o << "            PT_PROF(2);\n";
        o << "            if (threadIdx.x == 0) st_release_gpu(flags + tile * PERSIST_FLAG_STRIDE, base + 2ULL + step);   // release: cumulative over the barrier\n";
        o << "            PT_PROF(3);\n";
        o << "        }\n    }\n";
        // write the tile back (core cells of every class)
        for (int q = 0; q < nq; ++q) {
            o << "    for (int e = threadIdx.x; e < w0 * " << (S1 * S2) << "; e += PT_TPB) {\n";
            o << "        const int l2 = e % " << S2 << ", col = (e / " << S2 << ") % " << S1 << ", l0 = e / " << (S1 * S2) << ", l1 = col - " << H[2] << ";\n";
            o << "        if (l1 < 0 || l1 >= w1) continue;\n";
            o << "        const int g = Seg" << q << "::pos(o0 + l0, o1 + l1, l2), j = " << ((long)q * CELLS + (long)H[0] * S1 * S2) << " + e;\n";
            o << "        gx[g] = shx[j]; gy[g] = shy[j]; gz[g] = shz[j];\n    }\n";
        }
        o << "    if (pa.prof && threadIdx.x == 0) for (int k = 0; k < 4; ++k) pa.prof[((size_t)blockIdx.y * gridDim.x + tile) * 8 + k] = (unsigned long long)pf[k];\n";
        o << "    { const int w = __reduce_add_sync(0xffffffffu, n_acc); if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sh_acc, w); }\n";
        o << "    __syncthreads();\n";
        o << "    if (threadIdx.x == 0 && sh_acc) atomicAdd(a.accepted + (size_t)rep * ACC_STRIPE + (tile & (ACC_STRIPE - 1)), (unsigned long long)sh_acc);\n";
        o << "}\n";
        return true;
    }

    std::string run(JitPlan &plan) {
        {
            std::ostringstream head;
            head.swap(o);
            for (size_t s = 0; s < hm.segs.size(); ++s) segment((int)s);   // fills ktab
            std::string segs = o.str();
            o.swap(head);
            {   // fp64 flops of one overrelaxation update, averaged over the sites: neighbour field + F = g - h (3) + s.F (5)
                // + F.F (5) + 2 * / (2) + the reflection (6); an on-site term adds 3 dot products and 3 doublings (18)
                double f = 0.0;
                long n = 0;
                for (size_t s = 0; s < hm.segs.size(); ++s) {
                    f += (double)hm.segs[s].count * ((double)seg_flops[s] + 21.0 + (hm.onsite_coef[hm.segs[s].basis] >= 0 ? 18.0 : 0.0));
                    n += hm.segs[s].count;
                }
                plan.flops_or_update = n ? f / (double)n : 0.0;
            }
            o << "#define NPAD " << hm.npad << "\n";
            o << "#define SPIN_S " << lit(hm.S) << "\n";
            o << kJitPrelude << "\n";
            o << "__constant__ double CSMC_K[" << std::max<size_t>(ktab.size(), 1) << "] = {";
            for (size_t k = 0; k < ktab.size(); ++k) o << (k ? ", " : "") << lit(ktab[k]);
            if (ktab.empty()) o << "0.0";
            o << "};\n\n" << segs;
        }
        if (plan.persist_only) {   // the tile-resident kernel lives in a module of its own
            emit_persist(plan);
            return o.str();
        }
        plan.tiles.assign(hm.n_colours, 1);
        plan.groups.assign(hm.n_colours, 1);
        for (int c = 0; c < hm.n_colours; ++c) {
            const int s0 = hm.colour_seg_begin[c], s1 = hm.colour_seg_begin[c + 1];
            const int nseg = s1 - s0;
            // CTA tile over supercell coordinates (multi-dimensional, for L1 reuse of shifted neighbour
            // runs): 256 threads, the fastest lattice dimension gets up to 32, the rest goes to the slower
            // dimensions.  Tiling is laid over the largest class extent of the colour; classes guard.
            int M[MAXD] = {1, 1, 1}, T[MAXD] = {1, 1, 1}, NT[MAXD] = {1, 1, 1};
            for (int s = s0; s < s1; ++s) for (int d = 0; d < MAXD; ++d) M[d] = std::max(M[d], hm.segs[s].M[d]);
            static const int sw_tpb_env = std::getenv("CSMC_JIT_TPB") ? std::atoi(std::getenv("CSMC_JIT_TPB")) : 128;
            const int sw_tpb = (sw_tpb_env == 64 || sw_tpb_env == 128 || sw_tpb_env == 256 || sw_tpb_env == 512) ? sw_tpb_env : 128;
            plan.sweep_tpb = sw_tpb;
            int left = sw_tpb;
            const int last = hm.D - 1;
            // CSMC_JIT_TLAST (A/B knob): tile extent along the fastest dimension; larger values make the tile rows that
            // time-skewed strips move by thinner (32 -> 4 supercell rows per tile row in 2-D, 128 -> 1)
            static const int tlast_env = std::getenv("CSMC_JIT_TLAST") ? std::atoi(std::getenv("CSMC_JIT_TLAST")) : 0;
            const int tlast_max = (tlast_env == 16 || tlast_env == 32 || tlast_env == 64 || tlast_env == 128) ? tlast_env : 32;
            T[last] = std::min(std::min(tlast_max, pow2_ceil(M[last])), left);
            if (hm.D == 3 && !tlast_env) T[last] = std::min(T[last], 16);
            left /= T[last];
            for (int d = last - 1; d >= 0; --d) {
                int want = (d == 0) ? left : std::min(left, std::max(1, (int)pow2_floor((int)std::max(1.0, std::sqrt((double)left)))));
                T[d] = std::min(want, pow2_ceil(M[d]));
                if (d == 0) T[d] = left;   // always 256 threads; out-of-range threads idle
                left /= T[d];
            }
            if (hm.D == 1) T[0] = sw_tpb;
            for (int d = 0; d < MAXD; ++d) NT[d] = (M[d] + T[d] - 1) / T[d];
            plan.tiles[c] = NT[0] * NT[1] * NT[2];
            // time-skewed strips need one tiling for all colours, whole tile rows along dimension 0 and every
            // class of the colour spanning the full extent (so that a tile row means the same sites everywhere)
            {
                bool ok = hm.D >= 2 && M[0] % T[0] == 0 && NT[0] >= 2;
                for (int s = s0; s < s1; ++s) ok = ok && hm.segs[s].M[0] == M[0];
                if (c == 0) { skew_ok = ok; skew_T0 = T[0]; skew_NT0 = NT[0]; skew_row = NT[1] * NT[2]; }
                else skew_ok = skew_ok && ok && skew_T0 == T[0] && skew_NT0 == NT[0] && skew_row == NT[1] * NT[2];
            }
            // classes of the colour are fused in pairs into one thread (shared neighbour loads, ILP);
            // tuning knobs (environment, for A/B runs): CSMC_JIT_FUSE / CSMC_JIT_FUSE_METRO (classes per
            // thread), CSMC_JIT_MB / CSMC_JIT_MB_METRO (min resident CTAs per SM in __launch_bounds__)
            auto env_int = [](const char *n, int dflt) { const char *v = std::getenv(n); return v ? std::atoi(v) : dflt; };
            const int fuse_or = std::max(1, env_int("CSMC_JIT_FUSE", 2)), fuse_mc = std::max(1, env_int("CSMC_JIT_FUSE_METRO", 1));
            const int mb_or = std::max(1, env_int("CSMC_JIT_MB", 1)), mb_mc = std::max(1, env_int("CSMC_JIT_MB_METRO", mb_or));
            int SPc = 1;
            for (int sg = s0; sg < s1; ++sg) SPc = std::max(SPc, seg_split[sg]);
            if (SPc > 1) skew_ok = false;
            if (SPc > 1) {
                // ---- heavy sites: SPc warps share one site (see segment()); tile = sw_tpb / SPc sites ----
                const int TS = sw_tpb / SPc;
                int Ts[MAXD] = {1, 1, 1}, NTs[MAXD] = {1, 1, 1};
                {
                    int left2 = TS;
                    Ts[last] = std::min(std::min(32, pow2_ceil(M[last])), left2);
                    left2 /= Ts[last];
                    for (int d = last - 1; d >= 0; --d) { Ts[d] = (d == 0) ? left2 : std::min(left2, pow2_ceil(M[d])); if (d == 0) Ts[d] = left2; left2 /= Ts[d]; }
                    if (hm.D == 1) Ts[0] = TS;
                    for (int d = 0; d < MAXD; ++d) NTs[d] = (M[d] + Ts[d] - 1) / Ts[d];
                }
                plan.tiles[c] = NTs[0] * NTs[1] * NTs[2];
                plan.groups[c] = nseg;
                plan.groups_metro.resize(hm.n_colours);
                plan.groups_metro[c] = nseg;
                for (int u = 0; u < 4; ++u) {
                    o << "extern \"C\" __global__ void __launch_bounds__(" << sw_tpb << ", " << (u >= 2 ? mb_mc : mb_or) << ") csmc_sweep_c" << c << "_u" << u << "(double *spins, const SweepArgs a) {\n";
                    o << "#ifdef CSMC_PDL\n#if CSMC_PDL == 1\n    pdl_launch_dependents();\n#endif\n    pdl_wait();\n#endif\n";
                    o << "    __shared__ double part_g[" << (SPc - 1) << "][3][" << TS << "];\n";
                    o << "    const int rep = blockIdx.z + a.rep0;\n    int t = blockIdx.x + CSMC_TILE_OFF(a);\n";
                    o << "    const int warp = threadIdx.x >> 5, part = warp % " << SPc << ", ls = (warp / " << SPc << ") * 32 + (threadIdx.x & 31);\n";
                    o << "    const int t2 = t % " << NTs[2] << "; t /= " << NTs[2] << "; const int t1 = t % " << NTs[1] << "; const int t0 = t / " << NTs[1] << ";\n";
                    o << "    const int l2 = ls & " << (Ts[2] - 1) << ", l1 = (ls >> " << ilog2(Ts[2]) << ") & " << (Ts[1] - 1) << ", l0 = ls >> " << (ilog2(Ts[2]) + ilog2(Ts[1])) << ";\n";
                    o << "    const int m0 = t0 * " << Ts[0] << " + l0, m1 = t1 * " << Ts[1] << " + l1, m2 = t2 * " << Ts[2] << " + l2;\n";
                    o << "    int n_acc = 0;\n    switch (blockIdx.y) {\n";
                    for (int sg = s0; sg < s1; ++sg) {
                        o << "    case " << (sg - s0) << ": {\n";
                        o << "        double g0 = 0.0, g1 = 0.0, g2 = 0.0;\n        Site<Seg" << sg << "> d;\n";
                        o << "        switch (part) {\n";
                        o << "        case 0: site_load<Seg" << sg << ", true, 0>(d, spins, rep, m0, m1, m2); break;\n";
                        for (int pp = 1; pp < SPc; ++pp)
                            o << "        case " << pp << ": site_partial<Seg" << sg << ", " << pp << ">(spins, rep, m0, m1, m2, g0, g1, g2); break;\n";
                        o << "        default: break;\n        }\n";
                        o << "        if (part > 0) { part_g[part - 1][0][ls] = g0; part_g[part - 1][1][ls] = g1; part_g[part - 1][2][ls] = g2; }\n";
                        o << "        __syncthreads();\n";
                        o << "        if (part == 0) {\n            double x0 = 0.0, x1 = 0.0, x2 = 0.0;\n";
                        o << "            for (int pp = 0; pp < " << (SPc - 1) << "; ++pp) { x0 += part_g[pp][0][ls]; x1 += part_g[pp][1][ls]; x2 += part_g[pp][2][ls]; }\n";
                        o << "            n_acc += site_finish<" << u << ", Seg" << sg << ", true, 0>(d, spins, rep, a, -1, 0ULL, x0, x1, x2) ? 1 : 0;\n        }\n";
                        o << "    } break;\n";
                    }
                    o << "    default: break;\n    }\n";
                    if (u >= 2) o << "    count_accepted(n_acc, rep, a);\n";
                    o << "#if defined(CSMC_PDL) && CSMC_PDL == 2\n    pdl_launch_dependents();\n#endif\n}\n";
                }
            } else {
            plan.groups[c] = 0;
            // CSMC_JIT_TPC (experiment): a CTA handles TPC consecutive tiles one after the other; with CSMC_JIT_PREFETCH=1 it
            // prefetches the next tile's operands into L1 right after issuing the current tile's loads, so that a pass
            // that needs two waves of CTAs pays the L2 latency once
            const int tpc_or = std::max(1, env_int("CSMC_JIT_TPC", 1)), tpc_mc = std::max(1, env_int("CSMC_JIT_TPC_METRO", 1));
            const bool do_pf = env_int("CSMC_JIT_PREFETCH", 1) != 0;
            plan.tiles_per_cta.resize(hm.n_colours * 4, 1);
            const int n_tiles_c = NT[0] * NT[1] * NT[2];
            for (int u = 0; u < 4; ++u) {
                const int G = std::min(nseg, u >= 2 ? fuse_mc : fuse_or);
                const int ngroups = (nseg + G - 1) / G;
                const int TPC = (SPc > 1) ? 1 : (u >= 2 ? tpc_mc : tpc_or);
                plan.tiles_per_cta[c * 4 + u] = TPC;
                if (u == 0) plan.groups[c] = ngroups;
                if (u == 2) plan.groups_metro.resize(hm.n_colours), plan.groups_metro[c] = ngroups;
                o << "extern \"C\" __global__ void __launch_bounds__(" << sw_tpb << ", " << (u >= 2 ? mb_mc : mb_or) << ") csmc_sweep_c" << c << "_u" << u << "(double *spins, const SweepArgs a) {\n";
                o << "#ifdef CSMC_PDL\n#if CSMC_PDL == 1\n    pdl_launch_dependents();\n#endif\n    pdl_wait();\n#endif\n";
                o << "    const int rep = blockIdx.z + a.rep0;\n";
                o << "    const int l2 = threadIdx.x & " << (T[2] - 1) << ", l1 = (threadIdx.x >> " << ilog2(T[2]) << ") & " << (T[1] - 1)
                  << ", l0 = threadIdx.x >> " << (ilog2(T[2]) + ilog2(T[1])) << ";\n";
                auto coords = [&](const std::string &tv, const std::string &sfx) {
                    o << "        int tq" << sfx << " = " << tv << ";\n";
                    o << "        const int t2" << sfx << " = tq" << sfx << " % " << NT[2] << "; tq" << sfx << " /= " << NT[2] << "; const int t1" << sfx << " = tq" << sfx << " % " << NT[1]
                      << "; const int t0" << sfx << " = tq" << sfx << " / " << NT[1] << ";\n";
                    o << "        const int m0" << sfx << " = t0" << sfx << " * " << T[0] << " + l0, m1" << sfx << " = t1" << sfx << " * " << T[1] << " + l1, m2" << sfx << " = t2" << sfx << " * " << T[2] << " + l2;\n";
                };
                if (TPC == 1) {
                    o << "    {\n";
                    coords("(int)blockIdx.x + CSMC_TILE_OFF(a)", "");
                } else {
                    o << "    const int t_end = CSMC_TILE_END(a, " << n_tiles_c << ");\n";
                    o << "    for (int it = 0; it < " << TPC << "; ++it) {\n";
                    o << "        const int t_cur = ((int)blockIdx.x * " << TPC << " + it) + CSMC_TILE_OFF(a);\n";
                    o << "        if (t_cur >= t_end) break;\n";
                    coords("t_cur", "");
                }
                o << "    switch (blockIdx.y) {\n";
                for (int g = 0; g < ngroups; ++g) {
                    o << "    case " << g << ": {\n";
                    for (int s = s0 + g * G; s < std::min(s1, s0 + (g + 1) * G); ++s) o << "        Site<Seg" << s << "> d" << s << "; site_load(d" << s << ", spins, rep, m0, m1, m2);\n";
                    if (TPC > 1 && do_pf) {
                        o << "        if (it + 1 < " << TPC << " && t_cur + 1 < t_end) {\n";
                        coords("t_cur + 1", "n");
                        for (int s = s0 + g * G; s < std::min(s1, s0 + (g + 1) * G); ++s) o << "        site_prefetch<Seg" << s << ">(spins, rep, m0n, m1n, m2n);\n";
                        o << "        }\n";
                    }
                    o << "        int n_acc = 0;\n";
                    for (int s = s0 + g * G; s < std::min(s1, s0 + (g + 1) * G); ++s) o << "        n_acc += site_finish<" << u << ">(d" << s << ", spins, rep, a) ? 1 : 0;\n";
                    if (u >= 2) o << "        count_accepted(n_acc, rep, a);\n";
                    o << "    } break;\n";
                }
                o << "    default: break;\n    }\n    }\n#if defined(CSMC_PDL) && CSMC_PDL == 2\n    pdl_launch_dependents();\n#endif\n}\n";
            }
            }
            o << "extern \"C\" __global__ void __launch_bounds__(TPB) csmc_energy_c" << c << "(const double *spins, double *__restrict__ partials, int n_partials, int partial_base) {\n";
            o << "    double v[4] = {0.0, 0.0, 0.0, 0.0};\n    switch (blockIdx.y) {\n";
            for (int s = s0; s < s1; ++s) o << "    case " << (s - s0) << ": energy_site<Seg" << s << ">(spins, v); break;\n";
            o << "    default: break;\n    }\n    energy_block_reduce(v, partials, n_partials, partial_base);\n}\n";
        }
        plan.skew = plan.want_skew && skew_ok;
        plan.skew_rows = skew_NT0; plan.skew_tiles_per_row = skew_row;
        plan.skew_reach = std::max(1, (max_delta0 + std::max(1, skew_T0) - 1) / std::max(1, skew_T0));
        emit_fused(plan);
        // ---- resident kernel: one CTA per replica keeps the whole lattice in shared memory and runs
        // whole sweep schedules (n_cycles x (or_per_cycle OR + metro_per_cycle Metropolis), then det_sweeps
        // deterministic sweeps) with __syncthreads() between colour passes: one launch instead of
        // 2 * colours * sweeps launches for lattices that are launch-latency bound.
        plan.resident = (size_t)hm.npad * 24 <= 200 * 1024;
        if (plan.resident) {
            auto sweep_code = [&](int u, const char *ctr_extra) {
                for (int c = 0; c < hm.n_colours; ++c) {
                    for (int s = hm.colour_seg_begin[c]; s < hm.colour_seg_begin[c + 1]; ++s)
                        o << "            for (int idx = threadIdx.x; idx < Seg" << s << "::COUNT; idx += blockDim.x) n_acc += resident_site<" << u << ", Seg" << s
                          << ">(sh, idx, rep, a, " << ctr_extra << ");\n";
                    o << "            __syncthreads();\n";
                }
            };
            o << "extern \"C\" __global__ void __launch_bounds__(RES_TPB) csmc_resident(double *spins, const SweepArgs a, int n_cycles, int or_per_cycle,\n"
                 "        int metro_per_cycle, int cone, int adapt, int det_sweeps, double *meas, int write_energy) {\n";
            o << "    extern __shared__ double sh[];\n    const int rep = blockIdx.x;\n";
            o << "    double *g = spins + (size_t)rep * (3ull * NPAD);\n";
            o << "    for (int i = threadIdx.x; i < 3 * NPAD; i += blockDim.x) sh[i] = g[i];\n    __syncthreads();\n";
            o << "    int n_acc = 0;\n";
            o << "    for (int cyc = 0; cyc < n_cycles; ++cyc) {\n";
            o << "        for (int k = 0; k < or_per_cycle; ++k) {\n";
            sweep_code(0, "0ULL");
            o << "        }\n        for (int k = 0; k < metro_per_cycle; ++k) {\n";
            o << "            const unsigned long long ce = (unsigned long long)cyc * metro_per_cycle + k;\n";
            o << "            if (cone) {\n                const int acc_before = n_acc;\n";
            sweep_code(3, "ce");
            o << "                if (adapt) resident_adapt_sigma(n_acc - acc_before, " << (double)hm.N << ", rep, a);\n";
            o << "            } else {\n";
            sweep_code(2, "ce");
            o << "            }\n        }\n    }\n";
            o << "    for (int k = 0; k < det_sweeps; ++k) {\n";
            sweep_code(1, "0ULL");
            o << "    }\n";
            o << "    for (int i = threadIdx.x; i < 3 * NPAD; i += blockDim.x) g[i] = sh[i];\n";
            o << "    __shared__ int sh_acc;\n    __shared__ double red[4][RES_TPB / 32];\n";
            o << "    if (threadIdx.x == 0) sh_acc = 0;\n    __syncthreads();\n";
            o << "    { const int w = __reduce_add_sync(0xffffffffu, n_acc); if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sh_acc, w); }\n";
            o << "    __syncthreads();\n";
            o << "    if (threadIdx.x == 0 && sh_acc) a.accepted[(size_t)rep * ACC_STRIPE] += (unsigned long long)sh_acc;   // this CTA owns the replica\n";
            o << "    if (meas) {\n        double v[4] = {0.0, 0.0, 0.0, 0.0};\n";
            for (size_t s = 0; s < hm.segs.size(); ++s)
                o << "        for (int idx = threadIdx.x; idx < Seg" << s << "::COUNT; idx += blockDim.x) energy_site_at<Seg" << s << ">(sh, idx, v);\n";
            o << "        for (int k = 0; k < 4; ++k) { const double w = warp_sum(v[k]); if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = w; }\n";
            o << "        __syncthreads();\n";
            o << "        if (threadIdx.x < 4) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];\n";
            o << "            if (threadIdx.x > 0 || write_energy) meas[(size_t)rep * 8 + threadIdx.x] = t; }\n";
            o << "        if (threadIdx.x == 4) { unsigned long long acc = 0; for (int k = 0; k < ACC_STRIPE; ++k) acc += a.accepted[(size_t)rep * ACC_STRIPE + k];\n";
            o << "            meas[(size_t)rep * 8 + 4] = (double)acc; meas[(size_t)rep * 8 + 5] = a.sigma[rep]; }\n    }\n}\n";
        }
        return o.str();
    }
};

// ---- NVRTC through dlopen ------------------------------------------------------------------------------
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *lib = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};
Nvrtc g_rtc;
std::mutex g_rtc_mu;

bool load_nvrtc() {
    if (g_rtc.lib) return true;
    const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char *n : names) {
        g_rtc.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_rtc.lib) break;
    }
    if (!g_rtc.lib) { g_rtc.err = std::string("cannot load libnvrtc: ") + dlerror(); return false; }
#define SYM(f) g_rtc.f = (decltype(g_rtc.f))dlsym(g_rtc.lib, "nvrtc" #f)
    SYM(CreateProgram); SYM(CompileProgram); SYM(GetCUBINSize); SYM(GetCUBIN); SYM(GetProgramLogSize);
    SYM(GetProgramLog); SYM(DestroyProgram); SYM(GetErrorString);
#undef SYM
    if (!g_rtc.CreateProgram || !g_rtc.CompileProgram || !g_rtc.GetCUBINSize || !g_rtc.GetCUBIN) {
        g_rtc.err = "libnvrtc lacks required symbols";
        g_rtc.lib = nullptr;
        return false;
    }
    return true;
}

}  // namespace

std::string jit_generate_source(const HostModel &hm, bool pdl, JitPlan *plan) {
    Gen g(hm);
    JitPlan local;
    std::string src = g.run(plan ? *plan : local);
    const char *mode = std::getenv("CSMC_JIT_PDL_MODE");   // 1: trigger dependents at kernel start, 2: at kernel end
    if (plan && plan->skew) src = "#define CSMC_SKEW 1\n" + src;
    return pdl ? std::string("#define CSMC_PDL ") + (mode && mode[0] == '2' ? "2" : "1") + "\n" + src : src;
}

// Time-skewed strips.  P colour passes over a lattice whose spins exceed L2 normally stream the whole lattice
// through L2 P times.  Dependencies only reach `reach` tile rows, so the passes can instead be run strip by
// strip -- all P passes on one strip of tile rows while it is L2-resident -- if the strip moves up by `reach`
// rows per pass: a row then sees its neighbours exactly as the pass-by-pass order would show them (each
// neighbour row has finished the previous pass and has not started the next one), so every site update reads
// the same values and the result is bit-identical.  With periodic wrap the first strip shrinks from both sides
// (nothing outside it has been updated yet), the following strips are parallelograms, and the last one grows
// on both sides across the wrap.  Rows are CTA-tile rows along lattice dimension 0.
std::vector<SkewLaunch> skew_schedule(int n_rows, int n_passes, int reach, int budget_rows) {
    std::vector<SkewLaunch> out;
    const int P = n_passes, d = std::max(1, reach);
    const int shift = (P - 1) * d;                       // total movement of a strip boundary
    // first strip W0 rows (shrinks to W0 - 2 shift), last strip Wc rows (grows to Wc + 2 shift = W0); on a lattice
    // only slightly larger than the budget the two share it
    const int W0 = std::min(budget_rows, (n_rows + 2 * shift) / 2), Wc = W0 - 2 * shift;
    if (P < 2 || n_rows < 2 || budget_rows < 1 || Wc < 1) return out;
    const int middle = n_rows - W0 - Wc;
    const int n_mid = (middle + budget_rows - 1) / budget_rows;
    std::vector<int> B;                                  // boundaries of the parallelogram strips at pass 0
    for (int s = 0; s <= n_mid; ++s) B.push_back(W0 + (int)((long long)middle * s / std::max(1, n_mid)));
    for (int p = 0; p < P; ++p) out.push_back({p, p * d, W0 - 2 * p * d});                 // shrinking first strip
    for (int s = 0; s < n_mid; ++s)
        for (int p = 0; p < P; ++p) out.push_back({p, B[s] - p * d, B[s + 1] - B[s]});    // parallelograms
    const int BS = B.back();
    for (int p = 0; p < P; ++p) {                                                          // growing last strip, wraps
        out.push_back({p, BS - p * d, n_rows - (BS - p * d)});
        if (p > 0) out.push_back({p, 0, p * d});
    }
    return out;
}

// returns "" on success
// Optional on-disk cache of compiled cubins (CSMC_CACHE_DIR=<dir>), keyed by a hash of the generated
// source: repeated runs of the same model skip NVRTC (seconds for cubic / quartic models).
static uint64_t fnv1a(const char *p, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ULL; }
    return h;
}

static std::string cache_path(const std::string &src) {
    const char *dir = std::getenv("CSMC_CACHE_DIR");
    if (!dir || !*dir) return "";
    char buf[80];
    std::snprintf(buf, sizeof buf, "/csmc_%016zx_%zu.sm_100a.cubin", std::hash<std::string>{}(src), src.size());
    return std::string(dir) + buf;
}

std::string jit_compile(const std::string &src, std::vector<char> &cubin, std::string &log) {
    std::lock_guard<std::mutex> lk(g_rtc_mu);
    const std::string cpath = cache_path(src);
    if (!cpath.empty()) {
        // cache file = cubin followed by a 16-byte trailer {length, FNV-1a of the cubin}: a torn or foreign file is
        // ignored (and recompiled) instead of being handed to the module loader
        if (FILE *f = std::fopen(cpath.c_str(), "rb")) {
            std::fseek(f, 0, SEEK_END);
            const long n = std::ftell(f);
            std::fseek(f, 0, SEEK_SET);
            std::vector<char> blob;
            if (n > 16) { blob.resize((size_t)n); if (std::fread(blob.data(), 1, (size_t)n, f) != (size_t)n) blob.clear(); }
            std::fclose(f);
            if (!blob.empty()) {
                uint64_t trailer[2];
                std::memcpy(trailer, blob.data() + blob.size() - 16, 16);
                if (trailer[0] == blob.size() - 16 && trailer[1] == fnv1a(blob.data(), blob.size() - 16)) {
                    cubin.assign(blob.begin(), blob.end() - 16);
                    return "";
                }
            }
        }
    }
    if (!load_nvrtc()) return g_rtc.err;
    nvrtcProgram prog = nullptr;
    // CSMC_JIT_DUMP=<dir>: keep the generated source on disk under the name the cubin's line table
    // refers to, so `ncu --import-source on` / cuobjdump can map SASS back to it
    std::string name = "csmc_jit.cu";
    if (const char *dir = std::getenv("CSMC_JIT_DUMP")) {
        size_t hsh = std::hash<std::string>{}(src);
        char buf[64];
        std::snprintf(buf, sizeof buf, "/csmc_jit_%016zx.cu", hsh);
        name = std::string(dir) + buf;
        if (FILE *f = std::fopen(name.c_str(), "w")) { std::fwrite(src.data(), 1, src.size(), f); std::fclose(f); }
    }
    int rc = g_rtc.CreateProgram(&prog, src.c_str(), name.c_str(), 0, nullptr, nullptr);
    if (rc != 0) return "nvrtcCreateProgram failed";
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    rc = g_rtc.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    if (g_rtc.GetProgramLogSize && g_rtc.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
        log.resize(ls);
        g_rtc.GetProgramLog(prog, &log[0]);
    }
    if (rc != 0) {
        std::string e = std::string("nvrtcCompileProgram: ") + (g_rtc.GetErrorString ? g_rtc.GetErrorString(rc) : "error") + "\n" + log.substr(0, 4000);
        g_rtc.DestroyProgram(&prog);
        return e;
    }
    size_t n = 0;
    if (g_rtc.GetCUBINSize(prog, &n) != 0 || n == 0) { g_rtc.DestroyProgram(&prog); return "nvrtcGetCUBINSize failed"; }
    cubin.resize(n);
    rc = g_rtc.GetCUBIN(prog, cubin.data());
    g_rtc.DestroyProgram(&prog);
    if (rc == 0 && !cpath.empty()) {
        // per-process temporary name: the ranks of a job compile the same model at the same time
        const std::string tmp = cpath + "." + std::to_string((long)getpid()) + ".tmp";
        if (FILE *f = std::fopen(tmp.c_str(), "wb")) {
            const uint64_t trailer[2] = {(uint64_t)cubin.size(), fnv1a(cubin.data(), cubin.size())};
            const bool ok = std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size() && std::fwrite(trailer, 1, 16, f) == 16;
            std::fclose(f);
            if (ok) std::rename(tmp.c_str(), cpath.c_str()); else std::remove(tmp.c_str());
        }
    }
    return rc == 0 ? "" : "nvrtcGetCUBIN failed";
}

}  // namespace csmc
