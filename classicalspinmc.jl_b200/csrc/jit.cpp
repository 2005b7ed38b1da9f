// jit.cpp — runtime specialisation of the pass kernels: generates CUDA C++ for one lattice model
// (unrolled interaction terms, literal coefficients, constant geometry) and compiles it to an
// sm_100a cubin with NVRTC (loaded with dlopen so libcsmc.so has no link-time dependency on it).
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
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

struct Gen {
    const HostModel &hm;
    std::ostringstream o;
    std::map<std::tuple<int, int, int, int>, int> seg_of_class;
    explicit Gen(const HostModel &h) : hm(h) {
        for (size_t s = 0; s < hm.segs.size(); ++s)
            seg_of_class[std::make_tuple(hm.segs[s].basis, hm.segs[s].r[0], hm.segs[s].r[1], hm.segs[s].r[2])] = (int)s;
    }

    // emits code computing `int j<tag>` (storage position of neighbour k of term t) and, for open
    // boundaries, updates `ok`.  Returns false when the neighbour class does not exist at all.
    bool neighbour(const HostSeg &hs, const HostTerm &t, int k, const std::string &tag) {
        int r2[MAXD] = {0, 0, 0}, delta[MAXD] = {0, 0, 0};
        for (int d = 0; d < hm.D; ++d) {
            r2[d] = posmod(hs.r[d] + t.off[k][d], hm.P[d]);
            delta[d] = floordiv(hs.r[d] + t.off[k][d], hm.P[d]);
        }
        auto it = seg_of_class.find(std::make_tuple(t.nb_basis[k], r2[0], r2[1], r2[2]));
        if (it == seg_of_class.end()) return false;
        const HostSeg &ns = hm.segs[it->second];
        o << "        int j" << tag << ";\n        {\n";
        for (int d = 0; d < MAXD; ++d) {
            if (d >= hm.D) { o << "            const int n" << d << " = 0;\n"; continue; }
            o << "            int n" << d << " = m" << d << " + (" << delta[d] << ");\n";
            if (hm.periodic) {
                if (delta[d] > 0) o << "            n" << d << " = (n" << d << " >= " << ns.M[d] << ") ? n" << d << " - " << ns.M[d] << " : n" << d << ";\n";
                if (delta[d] < 0) o << "            n" << d << " = (n" << d << " < 0) ? n" << d << " + " << ns.M[d] << " : n" << d << ";\n";
            } else {
                if (delta[d] < 0) o << "            ok = ok && (n" << d << " >= 0);\n";
                if (hs.M[d] - 1 + delta[d] >= ns.M[d]) o << "            ok = ok && (n" << d << " < " << ns.M[d] << ");\n";
            }
        }
        o << "            j" << tag << " = " << ns.start << " + (n0 * " << ns.M[1] << " + n1) * " << ns.M[2] << " + n2;\n        }\n";
        return true;
    }

    void segment(int s) {
        const HostSeg &hs = hm.segs[s];
        const int b = hs.basis;
        o << "struct Seg" << s << " {\n";
        o << "    static constexpr int START = " << hs.start << ", COUNT = " << hs.count << ";\n";
        o << "    static constexpr double H0 = " << lit(hm.field[3 * b]) << ", H1 = " << lit(hm.field[3 * b + 1]) << ", H2 = " << lit(hm.field[3 * b + 2]) << ";\n";
        const bool ons = hm.onsite_coef[b] >= 0;
        o << "    static constexpr bool ONSITE = " << (ons ? "true" : "false") << ";\n";
        for (int k = 0; k < 9; ++k) o << "    static constexpr double O" << k << " = " << lit(hm.onsite[9 * b + k]) << ";\n";
        // idx -> supercell coordinates
        o << "    static __device__ __forceinline__ void locate(int idx, int &m0, int &m1, int &m2) {\n";
        o << "        m2 = idx % " << hs.M[2] << "; const int t = idx / " << hs.M[2] << "; m1 = t % " << hs.M[1] << "; m0 = t / " << hs.M[1] << ";\n    }\n";
        // reference site index (Philox counter)
        o << "    static __device__ __forceinline__ unsigned site(int m0, int m1, int m2) {\n";
        o << "        return (unsigned)(((" << b << " * " << hm.L[0] << " + (m0 * " << hs.P[0] << " + " << hs.r[0] << ")) * " << hm.L[1]
          << " + (m1 * " << hs.P[1] << " + " << hs.r[1] << ")) * " << hm.L[2] << " + (m2 * " << hs.P[2] << " + " << hs.r[2] << "));\n    }\n";
        // unrolled neighbour field: a* bilinear, b* cubic, c* quartic accumulators
        o << "    static __device__ __forceinline__ void field(const double *__restrict__ sx, const double *__restrict__ sy, const double *__restrict__ sz,\n"
             "            int m0, int m1, int m2, double &a0, double &a1, double &a2, double &b0, double &b1, double &b2, double &c0, double &c1, double &c2) {\n";
        int tn = 0;
        for (const auto &t : hm.basis_terms[b]) {
            const double *C = hm.coefs.data() + t.coef;
            const int nn = t.kind - 1;
            const int ncoef = t.kind == 2 ? 9 : t.kind == 3 ? 27 : 81;
            bool any = false;
            for (int k = 0; k < ncoef; ++k) any |= (C[k] != 0.0);
            if (!any) { ++tn; continue; }
            o << "      { // term " << tn << " kind " << t.kind << "\n";
            if (!hm.periodic) o << "        bool ok = true;\n";
            bool exists = true;
            std::ostringstream saved;
            saved.swap(o);
            for (int k = 0; k < nn && exists; ++k) exists = neighbour(hs, t, k, std::to_string(k));
            std::string body = o.str();
            o.swap(saved);
            if (!exists) { o << "      }\n"; ++tn; continue; }
            o << body;
            if (!hm.periodic) o << "        if (ok) {\n";
            const char *nm[3] = {"p", "q", "w"};
            for (int k = 0; k < nn; ++k)
                o << "        const double " << nm[k] << "0 = sx[j" << k << "], " << nm[k] << "1 = sy[j" << k << "], " << nm[k] << "2 = sz[j" << k << "];\n";
            if (t.kind == 2) {
                for (int a = 0; a < 3; ++a) {
                    std::string e;
                    for (int c = 0; c < 3; ++c)
                        if (C[3 * a + c] != 0.0) e += (e.empty() ? "" : " + ") + lit(C[3 * a + c]) + " * p" + std::to_string(c);
                    if (!e.empty()) o << "        a" << a << " += " << e << ";\n";
                }
            } else if (t.kind == 3) {
                for (int bb = 0; bb < 3; ++bb)
                    for (int c = 0; c < 3; ++c) {
                        bool used = false;
                        for (int a = 0; a < 3; ++a) used |= (C[a * 9 + bb * 3 + c] != 0.0);
                        if (!used) continue;
                        o << "        { const double v = p" << bb << " * q" << c << ";";
                        for (int a = 0; a < 3; ++a)
                            if (C[a * 9 + bb * 3 + c] != 0.0) o << " b" << a << " += " << lit(C[a * 9 + bb * 3 + c]) << " * v;";
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
                                if (C[a * 27 + bb * 9 + c * 3 + d] != 0.0) o << " c" << a << " += " << lit(C[a * 27 + bb * 9 + c * 3 + d]) << " * v;";
                            o << " }\n";
                        }
                        o << "        }\n";
                    }
            }
            if (!hm.periodic) o << "        }\n";
            o << "      }\n";
            ++tn;
        }
        o << "    }\n};\n\n";
    }

    std::string run() {
        o << "#define NPAD " << hm.npad << "\n";
        o << "#define SPIN_S " << lit(hm.S) << "\n";
        o << kJitPrelude << "\n";
        for (size_t s = 0; s < hm.segs.size(); ++s) segment((int)s);
        for (int c = 0; c < hm.n_colours; ++c) {
            const int s0 = hm.colour_seg_begin[c], s1 = hm.colour_seg_begin[c + 1];
            const int nseg = s1 - s0;
            int mx = 1;
            for (int s = s0; s < s1; ++s) mx = std::max(mx, hm.segs[s].count);
            const int tiles_per_seg = (mx + 255) / 256;
            // grid = (tiles_per_seg * nseg, 1, replicas).  The segments (classes) of the colour are
            // interleaved along blockIdx.x, so classes that gather from the same neighbour classes sweep
            // the same region of the lattice at the same time: the other colours are read from DRAM once
            // per pass instead of once per class (measured at L=4096: 604 MB -> 402 MB read per pass).
            (void)tiles_per_seg;
            for (int u = 0; u < 4; ++u) {
                o << "extern \"C\" __global__ void __launch_bounds__(TPB) csmc_sweep_c" << c << "_u" << u << "(double *__restrict__ spins, const SweepArgs a) {\n";
                o << "#ifdef CSMC_PDL\n    pdl_launch_dependents();\n    pdl_wait();\n#endif\n";
                o << "    const int rep = blockIdx.z, tile = blockIdx.x;\n";
                o << "    const int idx = (tile / " << nseg << ") * TPB + threadIdx.x;\n";
                o << "    switch (tile % " << nseg << ") {\n";
                for (int s = s0; s < s1; ++s) o << "    case " << (s - s0) << ": sweep_site<" << u << ", Seg" << s << ">(spins, a, idx, rep); break;\n";
                o << "    default: break;\n    }\n}\n";
            }
            o << "extern \"C\" __global__ void __launch_bounds__(TPB) csmc_energy_c" << c << "(const double *__restrict__ spins, double *__restrict__ partials, int n_partials, int partial_base) {\n";
            o << "    double v[4] = {0.0, 0.0, 0.0, 0.0};\n    switch (blockIdx.y) {\n";
            for (int s = s0; s < s1; ++s) o << "    case " << (s - s0) << ": energy_site<Seg" << s << ">(spins, v); break;\n";
            o << "    default: break;\n    }\n    energy_block_reduce(v, partials, n_partials, partial_base);\n}\n";
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

std::string jit_generate_source(const HostModel &hm, bool pdl) {
    Gen g(hm);
    std::string src = g.run();
    return pdl ? "#define CSMC_PDL 1\n" + src : src;
}

// returns "" on success
std::string jit_compile(const std::string &src, std::vector<char> &cubin, std::string &log) {
    std::lock_guard<std::mutex> lk(g_rtc_mu);
    if (!load_nvrtc()) return g_rtc.err;
    nvrtcProgram prog = nullptr;
    int rc = g_rtc.CreateProgram(&prog, src.c_str(), "csmc_jit.cu", 0, nullptr, nullptr);
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
    return rc == 0 ? "" : "nvrtcGetCUBIN failed";
}

}  // namespace csmc
