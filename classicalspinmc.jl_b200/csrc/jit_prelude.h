// jit_prelude.h — static part of the runtime-specialised kernel source (compiled with NVRTC for
// sm_100a at csmc_create time, see jit.cpp).  The model-specific part (one `struct SegK` per
// colouring class with fully unrolled interaction terms, literal coefficients and constant
// geometry, plus the extern "C" kernels) is generated and appended to this text.
//
// The arithmetic is the same as the ahead-of-time kernels in kernels.cuh; what changes is that
// every per-class constant is a compile-time literal, so the generated SASS carries no index-table
// or coefficient loads and no term loops.
#pragma once

static const char *kJitPrelude = R"CSMCJIT(
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;

enum { UPD_OR = 0, UPD_DET = 1, UPD_METRO = 2, UPD_CONE = 3 };
enum { TAG_PROPOSE = 0, TAG_ACCEPT = 1, TAG_INIT = 2, TAG_EXCHANGE = 3 };
#define TPB 256
#define ACC_STRIPE 32
#define RES_TPB 256

template <bool NC> __device__ __forceinline__ double ld(const double *p) { return NC ? __ldg(p) : *p; }

struct SweepArgs {
    const double *beta;
    const double *sigma;
    unsigned long long *accepted;
    const unsigned long long *ctr_base;
    unsigned long long ctr_off;
    unsigned long long seed;
    int replica_base;
    int rep0;
#ifdef CSMC_SKEW
    int tile_off;   // first CTA tile of this launch (time-skewed strips, api.cu); the host struct always carries it
    int tile_end;   // one past its last tile
#endif
};
#ifdef CSMC_SKEW
#define CSMC_TILE_OFF(a) ((a).tile_off)
#define CSMC_TILE_END(a, total) ((a).tile_end)
#else
#define CSMC_TILE_OFF(a) 0
#define CSMC_TILE_END(a, total) (total)
#endif

// by-value argument of csmc_persist (csmc_internal.h mirrors it)
#define PERSIST_MAX_OPS 48
#define PERSIST_FLAG_STRIDE 4
struct PersistArgs {
    unsigned long long *flags;
    int *err;
    unsigned long long timeout_cycles;
    int n_ops;
    unsigned char upd[PERSIST_MAX_OPS];
    unsigned short ctr_rel[PERSIST_MAX_OPS];
    int pad_;
    unsigned long long *prof;
};
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spins until *p >= want; gives up (and raises the sticky error word) after `budget` cycles or when another CTA already
// gave up, so a tile that never arrives fails the call instead of hanging the GPU
__device__ __forceinline__ void persist_wait(const unsigned long long *p, unsigned long long want, int *err, unsigned long long budget) {
    if (ld_acquire_gpu(p) >= want) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_gpu(p) < want) {
        if ((++spins & 255u) == 0u) {
            if (*(volatile int *)err != 0) return;
            if ((unsigned long long)(clock64() - t0) > budget) { *(volatile int *)err = 1; __threadfence_system(); return; }
        }
    }
}

// asynchronous 8-byte global -> shared copy (LDGSTS): no register staging, every copy of a thread in flight at once
__device__ __forceinline__ void cp_async8(double *smem, const double *g) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct u4 { uint32_t x, y, z, w; };
__device__ __forceinline__ u4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
    return r;
}
__device__ __forceinline__ u4 philox_stream(unsigned long long seed, uint32_t c0, uint32_t c1, unsigned long long ctr, uint32_t tag) {
    return philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, (uint32_t)ctr, (uint32_t)((ctr >> 32) << 8) | tag);
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * 0x1.0p-53;
}
__device__ __forceinline__ void philox_to_3_uniforms(const u4 &r, double &u1, double &u2, double &u3) {
    u1 = (double)(((unsigned long long)r.x << 11) | (r.y >> 21)) * 0x1.0p-43;
    u2 = (double)(((unsigned long long)(r.y & 0x1FFFFFu) << 22) | (r.z >> 10)) * 0x1.0p-43;
    u3 = (double)(((unsigned long long)(r.z & 0x3FFu) << 32) | r.w) * 0x1.0p-42;
}
__device__ __forceinline__ void random_orientation(double S, double u1, double u2, double &x, double &y, double &z) {
    double sn, cs;
    sincospi(2.0 * u1, &sn, &cs);
    const double zz = 2.0 * u2 - 1.0;
    const double r = sqrt(1.0 - zz * zz);
    x = S * (r * cs); y = S * (r * sn); z = S * zz;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch: a pass may be scheduled while the previous pass drains (its CTAs
// fill SMs as they free up and run their prologue); it must not touch spins before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Per-site working set.  A pass first *loads* everything a thread needs (own spin and all neighbour
// spins, for every class the thread handles), then computes, then stores: with all loads issued up
// front the Metropolis RNG / proposal arithmetic overlaps the memory latency, and the sites of
// different classes handled by one thread are independent instruction streams (ILP).
template <class SEG>
struct Site {
    bool valid;
    int pos, m0, m1, m2;
    unsigned ok;                      // open boundaries: bit t set <=> term t has all its neighbours
    double s0, s1, s2;
    double nb[SEG::PRELOAD ? 3 * SEG::NNB + 1 : 1];   // neighbour spins, compile-time indexed (registers)
};

template <class SEG, bool NC = true, int PART = -1>
__device__ __forceinline__ void site_load(Site<SEG> &d, const double *spins, int rep, int m0, int m1, int m2) {
    const double *sx = spins + (size_t)rep * (3ull * NPAD), *sy = sx + NPAD, *sz = sy + NPAD;
    d.valid = SEG::valid(m0, m1, m2);
    d.m0 = m0; d.m1 = m1; d.m2 = m2;
    d.ok = 0xffffffffu;
    if (d.valid) {
        d.pos = SEG::pos(m0, m1, m2);
        d.s0 = sx[d.pos]; d.s1 = sy[d.pos]; d.s2 = sz[d.pos];
        if (SEG::PRELOAD) SEG::template load<NC, PART>(sx, sy, sz, m0, m1, m2, d.nb, d.ok);
    }
}

__device__ __forceinline__ void pf_l1(const double *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// the operands site_load would read for (m0, m1, m2), as L1 prefetches
template <class SEG>
__device__ __forceinline__ void site_prefetch(const double *spins, int rep, int m0, int m1, int m2) {
    const double *sx = spins + (size_t)rep * (3ull * NPAD), *sy = sx + NPAD, *sz = sy + NPAD;
    if (SEG::valid(m0, m1, m2)) {
        const int pos = SEG::pos(m0, m1, m2);
        pf_l1(sx + pos); pf_l1(sy + pos); pf_l1(sz + pos);
        if (SEG::PRELOAD) SEG::prefetch(sx, sy, sz, m0, m1, m2);
    }
}

// Dot product with the fused-multiply-add order written out: left to the compiler, the contraction of
// a0*b0 + a1*b1 + a2*b2 depends on the code around it, and the same update inlined into different kernels
// (per-colour pass, resident, fused sweep) must round identically.
__device__ __forceinline__ double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return fma(a2, b2, fma(a1, b1, a0 * b0));
}

// PART >= 0: this thread evaluated only the slots of its part; (x0, x1, x2) carries the other parts' sums.
// (sx, sy, sz) are the component arrays d.pos indexes (global memory or a shared-memory tile); prep is the
// local replica index used for beta / sigma / the Philox stream.
template <int UPD, class SEG, bool NC, int PART>
__device__ __forceinline__ bool site_finish_ptr(Site<SEG> &d, double *sx, double *sy, double *sz, int prep, const SweepArgs &a,
                                                unsigned long long ctr_extra, double x0, double x1, double x2) {
    const int rep = prep, grep_override = -1;
    bool accepted = false;
    if (d.valid) {
        const int pos = d.pos;
        const double s0 = d.s0, s1 = d.s1, s2 = d.s2;
        double n0 = 0.0, n1 = 0.0, n2 = 0.0, u3 = 0.0;
        if (UPD == UPD_METRO || UPD == UPD_CONE) {
            // the Philox call and the proposal are independent of the loads issued in site_load
            const unsigned long long ctr = (a.ctr_base ? *a.ctr_base : 0ULL) + a.ctr_off + ctr_extra;
            const uint32_t site = SEG::site(d.m0, d.m1, d.m2);
            const int rr = grep_override >= 0 ? grep_override : rep;   // resident kernels index shared memory with rep == 0
            const uint32_t grep = (uint32_t)(a.replica_base + rr);
            const u4 r = philox_stream(a.seed, site, grep, ctr, TAG_PROPOSE);
            double u1, u2;
            philox_to_3_uniforms(r, u1, u2, u3);
            random_orientation(SPIN_S, u1, u2, n0, n1, n2);
            if (UPD == UPD_CONE) {
                const double sg = a.sigma[grep_override >= 0 ? grep_override : rep];
                n0 = s0 + sg * n0; n1 = s1 + sg * n1; n2 = s2 + sg * n2;
                const double nrm = sqrt(dot3(n0, n1, n2, n0, n1, n2));
                n0 = n0 / nrm * SPIN_S; n1 = n1 / nrm * SPIN_S; n2 = n2 / nrm * SPIN_S;
            }
        }
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        if (UPD == UPD_OR || UPD == UPD_DET) {
            if (SEG::ONSITE) {
                g0 = 2 * dot3(SEG::O0, SEG::O1, SEG::O2, s0, s1, s2);
                g1 = 2 * dot3(SEG::O3, SEG::O4, SEG::O5, s0, s1, s2);
                g2 = 2 * dot3(SEG::O6, SEG::O7, SEG::O8, s0, s1, s2);
            }
        }
        if (SEG::PRELOAD) SEG::template field<PART>(d.nb, d.ok, g0, g1, g2, g0, g1, g2, g0, g1, g2);
        else SEG::template field_stream<NC>(sx, sy, sz, d.m0, d.m1, d.m2, g0, g1, g2, g0, g1, g2, g0, g1, g2);
        g0 += x0; g1 += x1; g2 += x2;
        const double F0 = g0 - SEG::H0, F1 = g1 - SEG::H1, F2 = g2 - SEG::H2;
        if (UPD == UPD_OR) {
            if (!(F0 == 0.0 && F1 == 0.0 && F2 == 0.0)) {
                const double proj = 2.0 * dot3(s0, s1, s2, F0, F1, F2) / dot3(F0, F1, F2, F0, F1, F2);
                sx[pos] = fma(proj, F0, -s0); sy[pos] = fma(proj, F1, -s1); sz[pos] = fma(proj, F2, -s2);
            }
        } else if (UPD == UPD_DET) {
            if (!(F0 == 0.0 && F1 == 0.0 && F2 == 0.0)) {
                const double nrm = sqrt(dot3(F0, F1, F2, F0, F1, F2));
                sx[pos] = -F0 / nrm * SPIN_S; sy[pos] = -F1 / nrm * SPIN_S; sz[pos] = -F2 / nrm * SPIN_S;
            }
        } else {
            double dE = dot3(n0 - s0, n1 - s1, n2 - s2, F0, F1, F2);
            if (SEG::ONSITE) {
                const double en = dot3(n0, n1, n2, dot3(SEG::O0, SEG::O1, SEG::O2, n0, n1, n2), dot3(SEG::O3, SEG::O4, SEG::O5, n0, n1, n2),
                                       dot3(SEG::O6, SEG::O7, SEG::O8, n0, n1, n2));
                const double eo = dot3(s0, s1, s2, dot3(SEG::O0, SEG::O1, SEG::O2, s0, s1, s2), dot3(SEG::O3, SEG::O4, SEG::O5, s0, s1, s2),
                                       dot3(SEG::O6, SEG::O7, SEG::O8, s0, s1, s2));
                dE += en - eo;
            }
            accepted = dE < 0.0 || u3 < exp(-dE * a.beta[grep_override >= 0 ? grep_override : rep]);
            if (accepted) { sx[pos] = n0; sy[pos] = n1; sz[pos] = n2; }
        }
    }
    return accepted;
}

template <int UPD, class SEG, bool NC = true, int PART = -1>
__device__ __forceinline__ bool site_finish(Site<SEG> &d, double *spins, int rep, const SweepArgs &a, int grep_override = -1,
                                            unsigned long long ctr_extra = 0ULL, double x0 = 0.0, double x1 = 0.0, double x2 = 0.0) {
    double *sx = spins + (size_t)rep * (3ull * NPAD), *sy = sx + NPAD, *sz = sy + NPAD;
    return site_finish_ptr<UPD, SEG, NC, PART>(d, sx, sy, sz, grep_override >= 0 ? grep_override : rep, a, ctr_extra, x0, x1, x2);
}

// a non-zero part: evaluate its slots' contribution to the neighbour field of the site
template <class SEG, int PART>
__device__ __forceinline__ void site_partial(const double *spins, int rep, int m0, int m1, int m2, double &g0, double &g1, double &g2) {
    Site<SEG> d;
    site_load<SEG, true, PART>(d, spins, rep, m0, m1, m2);
    g0 = g1 = g2 = 0.0;
    if (d.valid) SEG::template field<PART>(d.nb, d.ok, g0, g1, g2, g0, g1, g2, g0, g1, g2);
}

// one atomic per CTA, striped over ACC_STRIPE addresses per replica (same-address L2 atomics serialise)
__device__ __forceinline__ void count_accepted(int n_mine, int rep, const SweepArgs &a) {
    __shared__ int sh_acc;
    if (threadIdx.x == 0) sh_acc = 0;
    __syncthreads();
    const int w = __reduce_add_sync(0xffffffffu, n_mine);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sh_acc, w);
    __syncthreads();
    if (threadIdx.x == 0 && sh_acc)
        atomicAdd(a.accepted + (size_t)rep * ACC_STRIPE + (blockIdx.x & (ACC_STRIPE - 1)), (unsigned long long)sh_acc);
}

template <class SEG>
__device__ __forceinline__ void energy_site(const double *spins, double (&v)[4]) {
    const int idx = blockIdx.x * TPB + threadIdx.x;
    const int rep = blockIdx.z;
    if (idx < SEG::COUNT) {
        int m0, m1, m2;
        SEG::locate(idx, m0, m1, m2);
        Site<SEG> d;
        site_load(d, spins, rep, m0, m1, m2);
        const double s0 = d.s0, s1 = d.s1, s2 = d.s2;
        double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0, c0 = 0, c1 = 0, c2 = 0;
        if (SEG::PRELOAD) SEG::template field<-1>(d.nb, d.ok, a0, a1, a2, b0, b1, b2, c0, c1, c2);
        else {
            const double *sx = spins + (size_t)rep * (3ull * NPAD), *sy = sx + NPAD, *sz = sy + NPAD;
            SEG::template field_stream<true>(sx, sy, sz, m0, m1, m2, a0, a1, a2, b0, b1, b2, c0, c1, c2);
        }
        double e = (s0 * a0 + s1 * a1 + s2 * a2) / 2 + (s0 * b0 + s1 * b1 + s2 * b2) / 3 +
                   (s0 * c0 + s1 * c1 + s2 * c2) / 4 - (s0 * SEG::H0 + s1 * SEG::H1 + s2 * SEG::H2);
        if (SEG::ONSITE)
            e += s0 * (SEG::O0 * s0 + SEG::O1 * s1 + SEG::O2 * s2) + s1 * (SEG::O3 * s0 + SEG::O4 * s1 + SEG::O5 * s2) +
                 s2 * (SEG::O6 * s0 + SEG::O7 * s1 + SEG::O8 * s2);
        v[0] = e; v[1] = s0; v[2] = s1; v[3] = s2;
    }
}

// weighted site energy (total_energy convention) from shared-memory spins, resident kernel
template <class SEG>
__device__ __forceinline__ void energy_site_at(const double *spins, int idx, double (&v)[4]) {
    if (idx < SEG::COUNT) {
        int m0, m1, m2;
        SEG::locate(idx, m0, m1, m2);
        Site<SEG> d;
        site_load<SEG, false>(d, spins, 0, m0, m1, m2);
        const double s0 = d.s0, s1 = d.s1, s2 = d.s2;
        double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0, c0 = 0, c1 = 0, c2 = 0;
        if (SEG::PRELOAD) SEG::template field<-1>(d.nb, d.ok, a0, a1, a2, b0, b1, b2, c0, c1, c2);
        else SEG::template field_stream<false>(spins, spins + NPAD, spins + 2 * NPAD, m0, m1, m2, a0, a1, a2, b0, b1, b2, c0, c1, c2);
        double e = (s0 * a0 + s1 * a1 + s2 * a2) / 2 + (s0 * b0 + s1 * b1 + s2 * b2) / 3 +
                   (s0 * c0 + s1 * c1 + s2 * c2) / 4 - (s0 * SEG::H0 + s1 * SEG::H1 + s2 * SEG::H2);
        if (SEG::ONSITE)
            e += s0 * (SEG::O0 * s0 + SEG::O1 * s1 + SEG::O2 * s2) + s1 * (SEG::O3 * s0 + SEG::O4 * s1 + SEG::O5 * s2) +
                 s2 * (SEG::O6 * s0 + SEG::O7 * s1 + SEG::O8 * s2);
        v[0] += e; v[1] += s0; v[2] += s1; v[3] += s2;
    }
}

// one site of a resident sweep: spins live in shared memory (plain loads), rep indexes beta/sigma/RNG
template <int UPD, class SEG>
__device__ __forceinline__ int resident_site(double *sh, int idx, int rep, const SweepArgs &a, unsigned long long ctr_extra) {
    if (idx >= SEG::COUNT) return 0;
    int m0, m1, m2;
    SEG::locate(idx, m0, m1, m2);
    Site<SEG> d;
    site_load<SEG, false>(d, sh, 0, m0, m1, m2);
    return site_finish<UPD, SEG, false>(d, sh, 0, a, rep, ctr_extra) ? 1 : 0;
}

// resident kernel, MetropolisAdaptive (src/metropolis.jl:129-131): block-wide acceptance of the sweep
// that just finished, then sigma <- clamp(sigma * 0.5 / max(1 - a, 0.05), 0, 100) for this replica
__device__ __forceinline__ void resident_adapt_sigma(int n_acc_sweep, double n_sites, int rep, const SweepArgs &a) {
    __shared__ int sh_sweep_acc;
    if (threadIdx.x == 0) sh_sweep_acc = 0;
    __syncthreads();
    const int w = __reduce_add_sync(0xffffffffu, n_acc_sweep);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&sh_sweep_acc, w);
    __syncthreads();
    if (threadIdx.x == 0) {
        const double acc = (double)sh_sweep_acc / n_sites, f = 0.5 / fmax(1.0 - acc, 0.05);
        double *sig = const_cast<double *>(a.sigma);
        sig[rep] = fmin(fmax(sig[rep] * f, 0.0), 100.0);
    }
    __syncthreads();
}

__device__ __forceinline__ void energy_block_reduce(const double (&v)[4], double *__restrict__ partials, int n_partials, int partial_base) {
    __shared__ double sh[4][TPB / 32];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double w = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < TPB / 32; ++w) t += sh[threadIdx.x][w];
        const size_t slot = (size_t)blockIdx.z * n_partials + partial_base + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        partials[slot * 4 + threadIdx.x] = t;
    }
}
)CSMCJIT";
