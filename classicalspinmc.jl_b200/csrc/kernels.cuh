// kernels.cuh — sm_100a device code of libcsmc: colour-pass sweep kernels (overrelaxation,
// deterministic, Metropolis uniform / cone), energy + magnetisation reduction, evaluation kernels,
// layout conversion, Philox4x32-10, replica exchange.
//
// Layout: spins are fp64 SoA per replica, [replica][x|y|z][npad], stored colour-major and, inside a
// colour, class-major (one dense sub-lattice per colouring class) so that a colour pass reads and
// writes contiguous, coalesced runs.  All per-class constants (Zeeman vector, coupling matrices /
// tensors, neighbour-class geometry) arrive in the kernel parameter block, i.e. the constant bank.
#pragma once
#include <cuda_runtime.h>

#include "csmc_internal.h"

namespace csmc {

constexpr int TPB = 256;  // threads per block of the pass kernels
// accepted-proposal counters are striped: [replica][ACC_STRIPE]; one atomic per CTA lands on stripe
// blockIdx.x % ACC_STRIPE (same-address L2 atomics serialise: one per warp cost ~10 us per pass at C2)
constexpr int ACC_STRIPE = 32;

enum { UPD_OR = 0, UPD_DET = 1, UPD_METRO = 2, UPD_CONE = 3 };
enum { TAG_PROPOSE = 0, TAG_ACCEPT = 1, TAG_INIT = 2, TAG_EXCHANGE = 3 };

struct SweepArgs {
    const double *beta;                  // [R] per-replica inverse temperature   (Metropolis)
    const double *sigma;                 // [R] per-replica cone width            (cone moves)
    unsigned long long *accepted;        // [R] accepted-proposal counters        (Metropolis)
    const unsigned long long *ctr_base;  // optional device-resident sweep counter (graph replay)
    unsigned long long ctr_off;          // + by-value offset
    unsigned long long seed;
    int replica_base;                    // global index of local replica 0
    int rep0;                            // first local replica of this launch (replica groups on separate streams)
    int tile_off;                        // first CTA tile of this launch: read only by runtime-specialised kernels built
    int tile_end;                        // with CSMC_SKEW (time-skewed strips): [tile_off, tile_end); every other kernel's SweepArgs ends at rep0
};
static_assert(sizeof(SweepArgs) == 64, "SweepArgs layout (jit_prelude.h mirrors it)");

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) --------------------------------------------
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
    return u4{c0, c1, c2, c3};
}
__device__ __forceinline__ u4 philox_stream(unsigned long long seed, uint32_t c0, uint32_t c1, unsigned long long ctr, uint32_t tag) {
    return philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, (uint32_t)ctr, (uint32_t)((ctr >> 32) << 8) | tag);
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * 0x1.0p-53;
}
// One Philox call per Metropolis proposal: its 128 bits are split into three uniforms in [0,1):
// u1 (43 bits: azimuth), u2 (43 bits: z), u3 (42 bits: acceptance).
__device__ __forceinline__ void philox_to_3_uniforms(const u4 &r, double &u1, double &u2, double &u3) {
    u1 = (double)(((unsigned long long)r.x << 11) | (r.y >> 21)) * 0x1.0p-43;
    u2 = (double)(((unsigned long long)(r.y & 0x1FFFFFu) << 22) | (r.z >> 10)) * 0x1.0p-43;
    u3 = (double)(((unsigned long long)(r.z & 0x3FFu) << 32) | r.w) * 0x1.0p-42;
}

// random_spin_orientation (src/lattice.jl:306-311): phi = 2 pi u1, z = 2 u2 - 1
__device__ __forceinline__ void random_orientation(double S, double u1, double u2, double &x, double &y, double &z) {
    double sn, cs;
    sincospi(2.0 * u1, &sn, &cs);
    const double zz = 2.0 * u2 - 1.0;
    const double r = sqrt(1.0 - zz * zz);
    x = S * (r * cs); y = S * (r * sn); z = S * zz;
}

// ---- neighbour lookup --------------------------------------------------------------------------------
template <class P, bool STRUCT>
__device__ __forceinline__ int nbr_pos(const P &p, const DevTerm &t, int k, int pos, const int (&m)[MAXD]) {
    if (!STRUCT) {
        return __ldg(p.nbr + (size_t)(t.row + k) * p.npad + pos);
    } else {
        const int g = t.nseg[k];
        if (g < 0) return -1;
        const SegGeom &G = p.geom[g];
        int lin = 0;
#pragma unroll
        for (int d = 0; d < MAXD; ++d) {
            int mm = m[d] + t.d[k][d];
            const int Md = G.M[d];
            if (p.periodic) {
                if (mm >= Md) mm -= Md;
                if (mm < 0) mm += Md;
            } else if (mm < 0 || mm >= Md) {
                return -1;
            }
            lin = lin * Md + mm;
        }
        return G.start + lin;
    }
}

// Accumulates the neighbour part of the local field (src/hamiltonian.jl:25-65) into g2/g3/g4
// (bilinear / cubic / quartic).  Callers that do not need the split pass the same array thrice.
template <class P, bool STRUCT>
__device__ __forceinline__ void accumulate_field(const P &p, const DevSeg &seg, const double *__restrict__ sx,
                                                 const double *__restrict__ sy, const double *__restrict__ sz,
                                                 int pos, const int (&m)[MAXD], double (&g2)[3], double (&g3)[3], double (&g4)[3]) {
    const DevTerm *t = p.terms + seg.term_begin;
    for (int n = 0; n < seg.n2; ++n, ++t) {
        const int j = nbr_pos<P, STRUCT>(p, *t, 0, pos, m);
        if (j < 0) continue;
        const double x = sx[j], y = sy[j], z = sz[j];
        const double *J = p.coefs + t->coef;
        g2[0] += J[0] * x + J[1] * y + J[2] * z;
        g2[1] += J[3] * x + J[4] * y + J[5] * z;
        g2[2] += J[6] * x + J[7] * y + J[8] * z;
    }
    for (int n = 0; n < seg.n3; ++n, ++t) {
        const int j = nbr_pos<P, STRUCT>(p, *t, 0, pos, m);
        const int k = nbr_pos<P, STRUCT>(p, *t, 1, pos, m);
        if (j < 0 || k < 0) continue;
        const double sj[3] = {sx[j], sy[j], sz[j]}, sk[3] = {sx[k], sy[k], sz[k]};
        const double *C = p.coefs + t->coef;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double w = sj[b] * sk[c];
                g3[0] += C[0 * 9 + b * 3 + c] * w;
                g3[1] += C[1 * 9 + b * 3 + c] * w;
                g3[2] += C[2 * 9 + b * 3 + c] * w;
            }
    }
    for (int n = 0; n < seg.n4; ++n, ++t) {
        const int j = nbr_pos<P, STRUCT>(p, *t, 0, pos, m);
        const int k = nbr_pos<P, STRUCT>(p, *t, 1, pos, m);
        const int l = nbr_pos<P, STRUCT>(p, *t, 2, pos, m);
        if (j < 0 || k < 0 || l < 0) continue;
        const double sj[3] = {sx[j], sy[j], sz[j]}, sk[3] = {sx[k], sy[k], sz[k]}, sl[3] = {sx[l], sy[l], sz[l]};
        const double *R = p.coefs + t->coef;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double wbc = sj[b] * sk[c];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double w = wbc * sl[d];
                    g4[0] += R[0 * 27 + b * 9 + c * 3 + d] * w;
                    g4[1] += R[1 * 27 + b * 9 + c * 3 + d] * w;
                    g4[2] += R[2 * 27 + b * 9 + c * 3 + d] * w;
                }
            }
    }
}

// thread -> (segment, position, supercell coordinates); false when out of range
template <class P, bool STRUCT>
__device__ __forceinline__ bool locate(const P &p, const DevSeg &seg, int idx, int &pos, int (&m)[MAXD]) {
    if (idx >= seg.count) return false;
    pos = seg.start + idx;
    if (STRUCT) {
        int t = idx;
        m[2] = t % seg.M[2]; t /= seg.M[2];
        m[1] = t % seg.M[1]; m[0] = t / seg.M[1];
    } else {
        m[0] = m[1] = m[2] = 0;
    }
    return true;
}

// reference (0-based) site index of the thread's site: the Philox counter, layout independent
template <class P, bool STRUCT>
__device__ __forceinline__ uint32_t ref_index(const P &p, const DevSeg &seg, int pos, const int (&m)[MAXD]) {
    if (!STRUCT) return (uint32_t)__ldg(p.ref_of_pos + pos);
    int idx = seg.basis;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) idx = idx * p.L[d] + (m[d] * seg.P[d] + seg.r[d]);
    return (uint32_t)idx;
}

// ---- colour-pass sweep kernel --------------------------------------------------------------------------
// One thread per site of the pass; grid = (blocks over the longest segment, segments of the colour,
// replicas).  Same-colour sites share no interaction term, so the pass is race-free.
template <class P, bool STRUCT, int UPD>
__global__ void __launch_bounds__(TPB) k_sweep(const __grid_constant__ P p, const SweepArgs a) {
    const DevSeg &seg = p.segs[blockIdx.y];
    int pos, m[MAXD];
    const bool active = locate<P, STRUCT>(p, seg, blockIdx.x * TPB + threadIdx.x, pos, m);
    const int rep = blockIdx.z + a.rep0;
    double *sx = p.spins + (size_t)rep * p.rep_stride, *sy = sx + p.npad, *sz = sy + p.npad;
    bool accepted = false;
    if (active) {
        const double s0 = sx[pos], s1 = sy[pos], s2 = sz[pos];
        double g[3] = {0.0, 0.0, 0.0};
        if (UPD == UPD_OR || UPD == UPD_DET) {
            // on-site part of the field first, as src/hamiltonian.jl:19-22
            if (seg.onsite >= 0) {
                const double *O = p.coefs + seg.onsite;
                g[0] = 2 * (O[0] * s0 + O[1] * s1 + O[2] * s2);
                g[1] = 2 * (O[3] * s0 + O[4] * s1 + O[5] * s2);
                g[2] = 2 * (O[6] * s0 + O[7] * s1 + O[8] * s2);
            }
        }
        accumulate_field<P, STRUCT>(p, seg, sx, sy, sz, pos, m, g, g, g);
        const double F0 = g[0] - seg.h[0], F1 = g[1] - seg.h[1], F2 = g[2] - seg.h[2];  // :66
        if (UPD == UPD_OR) {
            // src/monte_carlo.jl:131-137
            if (!(F0 == 0.0 && F1 == 0.0 && F2 == 0.0)) {
                const double proj = 2.0 * (s0 * F0 + s1 * F1 + s2 * F2) / (F0 * F0 + F1 * F1 + F2 * F2);
                sx[pos] = -s0 + proj * F0; sy[pos] = -s1 + proj * F1; sz[pos] = -s2 + proj * F2;
            }
        } else if (UPD == UPD_DET) {
            // src/monte_carlo.jl:206-210
            if (!(F0 == 0.0 && F1 == 0.0 && F2 == 0.0)) {
                const double nrm = sqrt(F0 * F0 + F1 * F1 + F2 * F2);
                sx[pos] = -F0 / nrm * p.S; sy[pos] = -F1 / nrm * p.S; sz[pos] = -F2 / nrm * p.S;
            }
        } else {
            // src/metropolis.jl:65-101 with dE from one field evaluation:
            // e(s) = s.O.s + s.(G - h)  =>  dE = (s'-s).(G-h) + s'.O.s' - s.O.s
            const unsigned long long ctr = (a.ctr_base ? *a.ctr_base : 0ULL) + a.ctr_off;
            const uint32_t site = ref_index<P, STRUCT>(p, seg, pos, m);
            const uint32_t grep = (uint32_t)(a.replica_base + rep);
            const u4 r = philox_stream(a.seed, site, grep, ctr, TAG_PROPOSE);
            double n0, n1, n2, u1, u2, u3;
            philox_to_3_uniforms(r, u1, u2, u3);
            random_orientation(p.S, u1, u2, n0, n1, n2);
            if (UPD == UPD_CONE) {
                // gaussian_move, src/metropolis.jl:84-87
                const double sg = a.sigma[rep];
                n0 = s0 + sg * n0; n1 = s1 + sg * n1; n2 = s2 + sg * n2;
                const double nrm = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
                n0 = n0 / nrm * p.S; n1 = n1 / nrm * p.S; n2 = n2 / nrm * p.S;
            }
            double dE = (n0 - s0) * F0 + (n1 - s1) * F1 + (n2 - s2) * F2;
            if (seg.onsite >= 0) {
                const double *O = p.coefs + seg.onsite;
                const double en = n0 * (O[0] * n0 + O[1] * n1 + O[2] * n2) + n1 * (O[3] * n0 + O[4] * n1 + O[5] * n2) +
                                  n2 * (O[6] * n0 + O[7] * n1 + O[8] * n2);
                const double eo = s0 * (O[0] * s0 + O[1] * s1 + O[2] * s2) + s1 * (O[3] * s0 + O[4] * s1 + O[5] * s2) +
                                  s2 * (O[6] * s0 + O[7] * s1 + O[8] * s2);
                dE += en - eo;
            }
            accepted = dE < 0.0 || u3 < exp(-dE * a.beta[rep]);  // src/metropolis.jl:73
            if (accepted) { sx[pos] = n0; sy[pos] = n1; sz[pos] = n2; }
        }
    }
    if (UPD == UPD_METRO || UPD == UPD_CONE) {
        const int n_acc = __syncthreads_count(accepted);
        if (threadIdx.x == 0 && n_acc)
            atomicAdd(a.accepted + (size_t)rep * ACC_STRIPE + (blockIdx.x & (ACC_STRIPE - 1)), (unsigned long long)n_acc);
    }
}

// ---- energy + magnetisation ----------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// per-site weighted energy (src/hamiltonian.jl:70-131: e2/2 + e3/3 + e4/4 - s.h + s.O.s) and spin sum,
// warp-shuffle + shared-memory block reduction, one partial (E, Mx, My, Mz) per block, fixed order.
template <class P, bool STRUCT>
__global__ void __launch_bounds__(TPB) k_energy(const __grid_constant__ P p, double *__restrict__ partials,
                                                int n_partials, int partial_base) {
    const DevSeg &seg = p.segs[blockIdx.y];
    int pos, m[MAXD];
    const bool active = locate<P, STRUCT>(p, seg, blockIdx.x * TPB + threadIdx.x, pos, m);
    const int rep = blockIdx.z;
    const double *sx = p.spins + (size_t)rep * p.rep_stride, *sy = sx + p.npad, *sz = sy + p.npad;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (active) {
        const double s0 = sx[pos], s1 = sy[pos], s2 = sz[pos];
        double g2[3] = {0, 0, 0}, g3[3] = {0, 0, 0}, g4[3] = {0, 0, 0};
        accumulate_field<P, STRUCT>(p, seg, sx, sy, sz, pos, m, g2, g3, g4);
        double e = (s0 * g2[0] + s1 * g2[1] + s2 * g2[2]) / 2 + (s0 * g3[0] + s1 * g3[1] + s2 * g3[2]) / 3 +
                   (s0 * g4[0] + s1 * g4[1] + s2 * g4[2]) / 4 - (s0 * seg.h[0] + s1 * seg.h[1] + s2 * seg.h[2]);
        if (seg.onsite >= 0) {
            const double *O = p.coefs + seg.onsite;
            e += s0 * (O[0] * s0 + O[1] * s1 + O[2] * s2) + s1 * (O[3] * s0 + O[4] * s1 + O[5] * s2) +
                 s2 * (O[6] * s0 + O[7] * s1 + O[8] * s2);
        }
        v[0] = e; v[1] = s0; v[2] = s1; v[3] = s2;
    }
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
        const size_t slot = (size_t)rep * n_partials + partial_base + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        partials[slot * 4 + threadIdx.x] = t;
    }
}

// second stage: one block per replica sums its partials in a fixed order (deterministic) and writes
// the 8-double measurement record {E, Mx, My, Mz, accepted, 0, 0, 0}.
__device__ __forceinline__ void reduce_record(const double *__restrict__ partials, int n_partials,
                                              const unsigned long long *__restrict__ accepted,
                                              const double *__restrict__ sigma, double *meas, int write_energy) {
    const int rep = blockIdx.x;
    double v[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < n_partials; i += 256) {
        const double *q = partials + ((size_t)rep * n_partials + i) * 4;
        v[0] += q[0]; v[1] += q[1]; v[2] += q[2]; v[3] += q[3];
    }
    __shared__ double sh[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double w = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        if (threadIdx.x > 0 || write_energy) meas[(size_t)rep * 8 + threadIdx.x] = t;
    }
    if (threadIdx.x == 4) {
        unsigned long long acc = 0;
        for (int k = 0; k < ACC_STRIPE; ++k) acc += accepted[(size_t)rep * ACC_STRIPE + k];
        meas[(size_t)rep * 8 + 4] = (double)acc;
        meas[(size_t)rep * 8 + 5] = sigma[rep];   // cone width travels with the temperature slot
    }
}

__global__ void __launch_bounds__(256) k_reduce_partials(const double *__restrict__ partials, int n_partials,
                                                         const unsigned long long *__restrict__ accepted,
                                                         const double *__restrict__ sigma,
                                                         double *__restrict__ meas, int write_energy) {
    reduce_record(partials, n_partials, accepted, sigma, meas, write_energy);
}

// ---- evaluation kernels (API / parity tests): outputs in reference site order -------------------------
// what == 0: local field H - h (src/hamiltonian.jl:3-67) -> out[N x 3];  what == 1: site energy
// (src/hamiltonian.jl:139-196) -> out[N]
template <class P, bool STRUCT>
__global__ void __launch_bounds__(TPB) k_eval(const __grid_constant__ P p, int rep, int what, double *__restrict__ out,
                                              int seg_off, int block_off) {
    // (seg_off, block_off) != 0: single-site queries launch only the block that holds the site
    const DevSeg &seg = p.segs[blockIdx.y + seg_off];
    int pos, m[MAXD];
    if (!locate<P, STRUCT>(p, seg, (blockIdx.x + block_off) * TPB + threadIdx.x, pos, m)) return;
    const double *sx = p.spins + (size_t)rep * p.rep_stride, *sy = sx + p.npad, *sz = sy + p.npad;
    const double s0 = sx[pos], s1 = sy[pos], s2 = sz[pos];
    const int ref = __ldg(p.ref_of_pos + pos);
    double g[3] = {0.0, 0.0, 0.0};
    const double *O = seg.onsite >= 0 ? p.coefs + seg.onsite : nullptr;
    if (what == 0 && O) {
        g[0] = 2 * (O[0] * s0 + O[1] * s1 + O[2] * s2);
        g[1] = 2 * (O[3] * s0 + O[4] * s1 + O[5] * s2);
        g[2] = 2 * (O[6] * s0 + O[7] * s1 + O[8] * s2);
    }
    accumulate_field<P, STRUCT>(p, seg, sx, sy, sz, pos, m, g, g, g);
    if (what == 0) {
        out[3 * (size_t)ref + 0] = g[0] - seg.h[0];
        out[3 * (size_t)ref + 1] = g[1] - seg.h[1];
        out[3 * (size_t)ref + 2] = g[2] - seg.h[2];
    } else {
        double e = s0 * g[0] + s1 * g[1] + s2 * g[2] - (s0 * seg.h[0] + s1 * seg.h[1] + s2 * seg.h[2]);
        if (O)
            e += s0 * (O[0] * s0 + O[1] * s1 + O[2] * s2) + s1 * (O[3] * s0 + O[4] * s1 + O[5] * s2) +
                 s2 * (O[6] * s0 + O[7] * s1 + O[8] * s2);
        out[ref] = e;
    }
}

// ---- layout conversion: reference AoS (N x 3) <-> colour-major SoA ---------------------------------------
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ spins, const int32_t *__restrict__ ref_of_pos, int npad) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npad) return;
    const int ref = ref_of_pos[pos];
    double x = 0.0, y = 0.0, z = 0.0;
    if (ref >= 0) { x = aos[3 * (size_t)ref]; y = aos[3 * (size_t)ref + 1]; z = aos[3 * (size_t)ref + 2]; }
    spins[pos] = x; spins[npad + pos] = y; spins[2 * (size_t)npad + pos] = z;
}
__global__ void k_soa_to_aos(const double *__restrict__ spins, double *__restrict__ aos, const int32_t *__restrict__ ref_of_pos, int npad) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npad) return;
    const int ref = ref_of_pos[pos];
    if (ref < 0) return;
    aos[3 * (size_t)ref] = spins[pos]; aos[3 * (size_t)ref + 1] = spins[npad + pos]; aos[3 * (size_t)ref + 2] = spins[2 * (size_t)npad + pos];
}

// Lattice(...; initialCondition=:random), src/lattice.jl:76-79, Philox stream (seed, replica, site)
__global__ void k_randomize(double *__restrict__ spins, const int32_t *__restrict__ ref_of_pos, int npad, int64_t rep_stride,
                            double S, unsigned long long seed, int replica_base) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npad) return;
    const int ref = ref_of_pos[pos];
    const int rep = blockIdx.y;
    double x = 0.0, y = 0.0, z = 0.0;
    if (ref >= 0) {
        const u4 r = philox_stream(seed, (uint32_t)ref, (uint32_t)(replica_base + rep), 0ULL, TAG_INIT);
        random_orientation(S, u53(r.x, r.y), u53(r.z, r.w), x, y, z);
    }
    double *s = spins + (size_t)rep * rep_stride;
    s[pos] = x; s[npad + pos] = y; s[2 * (size_t)npad + pos] = z;
}

__global__ void k_add_u64(unsigned long long *ctr, unsigned long long v) { *ctr += v; }

// ---- equal-time structure factor (src/spin_correlations.jl:6-43) ----------------------------------------
// A_u(k) = sum_i exp(-i k.r_i) s_i^u with r_i = sum_d i_d a_d + basis_b: the phase factorises into one
// factor per lattice dimension and one per basis site, so each CTA tabulates exp(-i i_d k.a_d) for its
// tile of KT wavevectors in shared memory (sum_d L_d + n_basis sincos per wavevector instead of N) and
// every (site, k) costs D complex multiplications.  Partial sums per (site chunk, k) are reduced in a
// fixed order by k_ssf_finish (deterministic).
struct SsfGeom {
    int D, n_basis, L[MAXD];
    int npad, n_k, KT, chunk;      // chunk = storage positions per blockIdx.y
    int table_len;                 // sum_d L_d + n_basis
};

__global__ void __launch_bounds__(256) k_ssf_partial(const double *__restrict__ spins, const int32_t *__restrict__ ref_of_pos,
                                                     const double *__restrict__ theta /* [n_k][MAXD] k.a_d */,
                                                     const double *__restrict__ phib /* [n_k][n_basis] k.basis_b */,
                                                     SsfGeom g, double *__restrict__ partial /* [chunks][n_k][6] */) {
    extern __shared__ double2 tab[];          // [KT][table_len]
    const int k0 = blockIdx.x * g.KT;
    const int nk = min(g.KT, g.n_k - k0);
    for (int e = threadIdx.x; e < nk * g.table_len; e += blockDim.x) {
        const int kk = e / g.table_len, j = e % g.table_len;
        double ang;
        int off = 0, d = 0;
        for (; d < g.D; ++d) { if (j < off + g.L[d]) break; off += g.L[d]; }
        if (d < g.D) ang = -theta[(size_t)(k0 + kk) * MAXD + d] * (double)(j - off);
        else ang = -phib[(size_t)(k0 + kk) * g.n_basis + (j - off)];
        double sn, cs;
        sincos(ang, &sn, &cs);
        tab[e] = make_double2(cs, sn);
    }
    __syncthreads();
    constexpr int KTMAX = 8;
    double acc[KTMAX][6];
#pragma unroll
    for (int kk = 0; kk < KTMAX; ++kk)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[kk][c] = 0.0;
    const double *sx = spins, *sy = sx + g.npad, *sz = sy + g.npad;
    const int p_begin = blockIdx.y * g.chunk, p_end = min(g.npad, p_begin + g.chunk);
    const int off1 = g.L[0], off2 = g.L[0] + (g.D > 1 ? g.L[1] : 0), offb = off2 + (g.D > 2 ? g.L[2] : 0);
    for (int pos = p_begin + threadIdx.x; pos < p_end; pos += blockDim.x) {
        int ref = __ldg(ref_of_pos + pos);
        if (ref < 0) continue;
        int i2 = 0, i1 = 0;
        if (g.D > 2) { i2 = ref % g.L[2]; ref /= g.L[2]; }
        if (g.D > 1) { i1 = ref % g.L[1]; ref /= g.L[1]; }
        const int i0 = ref % g.L[0], b = ref / g.L[0];
        const double s0 = sx[pos], s1 = sy[pos], s2 = sz[pos];
#pragma unroll
        for (int kk = 0; kk < KTMAX; ++kk) {
            if (kk < nk) {
                const double2 *t = tab + kk * g.table_len;
                double2 z = t[i0];
                if (g.D > 1) { const double2 w = t[off1 + i1]; z = make_double2(z.x * w.x - z.y * w.y, z.x * w.y + z.y * w.x); }
                if (g.D > 2) { const double2 w = t[off2 + i2]; z = make_double2(z.x * w.x - z.y * w.y, z.x * w.y + z.y * w.x); }
                { const double2 w = t[offb + b]; z = make_double2(z.x * w.x - z.y * w.y, z.x * w.y + z.y * w.x); }
                acc[kk][0] += z.x * s0; acc[kk][1] += z.y * s0;
                acc[kk][2] += z.x * s1; acc[kk][3] += z.y * s1;
                acc[kk][4] += z.x * s2; acc[kk][5] += z.y * s2;
            }
        }
    }
    __shared__ double red[8][6 * KTMAX];
#pragma unroll
    for (int kk = 0; kk < KTMAX; ++kk)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const double w = warp_sum(acc[kk][c]);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][kk * 6 + c] = w;
        }
    __syncthreads();
    if (threadIdx.x < 6 * nk) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        partial[((size_t)blockIdx.y * g.n_k + k0 + threadIdx.x / 6) * 6 + threadIdx.x % 6] = t;
    }
}

// Suv[3u+v, k] = Re(A_u conj(A_v)) / N (src/spin_correlations.jl:31-42); out is 9 x n_k column-major;
// accumulate != 0 adds to out (running sum for the PT loop's mean)
// slot_of_rep != nullptr: out is the per-temperature-slot accumulator block of the replica's current slot
__global__ void k_ssf_finish(const double *__restrict__ partial, int n_chunks, int n_k, double inv_n, double *__restrict__ out, int accumulate,
                             const int *__restrict__ slot_of_rep, int rep_global) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_k) return;
    if (slot_of_rep) out += (size_t)slot_of_rep[rep_global] * 9 * n_k;
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < n_chunks; ++c)
        for (int j = 0; j < 6; ++j) a[j] += partial[((size_t)c * n_k + k) * 6 + j];
    for (int u = 0; u < 3; ++u)
        for (int v = 0; v < 3; ++v) {
            const double val = (a[2 * u] * a[2 * v] + a[2 * u + 1] * a[2 * v + 1]) * inv_n;
            double *o = out + (size_t)k * 9 + 3 * u + v;
            *o = accumulate ? *o + val : val;
        }
}

// MetropolisAdaptive rule after a sweep, src/metropolis.jl:129-131, per replica
__global__ void k_adapt_sigma(double *sigma, const unsigned long long *accepted, unsigned long long *prev, double n_sites, int R) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    unsigned long long now = 0;
    for (int k = 0; k < ACC_STRIPE; ++k) now += accepted[(size_t)r * ACC_STRIPE + k];
    const double acc = (double)(now - prev[r]);
    prev[r] = now;
    const double a = acc / n_sites, f = 0.5 / fmax(1.0 - a, 0.05);
    sigma[r] = fmin(fmax(sigma[r] * f, 0.0), 100.0);
}

// ---- parallel tempering ------------------------------------------------------------------------------------
struct PtState {
    int n_slots, n_local, replica_base;
    const double *T_slot;          // [n_slots] temperature of a slot (fixed)
    int *slot_of_rep, *rep_of_slot;  // [n_slots]
    const double *meas_all;        // [n_slots][8] gathered measurement records (by global replica)
    double *E_last;                // [n_slots] energy after the replica's last Metropolis sweep
    double *acc_prev;              // [n_slots] accepted counter at the last flush
    double *acc_slot, *exch_slot;  // [n_slots] statistics attributed to temperature slots
    double *beta_local;            // [n_local] inverse temperatures of the local replicas (kernel input)
    double *sigma_local;           // [n_local] cone widths of the local replicas (kernel input)
    int *prev_rep_of_slot;         // [n_slots] scratch: occupancy before the exchange step
    int *accepted_pairs;           // [n_slots] decisions of the last exchange step
};

// after a Metropolis sweep + energy gather: E = total_energy (src/monte_carlo.jl:305) and
// accepted_local += ... (:304), attributed to the slot the replica currently occupies
__global__ void k_pt_update(PtState st, int update_energy) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= st.n_slots) return;
    const double *mrec = st.meas_all + (size_t)rep * 8;
    if (update_energy) st.E_last[rep] = mrec[0];
    const double acc = mrec[4];
    st.acc_slot[st.slot_of_rep[rep]] += acc - st.acc_prev[rep];
    st.acc_prev[rep] = acc;
}

// replica exchange, src/monte_carlo.jl:308-349: pairs (first, first+1), (first+2, first+3), ...;
// accept iff u < min(1, exp((1/T_b - 1/T_a)(E_b - E_a))); on accept the two replicas trade slots
// (temperatures move, configurations stay).  Every rank runs this redundantly on identical inputs.
__global__ void k_pt_exchange(PtState st, int first, unsigned long long exch_ctr, unsigned long long seed) {
    for (int k = threadIdx.x; k < st.n_slots; k += blockDim.x) { st.accepted_pairs[k] = 0; st.prev_rep_of_slot[k] = st.rep_of_slot[k]; }
    __syncthreads();
    for (int a = first + 2 * (int)threadIdx.x; a + 1 < st.n_slots; a += 2 * blockDim.x) {
        const int b = a + 1;
        const int ra = st.rep_of_slot[a], rb = st.rep_of_slot[b];
        const double Ta = st.T_slot[a], Tb = st.T_slot[b];
        const double Ea = st.E_last[ra], Eb = st.E_last[rb];
        const double w = exp((1.0 / Tb - 1.0 / Ta) * (Eb - Ea));
        const u4 r = philox_stream(seed, (uint32_t)a, 0xFFFFFFFFu, exch_ctr, TAG_EXCHANGE);
        if (u53(r.x, r.y) < fmin(1.0, w)) {
            st.rep_of_slot[a] = rb; st.rep_of_slot[b] = ra;
            st.slot_of_rep[ra] = b; st.slot_of_rep[rb] = a;
            st.exch_slot[a] += 1.0; st.exch_slot[b] += 1.0;
            st.accepted_pairs[a] = 1;
        }
    }
    __syncthreads();
    // a replica that moved to another temperature slot takes over that slot's beta and cone width
    // (the reference keeps mc.T and mc.sigma on the rank and moves the configuration, :336-347)
    for (int r = threadIdx.x; r < st.n_local; r += blockDim.x) {
        const int slot = st.slot_of_rep[st.replica_base + r];
        st.beta_local[r] = 1.0 / st.T_slot[slot];
        st.sigma_local[r] = st.meas_all[(size_t)st.prev_rep_of_slot[slot] * 8 + 5];
    }
}

// probe, src/monte_carlo.jl:368-370: (E, |M|) of every slot appended to the series
__global__ void k_pt_probe(PtState st, double *series_E, double *series_M, long long index) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= st.n_slots) return;
    const int rep = st.rep_of_slot[slot];
    const double *mrec = st.meas_all + (size_t)rep * 8;
    series_E[index * st.n_slots + slot] = st.E_last[rep];
    series_M[index * st.n_slots + slot] = sqrt(mrec[1] * mrec[1] + mrec[2] * mrec[2] + mrec[3] * mrec[3]);
}

// ---- measurement records gathered by stores into peer memory (NVLink / NVSwitch) --------------------------
// Opt-in alternative (CSMC_PEER_GATHER) to the ncclAllGather of the per-replica energies before an exchange
// (src/monte_carlo.jl:321-343 sends them with MPI.Sendrecv!).  Every rank owns a mailbox
// [2 parities][cap8 doubles] and one arrival flag per rank, mapped into every other process with CUDA IPC.
// A gather with sequence number q: each rank stores its block of records into mailbox parity q & 1 of EVERY
// rank, fences at system scope and publishes q in its flag there; each rank then waits until all of its own
// flags reached q and copies the mailbox into meas_all.  A rank can be at most one gather ahead of another
// (its next push comes after its own wait), so two parities suffice and flags only grow.
constexpr int PEER_MAX_RANKS = 16;
struct PeerPorts {
    double *mail[PEER_MAX_RANKS];              // mailbox of every rank (own rank: the local pointer)
    unsigned long long *flag[PEER_MAX_RANKS];  // arrival flags of every rank, [n_ranks] sequence numbers
    int n_ranks, rank;
    long long cap8;                            // doubles per mailbox parity
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// block g stores this rank's records mine[n_mine] at offset base8 of rank g's mailbox, then raises the flag
__global__ void __launch_bounds__(256) k_peer_push(PeerPorts pp, const double *mine, int n_mine, long long base8, unsigned long long seq) {
    const int g = blockIdx.x;
    double *dst = pp.mail[g] + (seq & 1ULL) * pp.cap8 + base8;
    for (int i = threadIdx.x; i < n_mine; i += blockDim.x) dst[i] = mine[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(pp.flag[g] + pp.rank, seq);
}

// k_reduce_partials with the push folded in: block `rep` stores its 8-double record into every rank's mailbox;
// the block that arrives last (counter *arrive, reset for the next launch) raises this rank's flag everywhere.
__global__ void __launch_bounds__(256) k_reduce_partials_push(const double *__restrict__ partials, int n_partials,
                                                              const unsigned long long *__restrict__ accepted,
                                                              const double *__restrict__ sigma, double *meas, int write_energy,
                                                              PeerPorts pp, long long base8, unsigned long long seq, unsigned int *arrive) {
    reduce_record(partials, n_partials, accepted, sigma, meas, write_energy);
    __syncthreads();
    const int rep = blockIdx.x;
    if ((int)threadIdx.x < pp.n_ranks * 8) {
        const int g = threadIdx.x >> 3, c = threadIdx.x & 7;
        pp.mail[g][(seq & 1ULL) * pp.cap8 + base8 + (long long)rep * 8 + c] = meas[(size_t)rep * 8 + c];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int sh_last;
    if (threadIdx.x == 0) sh_last = (atomicAdd(arrive, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (sh_last) {
        __threadfence_system();
        if (threadIdx.x == 0) *arrive = 0u;
        if ((int)threadIdx.x < pp.n_ranks) st_release_sys(pp.flag[threadIdx.x] + pp.rank, seq);
    }
}

// waits until every rank's flag reached seq, then copies the local mailbox parity into meas_all[n8].  A peer
// that never arrives (dead process) sets *err after timeout_ns instead of hanging the GPU.
__global__ void __launch_bounds__(256) k_peer_wait(const unsigned long long *flags, int n_ranks, unsigned long long seq,
                                                   const double *mail, double *__restrict__ meas_all, int n8,
                                                   volatile int *err, unsigned long long timeout_ns) {
    if ((int)threadIdx.x < n_ranks) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(flags + threadIdx.x) < seq) {
            if (global_timer_ns() - t0 > timeout_ns) { *err = 1 + (int)threadIdx.x; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n8; i += blockDim.x) meas_all[i] = __ldcv(mail + i);
}

}  // namespace csmc
