// csmc_internal.h — shared between the host model builder, the kernels and the C-ABI layer.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/csmc.h"

namespace csmc {

constexpr int MAXD = CSMC_MAX_DIM;

// ---- device-visible descriptors (passed by value as kernel parameters: they live in the
// ---- constant bank, so per-CTA-uniform lookups cost no global-memory traffic) --------------
struct SegGeom {          // where a segment lives in storage and its dense shape
    int32_t start;        // first storage position
    int32_t M[MAXD];      // supercell extents (structured); M[0]*M[1]*M[2] == count
};

struct DevSeg {           // one contiguous run of same-colour, same-basis sites
    int32_t start, count;
    int32_t M[MAXD];
    int32_t term_begin;   // first entry in terms[]
    int16_t n2, n3, n4;   // active bilinear / cubic / quartic entries (in that order)
    int16_t basis;        // 0-based
    int32_t onsite;       // offset into coefs[] of the 3x3 on-site matrix, -1 if zero
    int8_t P[MAXD];       // colouring period per dimension      (structured only)
    int8_t r[MAXD];       // residue of this class per dimension (structured only)
    int16_t geom;         // index of this segment in geom[]
    double h[3];          // Zeeman vector of the basis site
};
static_assert(sizeof(DevSeg) == 72, "DevSeg layout");

struct DevTerm {          // one active interaction slot of a segment
    int32_t coef;         // offset into coefs[]: 9 / 27 / 81 doubles, centre index first
    int16_t row;          // first row of this slot in the explicit neighbour table
    int16_t nseg[3];      // structured: geom[] index of each neighbour's segment
    int8_t d[3][MAXD];    // structured: supercell shift of each neighbour
    int8_t pad[3];
};
static_assert(sizeof(DevTerm) == 24, "DevTerm layout");

template <int NG_, int NS_, int NT_, int NC_>
struct PassParamsT {
    static constexpr int NG = NG_, NS = NS_, NT = NT_, NC = NC_;
    double *spins;           // [replica][3][npad]
    const int32_t *nbr;      // [rows][npad] explicit neighbour positions, -1 == null
    const int32_t *ref_of_pos; // [npad] reference (0-based) site index of a storage position
    int64_t rep_stride;      // 3 * npad
    int32_t npad;
    int32_t n_segs;          // segments in this pass
    int32_t L[MAXD];         // lattice shape (structured site-index arithmetic)
    int32_t periodic;
    double S;
    SegGeom geom[NG];        // all segments of the model
    DevSeg segs[NS];         // segments of this pass
    DevTerm terms[NT];
    double coefs[NC];
};
using PassSmall = PassParamsT<16, 8, 64, 384>;
using PassLarge = PassParamsT<128, 64, 320, 2048>;
static_assert(sizeof(PassLarge) < 32000, "kernel parameter limit");

// ---- host-side model -------------------------------------------------------------------------
struct HostSeg {
    int colour, basis;
    int start, count;
    int M[MAXD];
    int P[MAXD], r[MAXD];
    std::vector<int> sites;  // reference indices (0-based) in storage order (generic fallback)
};

struct HostTerm {
    int kind;             // 2, 3, 4
    int row;              // explicit-table row of the first neighbour
    int coef;             // offset into coefs
    int nb_basis[3];      // neighbour basis (0-based)
    int off[3][MAXD];     // neighbour cell offsets relative to the centre cell
};

struct HostModel {
    int D = 0, n_basis = 0, periodic = 1;
    int L[MAXD] = {1, 1, 1};
    int64_t N = 0;
    int N2 = 0, N3 = 0, N4 = 0;
    double S = 0.5;
    std::vector<double> field, onsite;   // per basis
    int n_colours = 0;
    bool structured = false;             // arithmetic-neighbour kernels applicable
    bool pattern = false;                // periodic colouring pattern found
    bool self_loop = false;              // some site interacts with itself
    int P[MAXD] = {1, 1, 1};
    int npad = 0;
    int n_rows = 0;                      // rows in the explicit neighbour table
    std::vector<int32_t> colour_of_site; // [N] reference order
    std::vector<int32_t> pos_of_ref;     // [N]
    std::vector<int32_t> ref_of_pos;     // [npad], -1 in alignment holes
    std::vector<int32_t> nbr;            // [n_rows][npad]
    std::vector<HostSeg> segs;           // sorted by colour
    std::vector<int> colour_seg_begin;   // [n_colours + 1]
    std::vector<std::vector<HostTerm>> basis_terms; // per basis: active slots (bilinear, cubic, quartic)
    std::vector<double> coefs;
    std::vector<int> onsite_coef;        // per basis offset or -1
    bool need_large = false;
};

// builds everything above; returns "" or an error message
std::string build_host_model(const csmc_model *m, int flags, HostModel &out);
// reference-layout tables (1-based, 0 == null), for csmc_get_tables
void reference_tables(const csmc_model *m, int64_t *bil, int64_t *cub, int64_t *quar);

// runtime specialisation (jit.cpp): model -> CUDA C++ source -> sm_100a cubin (NVRTC); "" on success
struct JitPlan {
    std::vector<int> tiles, groups, groups_metro;
    std::vector<int> tiles_per_cta;   // [colour * 4 + update kind]: consecutive tiles one CTA handles (CSMC_JIT_TPC experiment)
    bool resident = false;
    int sweep_tpb = 256;
    bool want_fused = false;     // in: also generate the fused full-sweep kernels (CSMC_FLAG_FUSED)
    bool fused = false;          // full-sweep kernel with shared-memory tiles (two-colour periodic models)
    int fused_tiles = 0, fused_smem = 0, fused_tpb = 256;
    // time-skewed strips (api.cu, enqueue_skewed): launches over a range of CTA-tile rows along lattice dimension 0
    bool want_skew = false;      // in: emit the sweep kernels with a tile offset (CSMC_SKEW)
    bool skew = false;           // out: usable -- every colour has the same tiling, whole tile rows, no split sites
    int skew_rows = 0;           // CTA-tile rows along dimension 0
    int skew_tiles_per_row = 0;  // CTA tiles per tile row (grid.x = rows * tiles_per_row)
    int skew_reach = 0;          // tile rows a site's neighbours can be away (>= 1)
    double flops_or_update = 0.0;   // out: fp64 flops of one overrelaxation site update (fma = 2), averaged over the sites
    // tile-resident persistent kernel (jit.cpp emit_persist, api.cu enqueue_persist): one CTA per SM keeps a tile of the
    // lattice (plus halo) in shared memory for a whole sequence of sweeps and exchanges only halos through L2
    int persist_replicas = 0;    // in: replicas of the handle (0: do not generate the kernel)
    bool persist_only = false;   // in: emit nothing but the persistent kernel (it lives in a module of its own)
    int persist_sms = 148;       // in: SMs of the device (CTAs that can be co-resident at one CTA per SM)
    int persist_smem_max = 227 * 1024;   // in: opt-in shared memory per CTA
    bool persist = false;        // out: usable
    int persist_tiles = 0;       // CTA tiles per replica (grid.x)
    int persist_nrep = 0;        // replicas per launch (grid.y <= this)
    int persist_smem = 0, persist_tpb = 512;
    int persist_g[2] = {1, 1}, persist_w[2] = {1, 1};   // tiles / tile extent (supercells) along dimensions 0 and 1
};   // per colour: grid.x (CTA tiles), grid.y (class groups)

// by-value argument of csmc_persist (mirrored in jit_prelude.h)
constexpr int PERSIST_MAX_OPS = 48;
struct PersistArgs {
    unsigned long long *flags;          // [replica][tile * PERSIST_FLAG_STRIDE]: passes completed, monotonic across launches
    int *err;                           // sticky error word (mapped host memory): a neighbour tile did not arrive in time
    unsigned long long timeout_cycles;  // spin budget per wait
    int n_ops;                          // sweeps of this launch
    unsigned char upd[PERSIST_MAX_OPS];        // update kind of sweep k
    unsigned short ctr_rel[PERSIST_MAX_OPS];   // Metropolis counter of sweep k relative to SweepArgs::ctr_off
    int pad_;
    unsigned long long *prof;           // optional (CSMC_PERSIST_PROF=1): [tile][8] cycles spent waiting / reloading / updating / publishing
};
constexpr int PERSIST_FLAG_STRIDE = 4;   // 32 bytes per flag: one L2 sector each
static_assert(sizeof(PersistArgs) == 184, "PersistArgs layout (jit_prelude.h mirrors it)");
// launches (pass, first tile row, tile rows) that run P colour passes strip by strip without changing any result;
// empty when the lattice is too small for the budget to matter (csmc_skew_schedule exports it for the tests)
struct SkewLaunch { int pass, row0, nrows; };
std::vector<SkewLaunch> skew_schedule(int n_rows, int n_passes, int reach, int budget_rows);
std::string jit_generate_source(const HostModel &hm, bool pdl = false, JitPlan *plan = nullptr);
std::string jit_compile(const std::string &src, std::vector<char> &cubin, std::string &log);

template <class P>
std::string fill_pass_params(const HostModel &hm, int colour, P &p);

}  // namespace csmc
