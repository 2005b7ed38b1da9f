/*
 * csmc.h — C-ABI of libcsmc.so, the B200 (sm_100a) sweep engine that replaces the bodies of
 * ClassicalSpinMC.jl's hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference (pure Julia) has no FFI today; each entry point below names the reference
 * function whose *body* it replaces (paths relative to the reference repo).  A Julia `ccall`
 * stub for every export is shown in INTEGRATION.md; `julia/ClassicalSpinMC/src/libcsmc.jl`
 * holds the real bindings.
 *
 * Conventions
 *   - every function returns int32 status (CSMC_OK == 0); no exception crosses the ABI.
 *     A human-readable message is available through csmc_last_error().
 *   - site indices are 1-based on the ABI (Julia convention), basis indices too.
 *   - spins: N x 3 doubles, row-major  (== Julia `Array{Float64,2}` of size 3 x N).
 *   - 3x3 matrices: 9 doubles row-major, m11 m12 m13 m21 ... (== `InteractionMatrix`,
 *     src/interaction_matrix.jl:2-12).
 *   - cubic / quartic tensors: Julia column-major, element [a,b,c(,d)] (0-based) at
 *     a + 3*b + 9*c (+ 27*d)   (== `Array{Float64,3}` / `Array{Float64,4}` memory).
 *   - site order: basis slowest, last lattice dimension fastest (src/lattice.jl:29-33):
 *     p = 1 + i_D + L_D*(i_{D-1} + ... + L_1*(b-1)),  i_d 0-based cell coordinate.
 *   - the caller owns every host buffer; the library copies during the call and never
 *     retains host pointers.  One controlling host thread per handle.
 *   - all calls are synchronous at return unless the name ends in `_async`.
 */
#ifndef CSMC_H
#define CSMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSMC_VERSION 100 /* 0.1.0 */

enum {
    CSMC_OK = 0,
    CSMC_ERR_INVALID = 1, /* bad argument / inconsistent model            */
    CSMC_ERR_CUDA = 2,    /* CUDA runtime failure (message has the detail) */
    CSMC_ERR_NCCL = 3,    /* NCCL failure or NCCL not loadable             */
    CSMC_ERR_UNSUPPORTED = 4,
    CSMC_ERR_NOMEM = 5
};

#define CSMC_MAX_DIM 3

/*
 * Unit-cell level description of the Hamiltonian and lattice: what `Lattice(shape, uc, S; bc)`
 * consumes (src/lattice.jl:65-291, src/unit_cell.jl:3-75).  The library derives the per-site
 * neighbour / coupling tables of the reference in closed form (O(N * terms) instead of the
 * reference's O(N^2 * terms) `findfirst` scans, src/lattice.jl:196,228-229,273-275).
 */
typedef struct csmc_model {
    int32_t dim;                 /* D, 1..3                                              */
    int32_t shape[CSMC_MAX_DIM]; /* unit cells per dimension                             */
    int32_t n_basis;             /* >= 1                                                 */
    int32_t periodic;            /* 1: bc="periodic", 0: bc="open" (src/lattice.jl:101)  */
    double S;                    /* spin length                                          */
    const double *field;         /* n_basis x 3, per-basis h (resolved as lattice.jl:117-140) */
    const double *onsite;        /* n_basis x 9, per-basis on-site matrix (row-major)    */

    int32_t n_bilinear;          /* N2 = length(uc.bilinear)                             */
    const int32_t *bil_basis;    /* N2 x 2  (b1,b2), 1-based                             */
    const int32_t *bil_offset;   /* N2 x D  unit-cell offset of site 2                   */
    const double *bil_matrix;    /* N2 x 9  row-major                                    */

    int32_t n_cubic;             /* N3                                                   */
    const int32_t *cub_basis;    /* N3 x 3                                               */
    const int32_t *cub_offset;   /* N3 x 2 x D  (o2, o3)                                 */
    const double *cub_tensor;    /* N3 x 27 column-major                                 */

    int32_t n_quartic;           /* N4                                                   */
    const int32_t *quar_basis;   /* N4 x 4                                               */
    const int32_t *quar_offset;  /* N4 x 3 x D  (o2, o3, o4)                             */
    const double *quar_tensor;   /* N4 x 81 column-major                                 */
} csmc_model;

typedef struct csmc_opts {
    int32_t device;       /* CUDA device ordinal                                          */
    int32_t n_replicas;   /* replicas (temperatures) held by THIS handle / GPU, >= 1      */
    uint64_t seed;        /* Philox key; same seed on every rank of a PT job              */
    void *stream;         /* cudaStream_t to launch on, or NULL: library creates its own  */
    int32_t replica_base; /* global index of local replica 0 (PT sharding), else 0        */
    int32_t flags;        /* CSMC_FLAG_*                                                  */
} csmc_opts;

#define CSMC_FLAG_FORCE_GENERIC 1 /* use the explicit-index-table kernels even when the      \
                                     arithmetic-neighbour (structured) kernels apply        */
#define CSMC_FLAG_NO_GRAPH 2      /* plain stream launches instead of CUDA-graph replay      */
#define CSMC_FLAG_JIT 4           /* require the runtime-specialised (NVRTC) kernels: fail     \
                                     csmc_create if they cannot be built                      */
#define CSMC_FLAG_NO_JIT 8        /* never specialise at run time (ahead-of-time kernels only) */
#define CSMC_FLAG_PDL 16          /* always launch the specialised passes with programmatic      \
                                     dependent launch (griddepcontrol); by default csmc_create    \
                                     times both modes on the model and keeps the faster          */
#define CSMC_FLAG_NO_RESIDENT 32  /* never use the resident (one CTA per replica) kernel           */
#define CSMC_FLAG_NO_AUTOTUNE 64  /* skip the launch-mode autotune at csmc_create (eager path):      \
                                     programmatic dependent launch is then off unless CSMC_FLAG_PDL */
#define CSMC_FLAG_FUSED 128       /* experimental, off by default: run pairs of consecutive sweeps on   \
                                     the fused full-sweep kernels (one launch per sweep: tile + halo  \
                                     in shared memory, both colours, ping-pong spin buffers;          \
                                     two-colour periodic 1-D/2-D models).  Measured slower than the   \
                                     per-colour passes on B200 (shared-memory wavefront bound, see    \
                                     DESIGN.md section 5), kept for that comparison                   */
#define CSMC_FLAG_SKEW 256        /* build the tile-offset kernels even for small lattices (also CSMC_SKEW=1;     \
                                     CSMC_SKEW=0 never builds them; default: built when one replica's spins exceed    \
                                     the L2 budget).  With them a single lattice whose                                \
                                     spins exceed the L2 budget (64 MiB; CSMC_L2_BLOCK_MB) runs sequences of sweeps  \
                                     as time-skewed strips of CTA-tile rows -- all colour passes of the sequence on   \
                                     one strip while it is L2-resident, the strip moving one dependency reach per     \
                                     pass -- instead of streaming the lattice through L2 once per pass.  Results are  \
                                     bit-identical (csmc_skew_schedule, csmc_skew_info)                               */
#define CSMC_FLAG_NO_PERSIST 512  /* never use the tile-resident persistent kernel (also CSMC_PERSIST=0)              */
#define CSMC_FLAG_PERSIST 1024    /* build and use it whenever it applies (also CSMC_PERSIST=1; CSMC_PERSIST=probe builds it \
                                     and lets the create-time probe choose between it and the per-colour pass kernels;   \
                                     default: not built -- it lost to the pass kernels on every measured workload).  The tile-resident kernel (csmc_persist_info)    \
                                     runs a whole sequence of sweeps in ONE cooperative launch: every CTA (one per SM)    \
                                     keeps a tile of the lattice plus its halo in shared memory for all colour passes of  \
                                     the sequence, stores the sites other tiles read to the global spin array after each  \
                                     pass and synchronises with its <= 8 neighbour tiles only (release / acquire progress \
                                     counters in L2, no grid-wide barrier).  Applies to periodic pattern-coloured models  \
                                     whose replicas fit the SMs' shared memory (<= ~28 MB of spins per launch; more        \
                                     replicas run batch by batch).  Results are bit-identical to the pass kernels.         */
/* Default: models whose colouring is a periodic pattern get kernels specialised for that model
 * (unrolled terms, literal coefficients, constant geometry; compiled for sm_100a with NVRTC at
 * csmc_create) when n_sites * n_replicas >= 32768; smaller problems are launch-latency bound and
 * use the ahead-of-time kernels for short requests; the first request of >= 32 sweeps on a lattice
 * of <= 4096 sites builds the specialised kernels lazily and from then on runs whole sweep schedules
 * in ONE launch on the "resident" kernel: one CTA per replica, the lattice held in shared memory,
 * __syncthreads() between colour passes (csmc_kernel_mode == 3). */

typedef struct csmc_handle csmc_handle;

/* ---- library / handle ------------------------------------------------------------------ */
int32_t csmc_version(void);
/* message of the last failure on `h` (or, with h == NULL, of the last failed csmc_create). */
const char *csmc_last_error(const csmc_handle *h);

/* Replaces: table-building body of `Lattice(...)`, src/lattice.jl:168-288, plus the colouring
 * of the interaction hypergraph that makes a colour class updatable race-free. */
int32_t csmc_create(const csmc_model *model, const csmc_opts *opts, csmc_handle **out);
int32_t csmc_destroy(csmc_handle *h);

/* Host-only planning step of csmc_create (no CUDA call, usable without a GPU): colours the
 * interaction hypergraph and reports what csmc_create would set up.  colour[N] (reference site
 * order, may be NULL), *n_colours, *structured (1: arithmetic-neighbour kernels apply),
 * storage_pos[N] (may be NULL): position of each site in the colour-major device layout. */
int32_t csmc_plan(const csmc_model *model, int32_t flags, int32_t *colour, int32_t *n_colours,
                  int32_t *structured, int32_t *storage_pos);

/* Host-only: the per-site neighbour tables of the reference's Lattice constructor in closed form
 * (lat.bilinear_sites / cubic_sites / quartic_sites, src/lattice.jl:176-286): bil[N x N2],
 * cub[N x N3 x 2], quar[N x N4 x 3]; 1-based, 0 == null slot.  Any pointer may be NULL. */
int32_t csmc_reference_tables(const csmc_model *model, int64_t *bil, int64_t *cub, int64_t *quar);

int32_t csmc_n_sites(const csmc_handle *h, int64_t *n);
int32_t csmc_n_replicas(const csmc_handle *h, int32_t *r);
int32_t csmc_n_colours(const csmc_handle *h, int32_t *c);
/* colour[N] (0-based colour of site p, reference order) */
int32_t csmc_get_colouring(const csmc_handle *h, int32_t *colour);
/* 1 if the arithmetic-neighbour kernels are in use, 0 if the explicit-table kernels are. */
int32_t csmc_is_structured(const csmc_handle *h, int32_t *flag);
/* which pass kernels this handle launches: 0 explicit-table, 1 arithmetic-neighbour (both ahead of
 * time), 2 runtime-specialised for this model (NVRTC, sm_100a), 3 runtime-specialised with the
 * resident small-lattice kernel for sweep schedules. */
int32_t csmc_kernel_mode(const csmc_handle *h, int32_t *mode);
/* Result of the launch-mode autotune of csmc_create: ms[0] / ms[1] = time of the probe run without /
 * with programmatic dependent launch (0 when the autotune did not run), *pdl_selected = mode in use. */
int32_t csmc_autotune_report(const csmc_handle *h, float ms[2], int32_t *pdl_selected);
/* Replica groups: with several replicas in a handle, sequences of >= 2 sweeps run as `groups`
 * independent chains (n_replicas / groups replicas each) on separate streams -- parallel branches of
 * the replayed CUDA graph -- so that the fixed latencies of the single-wave colour passes overlap.
 * csmc_create times 1, 2 and 4 groups on the model (same probe run as above) and keeps the fastest;
 * ms[0..2] are those times (0 when not measured).  The environment variable CSMC_SWEEP_GROUPS
 * overrides.  Results do not depend on the group count (replicas are independent chains). */
int32_t csmc_sweep_groups(const csmc_handle *h, int32_t *groups, float ms[3]);
/* Replica blocks: when the spins of all replicas of the handle exceed the L2 budget (64 MiB; CSMC_L2_BLOCK_MB),
 * a sequence of sweeps enqueued at once (the OR block + Metropolis sweep between two exchanges) can run block
 * by block -- all sweeps for the first replicas, then for the next -- so that each block is read from HBM once
 * and stays L2-resident for every colour pass of the sequence.  csmc_create times the unblocked order against the
 * block count the budget asks for and one more, and keeps the fastest (ms[0] all replicas per pass, ms[1] the
 * faster blocked candidate; 0 when not probed); CSMC_REPLICA_BLOCKS=n forces n blocks.
 * Results do not depend on it (replicas are independent between exchanges). */
int32_t csmc_replica_blocks(const csmc_handle *h, int32_t *blocks, float ms[2]);
/* Time-skewed strips (CSMC_FLAG_SKEW).  csmc_skew_schedule is host-only: the launch plan for n_passes colour passes
 * over n_rows CTA-tile rows when a site's neighbours are at most `reach` rows away and budget_rows rows fit in L2:
 * launches[3 * i] = {pass, first row, rows}; *n = number of launches (0: not applicable, run pass by pass);
 * at most cap triples are written.  Executed in that order every site update reads exactly the values the
 * pass-by-pass order would give it (tests/test_host_plan.py checks this for the plan itself, the GPU tests for the
 * spins).  csmc_skew_info: whether the handle runs such plans (its kernels carry the tile offset and the create-time
 * probe did not find the pass-by-pass order faster), and the geometry they use. */
int32_t csmc_skew_schedule(int32_t n_rows, int32_t n_passes, int32_t reach, int32_t budget_rows,
                           int32_t *launches, int64_t cap, int64_t *n);
/* Tile-resident persistent kernel: *tiles = CTA tiles per replica in use (0: not in use for this handle), grid[2] =
 * tiles along lattice dimensions 0 and 1, *replicas_per_launch, *smem_bytes per CTA; ms[2] = create-time probe
 * (pass kernels / persistent kernel; zeros if not measured).  Any pointer may be NULL. */
int32_t csmc_persist_info(const csmc_handle *h, int32_t *tiles, int32_t grid[2], int32_t *replicas_per_launch,
                          int32_t *smem_bytes, float ms[2]);
/* Roofline inputs of the handle's runtime-specialised kernels, counted by the code generator: fp64 flops of one
 * overrelaxation site update (neighbour field over the unrolled bilinear / cubic / quartic terms + the reflection
 * s <- -s + 2 (s.F)/(F.F) F, src/monte_carlo.jl:126-139; fma = 2 flops, averaged over the sites; 0 without specialised
 * kernels) and the algorithmic bytes per single-spin update, 24 (C + 1) for C colours.  Either pointer may be NULL. */
int32_t csmc_kernel_costs(const csmc_handle *h, double *flops_per_or_update, double *bytes_per_update);
/* host-only (no GPU needed): plans -- and with compile != 0 compiles for sm_100a -- the tile-resident kernel of `model`
 * for n_replicas replicas on a device with n_sms SMs and smem_max bytes of opt-in shared memory per CTA (0: B200's 148 /
 * 227 KiB).  info = {usable, tiles per replica, tiles along dim 0, dim 1, tile extent (supercells) along dim 0, dim 1,
 * replicas per launch, shared memory per CTA}.  source / log as csmc_jit_check. */
int32_t csmc_persist_check(const csmc_model *model, int32_t n_replicas, int32_t n_sms, int32_t smem_max, int32_t compile,
                           char *source, int64_t source_cap, int64_t *source_len, char *log, int64_t log_cap, int32_t info[8]);
int32_t csmc_skew_info(const csmc_handle *h, int32_t *usable, int32_t *tile_rows, int32_t *reach,
                       int32_t *budget_rows);
/* host-only: the strip geometry the specialised kernels of `model` would use (CTA-tile rows along lattice
 * dimension 0, dependency reach in rows, CTA tiles per row); *usable = 0 when the model cannot be strip-mined */
int32_t csmc_skew_geometry(const csmc_model *model, int32_t *usable, int32_t *tile_rows, int32_t *reach,
                           int32_t *tiles_per_row);
/* Host-only (no GPU needed): generate the specialised kernel source for `model` and, if
 * compile != 0, compile it with NVRTC for sm_100a.  source/log may be NULL; *_cap are buffer sizes;
 * *source_len receives the full source length.  Used by build checks and tests. */
int32_t csmc_jit_check(const csmc_model *model, int32_t compile, char *source, int64_t source_cap,
                       int64_t *source_len, char *log, int64_t log_cap);
/* kernels launched on this handle since creation (bench.py's `gpu_launches`). */
int32_t csmc_launch_count(const csmc_handle *h, int64_t *n);

/* Reference-layout neighbour tables as the library derived them (for parity tests against
 * lat.bilinear_sites / cubic_sites / quartic_sites, src/lattice.jl:205,238,285):
 * bil[N x N2], cub[N x N3 x 2], quar[N x N4 x 3]; 1-based, 0 == null slot. Any may be NULL. */
int32_t csmc_get_tables(const csmc_handle *h, int64_t *bil, int64_t *cub, int64_t *quar);

/* ---- state ------------------------------------------------------------------------------ */
/* Replaces direct writes to `lattice.spins` (src/lattice.jl:297-301). replica is local, 0-based */
int32_t csmc_set_spins(csmc_handle *h, int32_t replica, const double *spins);
int32_t csmc_get_spins(csmc_handle *h, int32_t replica, double *spins);
/* Device-side `random_spin_orientation` for every site of every replica (src/lattice.jl:76-79,
 * 306-311), Philox stream (seed, global replica, site). */
int32_t csmc_randomize_spins(csmc_handle *h, uint64_t seed);

/* ---- Hamiltonian evaluation ------------------------------------------------------------- */
/* Replaces get_local_field(lattice, point), src/hamiltonian.jl:3-67 (returns H - h). */
int32_t csmc_local_field(csmc_handle *h, int32_t replica, int64_t site, double out[3]);
int32_t csmc_local_field_all(csmc_handle *h, int32_t replica, double *out /* N x 3 */);
/* Replaces energy(lattice, point), src/hamiltonian.jl:139-196, for every site. */
int32_t csmc_site_energy_all(csmc_handle *h, int32_t replica, double *out /* N */);
/* Replaces total_energy(lattice), src/hamiltonian.jl:70-132. E[n_replicas]. */
int32_t csmc_total_energy(csmc_handle *h, double *E);
/* Replaces the vector sum inside get_magnetization, src/observables.jl:12-18: M3[n_replicas x 3]
 * (the caller takes the norm). */
int32_t csmc_magnetization(csmc_handle *h, double *M3);

/* Replaces compute_equal_time_correlations(lat, ks), src/spin_correlations.jl:6-43:
 * Suv[3u+v, n] = Re(A_u(k_n) conj(A_v(k_n))) / N, A_u(k) = sum_i exp(-i k.r_i) s_i^u.
 * lattice_vectors: D x D column-major (column d = a_d, as unit_cell/lattice_vectors); basis: n_basis x D
 * row-major; ks: D x n_k column-major (Julia Matrix{Float64}(D, N_k)); Suv: 9 x n_k column-major. */
int32_t csmc_structure_factor(csmc_handle *h, int32_t replica, const double *lattice_vectors,
                              const double *basis, const double *ks, int64_t n_k, double *Suv);

/* ---- sweeps ------------------------------------------------------------------------------ */
/* Replaces overrelaxation!(lattice), src/monte_carlo.jl:126-139: n_sweeps colour-ordered sweeps. */
int32_t csmc_overrelax(csmc_handle *h, int32_t n_sweeps);
/* Replaces the body of deterministic_updates!, src/monte_carlo.jl:201-213, as colour-ordered
 * full sweeps s <- -F/|F| * S. */
int32_t csmc_deterministic(csmc_handle *h, int32_t n_sweeps);
/* Replaces metropolis!(mc, T), src/metropolis.jl:65-82 (+ calculate_energy_diff!, :94-101):
 * n_sweeps colour-ordered sweeps at per-replica temperatures T[n_replicas];
 * accepted[n_replicas] accumulates accepted proposals (may be NULL). */
int32_t csmc_metropolis(csmc_handle *h, const double *T, int32_t n_sweeps, double *accepted);
/* Cone-move variant (gaussian_move, src/metropolis.jl:84-87,103-153) with per-replica sigma.
 * adapt != 0 applies the MetropolisAdaptive rule (src/metropolis.jl:129-131) after each sweep
 * and writes the new sigma back. */
int32_t csmc_metropolis_cone(csmc_handle *h, const double *T, double *sigma, int32_t adapt,
                             int32_t n_sweeps, double *accepted);

/* Replaces the inner `while t < t_thermalization` loop of simulated_annealing!,
 * src/monte_carlo.jl:169-182, at per-replica temperatures T: for t = 1 .. t_thermalization-1:
 * overrelaxation sweep (if rate != 0), Metropolis sweep when t % rate == 0 (every t if rate == 0).
 * accepted[n_replicas] (may be NULL) receives the accepted-proposal totals. */
int32_t csmc_anneal_temperature(csmc_handle *h, const double *T, int64_t t_thermalization,
                                int32_t overrelaxation_rate, double *accepted);

/* Same loop with the cone-move algorithms (alg=MetropolisAdaptive() when adapt != 0, else
 * MetropolisFixedCone()): sigma[n_replicas] is in/out (the caller resets it to sigma0 per temperature,
 * src/monte_carlo.jl:171). */
int32_t csmc_anneal_temperature_cone(csmc_handle *h, const double *T, double *sigma, int32_t adapt,
                                     int64_t t_thermalization, int32_t overrelaxation_rate,
                                     double *accepted);

/* Steady-state throughput loop used by bench.py: n_cycles x (or_per_cycle overrelaxation sweeps +
 * metro_per_cycle Metropolis sweeps) at the current temperatures, no host sync inside.
 * The `_async` form returns after enqueueing on the handle's stream. */
int32_t csmc_set_temperatures(csmc_handle *h, const double *T);
/* per-replica cone width `mc.sigma` (src/metropolis.jl:24, default sigma0 = 60) used by the cone-move
 * variants inside csmc_pt_run; csmc_get_sigma reads the (possibly adapted) values back. */
int32_t csmc_set_sigma(csmc_handle *h, const double *sigma);
int32_t csmc_get_sigma(csmc_handle *h, double *sigma);
int32_t csmc_cycles_async(csmc_handle *h, int64_t n_cycles, int32_t or_per_cycle,
                          int32_t metro_per_cycle);
int32_t csmc_sync(csmc_handle *h);
/* accepted proposals per replica since the last call with reset != 0 */
int32_t csmc_get_accepted(csmc_handle *h, double *accepted, int32_t reset);

/* ---- parallel tempering (src/monte_carlo.jl:235-398) ------------------------------------ */
/* The job has n_slots temperature slots T_all[n_slots] (slot == the reference's MPI rank).
 * This handle's local replica r starts in slot replica_base + r.  Temperatures are exchanged,
 * configurations never move. */
int32_t csmc_pt_init(csmc_handle *h, int32_t n_slots, const double *T_all);
/* Multi-GPU: rank 0 calls csmc_comm_unique_id, the host broadcasts the 128 bytes (e.g. with
 * torch.distributed / MPI), every rank calls csmc_comm_init.  Single-GPU jobs skip both.
 * The ranks' replica blocks [replica_base, replica_base + n_replicas) must tile the slots in rank
 * order; they need not be equally long (equal blocks are gathered with one ncclAllGather, unequal
 * ones with one grouped ncclBroadcast per rank). */
int32_t csmc_comm_unique_id(uint8_t id[128]);
int32_t csmc_comm_init(csmc_handle *h, int32_t n_ranks, int32_t rank, const uint8_t id[128]);
/* How the per-replica measurement records (the energies of the exchange test, which the reference sends
 * with MPI.Sendrecv!, src/monte_carlo.jl:321-343) travel between the ranks' GPUs:
 * 0 no communicator, 1 NCCL collectives (default), 2 stores into peer memory (one push kernel + one wait
 * kernel instead of the collective), 3 the same with the push folded into the energy reduction kernel.
 * 2 / 3 are selected with the environment variable CSMC_PEER_GATHER=1 / 2, set identically for every rank
 * before csmc_comm_init: each rank's mailbox is mapped into the other processes with CUDA IPC (one node,
 * NVLink / NVSwitch peer access); if any rank cannot map a peer the job stays on 1.  Results do not depend
 * on the mode. */
int32_t csmc_comm_mode(const csmc_handle *h, int32_t *mode);

typedef struct csmc_pt_params {
    int64_t t_thermalization;
    int64_t t_measurement;
    int32_t probe_rate;
    int32_t swap_rate;
    int32_t overrelaxation_rate;
    int32_t algorithm; /* the `alg` kwarg: 0 Metropolis(), 1 MetropolisAdaptive(), 2 MetropolisFixedCone() */
} csmc_pt_params;

/* Replaces the body of the `while sweep < total_sweeps` loop, src/monte_carlo.jl:295-388, for
 * sweeps sweep_begin .. sweep_end-1 (host chunks the run at checkpoint / report boundaries):
 * overrelaxation (:298-300), Metropolis + total_energy when sweep % dosweep == 0 (:302-305),
 * replica exchange when sweep % swap_rate == 0 (:308-349), E/M probe when
 * sweep >= t_thermalization and sweep % probe_rate == 0 (:353,368-370). */
int32_t csmc_pt_run(csmc_handle *h, const csmc_pt_params *p, int64_t sweep_begin,
                    int64_t sweep_end);
/* mc.corr = true (src/monte_carlo.jl:371-375, src/observables.jl:27-30): from now on every probe also
 * computes the equal-time structure factor of each local replica at the given momenta and adds it to
 * the accumulator of the replica's current temperature slot.  csmc_pt_get_ssf returns the per-slot sums
 * [n_slots x n_k x 9] (this rank's contributions; the host adds the ranks and divides by *n_probes). */
int32_t csmc_pt_set_momenta(csmc_handle *h, const double *lattice_vectors, const double *basis,
                            const double *ks, int64_t n_k);
int32_t csmc_pt_get_ssf(csmc_handle *h, double *sums, int64_t *n_probes);
/* Replaces the exchange block alone (src/monte_carlo.jl:308-349) on fresh energies;
 * parity = (sweep / swap_rate) % 2 selects the pairing (:311-315) and nothing else: the uniforms of the
 * acceptance test (:327-330) come from the shared Philox stream at a per-handle call index that advances
 * with every call (reset by csmc_pt_init), so a caller's own loop over this export is a valid chain.
 * accepted_pairs[n_slots] (may be NULL): 1 where the pair starting at that slot swapped. */
int32_t csmc_pt_exchange(csmc_handle *h, int32_t parity, int32_t *accepted_pairs);
/* slot_of_replica[n_slots] for ALL global replicas (identical on every rank). */
int32_t csmc_pt_get_slots(csmc_handle *h, int32_t *slot_of_replica);
/* Probe series recorded so far, attributed to temperature slots:
 * E[n_probes x n_slots], M[n_probes x n_slots] (|sum s|). *n_probes is in/out (capacity/count).
 * Only slots whose replica lives on this rank are filled when no communicator is attached
 * across ranks; with a communicator every rank holds all slots. */
int32_t csmc_pt_get_series(csmc_handle *h, int64_t *n_probes, double *E, double *M);
/* cumulative statistics, per slot: accepted Metropolis proposals and accepted exchanges
 * (src/monte_carlo.jl:269-274, src/helper.jl:25-78). */
int32_t csmc_pt_get_stats(csmc_handle *h, double *accepted_local, double *exchanges);

#ifdef __cplusplus
}
#endif
#endif /* CSMC_H */
