// api.cu — the C-ABI of libcsmc.so (include/csmc.h): handle management, launch sequencing, NCCL.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace csmc;

// ---- NCCL, loaded lazily so single-GPU users need no NCCL at all ------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclFloat64 = 8 };
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { g_nccl.err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(g_nccl.lib, "ncclBroadcast");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(g_nccl.lib, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(g_nccl.lib, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.Broadcast ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        g_nccl.err = "libnccl lacks required symbols";
        dlclose(g_nccl.lib); g_nccl.lib = nullptr;
        return false;
    }
    return true;
}
std::string g_create_error;
}  // namespace

struct csmc_handle {
    HostModel hm;
    // deep copy of the model (for csmc_get_tables)
    csmc_model model;
    std::vector<double> m_field, m_onsite, m_bilJ, m_cubT, m_quarT;
    std::vector<int32_t> m_bilB, m_bilO, m_cubB, m_cubO, m_quarB, m_quarO;

    int device = 0, R = 1, replica_base = 0, flags = 0;
    unsigned long long seed = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    double *d_spins = nullptr, *d_stage = nullptr, *d_out = nullptr;
    double *d_spins_alt = nullptr;     // second spin buffer of the fused full-sweep kernels (ping-pong), allocated on first use
    int32_t *d_nbr = nullptr, *d_ref_of_pos = nullptr;
    double *d_beta = nullptr, *d_sigma = nullptr;
    unsigned long long *d_acc = nullptr, *d_acc_prev = nullptr, *d_ctr = nullptr;
    double *d_partials = nullptr, *d_meas = nullptr;
    int n_partials = 0;
    std::vector<int> pass_blocks, partial_base;
    std::vector<PassSmall> ps;
    std::vector<PassLarge> pl;
    bool large = false;
    unsigned long long metro_ctr = 0;  // Metropolis sweeps enqueued so far (Philox counter)
    std::vector<unsigned long long> acc_base;  // per-replica counter value at last reset

    // parallel tempering
    int n_slots = 0;
    double *d_T_slot = nullptr, *d_meas_all = nullptr, *d_E_last = nullptr, *d_acc_prev_pt = nullptr;
    double *d_acc_slot = nullptr, *d_exch_slot = nullptr, *d_series_E = nullptr, *d_series_M = nullptr;
    int *d_slot_of_rep = nullptr, *d_rep_of_slot = nullptr, *d_accepted_pairs = nullptr, *d_prev_rep_of_slot = nullptr;
    long long series_cap = 0, n_probes = 0;
    unsigned long long exchange_calls = 0;   // csmc_pt_exchange calls so far: Philox counter of their uniforms
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;

    // runtime-specialised kernels
    bool jit = false;
    cudaLibrary_t jit_lib = nullptr;
    std::vector<cudaKernel_t> jit_sweep[4], jit_energy;
    JitPlan jit_plan;
    cudaKernel_t jit_resident = nullptr;
    cudaKernel_t jit_fused[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaKernel_t jit_persist = nullptr;
    cudaLibrary_t persist_lib = nullptr;             // its own module: independent of the launch-mode (PDL) variants
    JitPlan persist_plan;
    bool jit_tried = false, jit_pdl = false;
    // tile-resident persistent kernel (jit.cpp emit_persist)
    unsigned long long *d_persist_flags = nullptr;   // [R][tiles * PERSIST_FLAG_STRIDE] progress counters
    int *persist_err = nullptr;                      // mapped host memory, sticky
    bool persist_off = false;                        // autotune / CSMC_PERSIST=0 / a failed launch: pass kernels instead
    int persist_min_sweeps = 2;
    unsigned long long *d_persist_prof = nullptr;    // CSMC_PERSIST_PROF=1: per-tile phase cycle counters of the last launch
    float tune_persist_ms[2] = {0.f, 0.f};           // autotune: ms per probe run with pass kernels / persistent kernel
    float tune_ms[2] = {0.f, 0.f};   // autotune: ms per probe run without / with programmatic dependent launch
    std::string jit_note;

    // equal-time structure factor
    SsfGeom ssf{};
    int ssf_chunks = 0;
    double *d_ssf_theta = nullptr, *d_ssf_phib = nullptr, *d_ssf_partial = nullptr, *d_ssf_out = nullptr, *d_ssf_sum = nullptr;
    long long ssf_probes = 0;

    // CUDA graph of one bench cycle
    struct CycleGraph { cudaGraphExec_t exec; long long launches; };
    std::map<std::pair<int, int>, CycleGraph> cycle_graphs;   // keyed by (OR sweeps, Metropolis sweeps) per cycle
    // CUDA graphs of n consecutive overrelaxation sweeps (parallel-tempering loop)
    std::map<int, cudaGraphExec_t> or_graphs;
    std::map<int, long long> or_graph_launches;
    // replica groups on separate streams (sweep_groups)
    std::map<int, std::vector<SkewLaunch>> skew_cache;   // time-skewed launch plans by number of passes
    bool skew_off = false;             // autotune found the pass-by-pass order faster on this model
    float tune_skew_ms[2] = {0.f, 0.f};   // autotune: ms per probe run pass by pass / strip by strip
    int n_blocks = 1;                  // replica blocks run one after the other (L2 residency), see enqueue_sweep_seq
    float tune_blocks_ms[2] = {0.f, 0.f};   // autotune: ms per probe run unblocked / with n_blocks_wanted blocks
    int n_groups = 0;                  // 0: not decided yet
    float tune_groups_ms[3] = {0.f, 0.f, 0.f};   // autotune: ms per probe run with 1 / 2 / 4 groups
    std::vector<cudaStream_t> aux_streams;
    std::vector<cudaEvent_t> aux_done;
    cudaEvent_t ev_fork = nullptr;
    // multi-GPU: replica block of every rank (gathered at csmc_comm_init)
    std::vector<long long> rank_base, rank_count;
    bool even_partition = true;
    // measurement records gathered by stores into peer memory instead of NCCL (opt-in, CSMC_PEER_GATHER)
    struct PeerGather {
        int mode = 0;                          // 0 off, 1 push + wait kernels, 2 push folded into the energy reduction
        unsigned char *base = nullptr;         // local allocation: [flags 128 B][arrival counter 128 B][mailbox 2 x cap8 doubles]
        void *peer[PEER_MAX_RANKS] = {};       // the other ranks' allocations (cudaIpcOpenMemHandle)
        PeerPorts ports{};
        unsigned long long seq = 0;            // gathers issued so far (same on every rank)
        bool pushed = false;                   // the measurement just enqueued already pushed its records (mode 2)
        int *err = nullptr;                    // mapped host memory: set by k_peer_wait on timeout
        unsigned long long timeout_ns = 30000000000ULL;
    } peer;

    bool capture_open = false;   // a stream capture this handle began has not been ended (a call failed in between)
    std::string err;
    long long launches = 0;
};

namespace {

int fail(csmc_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, CSMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)
#define CKN(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != 0)                                                                               \
            return fail(h, CSMC_ERR_NCCL, std::string(#call) + ": " +                              \
                                              (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
    } while (0)

template <class T> cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

// one colour pass over the local replicas [a.rep0, a.rep0 + nrep) on `stream`
template <int UPD>
void launch_sweep_pass(csmc_handle *h, int colour, const SweepArgs &a, cudaStream_t stream, int nrep, int n_tiles = -1) {
    const int nseg = h->hm.colour_seg_begin[colour + 1] - h->hm.colour_seg_begin[colour];
    dim3 grid(h->pass_blocks[colour], nseg, nrep), block(TPB);
    if (h->jit) {
        // multi-dimensional CTA tiles over supercell coordinates, classes fused per thread (jit.cpp)
        SweepArgs a_full = a;          // a whole-grid launch of kernels built with a tile range covers every tile
        if (n_tiles < 0) { a_full.tile_off = 0; a_full.tile_end = h->jit_plan.tiles[colour]; }
        void *args[] = {(void *)&h->d_spins, (void *)&a_full};
        cudaLaunchConfig_t cfg{};
        // n_tiles >= 0: only the CTA tiles [a.tile_off, a.tile_off + n_tiles) (time-skewed strips, kernels built with CSMC_SKEW)
        const int tpc = (size_t)(colour * 4 + UPD) < h->jit_plan.tiles_per_cta.size() ? std::max(1, h->jit_plan.tiles_per_cta[colour * 4 + UPD]) : 1;
        const int tiles = n_tiles >= 0 ? n_tiles : h->jit_plan.tiles[colour];
        cfg.gridDim = dim3((tiles + tpc - 1) / tpc, (UPD >= UPD_METRO ? h->jit_plan.groups_metro : h->jit_plan.groups)[colour], nrep);
        cfg.blockDim = dim3(h->jit_plan.sweep_tpb);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = h->jit_pdl ? 1 : 0;
        cudaLaunchKernelExC(&cfg, (const void *)h->jit_sweep[UPD][colour], args);
    } else if (h->large) {
        if (h->hm.structured) k_sweep<PassLarge, true, UPD><<<grid, block, 0, stream>>>(h->pl[colour], a);
        else k_sweep<PassLarge, false, UPD><<<grid, block, 0, stream>>>(h->pl[colour], a);
    } else {
        if (h->hm.structured) k_sweep<PassSmall, true, UPD><<<grid, block, 0, stream>>>(h->ps[colour], a);
        else k_sweep<PassSmall, false, UPD><<<grid, block, 0, stream>>>(h->ps[colour], a);
    }
    h->launches++;
}

SweepArgs sweep_args(csmc_handle *h, unsigned long long ctr_off, bool device_ctr) {
    SweepArgs a{};
    a.beta = h->d_beta; a.sigma = h->d_sigma; a.accepted = h->d_acc;
    a.ctr_base = device_ctr ? h->d_ctr : nullptr;
    a.ctr_off = ctr_off; a.seed = h->seed; a.replica_base = h->replica_base; a.rep0 = 0;
    return a;
}

// one full sweep = every colour once, in colour order
template <int UPD>
void enqueue_sweep(csmc_handle *h, unsigned long long ctr_off = 0, bool device_ctr = false, cudaStream_t stream = nullptr,
                   int rep0 = 0, int nrep = -1) {
    SweepArgs a = sweep_args(h, ctr_off, device_ctr);
    a.rep0 = rep0;
    for (int c = 0; c < h->hm.n_colours; ++c) launch_sweep_pass<UPD>(h, c, a, stream ? stream : h->stream, nrep < 0 ? h->R : nrep);
}

// ---- fused full-sweep kernels (jit.cpp emit_fused): one launch per sweep, ping-pong between d_spins and d_spins_alt
bool fused_ready(csmc_handle *h) {
    if (!h->jit || !h->jit_plan.fused || !(h->flags & CSMC_FLAG_FUSED)) return false;
    if (!h->d_spins_alt) {   // not during stream capture: callers prepare before cudaStreamBeginCapture
        if (dalloc(&h->d_spins_alt, (size_t)h->R * 3 * h->hm.npad) != cudaSuccess) {
            cudaGetLastError();
            h->d_spins_alt = nullptr;
            h->jit_plan.fused = false;
            return false;
        }
    }
    return true;
}

void launch_fused(csmc_handle *h, int upd, const double *in, double *out, const SweepArgs &a) {
    void *args[] = {(void *)&in, (void *)&out, (void *)&a};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(h->jit_plan.fused_tiles, 1, h->R);
    cfg.blockDim = dim3(h->jit_plan.fused_tpb);
    cfg.dynamicSmemBytes = h->jit_plan.fused_smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->jit_pdl ? 1 : 0;
    cudaLaunchKernelExC(&cfg, (const void *)h->jit_fused[upd], args);
    h->launches++;
}

struct SweepOp { int upd; unsigned long long ctr_off; bool device_ctr; };
bool enqueue_persist_seq(csmc_handle *h, const SweepOp *seq, int n);

void enqueue_pass_sweep(csmc_handle *h, const SweepOp &op, cudaStream_t stream = nullptr, int rep0 = 0, int nrep = -1) {
    switch (op.upd) {
    case UPD_OR: enqueue_sweep<UPD_OR>(h, op.ctr_off, op.device_ctr, stream, rep0, nrep); break;
    case UPD_DET: enqueue_sweep<UPD_DET>(h, op.ctr_off, op.device_ctr, stream, rep0, nrep); break;
    case UPD_METRO: enqueue_sweep<UPD_METRO>(h, op.ctr_off, op.device_ctr, stream, rep0, nrep); break;
    default: enqueue_sweep<UPD_CONE>(h, op.ctr_off, op.device_ctr, stream, rep0, nrep); break;
    }
}

// Replica groups: replicas are independent between exchanges, so a sequence of sweeps over R replicas can
// run as G independent chains (R/G replicas each) on G streams.  Each colour pass is a single partial wave
// whose duration is mostly fixed latency (launch ramp, load latency, tail); concurrent chains fill those
// gaps with each other's work.  Inside a stream capture this becomes G parallel branches of the graph.
// makes streams / events for g groups available; returns the number that can be used
int ensure_group_streams(csmc_handle *h, int g) {
    g = std::max(1, std::min(std::min(g, h->R), 8));
    while ((int)h->aux_streams.size() < g - 1) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); break; }
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); cudaStreamDestroy(st); break; }
        h->aux_streams.push_back(st); h->aux_done.push_back(ev);
    }
    g = std::min(g, (int)h->aux_streams.size() + 1);
    if (g > 1 && !h->ev_fork && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); g = 1; }
    return g;
}

// number of replica groups for a sequence of n sweeps.  The count is chosen at csmc_create by the launch
// autotune (1, 2 or 4; CSMC_SWEEP_GROUPS overrides); single sweeps stay on one stream, where the fork /
// join costs more than the overlap returns (measured: -15 % at 4 groups, +-0 at 2).
int sweep_groups(csmc_handle *h, int n) {
    if (h->n_groups == 0) {
        int g = 1;
        if (const char *e = std::getenv("CSMC_SWEEP_GROUPS")) g = std::atoi(e);
        h->n_groups = ensure_group_streams(h, g);
    }
    return n >= 2 ? h->n_groups : 1;
}

// spins that replica blocks / time-skewed strips keep L2-resident at a time: 64 MiB of the 126 MB L2 (measured: blocks
// of 64 MiB are resident, blocks of 96 MiB are not, profiles/r1c_replica_blocks_sweep.jsonl); CSMC_L2_BLOCK_MB overrides
double l2_budget_bytes() {
    double budget_mb = 64.0;
    if (const char *e = std::getenv("CSMC_L2_BLOCK_MB")) budget_mb = std::max(0.001, std::atof(e));
    return budget_mb * 1048576.0;
}

// Time-skewed strips (jit.cpp, skew_schedule): a lattice whose spins exceed L2 runs a sequence of n sweeps strip by
// strip instead of pass by pass, each strip of CTA-tile rows staying L2-resident for all n * colours passes.
// Opt-in (CSMC_FLAG_SKEW / CSMC_SKEW=1); results are bit-identical to the pass-by-pass order.
long skew_budget_rows(const csmc_handle *h) {
    const double row_bytes = 3.0 * sizeof(double) * (double)h->hm.N / std::max(1, h->jit_plan.skew_rows);
    return (long)std::min(1e9, l2_budget_bytes() / row_bytes);
}

// launch plan for a sequence of n sweeps (cached), or nullptr when the strips would vanish before the last pass
const std::vector<SkewLaunch> *skew_plan_for(csmc_handle *h, int n, long budget) {
    if (n < 2) return nullptr;
    const int P = n * h->hm.n_colours;
    auto it = h->skew_cache.find(P);
    if (it == h->skew_cache.end())
        it = h->skew_cache.emplace(P, skew_schedule(h->jit_plan.skew_rows, P, h->jit_plan.skew_reach, (int)budget)).first;
    return it->second.empty() ? nullptr : &it->second;
}

void enqueue_skewed(csmc_handle *h, const SweepOp *seq, const std::vector<SkewLaunch> &plan) {
    const int C = h->hm.n_colours, tpr = h->jit_plan.skew_tiles_per_row;
    for (int rep = 0; rep < h->R; ++rep)
        for (const SkewLaunch &L : plan) {
            const SweepOp &op = seq[L.pass / C];
            const int colour = L.pass % C;
            SweepArgs a = sweep_args(h, op.ctr_off, op.device_ctr);
            a.rep0 = rep;
            a.tile_off = L.row0 * tpr;
            a.tile_end = (L.row0 + L.nrows) * tpr;
            switch (op.upd) {
            case UPD_OR: launch_sweep_pass<UPD_OR>(h, colour, a, h->stream, 1, L.nrows * tpr); break;
            case UPD_DET: launch_sweep_pass<UPD_DET>(h, colour, a, h->stream, 1, L.nrows * tpr); break;
            case UPD_METRO: launch_sweep_pass<UPD_METRO>(h, colour, a, h->stream, 1, L.nrows * tpr); break;
            default: launch_sweep_pass<UPD_CONE>(h, colour, a, h->stream, 1, L.nrows * tpr); break;
            }
        }
}

// A sequence of n sweeps as time-skewed strips, in chunks short enough that a strip boundary moves over at most a
// quarter of the L2 budget (longer chunks leave too little of the first strip).  false: not applicable, nothing enqueued.
bool enqueue_skewed_seq(csmc_handle *h, const SweepOp *seq, int n) {
    if (!h->jit || !h->jit_plan.skew || h->skew_off || n < 2) return false;
    const long budget = skew_budget_rows(h);
    if (budget >= h->jit_plan.skew_rows) return false;            // the whole lattice fits in the budget: nothing to gain
    const int C = h->hm.n_colours, reach = std::max(1, h->jit_plan.skew_reach);
    int chunk = std::min<long>(n, std::max<long>(2, (budget / (4L * reach) + 1) / C));
    while (chunk >= 2 && !skew_plan_for(h, chunk, budget)) --chunk;
    if (chunk < 2) return false;
    for (int i = 0; i < n; i += chunk) {
        const int len = std::min(chunk, n - i);
        if (const std::vector<SkewLaunch> *plan = skew_plan_for(h, len, budget)) enqueue_skewed(h, seq + i, *plan);
        else for (int k = 0; k < len; ++k) enqueue_pass_sweep(h, seq[i + k]);
    }
    return true;
}

// Replica blocks: when the spins of all R replicas do not fit in L2 but a sequence of n sweeps is enqueued at once
// (a cycle graph: the OR block + the Metropolis sweep between two exchanges), the sequence runs block by block --
// all n sweeps for the first R/B replicas, then for the next -- so that a block's spins are read from HBM once
// and stay L2-resident for the 2n..4n colour passes of the sequence instead of streaming every replica through
// L2 once per pass.  Replicas are independent between exchanges, so results do not change.  The count is chosen
// at csmc_create (autotune: blocked vs unblocked timing; CSMC_REPLICA_BLOCKS overrides, CSMC_L2_BLOCK_MB sets
// the per-block budget, default 64 MiB of the 126 MB L2).
int replica_blocks_wanted(const csmc_handle *h) {
    if (const char *e = std::getenv("CSMC_REPLICA_BLOCKS")) return std::max(1, std::min(std::atoi(e), h->R));
    const double bytes = (double)h->R * 3.0 * h->hm.npad * sizeof(double);
    const int nb = (int)std::ceil(bytes / l2_budget_bytes());
    return std::max(1, std::min(nb, h->R));
}

// a sequence of n sweeps over the local replicas [rb, rb + rn), as G concurrent chains when replica groups are on
void enqueue_sweep_seq_range(csmc_handle *h, const SweepOp *seq, int n, int rb, int rn) {
    const int G = std::min(n < 1 ? 1 : sweep_groups(h, n), rn);
    if (G > 1) {
        cudaEventRecord(h->ev_fork, h->stream);
        for (int g = 0; g < G; ++g) {
            const int r0 = rb + (int)((long long)rn * g / G), r1 = rb + (int)((long long)rn * (g + 1) / G);
            cudaStream_t st = g == 0 ? h->stream : h->aux_streams[g - 1];
            if (g > 0) cudaStreamWaitEvent(st, h->ev_fork, 0);
            for (int k = 0; k < n; ++k) enqueue_pass_sweep(h, seq[k], st, r0, r1 - r0);
            if (g > 0) { cudaEventRecord(h->aux_done[g - 1], st); cudaStreamWaitEvent(h->stream, h->aux_done[g - 1], 0); }
        }
        return;
    }
    for (int i = 0; i < n; ++i) enqueue_pass_sweep(h, seq[i], nullptr, rb, rn);
}

void enqueue_sweep_seq(csmc_handle *h, const SweepOp *seq, int n, bool fused) {
    if (fused && n >= 2) {
        int i = 0;
        if (n & 1) enqueue_pass_sweep(h, seq[i++]);
        for (; i < n; i += 2) {
            launch_fused(h, seq[i].upd, h->d_spins, h->d_spins_alt, sweep_args(h, seq[i].ctr_off, seq[i].device_ctr));
            launch_fused(h, seq[i + 1].upd, h->d_spins_alt, h->d_spins, sweep_args(h, seq[i + 1].ctr_off, seq[i + 1].device_ctr));
        }
        return;
    }
    if (enqueue_persist_seq(h, seq, n)) return;
    if (enqueue_skewed_seq(h, seq, n)) return;
    const int B = n >= 2 ? std::max(1, std::min(h->n_blocks, h->R)) : 1;
    for (int b = 0; b < B; ++b) {
        const int r0 = (int)((long long)h->R * b / B), r1 = (int)((long long)h->R * (b + 1) / B);
        enqueue_sweep_seq_range(h, seq, n, r0, r1 - r0);
    }
}

void enqueue_metropolis(csmc_handle *h, bool cone) {
    const SweepOp op{cone ? UPD_CONE : UPD_METRO, h->metro_ctr, false};
    enqueue_sweep_seq(h, &op, 1, false);
    h->metro_ctr++;
}

// peer-memory gather usable for the current parallel-tempering state (all ranks decide alike: n_slots is global)
bool peer_gather_active(const csmc_handle *h) {
    return h->peer.mode != 0 && h->comm && h->n_slots > 0 && (long long)h->n_slots * 8 <= h->peer.ports.cap8;
}

// energy + magnetisation of every local replica into meas[(R) x 8]
void enqueue_measure(csmc_handle *h, double *meas, bool write_energy) {
    for (int c = 0; c < h->hm.n_colours; ++c) {
        const int nseg = h->hm.colour_seg_begin[c + 1] - h->hm.colour_seg_begin[c];
        dim3 grid(h->pass_blocks[c], nseg, h->R), block(TPB);
        if (h->jit) {
            void *args[] = {(void *)&h->d_spins, (void *)&h->d_partials, (void *)&h->n_partials, (void *)&h->partial_base[c]};
            cudaLaunchKernel((const void *)h->jit_energy[c], grid, block, args, 0, h->stream);
        } else if (h->large) {
            if (h->hm.structured) k_energy<PassLarge, true><<<grid, block, 0, h->stream>>>(h->pl[c], h->d_partials, h->n_partials, h->partial_base[c]);
            else k_energy<PassLarge, false><<<grid, block, 0, h->stream>>>(h->pl[c], h->d_partials, h->n_partials, h->partial_base[c]);
        } else {
            if (h->hm.structured) k_energy<PassSmall, true><<<grid, block, 0, h->stream>>>(h->ps[c], h->d_partials, h->n_partials, h->partial_base[c]);
            else k_energy<PassSmall, false><<<grid, block, 0, h->stream>>>(h->ps[c], h->d_partials, h->n_partials, h->partial_base[c]);
        }
        h->launches++;
    }
    if (peer_gather_active(h) && h->peer.mode == 2 && h->d_meas_all && meas == h->d_meas_all + (size_t)h->replica_base * 8) {
        // the records go straight from the reduction into every rank's mailbox (enqueue_gather_meas then only waits)
        unsigned int *arrive = (unsigned int *)(h->peer.base + 128);
        k_reduce_partials_push<<<h->R, 256, 0, h->stream>>>(h->d_partials, h->n_partials, h->d_acc, h->d_sigma, meas, write_energy ? 1 : 0,
                                                            h->peer.ports, (long long)h->replica_base * 8, ++h->peer.seq, arrive);
        h->peer.pushed = true;
    } else {
        k_reduce_partials<<<h->R, 256, 0, h->stream>>>(h->d_partials, h->n_partials, h->d_acc, h->d_sigma, meas, write_energy ? 1 : 0);
    }
    h->launches++;
}

// what == 0: local fields, 1: site energies, into out (reference order).  only_site >= 0 (0-based reference
// index): evaluate just the block holding that site (single-site queries stay O(1) in the lattice size).
void enqueue_eval(csmc_handle *h, int rep, int what, double *out, long long only_site = -1) {
    int c_lo = 0, c_hi = h->hm.n_colours, seg_off = 0, block_off = 0;
    if (only_site >= 0) {
        const int pos = h->hm.pos_of_ref[only_site];
        int s = 0;
        while (s + 1 < (int)h->hm.segs.size() && !(pos >= h->hm.segs[s].start && pos < h->hm.segs[s].start + h->hm.segs[s].count)) ++s;
        c_lo = h->hm.segs[s].colour; c_hi = c_lo + 1;
        seg_off = s - h->hm.colour_seg_begin[c_lo];
        block_off = (pos - h->hm.segs[s].start) / TPB;
    }
    for (int c = c_lo; c < c_hi; ++c) {
        const int nseg = h->hm.colour_seg_begin[c + 1] - h->hm.colour_seg_begin[c];
        dim3 grid(h->pass_blocks[c], nseg, 1), block(TPB);
        if (only_site >= 0) grid = dim3(1, 1, 1);
        if (h->large) {
            if (h->hm.structured) k_eval<PassLarge, true><<<grid, block, 0, h->stream>>>(h->pl[c], rep, what, out, seg_off, block_off);
            else k_eval<PassLarge, false><<<grid, block, 0, h->stream>>>(h->pl[c], rep, what, out, seg_off, block_off);
        } else {
            if (h->hm.structured) k_eval<PassSmall, true><<<grid, block, 0, h->stream>>>(h->ps[c], rep, what, out, seg_off, block_off);
            else k_eval<PassSmall, false><<<grid, block, 0, h->stream>>>(h->ps[c], rep, what, out, seg_off, block_off);
        }
        h->launches++;
    }
}


// Builds (once) the kernels specialised for this handle's model; "" on success.  On failure the
// ahead-of-time kernels stay in use and csmc_kernel_mode reports the reason.
std::map<size_t, std::vector<char>> g_cubin_cache;   // generated source hash -> cubin (same model, many handles)
std::mutex g_cubin_mu;

struct JitModule {
    cudaLibrary_t lib = nullptr;
    std::vector<cudaKernel_t> sweep[4], energy;
    cudaKernel_t resident = nullptr;
    cudaKernel_t fused[4] = {nullptr, nullptr, nullptr, nullptr};
    JitPlan plan;
    bool pdl = false;
};

// generate + compile (or fetch from the per-process cache) + load the kernels specialised for hm
std::string load_jit_module(const HostModel &hm, bool pdl, JitModule &m, bool want_fused = false, bool want_skew = false) {
    m.plan.want_fused = want_fused;
    m.plan.want_skew = want_skew;
    std::string err, log;
    std::vector<char> cubin;
    try {
        const std::string src = jit_generate_source(hm, pdl, &m.plan);
        const size_t key = std::hash<std::string>{}(src);
        {
            std::lock_guard<std::mutex> lk(g_cubin_mu);
            auto it = g_cubin_cache.find(key);
            if (it != g_cubin_cache.end()) cubin = it->second;
        }
        if (cubin.empty()) {
            err = jit_compile(src, cubin, log);
            if (err.empty()) {
                std::lock_guard<std::mutex> lk(g_cubin_mu);
                if (g_cubin_cache.size() > 64) g_cubin_cache.clear();
                g_cubin_cache[key] = cubin;
            }
        }
    } catch (const std::exception &ex) {
        err = std::string("code generation failed: ") + ex.what();
    }
    if (!err.empty()) return err;
    cudaError_t ce = cudaLibraryLoadData(&m.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce != cudaSuccess) return std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce);
    for (int u = 0; u < 4; ++u) {
        m.sweep[u].resize(hm.n_colours);
        for (int c = 0; c < hm.n_colours; ++c) {
            const std::string nm = "csmc_sweep_c" + std::to_string(c) + "_u" + std::to_string(u);
            if (cudaLibraryGetKernel(&m.sweep[u][c], m.lib, nm.c_str()) != cudaSuccess) return "kernel not found: " + nm;
        }
    }
    m.energy.resize(hm.n_colours);
    for (int c = 0; c < hm.n_colours; ++c) {
        const std::string nm = "csmc_energy_c" + std::to_string(c);
        if (cudaLibraryGetKernel(&m.energy[c], m.lib, nm.c_str()) != cudaSuccess) return "kernel not found: " + nm;
    }
    if (m.plan.resident) {
        if (cudaLibraryGetKernel(&m.resident, m.lib, "csmc_resident") != cudaSuccess) return "kernel not found: csmc_resident";
        const int smem = 3 * hm.npad * (int)sizeof(double);
        if (cudaFuncSetAttribute((const void *)m.resident, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            cudaGetLastError();
            m.resident = nullptr;     // pass kernels still work
        }
    }
    if (m.plan.fused) {
        for (int u = 0; u < 4 && m.plan.fused; ++u) {
            const std::string nm = "csmc_fused_u" + std::to_string(u);
            if (cudaLibraryGetKernel(&m.fused[u], m.lib, nm.c_str()) != cudaSuccess ||
                cudaFuncSetAttribute((const void *)m.fused[u], cudaFuncAttributeMaxDynamicSharedMemorySize, m.plan.fused_smem) != cudaSuccess) {
                cudaGetLastError();
                m.plan.fused = false;   // pass kernels still work
            }
        }
    }
    m.pdl = pdl;
    return "";
}

// kernels with a tile offset (time-skewed strips): on request (CSMC_FLAG_SKEW, CSMC_SKEW=1), never with CSMC_SKEW=0,
// otherwise exactly when one replica's spins exceed the L2 budget -- smaller lattices keep the offset-free kernels
bool want_skew(const csmc_handle *h) {
    const char *e = std::getenv("CSMC_SKEW");
    if (e && e[0] == '0') return false;
    if ((h->flags & CSMC_FLAG_SKEW) != 0 || (e && e[0] == '1')) return true;
    return 3.0 * sizeof(double) * (double)h->hm.npad > l2_budget_bytes();
}

void drop_graphs(csmc_handle *h) {
    for (auto &kv : h->cycle_graphs) cudaGraphExecDestroy(kv.second.exec);
    h->cycle_graphs.clear();
    for (auto &kv : h->or_graphs) cudaGraphExecDestroy(kv.second);
    h->or_graphs.clear();
}

void install_jit_module(csmc_handle *h, const JitModule &m) {
    drop_graphs(h);   // graphs captured with other kernels must not be replayed any more
    h->jit_lib = m.lib;
    for (int u = 0; u < 4; ++u) h->jit_sweep[u] = m.sweep[u];
    h->jit_energy = m.energy;
    h->jit_resident = m.resident;
    for (int u = 0; u < 4; ++u) h->jit_fused[u] = m.fused[u];
    h->jit_plan = m.plan;
    h->jit_pdl = m.pdl;
    h->jit = true;
}

// ---- tile-resident persistent kernel (jit.cpp emit_persist) ---------------------------------------------------
// Built on request only: CSMC_FLAG_PERSIST / CSMC_PERSIST=1 use it whenever it applies, CSMC_PERSIST=probe builds it and
// lets the create-time autotune time it against the pass kernels.  By default it is not even compiled: on every workload
// measured it lost to the pass kernels, and its NVRTC build costs 1.6 s (C2) to 6 s (C5) per model.
bool want_persist(const csmc_handle *h) {
    const char *e = std::getenv("CSMC_PERSIST");
    if ((h->flags & CSMC_FLAG_NO_PERSIST) || (e && e[0] == '0')) return false;
    if ((h->flags & CSMC_FLAG_PERSIST) || (e && e[0] == '1')) return true;
    // an explicit request for the time-skewed strips or the fused full-sweep kernels is not overridden
    const char *sk = std::getenv("CSMC_SKEW");
    if ((h->flags & (CSMC_FLAG_SKEW | CSMC_FLAG_FUSED)) || (sk && sk[0] == '1')) return false;
    return e && (e[0] == 'p' || e[0] == 'a');      // "probe" / "auto"
}

void persist_release(csmc_handle *h) {
    if (h->d_persist_prof) {   // debugging aid: dump the phase counters of the last launch
        std::vector<unsigned long long> pf(8 * 1024);
        if (cudaMemcpy(pf.data(), h->d_persist_prof, sizeof(unsigned long long) * pf.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const int nt = std::min(1024, h->persist_plan.persist_tiles * std::max(1, h->persist_plan.persist_nrep));
            double sum[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
            for (int t = 0; t < nt; ++t) for (int k = 0; k < 4; ++k) { sum[k] += (double)pf[t * 8 + k]; mx[k] = std::max(mx[k], (double)pf[t * 8 + k]); }
            std::fprintf(stderr, "csmc_persist phases, cycles per CTA over the last launch (mean / max over %d CTAs): wait %.0f / %.0f, reload %.0f / %.0f, update %.0f / %.0f, publish %.0f / %.0f\n",
                         nt, sum[0] / nt, mx[0], sum[1] / nt, mx[1], sum[2] / nt, mx[2], sum[3] / nt, mx[3]);
        }
        cudaFree(h->d_persist_prof); h->d_persist_prof = nullptr;
    }
    if (h->persist_lib) { cudaLibraryUnload(h->persist_lib); h->persist_lib = nullptr; }
    h->jit_persist = nullptr;
    h->persist_plan = JitPlan{};
    cudaFree(h->d_persist_flags); h->d_persist_flags = nullptr;
    if (h->persist_err) { cudaFreeHost(h->persist_err); h->persist_err = nullptr; }
}

// generate + compile + load csmc_persist for this handle's model and replica count; "" on success (or not applicable)
std::string load_persist_module(csmc_handle *h) {
    if (h->jit_persist || !want_persist(h)) return "";
    int dev = 0, sms = 0, smem = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) {
        cudaGetLastError();
        return "cooperative launch not available";
    }
    JitPlan plan;
    plan.persist_only = true;
    plan.persist_replicas = h->R;
    plan.persist_sms = sms;
    plan.persist_smem_max = smem;
    std::string err, log;
    std::vector<char> cubin;
    try {
        const std::string src = jit_generate_source(h->hm, false, &plan);
        if (!plan.persist) return "";          // no tiling fits (lattice too large for the SMs' shared memory, open boundaries, ...)
        const size_t key = std::hash<std::string>{}(src);
        {
            std::lock_guard<std::mutex> lk(g_cubin_mu);
            auto it = g_cubin_cache.find(key);
            if (it != g_cubin_cache.end()) cubin = it->second;
        }
        if (cubin.empty()) {
            err = jit_compile(src, cubin, log);
            if (err.empty()) {
                std::lock_guard<std::mutex> lk(g_cubin_mu);
                if (g_cubin_cache.size() > 64) g_cubin_cache.clear();
                g_cubin_cache[key] = cubin;
            }
        }
    } catch (const std::exception &ex) {
        err = std::string("code generation failed: ") + ex.what();
    }
    if (!err.empty()) return err;
    int per_sm = 0;
    if (cudaLibraryLoadData(&h->persist_lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&h->jit_persist, h->persist_lib, "csmc_persist") != cudaSuccess ||
        cudaFuncSetAttribute((const void *)h->jit_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.persist_smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)h->jit_persist, plan.persist_tpb, plan.persist_smem) != cudaSuccess ||
        per_sm < 1 ||
        cudaMalloc((void **)&h->d_persist_flags, sizeof(unsigned long long) * (size_t)h->R * plan.persist_tiles * PERSIST_FLAG_STRIDE) != cudaSuccess ||
        cudaMemsetAsync(h->d_persist_flags, 0, sizeof(unsigned long long) * (size_t)h->R * plan.persist_tiles * PERSIST_FLAG_STRIDE, h->stream) != cudaSuccess ||
        cudaHostAlloc((void **)&h->persist_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
        const std::string e = std::string("persistent kernel: ") + cudaGetErrorString(cudaGetLastError());
        persist_release(h);
        return e;
    }
    *h->persist_err = 0;
    {   // in use only when asked for or when the create-time probe finds it faster than the pass kernels (autotune_pdl)
        const char *e = std::getenv("CSMC_PERSIST");
        h->persist_off = !((h->flags & CSMC_FLAG_PERSIST) != 0 || (e && e[0] == '1'));
    }
    if (std::getenv("CSMC_PERSIST_PROF") && cudaMalloc((void **)&h->d_persist_prof, sizeof(unsigned long long) * 8 * 1024) != cudaSuccess) { cudaGetLastError(); h->d_persist_prof = nullptr; }
    h->persist_plan = plan;
    if (const char *e = std::getenv("CSMC_PERSIST_MIN_SWEEPS")) h->persist_min_sweeps = std::max(1, std::atoi(e));
    return "";
}

// A sequence of n sweeps on the tile-resident kernel: one cooperative launch per batch of replicas that fills the SMs
// and per <= PERSIST_MAX_OPS sweeps.  false: not applicable, nothing enqueued.
bool enqueue_persist_seq(csmc_handle *h, const SweepOp *seq, int n) {
    if (!h->jit_persist || h->persist_off || n < h->persist_min_sweeps) return false;
    // the Metropolis counter source (device-resident for graph replays, by value otherwise) must be the same for every
    // Metropolis sweep of the sequence; overrelaxation / deterministic sweeps do not read it
    int dc = -1;
    for (int i = 0; i < n; ++i)
        if (seq[i].upd >= UPD_METRO) {
            if (dc >= 0 && dc != (seq[i].device_ctr ? 1 : 0)) return false;
            dc = seq[i].device_ctr ? 1 : 0;
        }
    const bool device_ctr = dc == 1;
    const JitPlan &pl = h->persist_plan;
    int *d_err = nullptr;
    if (cudaHostGetDevicePointer((void **)&d_err, h->persist_err, 0) != cudaSuccess) { cudaGetLastError(); return false; }
    for (int i = 0; i < n; i += PERSIST_MAX_OPS) {
        const int len = std::min(PERSIST_MAX_OPS, n - i);
        unsigned long long ctr_min = ~0ULL;
        for (int k = 0; k < len; ++k) if (seq[i + k].upd >= UPD_METRO) ctr_min = std::min(ctr_min, seq[i + k].ctr_off);
        if (ctr_min == ~0ULL) ctr_min = 0;
        PersistArgs pa{};
        pa.flags = h->d_persist_flags;
        pa.err = d_err;
        pa.timeout_cycles = 4000000000ULL;       // ~2 s at 2 GHz
        pa.prof = h->d_persist_prof;
        pa.n_ops = len;
        for (int k = 0; k < len; ++k) {
            const unsigned long long rel = seq[i + k].upd >= UPD_METRO ? seq[i + k].ctr_off - ctr_min : 0ULL;
            if (rel > 65535ULL) return false;
            pa.upd[k] = (unsigned char)seq[i + k].upd;
            pa.ctr_rel[k] = (unsigned short)rel;
        }
        for (int r0 = 0; r0 < h->R; r0 += pl.persist_nrep) {
            SweepArgs a = sweep_args(h, ctr_min, device_ctr);
            a.rep0 = r0;
            void *args[] = {(void *)&h->d_spins, (void *)&a, (void *)&pa};
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(pl.persist_tiles, std::min(pl.persist_nrep, h->R - r0), 1);
            cfg.blockDim = dim3(pl.persist_tpb);
            cfg.dynamicSmemBytes = pl.persist_smem;
            cfg.stream = h->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (cudaLaunchKernelExC(&cfg, (const void *)h->jit_persist, args) != cudaSuccess) {
                // nothing of this launch ran; earlier launches of the sequence are complete sweeps on other replicas /
                // earlier sweeps, so the caller cannot simply redo the sequence: report through the error word
                cudaGetLastError();
                *h->persist_err = 2;
                return true;
            }
            h->launches++;
        }
    }
    return true;
}

// Builds (once) the kernels specialised for this handle's model; "" on success.  On failure the
// ahead-of-time kernels stay in use and csmc_kernel_mode reports the reason.
std::string build_jit(csmc_handle *h) {
    if (h->jit) return "";
    if (h->jit_tried) return h->jit_note;
    h->jit_tried = true;
    const HostModel &hm = h->hm;
    if ((h->flags & CSMC_FLAG_NO_JIT) || !hm.structured || hm.self_loop) {
        h->jit_note = "not applicable (no periodic colouring pattern, self-interaction, or CSMC_FLAG_NO_JIT)";
        return h->jit_note;
    }
    JitModule m;
    const std::string err = load_jit_module(hm, (h->flags & CSMC_FLAG_PDL) != 0, m, (h->flags & CSMC_FLAG_FUSED) != 0, want_skew(h));
    if (err.empty()) {
        install_jit_module(h, m);
        if ((int64_t)hm.N * h->R >= 32768) {
            const std::string perr = load_persist_module(h);
            if (!perr.empty()) h->jit_note = "persistent kernel unavailable: " + perr;
        }
    } else {
        if (m.lib) cudaLibraryUnload(m.lib);
        cudaGetLastError();
        h->jit_note = err;
    }
    return err;
}

// small lattices: one CTA per replica, lattice resident in shared memory (see jit.cpp)
bool use_resident(csmc_handle *h, long long sweeps_requested) {
    if ((h->flags & (CSMC_FLAG_NO_RESIDENT | CSMC_FLAG_NO_JIT)) || h->hm.N > 4096) return false;
    if (!h->jit && !h->jit_tried && (sweeps_requested >= 32 || (h->flags & CSMC_FLAG_JIT))) build_jit(h);
    return h->jit && h->jit_resident != nullptr;
}

// n_cycles x (orc OR + mc Metropolis sweeps) then det deterministic sweeps, optional measurement record
void enqueue_resident(csmc_handle *h, int n_cycles, int orc, int mc, int cone, int det, double *meas, int write_energy, int adapt = 0) {
    SweepArgs a = sweep_args(h, h->metro_ctr, false);
    void *args[] = {(void *)&h->d_spins, (void *)&a, (void *)&n_cycles, (void *)&orc, (void *)&mc, (void *)&cone, (void *)&adapt, (void *)&det, (void *)&meas, (void *)&write_energy};
    const size_t smem = (size_t)3 * h->hm.npad * sizeof(double);
    cudaLaunchKernel((const void *)h->jit_resident, dim3(h->R), dim3(256), args, smem, h->stream);
    h->metro_ctr += (unsigned long long)n_cycles * mc;
    h->launches++;
}

int finish(csmc_handle *h) {
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    if (h->persist_err && *h->persist_err) {
        const int code = *h->persist_err;
        *h->persist_err = 0;
        h->persist_off = true;
        return fail(h, CSMC_ERR_CUDA, code == 2 ? "tile-resident kernel: cooperative launch failed (spins are in an intermediate state)"
                                                 : "tile-resident kernel: a neighbour tile did not arrive in time (CTAs not co-resident?)");
    }
    if (h->peer.err && *h->peer.err) {
        const int who = *h->peer.err - 1;
        *h->peer.err = 0;
        return fail(h, CSMC_ERR_NCCL, "peer-memory gather: rank " + std::to_string(who) + " did not deliver its measurement records in time");
    }
    return CSMC_OK;
}

int check_metropolis(csmc_handle *h) {
    if (h->hm.self_loop)
        return fail(h, CSMC_ERR_UNSUPPORTED,
                    "Metropolis: a site interacts with itself on this lattice (offset is a multiple of the "
                    "lattice size); the single-field dE is not valid there");
    return CSMC_OK;
}

int upload_T(csmc_handle *h, const double *T) {
    for (int r = 0; r < h->R; ++r)
        if (!(T[r] > 0.0)) return fail(h, CSMC_ERR_INVALID, "temperatures must be > 0");
    std::vector<double> beta(h->R);
    for (int r = 0; r < h->R; ++r) beta[r] = 1.0 / T[r];
    CK(cudaMemcpyAsync(h->d_beta, beta.data(), sizeof(double) * h->R, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));  // host buffers are not retained
    return CSMC_OK;
}

PtState pt_state(csmc_handle *h) {
    PtState st{};
    st.n_slots = h->n_slots; st.n_local = h->R; st.replica_base = h->replica_base;
    st.T_slot = h->d_T_slot; st.slot_of_rep = h->d_slot_of_rep; st.rep_of_slot = h->d_rep_of_slot;
    st.meas_all = h->d_meas_all; st.E_last = h->d_E_last; st.acc_prev = h->d_acc_prev_pt;
    st.acc_slot = h->d_acc_slot; st.exch_slot = h->d_exch_slot; st.beta_local = h->d_beta; st.sigma_local = h->d_sigma; st.prev_rep_of_slot = h->d_prev_rep_of_slot;
    st.accepted_pairs = h->d_accepted_pairs;
    return st;
}

void free_pt(csmc_handle *h) {
    cudaFree(h->d_T_slot); cudaFree(h->d_meas_all); cudaFree(h->d_E_last); cudaFree(h->d_acc_prev_pt);
    cudaFree(h->d_acc_slot); cudaFree(h->d_exch_slot); cudaFree(h->d_series_E); cudaFree(h->d_series_M);
    cudaFree(h->d_slot_of_rep); cudaFree(h->d_rep_of_slot); cudaFree(h->d_accepted_pairs); cudaFree(h->d_prev_rep_of_slot);
    h->d_T_slot = h->d_meas_all = h->d_E_last = h->d_acc_prev_pt = h->d_acc_slot = h->d_exch_slot = nullptr;
    h->d_series_E = h->d_series_M = nullptr;
    h->d_slot_of_rep = h->d_rep_of_slot = h->d_accepted_pairs = h->d_prev_rep_of_slot = nullptr;
    h->series_cap = 0; h->n_probes = 0; h->n_slots = 0; h->exchange_calls = 0;
    cudaFree(h->d_ssf_sum); h->d_ssf_sum = nullptr; h->ssf_probes = 0;
}

void peer_gather_release(csmc_handle *h) {
    auto &pg = h->peer;
    for (int g = 0; g < PEER_MAX_RANKS; ++g) if (pg.peer[g]) { cudaIpcCloseMemHandle(pg.peer[g]); pg.peer[g] = nullptr; }
    if (pg.base) { cudaFree(pg.base); pg.base = nullptr; }
    if (pg.err) { cudaFreeHost(pg.err); pg.err = nullptr; }
    pg.mode = 0; pg.seq = 0; pg.pushed = false; pg.ports = PeerPorts{};
}

// CSMC_PEER_GATHER=1|2 at csmc_comm_init (set it for every rank of the job or for none): every rank allocates
// its mailbox, the CUDA IPC handles travel through the communicator, every rank maps every other rank's mailbox.
// The outcome of the mapping is gathered too: unless every rank mapped every mailbox the NCCL collectives stay
// in use (csmc_comm_mode tells).  Needs all ranks on one node with peer access (NVLink / NVSwitch on a B200 box).
int peer_gather_setup(csmc_handle *h, int mode) {
    peer_gather_release(h);
    if (mode == 0 || h->n_ranks > PEER_MAX_RANKS) return CSMC_OK;   // default: nothing changes, no extra collective
    auto &pg = h->peer;
    constexpr size_t BYTES = 2u << 20, HEADER = 256;
    struct Card { cudaIpcMemHandle_t handle; int32_t mode, ok; };   // what a rank tells the others
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    Card mine{};
    mine.mode = mode; mine.ok = 0;
    if (mode) {
        bool ok = cudaMalloc((void **)&pg.base, BYTES) == cudaSuccess && cudaMemsetAsync(pg.base, 0, BYTES, h->stream) == cudaSuccess &&
                  cudaIpcGetMemHandle(&mine.handle, pg.base) == cudaSuccess &&
                  cudaHostAlloc((void **)&pg.err, sizeof(int), cudaHostAllocMapped) == cudaSuccess;
        if (ok) *pg.err = 0;
        else cudaGetLastError();
        mine.ok = ok ? 1 : 0;
    }
    // exchange the cards (every rank takes part, whatever its own setting)
    const int n = h->n_ranks;
    std::vector<Card> cards(n);
    unsigned char *d_cards = nullptr;
    CK(dalloc(&d_cards, sizeof(Card) * n));
    auto exchange = [&]() -> int {
        cudaError_t ce = cudaMemcpyAsync(d_cards + sizeof(Card) * h->rank, &mine, sizeof(Card), cudaMemcpyHostToDevice, h->stream);
        ncclResult_t nr = 0;
        if (ce == cudaSuccess) nr = g_nccl.AllGather(d_cards + sizeof(Card) * h->rank, d_cards, sizeof(Card), ncclUint8, h->comm, h->stream);
        if (ce == cudaSuccess && nr == 0) ce = cudaMemcpyAsync(cards.data(), d_cards, sizeof(Card) * n, cudaMemcpyDeviceToHost, h->stream);
        if (ce == cudaSuccess && nr == 0) ce = cudaStreamSynchronize(h->stream);
        if (ce != cudaSuccess) return fail(h, CSMC_ERR_CUDA, std::string("peer gather setup: ") + cudaGetErrorString(ce));
        if (nr != 0) return fail(h, CSMC_ERR_NCCL, "peer gather setup: ncclAllGather failed");
        return CSMC_OK;
    };
    int rc = exchange();
    bool all = rc == CSMC_OK;
    for (int g = 0; all && g < n; ++g) all = cards[g].mode == mode && cards[g].ok == 1;
    if (all) {
        for (int g = 0; g < n && mine.ok; ++g) {
            if (g == h->rank) continue;
            if (cudaIpcOpenMemHandle(&pg.peer[g], cards[g].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                pg.peer[g] = nullptr;
                mine.ok = 0;
            }
        }
    }
    // second round: did every rank map every mailbox?  (also orders every rank's memset before any push)
    if (rc == CSMC_OK) rc = exchange();
    cudaFree(d_cards);
    if (rc != CSMC_OK) { peer_gather_release(h); return rc; }
    for (int g = 0; all && g < n; ++g) all = cards[g].ok == 1;
    if (!all || mode == 0) { peer_gather_release(h); return CSMC_OK; }
    pg.mode = mode;
    pg.ports.n_ranks = n; pg.ports.rank = h->rank;
    pg.ports.cap8 = (long long)((BYTES - HEADER) / 2 / sizeof(double));
    for (int g = 0; g < n; ++g) {
        unsigned char *b = g == h->rank ? pg.base : (unsigned char *)pg.peer[g];
        pg.ports.flag[g] = (unsigned long long *)b;
        pg.ports.mail[g] = (double *)(b + HEADER);
    }
    if (const char *e = std::getenv("CSMC_PEER_TIMEOUT_MS")) pg.timeout_ns = (unsigned long long)std::max(1L, std::atol(e)) * 1000000ULL;
    return CSMC_OK;
}

// the ranks' replica blocks must tile [0, n_slots) in rank order (csmc_comm_init gathered them)
int check_partition(csmc_handle *h) {
    if (!h->comm) return CSMC_OK;
    long long next = 0;
    for (int g = 0; g < h->n_ranks; ++g) {
        if (h->rank_base[g] != next)
            return fail(h, CSMC_ERR_INVALID, "multi-GPU parallel tempering: replica_base of rank " + std::to_string(g) +
                                                 " must be the sum of the replica counts of the ranks before it");
        next += h->rank_count[g];
    }
    if (next != h->n_slots)
        return fail(h, CSMC_ERR_INVALID, "multi-GPU parallel tempering: the ranks hold " + std::to_string(next) +
                                             " replicas but csmc_pt_init was given " + std::to_string(h->n_slots) + " temperatures");
    return CSMC_OK;
}

// every rank's 8-double records into d_meas_all on every rank, in place (the per-replica energies of the
// exchange test, src/monte_carlo.jl:321-343): one ncclAllGather when the ranks hold equally many replicas,
// otherwise one grouped ncclBroadcast per rank
int enqueue_gather_meas(csmc_handle *h) {
    if (!h->comm) return CSMC_OK;
    if (peer_gather_active(h)) {
        auto &pg = h->peer;
        if (!pg.pushed) {
            k_peer_push<<<h->n_ranks, 256, 0, h->stream>>>(pg.ports, h->d_meas_all + (size_t)h->replica_base * 8, h->R * 8,
                                                           (long long)h->replica_base * 8, ++pg.seq);
            h->launches++;
        }
        pg.pushed = false;
        const double *mail = pg.ports.mail[h->rank] + (pg.seq & 1ULL) * pg.ports.cap8;
        k_peer_wait<<<1, 256, 0, h->stream>>>(pg.ports.flag[h->rank], h->n_ranks, pg.seq, mail, h->d_meas_all, h->n_slots * 8, pg.err, pg.timeout_ns);
        h->launches++;
        return CSMC_OK;
    }
    if (h->even_partition) {
        CKN(g_nccl.AllGather(h->d_meas_all + (size_t)h->replica_base * 8, h->d_meas_all, (size_t)h->R * 8, ncclFloat64, h->comm, h->stream));
    } else {
        CKN(g_nccl.GroupStart());
        for (int g = 0; g < h->n_ranks; ++g) {
            double *blk = h->d_meas_all + (size_t)h->rank_base[g] * 8;
            ncclResult_t r = g_nccl.Broadcast(blk, blk, (size_t)h->rank_count[g] * 8, ncclFloat64, g, h->comm, h->stream);
            if (r != 0) { g_nccl.GroupEnd(); CKN(r); }
        }
        CKN(g_nccl.GroupEnd());
    }
    return CSMC_OK;
}

// measurement + gather across ranks into d_meas_all (in place)
int enqueue_measure_all(csmc_handle *h, bool write_energy) {
    int rc = check_partition(h); if (rc) return rc;
    enqueue_measure(h, h->d_meas_all + (size_t)h->replica_base * 8, write_energy);
    return enqueue_gather_meas(h);
}

}  // namespace

// =========================================================================================================
extern "C" {

int32_t csmc_version(void) { return CSMC_VERSION; }

const char *csmc_last_error(const csmc_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int autotune_pdl(csmc_handle *h);
static int enqueue_or_block(csmc_handle *h, int n);

int32_t csmc_create(const csmc_model *model, const csmc_opts *opts, csmc_handle **out) {
    csmc_handle *h = nullptr;
    if (!model || !opts || !out) return fail(nullptr, CSMC_ERR_INVALID, "csmc_create: NULL argument");
    *out = nullptr;
    if (opts->n_replicas < 1 || opts->n_replicas > 65535) return fail(nullptr, CSMC_ERR_INVALID, "n_replicas must be 1..65535");
    h = new (std::nothrow) csmc_handle();
    if (!h) return fail(nullptr, CSMC_ERR_NOMEM, "out of host memory");
    std::string e;
    try {
        e = build_host_model(model, opts->flags, h->hm);
    } catch (const std::exception &ex) {
        e = std::string("model build failed: ") + ex.what();
    }
    if (!e.empty()) { delete h; return fail(nullptr, CSMC_ERR_INVALID, e); }

    // deep copy of the model
    const HostModel &hm = h->hm;
    const int D = hm.D;
    h->model = *model;
    h->m_field.assign(model->field, model->field + 3 * hm.n_basis);
    h->m_onsite.assign(model->onsite, model->onsite + 9 * hm.n_basis);
    if (hm.N2) { h->m_bilB.assign(model->bil_basis, model->bil_basis + 2 * hm.N2); h->m_bilO.assign(model->bil_offset, model->bil_offset + D * hm.N2); h->m_bilJ.assign(model->bil_matrix, model->bil_matrix + 9 * hm.N2); }
    if (hm.N3) { h->m_cubB.assign(model->cub_basis, model->cub_basis + 3 * hm.N3); h->m_cubO.assign(model->cub_offset, model->cub_offset + 2 * D * hm.N3); h->m_cubT.assign(model->cub_tensor, model->cub_tensor + 27 * hm.N3); }
    if (hm.N4) { h->m_quarB.assign(model->quar_basis, model->quar_basis + 4 * hm.N4); h->m_quarO.assign(model->quar_offset, model->quar_offset + 3 * D * hm.N4); h->m_quarT.assign(model->quar_tensor, model->quar_tensor + 81 * hm.N4); }
    h->model.field = h->m_field.data(); h->model.onsite = h->m_onsite.data();
    h->model.bil_basis = h->m_bilB.data(); h->model.bil_offset = h->m_bilO.data(); h->model.bil_matrix = h->m_bilJ.data();
    h->model.cub_basis = h->m_cubB.data(); h->model.cub_offset = h->m_cubO.data(); h->model.cub_tensor = h->m_cubT.data();
    h->model.quar_basis = h->m_quarB.data(); h->model.quar_offset = h->m_quarO.data(); h->model.quar_tensor = h->m_quarT.data();

    h->device = opts->device; h->R = opts->n_replicas; h->replica_base = opts->replica_base;
    h->seed = opts->seed; h->flags = opts->flags;
    h->large = hm.need_large;

    auto bail = [&](int code, const std::string &msg) { csmc_destroy(h); return fail(nullptr, code, msg); };
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bail(CSMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
    CKC(cudaSetDevice(h->device));
    if (opts->stream) { h->stream = (cudaStream_t)opts->stream; h->own_stream = false; }
    else { CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }

    const size_t npad = hm.npad;
    CKC(dalloc(&h->d_spins, (size_t)h->R * 3 * npad));
    CKC(cudaMemsetAsync(h->d_spins, 0, sizeof(double) * h->R * 3 * npad, h->stream));
    CKC(dalloc(&h->d_stage, (size_t)3 * hm.N));
    CKC(dalloc(&h->d_out, (size_t)3 * hm.N));
    if (!hm.nbr.empty()) {   // explicit-table kernels only
        CKC(dalloc(&h->d_nbr, hm.nbr.size()));
        CKC(cudaMemcpyAsync(h->d_nbr, hm.nbr.data(), sizeof(int32_t) * hm.nbr.size(), cudaMemcpyHostToDevice, h->stream));
    }
    CKC(dalloc(&h->d_ref_of_pos, npad));
    CKC(cudaMemcpyAsync(h->d_ref_of_pos, hm.ref_of_pos.data(), sizeof(int32_t) * npad, cudaMemcpyHostToDevice, h->stream));
    CKC(dalloc(&h->d_beta, h->R)); CKC(dalloc(&h->d_sigma, h->R));
    CKC(dalloc(&h->d_acc, (size_t)h->R * ACC_STRIPE)); CKC(dalloc(&h->d_acc_prev, h->R)); CKC(dalloc(&h->d_ctr, 1));
    CKC(cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * h->R * ACC_STRIPE, h->stream));
    CKC(cudaMemsetAsync(h->d_acc_prev, 0, sizeof(unsigned long long) * h->R, h->stream));
    CKC(cudaMemsetAsync(h->d_ctr, 0, sizeof(unsigned long long), h->stream));
    {
        std::vector<double> ones(h->R, 1.0), sig(h->R, 60.0);
        CKC(cudaMemcpyAsync(h->d_beta, ones.data(), sizeof(double) * h->R, cudaMemcpyHostToDevice, h->stream));
        CKC(cudaMemcpyAsync(h->d_sigma, sig.data(), sizeof(double) * h->R, cudaMemcpyHostToDevice, h->stream));
        CKC(cudaStreamSynchronize(h->stream));
    }
    h->acc_base.assign(h->R, 0ULL);
    if (std::getenv("CSMC_REPLICA_BLOCKS")) h->n_blocks = replica_blocks_wanted(h);   // otherwise 1 until the autotune decides

    // per-colour launch geometry and parameter blocks
    h->pass_blocks.assign(hm.n_colours, 1);
    h->partial_base.assign(hm.n_colours, 0);
    h->n_partials = 0;
    for (int c = 0; c < hm.n_colours; ++c) {
        int mx = 1;
        for (int s = hm.colour_seg_begin[c]; s < hm.colour_seg_begin[c + 1]; ++s) mx = std::max(mx, hm.segs[s].count);
        h->pass_blocks[c] = (mx + TPB - 1) / TPB;
        h->partial_base[c] = h->n_partials;
        h->n_partials += h->pass_blocks[c] * (hm.colour_seg_begin[c + 1] - hm.colour_seg_begin[c]);
    }
    CKC(dalloc(&h->d_partials, (size_t)h->R * h->n_partials * 4));
    CKC(dalloc(&h->d_meas, (size_t)h->R * 8));
    if (h->large) h->pl.resize(hm.n_colours); else h->ps.resize(hm.n_colours);
    for (int c = 0; c < hm.n_colours; ++c) {
        std::string pe;
        if (h->large) {
            pe = fill_pass_params(hm, c, h->pl[c]);
            h->pl[c].spins = h->d_spins; h->pl[c].nbr = h->d_nbr; h->pl[c].ref_of_pos = h->d_ref_of_pos;
        } else {
            pe = fill_pass_params(hm, c, h->ps[c]);
            h->ps[c].spins = h->d_spins; h->ps[c].nbr = h->d_nbr; h->ps[c].ref_of_pos = h->d_ref_of_pos;
        }
        if (!pe.empty()) return bail(CSMC_ERR_UNSUPPORTED, pe);
    }
    // runtime specialisation: eager for problems large enough to be bandwidth-bound (or on request),
    // lazy (first long sweep request) for small lattices, which then run on the resident kernel
    if ((h->flags & CSMC_FLAG_JIT) && !hm.structured)
        return bail(CSMC_ERR_UNSUPPORTED, "CSMC_FLAG_JIT: the model has no periodic colouring pattern (explicit-table kernels only)");
    if ((h->flags & CSMC_FLAG_JIT) || (int64_t)hm.N * h->R >= 32768) {
        std::string jerr = build_jit(h);
        if (!jerr.empty() && (h->flags & CSMC_FLAG_JIT)) return bail(CSMC_ERR_UNSUPPORTED, "runtime specialisation failed: " + jerr);
        if (jerr.empty() && (int64_t)hm.N * h->R >= 32768 && autotune_pdl(h) != CSMC_OK) return bail(CSMC_ERR_CUDA, "autotune failed: " + h->err);
    }
#undef CKC
    *out = h;
    return CSMC_OK;
}

int32_t csmc_plan(const csmc_model *model, int32_t flags, int32_t *colour, int32_t *n_colours,
                  int32_t *structured, int32_t *storage_pos) {
    if (!model) return fail(nullptr, CSMC_ERR_INVALID, "csmc_plan: NULL model");
    HostModel hm;
    std::string e;
    try {
        e = build_host_model(model, flags, hm);
    } catch (const std::exception &ex) {
        e = std::string("model build failed: ") + ex.what();
    }
    if (!e.empty()) return fail(nullptr, CSMC_ERR_INVALID, e);
    if (colour) std::memcpy(colour, hm.colour_of_site.data(), sizeof(int32_t) * hm.N);
    if (storage_pos) std::memcpy(storage_pos, hm.pos_of_ref.data(), sizeof(int32_t) * hm.N);
    if (n_colours) *n_colours = hm.n_colours;
    if (structured) *structured = hm.structured ? 1 : 0;
    return CSMC_OK;
}

// Autotune (eagerly specialised handles only): programmatic dependent launch helps some models and
// hurts others on B200 (measured: +5 % square / 3-D pyrochlore L=64, -8 % honeycomb), so both variants
// are built and the faster one on a short probe run (5 cycles of 10 OR + 1 Metropolis on random spins)
// is kept.  Leaves the handle in its freshly created state.
static int autotune_pdl(csmc_handle *h) {
    if (!h->jit || (h->flags & (CSMC_FLAG_PDL | CSMC_FLAG_NO_AUTOTUNE | CSMC_FLAG_NO_GRAPH)) || h->hm.self_loop) return CSMC_OK;
    if (h->jit_resident && h->hm.N <= 4096 && !(h->flags & CSMC_FLAG_NO_RESIDENT)) return CSMC_OK;   // sweeps run on the resident kernel
    JitModule other;
    if (!load_jit_module(h->hm, true, other, (h->flags & CSMC_FLAG_FUSED) != 0, want_skew(h)).empty()) { if (other.lib) cudaLibraryUnload(other.lib); cudaGetLastError(); return CSMC_OK; }
    JitModule base;
    base.lib = h->jit_lib; for (int u = 0; u < 4; ++u) base.sweep[u] = h->jit_sweep[u];
    base.energy = h->jit_energy; base.resident = h->jit_resident; base.plan = h->jit_plan; base.pdl = false;
    for (int u = 0; u < 4; ++u) base.fused[u] = h->jit_fused[u];
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int npad = h->hm.npad;
    dim3 grid((npad + 255) / 256, h->R);
    k_randomize<<<grid, 256, 0, h->stream>>>(h->d_spins, h->d_ref_of_pos, npad, 3LL * npad, h->hm.S, 0x7e57ULL, h->replica_base);
    const JitModule *mods[2] = {&base, &other};
    // ms of the best of 3 probe runs (5 cycles of 10 OR + 1 Metropolis) in the current configuration
    auto probe = [&](float &best) -> int {
        best = 1e30f;
        int rc = csmc_cycles_async(h, 3, 10, 1); if (rc) return rc;   // builds the graph, warms up
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0, h->stream));
            rc = csmc_cycles_async(h, 5, 10, 1); if (rc) return rc;
            CK(cudaEventRecord(e1, h->stream));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        return CSMC_OK;
    };
    const bool groups_from_env = std::getenv("CSMC_SWEEP_GROUPS") != nullptr;
    if (!groups_from_env) h->n_groups = 1;
    // the tile-resident kernel is timed last, against the best configuration of the pass kernels
    const bool persist_candidate = h->jit_persist != nullptr;
    const bool persist_forced = persist_candidate && !h->persist_off;
    h->persist_off = true;
    for (int v = 0; v < 2; ++v) {
        install_jit_module(h, *mods[v]);
        int rc = probe(h->tune_ms[v]); if (rc) return rc;
    }
    const int keep = h->tune_ms[1] < 0.97f * h->tune_ms[0] ? 1 : 0;
    install_jit_module(h, *mods[keep]);
    cudaLibraryUnload(mods[1 - keep]->lib);
    float best_so_far = h->tune_ms[keep];
    // time-skewed strips (enqueue_skewed_seq) were in use during the probes above if they apply to this lattice; unless
    // they were asked for explicitly, time the pass-by-pass order too and keep the faster
    {
        const char *e = std::getenv("CSMC_SKEW");
        const bool forced = (h->flags & CSMC_FLAG_SKEW) != 0 || (e && e[0] == '1');
        if (!forced && h->jit_plan.skew && skew_budget_rows(h) < h->jit_plan.skew_rows) {
            h->tune_skew_ms[1] = best_so_far;
            h->skew_off = true;
            drop_graphs(h);
            int rc = probe(h->tune_skew_ms[0]); if (rc) return rc;
            if (h->tune_skew_ms[0] < 0.97f * best_so_far) best_so_far = h->tune_skew_ms[0];
            else h->skew_off = false;
            drop_graphs(h);
        }
    }
    // replica blocks (enqueue_sweep_seq): all replicas per pass, or block by block so that a block stays in L2
    if (!std::getenv("CSMC_REPLICA_BLOCKS") && replica_blocks_wanted(h) > 1) {
        // candidates: the count the L2 budget asks for and one more (slightly smaller blocks leave room for the other
        // colours' lines and the write-backs: C3 x 64 replicas runs 3 % faster in 4 blocks of 48 MiB than in 3 of 64 MiB)
        h->tune_blocks_ms[0] = best_so_far;
        h->tune_blocks_ms[1] = 1e30f;
        const int want = replica_blocks_wanted(h);
        int best_nb = 1;
        for (int cand = want; cand <= std::min(want + 1, h->R); ++cand) {
            float ms = 0.f;
            h->n_blocks = cand;
            drop_graphs(h);
            int rc = probe(ms); if (rc) return rc;
            h->tune_blocks_ms[1] = std::min(h->tune_blocks_ms[1], ms);
            if (ms < 0.97f * h->tune_blocks_ms[0] && ms < best_so_far) { best_so_far = ms; best_nb = cand; }
        }
        h->n_blocks = best_nb;
        drop_graphs(h);
    }
    // replica groups on separate streams (enqueue_sweep_seq): 1, 2 or 4 concurrent chains
    if (!groups_from_env && h->R >= 2) {
        h->tune_groups_ms[0] = best_so_far;
        int best_g = 1;
        float best_ms = h->tune_groups_ms[0];
        for (int gi = 1; gi <= 2; ++gi) {
            const int g = 1 << gi;
            if (h->R < g || ensure_group_streams(h, g) != g) break;
            h->n_groups = g;
            drop_graphs(h);
            int rc = probe(h->tune_groups_ms[gi]); if (rc) return rc;
            if (h->tune_groups_ms[gi] < 0.97f * best_ms) { best_ms = h->tune_groups_ms[gi]; best_g = g; }
        }
        h->n_groups = best_g;
        drop_graphs(h);
    }
    if (persist_candidate) {
        const bool forced = persist_forced;
        h->tune_persist_ms[0] = best_so_far;
        h->persist_off = false;
        drop_graphs(h);
        int rc = probe(h->tune_persist_ms[1]);
        if (rc) { h->persist_off = true; h->err.clear(); drop_graphs(h); }
        else if (!forced && !(h->tune_persist_ms[1] < 0.97f * best_so_far)) { h->persist_off = true; drop_graphs(h); }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    // back to the freshly created state
    CK(cudaMemsetAsync(h->d_spins, 0, sizeof(double) * h->R * 3 * npad, h->stream));
    CK(cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * h->R * ACC_STRIPE, h->stream));
    CK(cudaMemsetAsync(h->d_ctr, 0, sizeof(unsigned long long), h->stream));
    h->metro_ctr = 0;
    h->launches = 0;
    CK(cudaStreamSynchronize(h->stream));
    return CSMC_OK;
}

int32_t csmc_reference_tables(const csmc_model *model, int64_t *bil, int64_t *cub, int64_t *quar) {
    if (!model) return fail(nullptr, CSMC_ERR_INVALID, "csmc_reference_tables: NULL model");
    HostModel hm;
    std::string e;
    try {
        e = build_host_model(model, CSMC_FLAG_FORCE_GENERIC, hm);  // validates the model
        if (e.empty()) reference_tables(model, bil, cub, quar);
    } catch (const std::exception &ex) {
        e = std::string("model build failed: ") + ex.what();
    }
    if (!e.empty()) return fail(nullptr, CSMC_ERR_INVALID, e);
    return CSMC_OK;
}

int32_t csmc_destroy(csmc_handle *h) {
    if (!h) return CSMC_OK;
    cudaSetDevice(h->device);
    if (h->stream && h->capture_open) {   // a capture this handle began must not outlive it (never ends a caller's own capture)
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(h->stream, &g);
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        h->capture_open = false;
    }
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto &kv : h->cycle_graphs) cudaGraphExecDestroy(kv.second.exec);
    for (auto &kv : h->or_graphs) cudaGraphExecDestroy(kv.second);
    for (auto st : h->aux_streams) cudaStreamDestroy(st);
    for (auto ev : h->aux_done) cudaEventDestroy(ev);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->jit_lib) cudaLibraryUnload(h->jit_lib);
    persist_release(h);
    peer_gather_release(h);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    free_pt(h);
    cudaFree(h->d_spins); cudaFree(h->d_spins_alt); cudaFree(h->d_stage); cudaFree(h->d_out); cudaFree(h->d_nbr); cudaFree(h->d_ref_of_pos);
    cudaFree(h->d_beta); cudaFree(h->d_sigma); cudaFree(h->d_acc); cudaFree(h->d_acc_prev); cudaFree(h->d_ctr);
    cudaFree(h->d_partials); cudaFree(h->d_meas);
    cudaFree(h->d_ssf_theta); cudaFree(h->d_ssf_phib); cudaFree(h->d_ssf_partial); cudaFree(h->d_ssf_out); cudaFree(h->d_ssf_sum);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return CSMC_OK;
}

#define NEED(h_) do { if (!(h_)) return CSMC_ERR_INVALID; } while (0)
#define NEEDARG(h_, p_) do { if (!(p_)) return fail((csmc_handle *)(h_), CSMC_ERR_INVALID, std::string(__func__) + ": NULL argument"); } while (0)

int32_t csmc_n_sites(const csmc_handle *h, int64_t *n) { NEED(h); NEEDARG(h, n); *n = h->hm.N; return CSMC_OK; }
int32_t csmc_n_replicas(const csmc_handle *h, int32_t *r) { NEED(h); NEEDARG(h, r); *r = h->R; return CSMC_OK; }
int32_t csmc_n_colours(const csmc_handle *h, int32_t *c) { NEED(h); NEEDARG(h, c); *c = h->hm.n_colours; return CSMC_OK; }
int32_t csmc_is_structured(const csmc_handle *h, int32_t *f) { NEED(h); NEEDARG(h, f); *f = h->hm.structured ? 1 : 0; return CSMC_OK; }
int32_t csmc_kernel_mode(const csmc_handle *h, int32_t *mode) {
    NEED(h); NEEDARG(h, mode);
    *mode = h->jit ? (h->jit_resident && h->hm.N <= 4096 && !(h->flags & CSMC_FLAG_NO_RESIDENT) ? 3 : 2) : (h->hm.structured ? 1 : 0);
    if (!h->jit && !h->jit_note.empty()) const_cast<csmc_handle *>(h)->err = "runtime specialisation unavailable: " + h->jit_note;
    return CSMC_OK;
}

int32_t csmc_autotune_report(const csmc_handle *h, float ms[2], int32_t *pdl_selected) {
    NEED(h);
    if (ms) { ms[0] = h->tune_ms[0]; ms[1] = h->tune_ms[1]; }
    if (pdl_selected) *pdl_selected = h->jit_pdl ? 1 : 0;
    return CSMC_OK;
}

int32_t csmc_sweep_groups(const csmc_handle *h, int32_t *groups, float ms[3]) {
    NEED(h);
    if (groups) *groups = std::max(1, h->n_groups);
    if (ms) for (int i = 0; i < 3; ++i) ms[i] = h->tune_groups_ms[i];
    return CSMC_OK;
}

int32_t csmc_skew_schedule(int32_t n_rows, int32_t n_passes, int32_t reach, int32_t budget_rows, int32_t *launches, int64_t cap, int64_t *n) {
    if (!n) return fail(nullptr, CSMC_ERR_INVALID, "csmc_skew_schedule: NULL argument");
    std::vector<SkewLaunch> plan;
    try {
        plan = skew_schedule(n_rows, n_passes, reach, budget_rows);
    } catch (const std::exception &ex) {
        return fail(nullptr, CSMC_ERR_NOMEM, ex.what());
    }
    *n = (int64_t)plan.size();
    if (launches)
        for (int64_t i = 0; i < std::min<int64_t>(cap, *n); ++i) {
            launches[3 * i] = plan[i].pass; launches[3 * i + 1] = plan[i].row0; launches[3 * i + 2] = plan[i].nrows;
        }
    return CSMC_OK;
}

int32_t csmc_skew_geometry(const csmc_model *model, int32_t *usable, int32_t *tile_rows, int32_t *reach, int32_t *tiles_per_row) {
    if (!model || !usable) return fail(nullptr, CSMC_ERR_INVALID, "csmc_skew_geometry: NULL argument");
    HostModel hm;
    std::string e;
    JitPlan plan;
    plan.want_skew = true;
    try {
        e = build_host_model(model, 0, hm);
        if (e.empty() && hm.structured && !hm.self_loop) jit_generate_source(hm, false, &plan);
    } catch (const std::exception &ex) {
        e = std::string("csmc_skew_geometry: ") + ex.what();
    }
    if (!e.empty()) return fail(nullptr, CSMC_ERR_INVALID, e);
    *usable = plan.skew ? 1 : 0;
    if (tile_rows) *tile_rows = plan.skew_rows;
    if (reach) *reach = plan.skew_reach;
    if (tiles_per_row) *tiles_per_row = plan.skew_tiles_per_row;
    return CSMC_OK;
}

int32_t csmc_skew_info(const csmc_handle *h, int32_t *usable, int32_t *tile_rows, int32_t *reach, int32_t *budget_rows) {
    NEED(h); NEEDARG(h, usable);
    *usable = (h->jit && h->jit_plan.skew && !h->skew_off) ? 1 : 0;
    if (tile_rows) *tile_rows = h->jit_plan.skew_rows;
    if (reach) *reach = h->jit_plan.skew_reach;
    if (budget_rows) *budget_rows = *usable ? (int32_t)std::min<long>(skew_budget_rows(h), 1L << 30) : 0;
    return CSMC_OK;
}

int32_t csmc_persist_info(const csmc_handle *h, int32_t *tiles, int32_t grid[2], int32_t *replicas_per_launch, int32_t *smem_bytes, float ms[2]) {
    NEED(h);
    const bool on = h->jit_persist && !h->persist_off;
    if (tiles) *tiles = on ? h->persist_plan.persist_tiles : 0;
    if (grid) { grid[0] = on ? h->persist_plan.persist_g[0] : 0; grid[1] = on ? h->persist_plan.persist_g[1] : 0; }
    if (replicas_per_launch) *replicas_per_launch = on ? h->persist_plan.persist_nrep : 0;
    if (smem_bytes) *smem_bytes = on ? h->persist_plan.persist_smem : 0;
    if (ms) { ms[0] = h->tune_persist_ms[0]; ms[1] = h->tune_persist_ms[1]; }
    return CSMC_OK;
}

int32_t csmc_replica_blocks(const csmc_handle *h, int32_t *blocks, float ms[2]) {
    NEED(h); NEEDARG(h, blocks);
    *blocks = std::max(1, std::min(h->n_blocks, h->R));
    if (ms) { ms[0] = h->tune_blocks_ms[0]; ms[1] = h->tune_blocks_ms[1]; }
    return CSMC_OK;
}

int32_t csmc_jit_check(const csmc_model *model, int32_t compile, char *source, int64_t source_cap,
                       int64_t *source_len, char *log, int64_t log_cap) {
    if (!model) return fail(nullptr, CSMC_ERR_INVALID, "csmc_jit_check: NULL model");
    HostModel hm;
    std::string e, src, lg;
    try {
        e = build_host_model(model, 0, hm);
        if (e.empty() && !hm.structured) e = "model has no periodic colouring pattern: explicit-table kernels only";
        JitPlan plan;
        plan.want_fused = true;   // the build check covers the experimental fused kernels too
        { const char *sk = std::getenv("CSMC_SKEW"); plan.want_skew = sk && sk[0] == '1'; }   // and, on request, the tile-offset variant
        if (e.empty()) src = jit_generate_source(hm, false, &plan);
        if (e.empty() && compile) {
            std::vector<char> cubin;
            e = jit_compile(src, cubin, lg);
        }
    } catch (const std::exception &ex) {
        e = std::string("jit check failed: ") + ex.what();
    }
    if (source_len) *source_len = (int64_t)src.size();
    if (source && source_cap > 0) { std::strncpy(source, src.c_str(), (size_t)source_cap - 1); source[source_cap - 1] = 0; }
    if (log && log_cap > 0) { std::strncpy(log, lg.c_str(), (size_t)log_cap - 1); log[log_cap - 1] = 0; }
    if (!e.empty()) return fail(nullptr, CSMC_ERR_UNSUPPORTED, e);
    return CSMC_OK;
}

int32_t csmc_persist_check(const csmc_model *model, int32_t n_replicas, int32_t n_sms, int32_t smem_max, int32_t compile,
                           char *source, int64_t source_cap, int64_t *source_len, char *log, int64_t log_cap, int32_t info[8]) {
    if (!model || !info) return fail(nullptr, CSMC_ERR_INVALID, "csmc_persist_check: NULL argument");
    HostModel hm;
    std::string e, src, lg;
    JitPlan plan;
    plan.persist_only = true;
    plan.persist_replicas = n_replicas;
    plan.persist_sms = n_sms > 0 ? n_sms : 148;
    plan.persist_smem_max = smem_max > 0 ? smem_max : 227 * 1024;
    try {
        e = build_host_model(model, 0, hm);
        if (e.empty() && (!hm.structured || hm.self_loop)) e = "model has no periodic colouring pattern (or interacts with itself): pass kernels only";
        if (e.empty()) src = jit_generate_source(hm, false, &plan);
        if (e.empty() && plan.persist && compile) {
            std::vector<char> cubin;
            e = jit_compile(src, cubin, lg);
        }
    } catch (const std::exception &ex) {
        e = std::string("persist check failed: ") + ex.what();
    }
    info[0] = plan.persist ? 1 : 0; info[1] = plan.persist_tiles; info[2] = plan.persist_g[0]; info[3] = plan.persist_g[1];
    info[4] = plan.persist_w[0]; info[5] = plan.persist_w[1]; info[6] = plan.persist_nrep; info[7] = plan.persist_smem;
    if (source_len) *source_len = (int64_t)src.size();
    if (source && source_cap > 0) { std::strncpy(source, src.c_str(), (size_t)source_cap - 1); source[source_cap - 1] = 0; }
    if (log && log_cap > 0) { std::strncpy(log, lg.c_str(), (size_t)log_cap - 1); log[log_cap - 1] = 0; }
    if (!e.empty()) return fail(nullptr, CSMC_ERR_UNSUPPORTED, e);
    return CSMC_OK;
}

int32_t csmc_kernel_costs(const csmc_handle *h, double *flops_per_or_update, double *bytes_per_update) {
    NEED(h);
    if (flops_per_or_update) *flops_per_or_update = h->jit ? h->jit_plan.flops_or_update : 0.0;
    if (bytes_per_update) *bytes_per_update = 24.0 * (h->hm.n_colours + 1);
    return CSMC_OK;
}

int32_t csmc_launch_count(const csmc_handle *h, int64_t *n) { NEED(h); NEEDARG(h, n); *n = h->launches; return CSMC_OK; }

int32_t csmc_get_colouring(const csmc_handle *h, int32_t *colour) {
    NEED(h); NEEDARG(h, colour);
    std::memcpy(colour, h->hm.colour_of_site.data(), sizeof(int32_t) * h->hm.N);
    return CSMC_OK;
}

int32_t csmc_get_tables(const csmc_handle *h, int64_t *bil, int64_t *cub, int64_t *quar) {
    NEED(h);
    reference_tables(&h->model, bil, cub, quar);
    return CSMC_OK;
}

int32_t csmc_set_spins(csmc_handle *h, int32_t replica, const double *spins) {
    NEED(h); NEEDARG(h, spins);
    if (replica < 0 || replica >= h->R) return fail(h, CSMC_ERR_INVALID, "replica out of range");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d_stage, spins, sizeof(double) * 3 * h->hm.N, cudaMemcpyHostToDevice, h->stream));
    const int npad = h->hm.npad;
    k_aos_to_soa<<<(npad + 255) / 256, 256, 0, h->stream>>>(h->d_stage, h->d_spins + (size_t)replica * 3 * npad, h->d_ref_of_pos, npad);
    h->launches++;
    return finish(h);
}

int32_t csmc_get_spins(csmc_handle *h, int32_t replica, double *spins) {
    NEED(h); NEEDARG(h, spins);
    if (replica < 0 || replica >= h->R) return fail(h, CSMC_ERR_INVALID, "replica out of range");
    CK(cudaSetDevice(h->device));
    const int npad = h->hm.npad;
    k_soa_to_aos<<<(npad + 255) / 256, 256, 0, h->stream>>>(h->d_spins + (size_t)replica * 3 * npad, h->d_stage, h->d_ref_of_pos, npad);
    h->launches++;
    CK(cudaMemcpyAsync(spins, h->d_stage, sizeof(double) * 3 * h->hm.N, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_randomize_spins(csmc_handle *h, uint64_t seed) {
    NEED(h);
    CK(cudaSetDevice(h->device));
    const int npad = h->hm.npad;
    dim3 grid((npad + 255) / 256, h->R);
    k_randomize<<<grid, 256, 0, h->stream>>>(h->d_spins, h->d_ref_of_pos, npad, 3LL * npad, h->hm.S, seed, h->replica_base);
    h->launches++;
    return finish(h);
}

int32_t csmc_local_field_all(csmc_handle *h, int32_t replica, double *out) {
    NEED(h); NEEDARG(h, out);
    if (replica < 0 || replica >= h->R) return fail(h, CSMC_ERR_INVALID, "replica out of range");
    CK(cudaSetDevice(h->device));
    enqueue_eval(h, replica, 0, h->d_out);
    CK(cudaMemcpyAsync(out, h->d_out, sizeof(double) * 3 * h->hm.N, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_local_field(csmc_handle *h, int32_t replica, int64_t site, double out[3]) {
    NEED(h); NEEDARG(h, out);
    if (replica < 0 || replica >= h->R) return fail(h, CSMC_ERR_INVALID, "replica out of range");
    if (site < 1 || site > h->hm.N) return fail(h, CSMC_ERR_INVALID, "site out of range (1-based)");
    CK(cudaSetDevice(h->device));
    enqueue_eval(h, replica, 0, h->d_out, site - 1);
    CK(cudaMemcpyAsync(out, h->d_out + 3 * (site - 1), sizeof(double) * 3, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_site_energy_all(csmc_handle *h, int32_t replica, double *out) {
    NEED(h); NEEDARG(h, out);
    if (replica < 0 || replica >= h->R) return fail(h, CSMC_ERR_INVALID, "replica out of range");
    CK(cudaSetDevice(h->device));
    enqueue_eval(h, replica, 1, h->d_out);
    CK(cudaMemcpyAsync(out, h->d_out, sizeof(double) * h->hm.N, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

static int measure_to_host(csmc_handle *h, std::vector<double> &rec) {
    CK(cudaSetDevice(h->device));
    enqueue_measure(h, h->d_meas, true);
    rec.resize((size_t)h->R * 8);
    CK(cudaMemcpyAsync(rec.data(), h->d_meas, sizeof(double) * h->R * 8, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_total_energy(csmc_handle *h, double *E) {
    NEED(h); NEEDARG(h, E);
    std::vector<double> rec;
    int rc = measure_to_host(h, rec);
    if (rc) return rc;
    for (int r = 0; r < h->R; ++r) E[r] = rec[8 * r];
    return CSMC_OK;
}

int32_t csmc_magnetization(csmc_handle *h, double *M3) {
    NEED(h); NEEDARG(h, M3);
    std::vector<double> rec;
    int rc = measure_to_host(h, rec);
    if (rc) return rc;
    for (int r = 0; r < h->R; ++r) for (int k = 0; k < 3; ++k) M3[3 * r + k] = rec[8 * r + 1 + k];
    return CSMC_OK;
}

int32_t csmc_overrelax(csmc_handle *h, int32_t n_sweeps) {
    NEED(h);
    CK(cudaSetDevice(h->device));
    if (n_sweeps > 0 && use_resident(h, n_sweeps)) enqueue_resident(h, 1, n_sweeps, 0, 0, 0, nullptr, 0);
    else {
        // blocks of up to 64 sweeps: one graph replay each (replica groups on concurrent streams included)
        for (int left = n_sweeps; left > 0;) {
            const int k = std::min(left, 64);
            int rc = enqueue_or_block(h, k); if (rc) return rc;
            left -= k;
        }
    }
    return finish(h);
}

int32_t csmc_deterministic(csmc_handle *h, int32_t n_sweeps) {
    NEED(h);
    CK(cudaSetDevice(h->device));
    if (n_sweeps > 0 && use_resident(h, n_sweeps)) enqueue_resident(h, 0, 0, 0, 0, n_sweeps, nullptr, 0);
    else {
        // as a sequence (blocks of <= 48 sweeps), so that the tile-resident kernel / strips / replica blocks apply
        const std::vector<SweepOp> seq((size_t)std::min(std::max(n_sweeps, 0), 48), SweepOp{UPD_DET, 0ULL, false});
        for (int left = n_sweeps; left > 0; left -= 48) enqueue_sweep_seq(h, seq.data(), std::min(left, 48), false);
    }
    return finish(h);
}

int32_t csmc_set_temperatures(csmc_handle *h, const double *T) {
    NEED(h); NEEDARG(h, T);
    CK(cudaSetDevice(h->device));
    return upload_T(h, T);
}

int32_t csmc_set_sigma(csmc_handle *h, const double *sigma) {
    NEED(h); NEEDARG(h, sigma);
    for (int r = 0; r < h->R; ++r) if (!(sigma[r] >= 0.0)) return fail(h, CSMC_ERR_INVALID, "sigma must be >= 0");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d_sigma, sigma, sizeof(double) * h->R, cudaMemcpyHostToDevice, h->stream));
    return finish(h);
}

int32_t csmc_get_sigma(csmc_handle *h, double *sigma) {
    NEED(h); NEEDARG(h, sigma);
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(sigma, h->d_sigma, sizeof(double) * h->R, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_get_accepted(csmc_handle *h, double *accepted, int32_t reset) {
    NEED(h); NEEDARG(h, accepted);
    CK(cudaSetDevice(h->device));
    std::vector<unsigned long long> stripes((size_t)h->R * ACC_STRIPE), now(h->R, 0ULL);
    CK(cudaMemcpyAsync(stripes.data(), h->d_acc, sizeof(unsigned long long) * stripes.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < h->R; ++r) for (int k = 0; k < ACC_STRIPE; ++k) now[r] += stripes[(size_t)r * ACC_STRIPE + k];
    for (int r = 0; r < h->R; ++r) { accepted[r] = (double)(now[r] - h->acc_base[r]); if (reset) h->acc_base[r] = now[r]; }
    return CSMC_OK;
}

int32_t csmc_metropolis(csmc_handle *h, const double *T, int32_t n_sweeps, double *accepted) {
    NEED(h); NEEDARG(h, T);
    int rc = check_metropolis(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    rc = upload_T(h, T); if (rc) return rc;
    std::vector<double> before(h->R), after(h->R);
    if (accepted) { rc = csmc_get_accepted(h, before.data(), 0); if (rc) return rc; }
    if (n_sweeps > 0) { rc = csmc_cycles_async(h, n_sweeps, 0, 1); if (rc) return rc; }   // resident kernel or graph replay
    rc = finish(h); if (rc) return rc;
    if (accepted) {
        rc = csmc_get_accepted(h, after.data(), 0); if (rc) return rc;
        for (int r = 0; r < h->R; ++r) accepted[r] = after[r] - before[r];
    }
    return CSMC_OK;
}

int32_t csmc_metropolis_cone(csmc_handle *h, const double *T, double *sigma, int32_t adapt, int32_t n_sweeps, double *accepted) {
    NEED(h); NEEDARG(h, T); NEEDARG(h, sigma);
    int rc = check_metropolis(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    rc = upload_T(h, T); if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_sigma, sigma, sizeof(double) * h->R, cudaMemcpyHostToDevice, h->stream));
    std::vector<double> before(h->R), after(h->R);
    rc = csmc_get_accepted(h, before.data(), 0); if (rc) return rc;
    {
        // running totals as of now, for the per-sweep acceptance of the adaptive rule
        std::vector<unsigned long long> prev(h->R);
        for (int r = 0; r < h->R; ++r) prev[r] = (unsigned long long)before[r] + h->acc_base[r];
        CK(cudaMemcpyAsync(h->d_acc_prev, prev.data(), sizeof(unsigned long long) * h->R, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (n_sweeps > 0 && use_resident(h, n_sweeps)) enqueue_resident(h, 1, 0, n_sweeps, 1, 0, nullptr, 0, adapt ? 1 : 0);
    else if (!adapt && n_sweeps >= 2) {
        // fixed cone width: the sweeps are one sequence (tile-resident kernel / strips / replica blocks apply)
        std::vector<SweepOp> seq;
        for (int done = 0; done < n_sweeps;) {
            const int k = std::min(n_sweeps - done, 48);
            seq.clear();
            for (int s = 0; s < k; ++s) seq.push_back({UPD_CONE, h->metro_ctr + (unsigned long long)s, false});
            enqueue_sweep_seq(h, seq.data(), k, false);
            h->metro_ctr += (unsigned long long)k;
            done += k;
        }
    } else for (int s = 0; s < n_sweeps; ++s) {
        enqueue_metropolis(h, true);
        if (adapt) { k_adapt_sigma<<<(h->R + 127) / 128, 128, 0, h->stream>>>(h->d_sigma, h->d_acc, h->d_acc_prev, (double)h->hm.N, h->R); h->launches++; }
    }
    CK(cudaMemcpyAsync(sigma, h->d_sigma, sizeof(double) * h->R, cudaMemcpyDeviceToHost, h->stream));
    rc = finish(h); if (rc) return rc;
    if (accepted) {
        rc = csmc_get_accepted(h, after.data(), 0); if (rc) return rc;
        for (int r = 0; r < h->R; ++r) accepted[r] = after[r] - before[r];
    }
    return CSMC_OK;
}

// ---- cycles: CUDA-graph replay of (or_per_cycle OR sweeps + metro_per_cycle Metropolis sweeps) --------
static int build_cycle_graph(csmc_handle *h, int orc, int mc, const csmc_handle::CycleGraph **out) {
    const auto key = std::make_pair(orc, mc);
    auto it = h->cycle_graphs.find(key);
    if (it != h->cycle_graphs.end()) { *out = &it->second; return CSMC_OK; }
    cudaGraph_t graph = nullptr;
    const long long before = h->launches;
    const bool fused = fused_ready(h);
    sweep_groups(h, 2);   // streams / events of the replica groups exist before the capture starts
    std::vector<SweepOp> seq;
    for (int s = 0; s < orc; ++s) seq.push_back({UPD_OR, 0ULL, false});
    for (int s = 0; s < mc; ++s) seq.push_back({UPD_METRO, (unsigned long long)s, true});
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->capture_open = true;
    enqueue_sweep_seq(h, seq.data(), (int)seq.size(), fused);
    if (mc > 0) { k_add_u64<<<1, 1, 0, h->stream>>>(h->d_ctr, (unsigned long long)mc); h->launches++; }
    h->capture_open = false;
    CK(cudaStreamEndCapture(h->stream, &graph));
    const long long launches = h->launches - before;
    h->launches = before;  // capture enqueues nothing
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(h, CSMC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    if (h->cycle_graphs.size() >= 16) {
        CK(cudaStreamSynchronize(h->stream));   // replays in flight keep their executable graphs until here
        for (auto &kv : h->cycle_graphs) cudaGraphExecDestroy(kv.second.exec);
        h->cycle_graphs.clear();
    }
    *out = &h->cycle_graphs.emplace(key, csmc_handle::CycleGraph{exec, launches}).first->second;
    return CSMC_OK;
}

int32_t csmc_cycles_async(csmc_handle *h, int64_t n_cycles, int32_t orc, int32_t mc) {
    NEED(h);
    if (n_cycles < 0 || orc < 0 || mc < 0) return fail(h, CSMC_ERR_INVALID, "negative cycle counts");
    if (mc > 0) { int rc = check_metropolis(h); if (rc) return rc; }
    CK(cudaSetDevice(h->device));
    if (n_cycles > 0 && orc + mc > 0 && use_resident(h, n_cycles * (orc + mc))) {
        for (int64_t done = 0; done < n_cycles;) {
            const int chunk = (int)std::min<int64_t>(n_cycles - done, 1 << 20);
            enqueue_resident(h, chunk, orc, mc, 0, 0, nullptr, 0);
            done += chunk;
        }
        CK(cudaGetLastError());
        return CSMC_OK;
    }
    if (h->flags & CSMC_FLAG_NO_GRAPH) {
        const bool fused = fused_ready(h);
        std::vector<SweepOp> seq;
        for (int64_t c = 0; c < n_cycles; ++c) {
            seq.clear();
            for (int s = 0; s < orc; ++s) seq.push_back({UPD_OR, 0ULL, false});
            for (int s = 0; s < mc; ++s) seq.push_back({UPD_METRO, h->metro_ctr + s, false});
            enqueue_sweep_seq(h, seq.data(), (int)seq.size(), fused);
            h->metro_ctr += mc;
        }
        CK(cudaGetLastError());
        return CSMC_OK;
    }
    // the graph reads the sweep counter from device memory: bring it up to date first
    CK(cudaMemcpyAsync(h->d_ctr, &h->metro_ctr, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
    const csmc_handle::CycleGraph *cg = nullptr;
    int rc = build_cycle_graph(h, orc, mc, &cg); if (rc) return rc;
    for (int64_t c = 0; c < n_cycles; ++c) CK(cudaGraphLaunch(cg->exec, h->stream));
    h->metro_ctr += (unsigned long long)n_cycles * mc;
    h->launches += n_cycles * cg->launches;
    CK(cudaGetLastError());
    return CSMC_OK;
}

// n consecutive overrelaxation sweeps: one graph replay instead of 2*n*colours launches
static int enqueue_or_block(csmc_handle *h, int n) {
    if (n <= 0) return CSMC_OK;
    const bool fused = fused_ready(h);
    const std::vector<SweepOp> seq((size_t)n, SweepOp{UPD_OR, 0ULL, false});
    if ((h->flags & CSMC_FLAG_NO_GRAPH) || n < 2) {
        enqueue_sweep_seq(h, seq.data(), n, fused);
        return CSMC_OK;
    }
    auto it = h->or_graphs.find(n);
    if (it == h->or_graphs.end()) {
        cudaGraph_t graph = nullptr;
        const long long before = h->launches;
        sweep_groups(h, 2);   // streams / events of the replica groups exist before the capture starts
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->capture_open = true;
        enqueue_sweep_seq(h, seq.data(), n, fused);
        h->capture_open = false;
        CK(cudaStreamEndCapture(h->stream, &graph));
        h->or_graph_launches[n] = h->launches - before;
        h->launches = before;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(h, CSMC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        if (h->or_graphs.size() > 64) {
            CK(cudaStreamSynchronize(h->stream));   // replays in flight keep their executable graphs until here
            for (auto &kv : h->or_graphs) cudaGraphExecDestroy(kv.second);
            h->or_graphs.clear();
        }
        it = h->or_graphs.emplace(n, exec).first;
    }
    CK(cudaGraphLaunch(it->second, h->stream));
    h->launches += h->or_graph_launches[n];
    return CSMC_OK;
}

int32_t csmc_sync(csmc_handle *h) { NEED(h); CK(cudaSetDevice(h->device)); return finish(h); }

int32_t csmc_anneal_temperature(csmc_handle *h, const double *T, int64_t t_thermalization, int32_t rate, double *accepted) {
    NEED(h); NEEDARG(h, T);
    if (rate < 0) return fail(h, CSMC_ERR_INVALID, "overrelaxation_rate must be >= 0");
    int rc = check_metropolis(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    rc = upload_T(h, T); if (rc) return rc;
    std::vector<double> before(h->R), after(h->R);
    if (accepted) { rc = csmc_get_accepted(h, before.data(), 0); if (rc) return rc; }
    const int64_t iters = t_thermalization - 1;  // src/monte_carlo.jl:169-172: t = 1 .. t_thermalization-1
    if (iters > 0) {
        if (rate == 0) {
            rc = csmc_cycles_async(h, iters, 0, 1); if (rc) return rc;       // :178-180
        } else {
            // t % rate == 0 closes a block of `rate` OR sweeps followed by one Metropolis sweep (:173-177)
            rc = csmc_cycles_async(h, iters / rate, rate, 1); if (rc) return rc;
            rc = csmc_cycles_async(h, 1, (int32_t)(iters % rate), 0); if (rc) return rc;
        }
    }
    rc = finish(h); if (rc) return rc;
    if (accepted) {
        rc = csmc_get_accepted(h, after.data(), 0); if (rc) return rc;
        for (int r = 0; r < h->R; ++r) accepted[r] = after[r] - before[r];
    }
    return CSMC_OK;
}

int32_t csmc_anneal_temperature_cone(csmc_handle *h, const double *T, double *sigma, int32_t adapt, int64_t t_thermalization,
                                     int32_t rate, double *accepted) {
    NEED(h); NEEDARG(h, T); NEEDARG(h, sigma);
    if (rate < 0) return fail(h, CSMC_ERR_INVALID, "overrelaxation_rate must be >= 0");
    int rc = check_metropolis(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    rc = upload_T(h, T); if (rc) return rc;
    rc = csmc_set_sigma(h, sigma); if (rc) return rc;
    std::vector<double> before(h->R), after(h->R);
    rc = csmc_get_accepted(h, before.data(), 0); if (rc) return rc;
    {
        std::vector<unsigned long long> prev(h->R);
        for (int r = 0; r < h->R; ++r) prev[r] = (unsigned long long)before[r] + h->acc_base[r];
        CK(cudaMemcpyAsync(h->d_acc_prev, prev.data(), sizeof(unsigned long long) * h->R, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const int64_t iters = t_thermalization - 1;                      // src/monte_carlo.jl:169-172
    if (iters > 0) {
        const int orc = rate == 0 ? 0 : rate;
        const int64_t n_cycles = rate == 0 ? iters : iters / rate;
        const int tail = rate == 0 ? 0 : (int)(iters % rate);
        if (use_resident(h, iters)) {
            for (int64_t done = 0; done < n_cycles;) {
                const int chunk = (int)std::min<int64_t>(n_cycles - done, 1 << 20);
                enqueue_resident(h, chunk, orc, 1, 1, 0, nullptr, 0, adapt ? 1 : 0);
                done += chunk;
            }
            if (tail) enqueue_resident(h, 1, tail, 0, 0, 0, nullptr, 0);
        } else {
            for (int64_t c = 0; c < n_cycles; ++c) {
                rc = enqueue_or_block(h, orc); if (rc) return rc;
                enqueue_metropolis(h, true);
                if (adapt) { k_adapt_sigma<<<(h->R + 127) / 128, 128, 0, h->stream>>>(h->d_sigma, h->d_acc, h->d_acc_prev, (double)h->hm.N, h->R); h->launches++; }
                if ((c & 255) == 255) CK(cudaGetLastError());
            }
            rc = enqueue_or_block(h, tail); if (rc) return rc;
        }
    }
    rc = finish(h); if (rc) return rc;
    rc = csmc_get_sigma(h, sigma); if (rc) return rc;
    if (accepted) {
        rc = csmc_get_accepted(h, after.data(), 0); if (rc) return rc;
        for (int r = 0; r < h->R; ++r) accepted[r] = after[r] - before[r];
    }
    return CSMC_OK;
}

// ---- equal-time structure factor ---------------------------------------------------------------------------
static int ssf_prepare(csmc_handle *h, const double *lattice_vectors, const double *basis, const double *ks, int64_t n_k) {
    const HostModel &hm = h->hm;
    const int D = hm.D;
    if (n_k < 1 || n_k > (1 << 24)) return fail(h, CSMC_ERR_INVALID, "n_k out of range");
    std::vector<double> theta((size_t)n_k * MAXD, 0.0), phib((size_t)n_k * hm.n_basis, 0.0);
    for (int64_t k = 0; k < n_k; ++k) {
        for (int d = 0; d < D; ++d) {            // theta_d = k . a_d  (a_d = column d of lattice_vectors, D x D column-major)
            double t = 0.0;
            for (int c = 0; c < D; ++c) t += ks[k * D + c] * lattice_vectors[d * D + c];
            theta[k * MAXD + d] = t;
        }
        for (int b = 0; b < hm.n_basis; ++b) {   // basis: n_basis x D row-major
            double t = 0.0;
            for (int c = 0; c < D; ++c) t += ks[k * D + c] * basis[b * D + c];
            phib[k * hm.n_basis + b] = t;
        }
    }
    cudaFree(h->d_ssf_theta); cudaFree(h->d_ssf_phib); cudaFree(h->d_ssf_partial); cudaFree(h->d_ssf_out);
    h->d_ssf_theta = h->d_ssf_phib = h->d_ssf_partial = h->d_ssf_out = nullptr;
    SsfGeom &g = h->ssf;
    g.D = D; g.n_basis = hm.n_basis; g.npad = hm.npad; g.n_k = (int)n_k;
    g.table_len = hm.n_basis;
    for (int d = 0; d < MAXD; ++d) { g.L[d] = hm.L[d]; if (d < D) g.table_len += hm.L[d]; }
    const size_t per_k = (size_t)g.table_len * sizeof(double2);
    const size_t budget = 160 * 1024;
    if (per_k > budget) return fail(h, CSMC_ERR_UNSUPPORTED, "structure factor: lattice extents too large for the shared-memory phase tables");
    g.KT = (int)std::max<size_t>(1, std::min<size_t>(8, budget / per_k));
    const int k_tiles = (g.n_k + g.KT - 1) / g.KT;
    int chunks = std::max(1, (2 * 148 + k_tiles - 1) / k_tiles);
    chunks = std::min(chunks, std::max(1, hm.npad / 2048));
    g.chunk = ((hm.npad + chunks - 1) / chunks + 255) / 256 * 256;
    h->ssf_chunks = (hm.npad + g.chunk - 1) / g.chunk;
    CK(dalloc(&h->d_ssf_theta, theta.size())); CK(dalloc(&h->d_ssf_phib, phib.size()));
    CK(dalloc(&h->d_ssf_partial, (size_t)h->ssf_chunks * n_k * 6)); CK(dalloc(&h->d_ssf_out, (size_t)n_k * 9));
    CK(cudaMemcpyAsync(h->d_ssf_theta, theta.data(), sizeof(double) * theta.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_ssf_phib, phib.data(), sizeof(double) * phib.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaFuncSetAttribute(k_ssf_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(g.KT * per_k)));
    return CSMC_OK;
}

static void ssf_enqueue(csmc_handle *h, int replica, double *out, int accumulate, const int *slot_of_rep = nullptr) {
    const SsfGeom &g = h->ssf;
    const int k_tiles = (g.n_k + g.KT - 1) / g.KT;
    const size_t smem = (size_t)g.KT * g.table_len * sizeof(double2);
    k_ssf_partial<<<dim3(k_tiles, h->ssf_chunks), 256, smem, h->stream>>>(h->d_spins + (size_t)replica * 3 * h->hm.npad, h->d_ref_of_pos,
                                                                             h->d_ssf_theta, h->d_ssf_phib, g, h->d_ssf_partial);
    k_ssf_finish<<<(g.n_k + 127) / 128, 128, 0, h->stream>>>(h->d_ssf_partial, h->ssf_chunks, g.n_k, 1.0 / (double)h->hm.N, out, accumulate,
                                                              slot_of_rep, h->replica_base + replica);
    h->launches += 2;
}

int32_t csmc_structure_factor(csmc_handle *h, int32_t replica, const double *lattice_vectors, const double *basis,
                              const double *ks, int64_t n_k, double *Suv) {
    NEED(h); NEEDARG(h, lattice_vectors); NEEDARG(h, basis); NEEDARG(h, ks); NEEDARG(h, Suv);
    if (replica < 0 || replica >= h->R) return fail(h, CSMC_ERR_INVALID, "replica out of range");
    if (h->d_ssf_sum) return fail(h, CSMC_ERR_INVALID, "momenta of a running parallel-tempering measurement are attached (csmc_pt_set_momenta)");
    CK(cudaSetDevice(h->device));
    int rc = ssf_prepare(h, lattice_vectors, basis, ks, n_k); if (rc) return rc;
    ssf_enqueue(h, replica, h->d_ssf_out, 0);
    CK(cudaMemcpyAsync(Suv, h->d_ssf_out, sizeof(double) * 9 * n_k, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_pt_set_momenta(csmc_handle *h, const double *lattice_vectors, const double *basis, const double *ks, int64_t n_k) {
    NEED(h); NEEDARG(h, lattice_vectors); NEEDARG(h, basis); NEEDARG(h, ks);
    if (h->n_slots == 0) return fail(h, CSMC_ERR_INVALID, "csmc_pt_init has not been called");
    CK(cudaSetDevice(h->device));
    int rc = ssf_prepare(h, lattice_vectors, basis, ks, n_k); if (rc) return rc;
    cudaFree(h->d_ssf_sum); h->d_ssf_sum = nullptr;
    CK(dalloc(&h->d_ssf_sum, (size_t)h->n_slots * 9 * n_k));
    CK(cudaMemsetAsync(h->d_ssf_sum, 0, sizeof(double) * h->n_slots * 9 * n_k, h->stream));
    h->ssf_probes = 0;
    return finish(h);
}

int32_t csmc_pt_get_ssf(csmc_handle *h, double *sums, int64_t *n_probes) {
    NEED(h); NEEDARG(h, sums);
    if (!h->d_ssf_sum) return fail(h, CSMC_ERR_INVALID, "csmc_pt_set_momenta has not been called");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(sums, h->d_ssf_sum, sizeof(double) * h->n_slots * 9 * h->ssf.n_k, cudaMemcpyDeviceToHost, h->stream));
    if (n_probes) *n_probes = h->ssf_probes;
    return finish(h);
}

// ---- parallel tempering ----------------------------------------------------------------------------------
int32_t csmc_pt_init(csmc_handle *h, int32_t n_slots, const double *T_all) {
    NEED(h); NEEDARG(h, T_all);
    if (n_slots < h->replica_base + h->R) return fail(h, CSMC_ERR_INVALID, "n_slots smaller than replica_base + n_replicas");
    for (int s = 0; s < n_slots; ++s) if (!(T_all[s] > 0.0)) return fail(h, CSMC_ERR_INVALID, "temperatures must be > 0");
    CK(cudaSetDevice(h->device));
    free_pt(h);
    h->n_slots = n_slots;
    CK(dalloc(&h->d_T_slot, n_slots)); CK(dalloc(&h->d_meas_all, (size_t)n_slots * 8)); CK(dalloc(&h->d_E_last, n_slots));
    CK(dalloc(&h->d_acc_prev_pt, n_slots)); CK(dalloc(&h->d_acc_slot, n_slots)); CK(dalloc(&h->d_exch_slot, n_slots));
    CK(dalloc(&h->d_slot_of_rep, n_slots)); CK(dalloc(&h->d_rep_of_slot, n_slots)); CK(dalloc(&h->d_accepted_pairs, n_slots));
    CK(dalloc(&h->d_prev_rep_of_slot, n_slots));
    std::vector<int> ident(n_slots);
    for (int s = 0; s < n_slots; ++s) ident[s] = s;
    CK(cudaMemcpyAsync(h->d_T_slot, T_all, sizeof(double) * n_slots, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_slot_of_rep, ident.data(), sizeof(int) * n_slots, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_rep_of_slot, ident.data(), sizeof(int) * n_slots, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_meas_all, 0, sizeof(double) * 8 * n_slots, h->stream));
    CK(cudaMemsetAsync(h->d_E_last, 0, sizeof(double) * n_slots, h->stream));
    CK(cudaMemsetAsync(h->d_acc_prev_pt, 0, sizeof(double) * n_slots, h->stream));
    CK(cudaMemsetAsync(h->d_acc_slot, 0, sizeof(double) * n_slots, h->stream));
    CK(cudaMemsetAsync(h->d_exch_slot, 0, sizeof(double) * n_slots, h->stream));
    CK(cudaMemsetAsync(h->d_accepted_pairs, 0, sizeof(int) * n_slots, h->stream));
    {
        std::vector<double> beta(h->R);
        for (int r = 0; r < h->R; ++r) beta[r] = 1.0 / T_all[h->replica_base + r];
        CK(cudaMemcpyAsync(h->d_beta, beta.data(), sizeof(double) * h->R, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    // E = total_energy(mc.lattice) before the loop (src/monte_carlo.jl:265); acc_prev = current counters
    int rc = enqueue_measure_all(h, true); if (rc) return rc;
    PtState st = pt_state(h);
    k_pt_update<<<(n_slots + 127) / 128, 128, 0, h->stream>>>(st, 1); h->launches++;
    CK(cudaMemsetAsync(h->d_acc_slot, 0, sizeof(double) * n_slots, h->stream));
    return finish(h);
}

int32_t csmc_comm_unique_id(uint8_t id[128]) {
    if (!id) return CSMC_ERR_INVALID;
    if (!load_nccl()) return fail(nullptr, CSMC_ERR_NCCL, g_nccl.err);
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != 0) return fail(nullptr, CSMC_ERR_NCCL, "ncclGetUniqueId failed");
    std::memcpy(id, u.internal, 128);
    return CSMC_OK;
}

int32_t csmc_comm_init(csmc_handle *h, int32_t n_ranks, int32_t rank, const uint8_t id[128]) {
    NEED(h); NEEDARG(h, id);
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, CSMC_ERR_INVALID, "bad rank / n_ranks");
    if (!load_nccl()) return fail(h, CSMC_ERR_NCCL, g_nccl.err);
    CK(cudaSetDevice(h->device));
    peer_gather_release(h);
    if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    CKN(g_nccl.CommInitRank(&h->comm, n_ranks, u, rank));
    h->n_ranks = n_ranks; h->rank = rank;
    // every rank learns every rank's replica block (replica_base, n_replicas)
    double *d_part = nullptr;
    CK(dalloc(&d_part, (size_t)2 * n_ranks));
    const double mine[2] = {(double)h->replica_base, (double)h->R};
    std::vector<double> part((size_t)2 * n_ranks);
    cudaError_t ce = cudaMemcpyAsync(d_part + 2 * rank, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream);
    ncclResult_t nr = 0;
    if (ce == cudaSuccess) nr = g_nccl.AllGather(d_part + 2 * rank, d_part, 2, ncclFloat64, h->comm, h->stream);
    if (ce == cudaSuccess && nr == 0) ce = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess && nr == 0) ce = cudaStreamSynchronize(h->stream);
    cudaFree(d_part);
    CK(ce);
    CKN(nr);
    h->rank_base.resize(n_ranks); h->rank_count.resize(n_ranks);
    h->even_partition = true;
    for (int g = 0; g < n_ranks; ++g) {
        h->rank_base[g] = (long long)part[2 * g]; h->rank_count[g] = (long long)part[2 * g + 1];
        if (h->rank_count[g] != h->R || h->rank_base[g] != (long long)g * h->R) h->even_partition = false;
    }
    int peer_mode = 0;
    if (const char *e = std::getenv("CSMC_PEER_GATHER")) peer_mode = std::max(0, std::min(2, std::atoi(e)));
    return peer_gather_setup(h, peer_mode);
}

int32_t csmc_comm_mode(const csmc_handle *h, int32_t *mode) {
    NEED(h); NEEDARG(h, mode);
    *mode = !h->comm ? 0 : (h->peer.mode == 0 ? 1 : 1 + h->peer.mode);
    return CSMC_OK;
}

static int ensure_series(csmc_handle *h, long long need) {
    if (need <= h->series_cap) return CSMC_OK;
    long long cap = std::max<long long>(need, std::max<long long>(1024, 2 * h->series_cap));
    double *nE = nullptr, *nM = nullptr;
    CK(dalloc(&nE, (size_t)cap * h->n_slots)); CK(dalloc(&nM, (size_t)cap * h->n_slots));
    if (h->n_probes > 0) {
        CK(cudaMemcpyAsync(nE, h->d_series_E, sizeof(double) * h->n_probes * h->n_slots, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nM, h->d_series_M, sizeof(double) * h->n_probes * h->n_slots, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    cudaFree(h->d_series_E); cudaFree(h->d_series_M);
    h->d_series_E = nE; h->d_series_M = nM; h->series_cap = cap;
    return CSMC_OK;
}

int32_t csmc_pt_run(csmc_handle *h, const csmc_pt_params *p, int64_t sweep_begin, int64_t sweep_end) {
    NEED(h); NEEDARG(h, p);
    if (h->n_slots == 0) return fail(h, CSMC_ERR_INVALID, "csmc_pt_init has not been called");
    if (p->swap_rate < 1 || p->probe_rate < 1 || p->overrelaxation_rate < 0) return fail(h, CSMC_ERR_INVALID, "bad PT parameters");
    int rc = check_metropolis(h); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    const int rate = p->overrelaxation_rate;
    const int dosweep = rate == 0 ? 1 : rate;                                   // src/monte_carlo.jl:289-293
    const int alg = p->algorithm;                                               // the `alg` kwarg, :236
    if (alg < 0 || alg > 2) return fail(h, CSMC_ERR_INVALID, "PT algorithm must be 0 (Metropolis), 1 (adaptive) or 2 (fixed cone)");
    const int cone = alg != 0, adapt = alg == 1;
    if (adapt) {   // per-sweep acceptance of the adaptive rule is measured against the running totals
        std::vector<double> now(h->R);
        rc = csmc_get_accepted(h, now.data(), 0); if (rc) return rc;
        std::vector<unsigned long long> prev(h->R);
        for (int r = 0; r < h->R; ++r) prev[r] = (unsigned long long)now[r] + h->acc_base[r];
        CK(cudaMemcpyAsync(h->d_acc_prev, prev.data(), sizeof(unsigned long long) * h->R, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    // series capacity for the probes of this chunk
    long long probes = 0;
    for (int64_t s = std::max<int64_t>(sweep_begin, p->t_thermalization); s < sweep_end; ++s) if (s % p->probe_rate == 0) ++probes;
    rc = ensure_series(h, h->n_probes + probes); if (rc) return rc;
    PtState st = pt_state(h);
    const int nb = (h->n_slots + 127) / 128;
    const bool resident = use_resident(h, sweep_end - sweep_begin);
    rc = check_partition(h); if (rc) return rc;
    double *mine = h->d_meas_all + (size_t)h->replica_base * 8;
    auto gather = [&]() -> int { return enqueue_gather_meas(h); };
    int pending_or = 0;   // overrelaxation sweeps not yet enqueued (flushed as one graph replay)
    // plain Metropolis on the pass kernels: the pending OR sweeps and the Metropolis sweep are one graph (the
    // same cycle graphs csmc_cycles_async replays), so the replica groups run OR block + Metropolis as
    // uninterrupted concurrent chains; the sweep counter then lives on the device
    const bool fold = !resident && !cone && rate != 0 && !(h->flags & CSMC_FLAG_NO_GRAPH);
    if (fold) {
        CK(cudaMemcpyAsync(h->d_ctr, &h->metro_ctr, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    // The energies are consumed only by an exchange (same sweep) or by probes before the next Metropolis
    // sweep, so total_energy (src/monte_carlo.jl:305) is evaluated -- and gathered across GPUs -- only at
    // Metropolis sweeps where one of the two follows; the values that are consumed are unchanged.
    auto probe_at = [&](int64_t s) { return s >= p->t_thermalization && s % p->probe_rate == 0; };
    auto energy_needed = [&](int64_t s) {
        if (h->n_slots > 1 && s % p->swap_rate == 0) return true;
        for (int64_t q = s; q < s + dosweep; ++q) if (probe_at(q)) return true;
        return false;
    };
    for (int64_t sweep = sweep_begin; sweep < sweep_end; ++sweep) {
        if (rate != 0) ++pending_or;                                            // :298-300
        const bool metro = (sweep % dosweep == 0);
        const bool probe = probe_at(sweep);
        const bool meas = metro && energy_needed(sweep);
        if (resident && metro) {
            // one launch: the pending OR sweeps, the Metropolis sweep and (if consumed) E/M of every local replica
            enqueue_resident(h, 1, pending_or, 1, cone, 0, meas ? mine : nullptr, 1, adapt);
            pending_or = 0;
        } else if (fold && metro) {
            const csmc_handle::CycleGraph *cg = nullptr;
            rc = build_cycle_graph(h, pending_or, 1, &cg); if (rc) return rc;
            CK(cudaGraphLaunch(cg->exec, h->stream));
            h->launches += cg->launches;
            h->metro_ctr++;
            pending_or = 0;
        } else if (metro || probe || sweep + 1 == sweep_end) {
            if (resident) { if (pending_or) enqueue_resident(h, 1, pending_or, 0, 0, 0, nullptr, 0); }
            else { rc = enqueue_or_block(h, pending_or); if (rc) return rc; }
            pending_or = 0;
        }
        if (metro) {                                                            // :302-305
            if (!resident) {
                if (!fold) enqueue_metropolis(h, cone != 0);
                if (adapt) { k_adapt_sigma<<<(h->R + 127) / 128, 128, 0, h->stream>>>(h->d_sigma, h->d_acc, h->d_acc_prev, (double)h->hm.N, h->R); h->launches++; }
                if (meas) enqueue_measure(h, mine, true);
            }
            if (meas) {
                rc = gather(); if (rc) return rc;
                k_pt_update<<<nb, 128, 0, h->stream>>>(st, 1); h->launches++;
            }
            if (h->n_slots > 1 && sweep % p->swap_rate == 0) {                  // :308-349
                const long long k = sweep / p->swap_rate;
                k_pt_exchange<<<1, 128, 0, h->stream>>>(st, (int)(k % 2), (unsigned long long)k, h->seed); h->launches++;
            }
        }
        if (probe) {                                                            // :353,368-370
            if (!metro) {   // fresh magnetisation, energy of the last Metropolis sweep
                if (resident) enqueue_resident(h, 0, 0, 0, 0, 0, mine, 0);
                else enqueue_measure(h, mine, false);
                rc = gather(); if (rc) return rc;
            }
            k_pt_probe<<<nb, 128, 0, h->stream>>>(st, h->d_series_E, h->d_series_M, h->n_probes); h->launches++;
            h->n_probes++;
            if (h->d_ssf_sum) {   // mc.corr: structure factor of every local replica, summed per temperature slot (:371-375)
                for (int r = 0; r < h->R; ++r) ssf_enqueue(h, r, h->d_ssf_sum, 1, h->d_slot_of_rep);
                h->ssf_probes++;
            }
        }
        if ((sweep & 63) == 63) CK(cudaGetLastError());
    }
    // attribute the Metropolis acceptance counts since the last exchange to the current slots
    if (resident) enqueue_resident(h, 0, 0, 0, 0, 0, mine, 0);
    else enqueue_measure(h, mine, false);
    rc = gather(); if (rc) return rc;
    k_pt_update<<<nb, 128, 0, h->stream>>>(st, 0); h->launches++;
    return finish(h);
}

int32_t csmc_pt_exchange(csmc_handle *h, int32_t parity, int32_t *accepted_pairs) {
    NEED(h);
    if (h->n_slots == 0) return fail(h, CSMC_ERR_INVALID, "csmc_pt_init has not been called");
    CK(cudaSetDevice(h->device));
    int rc = enqueue_measure_all(h, true); if (rc) return rc;
    PtState st = pt_state(h);
    k_pt_update<<<(h->n_slots + 127) / 128, 128, 0, h->stream>>>(st, 1); h->launches++;
    // every call draws fresh uniforms: the Philox counter is a per-handle call index (offset past the range
    // csmc_pt_run uses, sweep / swap_rate), `parity` only selects the pairing
    k_pt_exchange<<<1, 128, 0, h->stream>>>(st, parity & 1, (1ULL << 48) + h->exchange_calls++, h->seed); h->launches++;
    if (accepted_pairs) CK(cudaMemcpyAsync(accepted_pairs, h->d_accepted_pairs, sizeof(int) * h->n_slots, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_pt_get_slots(csmc_handle *h, int32_t *slot_of_replica) {
    NEED(h); NEEDARG(h, slot_of_replica);
    if (h->n_slots == 0) return fail(h, CSMC_ERR_INVALID, "csmc_pt_init has not been called");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(slot_of_replica, h->d_slot_of_rep, sizeof(int) * h->n_slots, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

int32_t csmc_pt_get_series(csmc_handle *h, int64_t *n_probes, double *E, double *M) {
    NEED(h); NEEDARG(h, n_probes);
    if (h->n_slots == 0) return fail(h, CSMC_ERR_INVALID, "csmc_pt_init has not been called");
    CK(cudaSetDevice(h->device));
    const long long n = std::min<long long>(*n_probes, h->n_probes);
    if (E && n > 0) CK(cudaMemcpyAsync(E, h->d_series_E, sizeof(double) * n * h->n_slots, cudaMemcpyDeviceToHost, h->stream));
    if (M && n > 0) CK(cudaMemcpyAsync(M, h->d_series_M, sizeof(double) * n * h->n_slots, cudaMemcpyDeviceToHost, h->stream));
    *n_probes = h->n_probes;
    return finish(h);
}

int32_t csmc_pt_get_stats(csmc_handle *h, double *accepted_local, double *exchanges) {
    NEED(h);
    if (h->n_slots == 0) return fail(h, CSMC_ERR_INVALID, "csmc_pt_init has not been called");
    CK(cudaSetDevice(h->device));
    if (accepted_local) CK(cudaMemcpyAsync(accepted_local, h->d_acc_slot, sizeof(double) * h->n_slots, cudaMemcpyDeviceToHost, h->stream));
    if (exchanges) CK(cudaMemcpyAsync(exchanges, h->d_exch_slot, sizeof(double) * h->n_slots, cudaMemcpyDeviceToHost, h->stream));
    return finish(h);
}

}  // extern "C"
