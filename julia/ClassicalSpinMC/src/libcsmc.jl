# ccall layer over include/csmc.h.  Every wrapper names the C export it binds.

const libcsmc = get(ENV, "CSMC_LIB", "libcsmc.so")

struct CsmcModel
    dim::Int32
    shape::NTuple{3,Int32}
    n_basis::Int32
    periodic::Int32
    S::Float64
    field::Ptr{Float64}
    onsite::Ptr{Float64}
    n_bilinear::Int32
    bil_basis::Ptr{Int32}
    bil_offset::Ptr{Int32}
    bil_matrix::Ptr{Float64}
    n_cubic::Int32
    cub_basis::Ptr{Int32}
    cub_offset::Ptr{Int32}
    cub_tensor::Ptr{Float64}
    n_quartic::Int32
    quar_basis::Ptr{Int32}
    quar_offset::Ptr{Int32}
    quar_tensor::Ptr{Float64}
end

struct CsmcOpts
    device::Int32
    n_replicas::Int32
    seed::UInt64
    stream::Ptr{Cvoid}
    replica_base::Int32
    flags::Int32
end

struct CsmcPtParams
    t_thermalization::Int64
    t_measurement::Int64
    probe_rate::Int32
    swap_rate::Int32
    overrelaxation_rate::Int32
    algorithm::Int32      # 0 Metropolis(), 1 MetropolisAdaptive(), 2 MetropolisFixedCone()
end

"Owns a `csmc_handle*`; destroyed by the finalizer."
mutable struct Engine
    ptr::Ptr{Cvoid}
    n_sites::Int
    n_replicas::Int
    replica_base::Int
    buffers::Any            # host arrays the model struct pointed into (kept alive until create returns)
    function Engine(ptr, n_sites, n_replicas, replica_base, buffers)
        e = new(ptr, n_sites, n_replicas, replica_base, buffers)
        finalizer(x -> (x.ptr == C_NULL || ccall((:csmc_destroy, libcsmc), Int32, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), e)
        return e
    end
end

last_error(p::Ptr{Cvoid}) = unsafe_string(ccall((:csmc_last_error, libcsmc), Cstring, (Ptr{Cvoid},), p))
check(e::Engine, rc) = rc == 0 ? nothing : error("libcsmc: " * last_error(e.ptr))
check(::Nothing, rc) = rc == 0 ? nothing : error("libcsmc: " * last_error(C_NULL))

function create_engine(model::CsmcModel, buffers; n_replicas=1, seed=rand(UInt64) >> 1, device=0, replica_base=0, flags=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    opts = CsmcOpts(device, n_replicas, seed, C_NULL, replica_base, flags)
    GC.@preserve buffers begin
        check(nothing, ccall((:csmc_create, libcsmc), Int32, (Ref{CsmcModel}, Ref{CsmcOpts}, Ref{Ptr{Cvoid}}), Ref(model), Ref(opts), h))
    end
    n = Ref{Int64}(0)
    ccall((:csmc_n_sites, libcsmc), Int32, (Ptr{Cvoid}, Ref{Int64}), h[], n)
    return Engine(h[], Int(n[]), n_replicas, replica_base, nothing)
end

# --- state ---------------------------------------------------------------------------------------------
set_spins!(e::Engine, spins::Matrix{Float64}, replica=0) =
    check(e, ccall((:csmc_set_spins, libcsmc), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), e.ptr, replica, spins))
get_spins!(e::Engine, spins::Matrix{Float64}, replica=0) =
    check(e, ccall((:csmc_get_spins, libcsmc), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), e.ptr, replica, spins))

# --- Hamiltonian -----------------------------------------------------------------------------------------
function local_field(e::Engine, site::Integer, replica=0)
    out = zeros(3)
    check(e, ccall((:csmc_local_field, libcsmc), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}), e.ptr, replica, site, out))
    return (out[1], out[2], out[3])
end
function total_energies(e::Engine)
    E = zeros(e.n_replicas)
    check(e, ccall((:csmc_total_energy, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}), e.ptr, E))
    return E
end
function magnetization_vectors(e::Engine)
    M = zeros(3, e.n_replicas)
    check(e, ccall((:csmc_magnetization, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}), e.ptr, M))
    return M
end

"Suv (9 x N_k) of the current device spins; ks is D x N_k (src/spin_correlations.jl:6-43)"
function structure_factor(e::Engine, lattice_vectors::Matrix{Float64}, basis::Matrix{Float64}, ks::Matrix{Float64}, replica=0)
    Suv = zeros(9, size(ks, 2))
    check(e, ccall((:csmc_structure_factor, libcsmc), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                   e.ptr, replica, lattice_vectors, basis, ks, size(ks, 2), Suv))
    return Suv
end

# --- sweeps ------------------------------------------------------------------------------------------------
overrelax!(e::Engine, n=1) = check(e, ccall((:csmc_overrelax, libcsmc), Int32, (Ptr{Cvoid}, Int32), e.ptr, n))
deterministic!(e::Engine, n=1) = check(e, ccall((:csmc_deterministic, libcsmc), Int32, (Ptr{Cvoid}, Int32), e.ptr, n))
function metropolis_sweeps!(e::Engine, T::Vector{Float64}, n=1)
    acc = zeros(e.n_replicas)
    check(e, ccall((:csmc_metropolis, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}), e.ptr, T, n, acc))
    return acc
end
function metropolis_cone_sweeps!(e::Engine, T::Vector{Float64}, sigma::Vector{Float64}, adapt::Bool, n=1)
    acc = zeros(e.n_replicas)
    check(e, ccall((:csmc_metropolis_cone, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64}),
                   e.ptr, T, sigma, adapt ? 1 : 0, n, acc))
    return acc
end
function anneal_temperature!(e::Engine, T::Vector{Float64}, t_thermalization::Integer, rate::Integer)
    acc = zeros(e.n_replicas)
    check(e, ccall((:csmc_anneal_temperature, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Float64}),
                   e.ptr, T, t_thermalization, rate, acc))
    return acc
end

set_sigma!(e::Engine, sigma::Vector{Float64}) = check(e, ccall((:csmc_set_sigma, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}), e.ptr, sigma))
function get_sigma(e::Engine)
    s = zeros(e.n_replicas)
    check(e, ccall((:csmc_get_sigma, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}), e.ptr, s))
    return s
end

function anneal_temperature_cone!(e::Engine, T::Vector{Float64}, sigma::Vector{Float64}, adapt::Bool, t_thermalization::Integer, rate::Integer)
    acc = zeros(e.n_replicas)
    check(e, ccall((:csmc_anneal_temperature_cone, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int64, Int32, Ptr{Float64}),
                   e.ptr, T, sigma, adapt ? 1 : 0, t_thermalization, rate, acc))
    return acc
end

# --- parallel tempering --------------------------------------------------------------------------------------
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(nothing, ccall((:csmc_comm_unique_id, libcsmc), Int32, (Ptr{UInt8},), id))
    return id
end
comm_init!(e::Engine, n_ranks, rank, id::Vector{UInt8}) =
    check(e, ccall((:csmc_comm_init, libcsmc), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), e.ptr, n_ranks, rank, id))
function comm_mode(e::Engine)
    m = Ref{Int32}(0)
    check(e, ccall((:csmc_comm_mode, libcsmc), Int32, (Ptr{Cvoid}, Ref{Int32}), e.ptr, m))
    return Int(m[])
end
pt_init!(e::Engine, T_all::Vector{Float64}) =
    check(e, ccall((:csmc_pt_init, libcsmc), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), e.ptr, length(T_all), T_all))
pt_run!(e::Engine, p::CsmcPtParams, sweep_begin, sweep_end) =
    check(e, ccall((:csmc_pt_run, libcsmc), Int32, (Ptr{Cvoid}, Ref{CsmcPtParams}, Int64, Int64), e.ptr, Ref(p), sweep_begin, sweep_end))
function pt_slots(e::Engine, n_slots)
    s = zeros(Int32, n_slots)
    check(e, ccall((:csmc_pt_get_slots, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Int32}), e.ptr, s))
    return s
end
function pt_series(e::Engine, n_slots)
    n = Ref{Int64}(0)
    check(e, ccall((:csmc_pt_get_series, libcsmc), Int32, (Ptr{Cvoid}, Ref{Int64}, Ptr{Float64}, Ptr{Float64}), e.ptr, n, C_NULL, C_NULL))
    E = zeros(n_slots, n[]); M = zeros(n_slots, n[])      # column k = probe k (C layout [probe][slot])
    check(e, ccall((:csmc_pt_get_series, libcsmc), Int32, (Ptr{Cvoid}, Ref{Int64}, Ptr{Float64}, Ptr{Float64}), e.ptr, n, E, M))
    return E, M
end
function pt_stats(e::Engine, n_slots)
    a = zeros(n_slots); x = zeros(n_slots)
    check(e, ccall((:csmc_pt_get_stats, libcsmc), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), e.ptr, a, x))
    return a, x
end

# cone widths per temperature slot for the progress report (src/helper.jl:32,49-50): the local replicas' widths
# (they travel with the slot) placed at their current slots; slots held by other processes are gathered by the caller
# through `allgather_sigma` (identity in a single process)
function pt_sigma_by_slot(e::Engine, n_slots, base, R)
    sig = fill(NaN, n_slots)
    slots = pt_slots(e, n_slots)
    s = get_sigma(e)
    for r in 1:R
        sig[slots[base+r]+1] = s[r]
    end
    return allgather_sigma(sig)
end

# --- introspection added in round 2 ---------------------------------------------------------------------------
# (tiles per replica of the tile-resident persistent kernel (0: not in use), tiles along dims 0 / 1, replicas per launch,
#  shared memory per CTA, create-time probe ms: pass kernels / persistent kernel)
function persist_info(e::Engine)
    tiles = Ref{Int32}(0); nrep = Ref{Int32}(0); smem = Ref{Int32}(0)
    grid = zeros(Int32, 2); ms = zeros(Float32, 2)
    check(e, ccall((:csmc_persist_info, libcsmc), Int32, (Ptr{Cvoid}, Ref{Int32}, Ptr{Int32}, Ref{Int32}, Ref{Int32}, Ptr{Float32}),
                   e.ptr, tiles, grid, nrep, smem, ms))
    return (tiles=Int(tiles[]), grid=(Int(grid[1]), Int(grid[2])), replicas_per_launch=Int(nrep[]), smem=Int(smem[]), probe_ms=(ms[1], ms[2]))
end
# (fp64 flops per overrelaxation update counted by the code generator, algorithmic bytes per update) — roofline inputs
function kernel_costs(e::Engine)
    f = Ref{Float64}(0.0); b = Ref{Float64}(0.0)
    check(e, ccall((:csmc_kernel_costs, libcsmc), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), e.ptr, f, b))
    return f[], b[]
end
