# MonteCarlo object and the three drivers.  Schedules follow src/monte_carlo.jl:157-398 of the
# reference; sweeps run on the GPU.  MPI is optional: with MPI.jl loaded and initialised, each rank is
# one process per GPU holding `length(T)` temperature slots (T may be a scalar, as in the reference).

struct SimulationParameters
    t_thermalization::Int64
    t_deterministic::Int64
    t_measurement::Int64
    probe_rate::Int64
    swap_rate::Int64
    overrelaxation_rate::Int64
    report_interval::Int64
    checkpoint_rate::Int64
end

function MCParamsBuffer(dict::Dict{String,Int64})::SimulationParameters
    names = String.(collect(fieldnames(SimulationParameters)))
    defaults = (1, 1, 1, 1, 1, 10, 0, 0)
    for (k, v) in zip(names, defaults)
        haskey(dict, k) || (dict[k] = v)
    end
    for k in keys(dict)
        k in names || @warn "'$k' not a valid MC parameter; ignoring"
    end
    return SimulationParameters((dict[k] for k in names)...)
end

mutable struct MonteCarlo
    T::Float64
    temperatures::Vector{Float64}
    parameters::SimulationParameters
    lattice::Lattice
    observables::Observables
    observables_all::Vector{Observables}
    replica_spins::Vector{Matrix{Float64}}
    lambda::Float64
    weight::Float64
    constraint::Function
    outpath::String
    outdir::String
    outprefix::String
    sigma::Real
    sigma0::Real
    corr::Bool
    momentum_vectors::Array{Float64,2}
    seed::UInt64
    engine::Union{Nothing,Engine}
    rank::Int
    comm_size::Int
end

# hooks a host program overrides when it runs under MPI.jl (kept as plain functions so the package
# loads without MPI):  comm_rank(), comm_size(), allgather_temperatures(T), bcast_bytes(v), allgather_sigma(s), barrier()
comm_rank() = 0
comm_size() = 1
allgather_temperatures(T::Vector{Float64}) = T
bcast_bytes(v::Vector{UInt8}) = v
allgather_sigma(sig::Vector{Float64}) = sig          # with MPI: element-wise merge of the ranks' non-NaN entries
barrier() = nothing

function MonteCarlo(T::Union{Float64,Vector{Float64}}, lattice::Lattice{D}, parameters::Dict{String,Int64};
                    constraint::Function=x -> 0.0, weight::Float64=0.0, outpath::String="", outprefix::String="configuration",
                    inparams::Dict{String,<:Any}=Dict{String,Any}(), overwrite::Bool=true, sigma0::Real=60,
                    corr::Bool=false, ks::Matrix{Float64}=Matrix{Float64}(undef, D, 0), seed::UInt64=rand(UInt64) >> 1) where {D}
    corr && error("equal-time structure factor (corr=true) is not part of this drop-in")
    temps = T isa Float64 ? [T] : copy(T)
    lat = deepcopy(lattice)
    lat.engine = nothing
    mc = MonteCarlo(temps[1], temps, MCParamsBuffer(parameters), lat, Observables(), [Observables() for _ in temps],
                    [r == 1 ? lat.spins : copy(lat.spins) for r in eachindex(temps)], 0.0, weight, constraint, "", outpath, outprefix,
                    sigma0, sigma0, corr, ks, seed, nothing, comm_rank(), comm_size())
    mc.observables = mc.observables_all[1]
    if length(outpath) > 0
        mc.rank == 0 && !isdir(outpath) && mkdir(outpath)
        barrier()
        # files are named by global temperature slot (== rank with one temperature per process, as in the
        # reference); every process holds the same number of slots
        base = mc.rank * length(temps)
        mc.outpath = string(outpath, outprefix, "_", base, ".h5")
        paramsfile = string(outpath, outprefix, ".h5.params")
        if mc.rank == 0 && !isfile(paramsfile) && overwrite
            create_params_file(mc, paramsfile)
            isempty(inparams) || write_attributes(paramsfile, inparams)
        end
        barrier()
        for r in eachindex(temps)
            path = string(outpath, outprefix, "_", base + r - 1, ".h5")
            if !isfile(path) && overwrite
                println("Creating new file $(basename(path)) for output on rank $(mc.rank)")
                initialize_hdf5(path, temps[r], paramsfile, mc.replica_spins[r], lat.site_positions)
            end
        end
    end
    return mc
end

function device!(mc::MonteCarlo; replica_base=0)
    R = length(mc.temperatures)
    if mc.engine === nothing || mc.engine.n_replicas != R || mc.engine.replica_base != replica_base
        model, buffers = pack_model(mc.lattice.unit_cell, mc.lattice.shape, mc.lattice.S, mc.lattice.bc)
        mc.engine = create_engine(model, buffers; n_replicas=R, seed=mc.seed, replica_base=replica_base,
                                  device=parse(Int, get(ENV, "LOCAL_RANK", "0")))
    end
    return mc.engine
end
function upload!(mc::MonteCarlo; kw...)
    e = device!(mc; kw...)
    mc.replica_spins[1] = mc.lattice.spins
    for (r, s) in enumerate(mc.replica_spins)
        set_spins!(e, s, r - 1)
    end
    return e
end
function download!(mc::MonteCarlo)
    for (r, s) in enumerate(mc.replica_spins)
        get_spins!(mc.engine, s, r - 1)
    end
    mc.lattice.spins = mc.replica_spins[1]
end

# --- the `alg` plug-in seam (src/metropolis.jl:181-199) ------------------------------------------------------
metropolis!(mc::MonteCarlo, T::Float64)::Float64 = metropolis_sweeps!(device!(mc), fill(T, length(mc.temperatures)))[1]
function metropolis_adaptive!(mc::MonteCarlo, T::Float64)::Float64
    sig = fill(Float64(mc.sigma), length(mc.temperatures))
    acc = metropolis_cone_sweeps!(device!(mc), fill(T, length(sig)), sig, true)
    mc.sigma = sig[1]
    return acc[1]
end
metropolis_fixed_cone!(mc::MonteCarlo, T::Float64)::Float64 =
    metropolis_cone_sweeps!(device!(mc), fill(T, length(mc.temperatures)), fill(Float64(mc.sigma), length(mc.temperatures)), false)[1]
unsupported_constraint!(mc::MonteCarlo, T::Float64)::Float64 =
    error("MetropolisConstraint variants are not part of the device path (the reference's metropolis_constraint! calls an undefined e_diff)")

const AlgWrapper = FunctionWrapper{Float64,Tuple{MonteCarlo,Float64}}
Metropolis() = AlgWrapper(metropolis!)
MetropolisAdaptive() = AlgWrapper(metropolis_adaptive!)
MetropolisFixedCone() = AlgWrapper(metropolis_fixed_cone!)
MetropolisConstraint() = AlgWrapper(unsupported_constraint!)
MetropolisConstraintAdaptive() = AlgWrapper(unsupported_constraint!)
is_plain_metropolis(alg::AlgWrapper) = alg.obj[] === metropolis!

# --- simulated annealing (src/monte_carlo.jl:157-190) ---------------------------------------------------------
function simulated_annealing!(mc::MonteCarlo, schedule::Function, T0::Float64=1.0; alg::AlgWrapper=Metropolis())
    p = mc.parameters
    T, time = T0, 1
    out = length(mc.outpath) > 0
    accept_total = p.t_thermalization * mc.lattice.size
    p.overrelaxation_rate != 0 && (accept_total /= p.overrelaxation_rate)
    e = upload!(mc)
    while T > mc.T
        R = 0.0
        mc.sigma = mc.sigma0
        if is_plain_metropolis(alg)
            R = anneal_temperature!(e, fill(T, e.n_replicas), p.t_thermalization, p.overrelaxation_rate)[1]
        elseif alg.obj[] === metropolis_adaptive! || alg.obj[] === metropolis_fixed_cone!
            sig = fill(Float64(mc.sigma), e.n_replicas)
            R = anneal_temperature_cone!(e, fill(T, e.n_replicas), sig, alg.obj[] === metropolis_adaptive!, p.t_thermalization, p.overrelaxation_rate)[1]
            mc.sigma = sig[1]
        else
            for t in 1:(p.t_thermalization-1)
                if p.overrelaxation_rate != 0
                    overrelax!(e)
                    t % p.overrelaxation_rate == 0 && (R += alg(mc, T))
                else
                    R += alg(mc, T)
                end
            end
        end
        println("Acceptance rate at T=$T: $(round(R/accept_total*100, digits=5)) % ")
        T = schedule(time)
        time += 1
        if out
            download!(mc)
            write_MC_checkpoint(mc)
        end
    end
    download!(mc)
end

# --- deterministic updates (src/monte_carlo.jl:201-213) as colour-ordered full sweeps -----------------------
function deterministic_updates!(mc::MonteCarlo)
    n_sweeps = cld(max(mc.parameters.t_deterministic - 1, 0), mc.lattice.size)
    e = upload!(mc)
    done = 0
    while done < n_sweeps
        k = min(4096, n_sweeps - done)
        deterministic!(e, k)
        done += k
    end
    download!(mc)
end

# --- parallel tempering (src/monte_carlo.jl:235-398) ------------------------------------------------------------
function parallel_tempering!(mc::MonteCarlo, saveIC::Vector{Int64}=Int64[]; alg::AlgWrapper=Metropolis())
    algid = alg.obj[] === metropolis! ? 0 : alg.obj[] === metropolis_adaptive! ? 1 : alg.obj[] === metropolis_fixed_cone! ? 2 :
            error("parallel_tempering! on the device supports Metropolis(), MetropolisAdaptive() and MetropolisFixedCone()")
    p = mc.parameters
    rank, nranks = comm_rank(), comm_size()
    R = length(mc.temperatures)
    T_all = allgather_temperatures(mc.temperatures)
    n_slots = length(T_all)
    n_slots == 1 && @warn "a single temperature slot; no replica exchanges will occur!"
    base = rank * R
    out = length(mc.outdir) > 0
    if nranks > 1
        # exchange decisions are drawn redundantly on every rank from the shared Philox stream: one seed per job
        seed0 = reinterpret(UInt64, bcast_bytes(collect(reinterpret(UInt8, [mc.seed]))))[1]
        if seed0 != mc.seed
            mc.seed = seed0
            mc.engine = nothing
        end
    end
    e = upload!(mc; replica_base=base)
    if nranks > 1
        id = rank == 0 ? comm_unique_id() : zeros(UInt8, 128)
        comm_init!(e, nranks, rank, bcast_bytes(id))
    end
    pt_init!(e, T_all)
    set_sigma!(e, fill(Float64(mc.sigma), R))
    slotfile(s) = string(mc.outdir, mc.outprefix, "_", s, ".h5")
    if out && !isempty(saveIC)                                   # src/monte_carlo.jl:278-283
        if rank == 0
            for s in saveIC
                d = string(mc.outdir, "IC_", s)
                isdir(d) || (println("Initializing IC collection on rank $s"); mkpath(d))
            end
        end
        barrier()
    end
    acc_prev, exch_prev, t_prev = zeros(n_slots), zeros(n_slots), 0   # progress report state (src/helper.jl:26)
    rank == 0 && @printf("Running sweeps on %s.\n", Dates.format(Dates.now(), "dd u yyyy HH:MM:SS"))
    total = p.t_thermalization + p.t_measurement
    cp = CsmcPtParams(p.t_thermalization, p.t_measurement, p.probe_rate, p.swap_rate, p.overrelaxation_rate, algid)
    sweep = 0
    buf = similar(mc.lattice.spins)
    while sweep < total
        nxt = total
        if p.checkpoint_rate > 0
            c = cld(max(sweep, p.t_thermalization), p.checkpoint_rate) * p.checkpoint_rate
            c < total && (nxt = min(nxt, c + 1))
        end
        p.report_interval > 0 && (nxt = min(nxt, (sweep ÷ p.report_interval + 1) * p.report_interval))
        pt_run!(e, cp, sweep, nxt)
        last = nxt - 1
        if out && p.checkpoint_rate > 0 && last >= p.t_thermalization && last % p.checkpoint_rate == 0
            slots = pt_slots(e, n_slots)
            for r in 1:R
                get_spins!(e, buf, r - 1)
                s = slots[base+r]
                write_spins(slotfile(s), buf)
                if s in saveIC
                    step = (last - p.t_thermalization) ÷ p.checkpoint_rate
                    write_initial_configuration(string(mc.outdir, "IC_", s, "/IC_", step, ".h5"), T_all[s+1], slotfile(s), buf)
                end
            end
        end
        sweep = nxt
        if p.report_interval > 0 && sweep % p.report_interval == 0
            # print_runtime_statistics!, src/helper.jl:25-78: rates since the previous report, per temperature slot
            # (== per MPI rank in the reference); the device keeps the counters per slot
            acc, exch = pt_stats(e, n_slots)
            sig = pt_sigma_by_slot(e, n_slots, base, R)
            if rank == 0
                dt = sweep - t_prev
                rate = max(p.overrelaxation_rate, 1)
                str = @sprintf("Sweep %d / %d (%.1f%%)\n", sweep, total, 100.0 * sweep / total)
                str *= @sprintf("\t\tthermalized : %s\n", sweep >= p.t_thermalization ? "YES" : "NO")
                for k in 1:n_slots
                    local_rate = (acc[k] - acc_prev[k]) / (dt * mc.lattice.size / rate) * 100.0          # :30-31
                    if n_slots == 1
                        str *= @sprintf("\t\tupdate acceptance rate : %.2f%%\tsigma : %.2f\n", local_rate, sig[k])
                    else
                        attempted = (k == 1 || k == n_slots) ? dt / p.swap_rate / 2.0 : dt / p.swap_rate   # :39
                        str *= @sprintf("\t\tsimulation %d update acceptance rate : %.2f%%\tsigma : %.2f\n", k - 1, local_rate, sig[k])
                        str *= @sprintf("\t\tsimulation %d replica exchange acceptance rate : %.2f%%\n", k - 1,
                                        (exch[k] - exch_prev[k]) / attempted * 100.0)
                    end
                end
                print(str * "\n")
            end
            acc_prev, exch_prev, t_prev = copy(acc), copy(exch), sweep
        end
    end
    mc.sigma = get_sigma(e)[1]
    E, M = pt_series(e, n_slots)
    for r in 1:R, k in 1:size(E, 2)
        update_observables!(mc.observables_all[r], E[base+r, k], M[base+r, k])
    end
    download!(mc)          # configurations stay with their replicas; replica r sits in slot slots[base+r]
    if out
        slots = pt_slots(e, n_slots)
        for r in 1:R
            # spins: written by the process that holds the slot's replica
            write_spins(slotfile(slots[base+r]), mc.replica_spins[r])
        end
        barrier()
        for r in 1:R
            # observables: accumulated per temperature slot; slot base+r-1 was measured on this process
            write_observables(slotfile(base + r - 1), mc.observables_all[r], T_all[base+r], mc.lattice.size)
        end
    end
    # the reference leaves at every rank the configuration that sits at that rank's temperature (it swaps
    # configurations, src/monte_carlo.jl:336-347): re-order the local copies by slot where the slot is local
    let slots = pt_slots(e, n_slots), byslot = Dict(slots[base+r] => mc.replica_spins[r] for r in 1:R)
        if all(haskey(byslot, base + r - 1) for r in 1:R)      # single process: every slot is local
            mc.replica_spins = [byslot[base+r-1] for r in 1:R]
            mc.lattice.spins = mc.replica_spins[1]
        end
    end
    rank == 0 && @printf("Simulation finished on %s.\n", Dates.format(Dates.now(), "dd u yyyy HH:MM:SS"))
    return
end
