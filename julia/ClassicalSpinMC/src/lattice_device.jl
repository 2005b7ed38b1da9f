# Lattice: host-visible state (`spins`, 3 x N) plus a lazily created device engine.  Public field
# names follow src/lattice.jl:4-24; the per-site tables are derived in closed form by the library
# (csmc_reference_tables) on first access instead of the reference's O(N^2) findfirst scans.

mutable struct Lattice{D}
    S::Real
    spins::Array{Float64,2}
    unit_cell::UnitCell{D}
    bc::String
    size::Int64
    shape::NTuple{D,Int64}
    site_positions::Array{Float64,2}
    engine::Union{Nothing,Engine}
end

function random_spin_orientation(S::Real, rng=Random.GLOBAL_RNG)::NTuple{3,Float64}
    phi = 2.0 * pi * rand(rng)
    z = 2.0 * rand(rng) - 1.0
    r = sqrt(1.0 - z * z)
    return S .* (r * cos(phi), r * sin(phi), z)
end

set_spin!(spins::Array{Float64,2}, s::NTuple{3,Float64}, point::Int64) = (spins[1, point], spins[2, point], spins[3, point]) = s
get_spin(spins::Array{Float64,2}, point::Int64) = (spins[1, point], spins[2, point], spins[3, point])

function site_positions(uc::UnitCell{D}, shape::NTuple{D,Int64}) where {D}
    nb = length(uc.basis)
    pos = Array{Float64,2}(undef, D, prod(shape) * nb)
    p = 0
    for b in 1:nb, cell in Iterators.product(ntuple(d -> 1:shape[D+1-d], D)...)   # last lattice index fastest
        idx = reverse(cell)
        p += 1
        pos[:, p] = sum((idx[d] - 1) .* uc.lattice_vectors[d] for d in 1:D) .+ uc.basis[b]
    end
    return pos
end

function Lattice(shape::NTuple{D,Int64}, uc::UnitCell{D}, S::Real=1 / 2; bc::String="periodic", initialCondition::Symbol=:random) where {D}
    isempty(uc.basis) && addBasisSite!(uc, zeros(Float64, D))
    bc in ("periodic", "open") || error("Invalid boundary condition option")
    N = prod(shape) * length(uc.basis)
    spins = Array{Float64,2}(undef, 3, N)
    if initialCondition == :random
        for i in 1:N
            set_spin!(spins, random_spin_orientation(S), i)
        end
    elseif initialCondition == :fm
        s = random_spin_orientation(S)
        for i in 1:N
            set_spin!(spins, s, i)
        end
    end
    return Lattice{D}(S, spins, uc, bc, N, shape, site_positions(uc, shape), nothing)
end

function engine!(lat::Lattice; kw...)
    if lat.engine === nothing
        model, buffers = pack_model(lat.unit_cell, lat.shape, lat.S, lat.bc)
        lat.engine = create_engine(model, buffers; kw...)
    end
    return lat.engine
end
upload!(lat::Lattice) = set_spins!(engine!(lat), lat.spins)
download!(lat::Lattice) = get_spins!(engine!(lat), lat.spins)

"Reference-layout neighbour tables (lat.bilinear_sites etc.), csmc_reference_tables."
function neighbour_tables(lat::Lattice)
    model, buffers = pack_model(lat.unit_cell, lat.shape, lat.S, lat.bc)
    n2, n3, n4 = length(lat.unit_cell.bilinear), length(lat.unit_cell.cubic), length(lat.unit_cell.quartic)
    bil = zeros(Int64, n2, lat.size); cub = zeros(Int64, 2, n3, lat.size); quar = zeros(Int64, 3, n4, lat.size)
    GC.@preserve buffers check(nothing, ccall((:csmc_reference_tables, libcsmc), Int32,
        (Ref{CsmcModel}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), Ref(model), bil, cub, quar))
    return bil, cub, quar
end
function Base.getproperty(lat::Lattice, name::Symbol)
    name === :bilinear_sites && return [Tuple(c) for c in eachcol(neighbour_tables(lat)[1])]
    name === :cubic_sites && return [Tuple(Tuple.(eachcol(c))) for c in eachslice(neighbour_tables(lat)[2], dims=3)]
    name === :quartic_sites && return [Tuple(Tuple.(eachcol(c))) for c in eachslice(neighbour_tables(lat)[3], dims=3)]
    return getfield(lat, name)
end

# Hamiltonian evaluation: bodies of src/hamiltonian.jl:3-136 and src/observables.jl:12-18 on the GPU
function get_local_field(lat::Lattice, point::Int64)
    upload!(lat)
    return local_field(lat.engine, point)
end
function total_energy(lat::Lattice)
    upload!(lat)
    return total_energies(lat.engine)[1]
end
energy_density(lat::Lattice) = total_energy(lat) / lat.size
function get_magnetization(lat::Lattice)::Float64
    upload!(lat)
    return norm(magnetization_vectors(lat.engine)[:, 1])
end

# equal-time structure factor, body of src/spin_correlations.jl:6-43 on the GPU
function compute_equal_time_correlations(lat::Lattice{D}, ks::Array{Float64,2}) where {D}
    upload!(lat)
    A = reduce(hcat, lat.unit_cell.lattice_vectors)                  # D x D, column d = a_d
    B = permutedims(reduce(hcat, lat.unit_cell.basis))               # n_basis x D; the ABI wants row-major n_basis x D
    return structure_factor(lat.engine, A, collect(transpose(B)), ks)  # column-major D x n_basis == row-major n_basis x D
end
