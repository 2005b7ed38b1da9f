# UnitCell, its builders and the Bravais presets.  API contract: src/unit_cell.jl:3-75 and
# src/bravais.jl:1-51 of the reference (names, argument order, strict argument types, silent
# dropping of all-zero couplings).  Couplings are kept as plain arrays; the library receives them as
# packed buffers (include/csmc.h: matrices row-major, tensors column-major).

struct UnitCell{D}
    lattice_vectors::NTuple{D,Vector{Float64}}
    basis::Vector{Vector{Float64}}
    field::Vector{Tuple{Int64,Vector{Float64}}}
    onsite::Vector{Tuple{Int64,Matrix{Float64}}}
    bilinear::Vector{Tuple{Int64,Int64,Matrix{Float64},NTuple{D,Int64}}}
    cubic::Vector{Tuple{Int64,Int64,Int64,Array{Float64,3},NTuple{D,Int64},NTuple{D,Int64}}}
    quartic::Vector{Tuple{Int64,Int64,Int64,Int64,Array{Float64,4},NTuple{D,Int64},NTuple{D,Int64},NTuple{D,Int64}}}
    function UnitCell(as...)
        D = length(as)
        new{D}(Tuple(Vector{Float64}(a) for a in as), Vector{Float64}[], [], [], [], [], [])
    end
end

nooffset(::UnitCell{D}) where {D} = ntuple(_ -> 0, D)

addBasisSite!(uc::UnitCell, site::Vector{Float64}) = push!(uc.basis, site)
addZeemanCoupling!(uc::UnitCell, b1::Int64, h::Vector{Float64}) = push!(uc.field, (b1, h))
addOnSite!(uc::UnitCell, b1::Int64, M::Matrix{Float64}) = iszero(M) ? nothing : push!(uc.onsite, (b1, M))
addBilinear!(uc::UnitCell{D}, b1::Int64, b2::Int64, M::Matrix{Float64}, offset::NTuple{D,Int64}=nooffset(uc)) where {D} =
    iszero(M) ? nothing : push!(uc.bilinear, (b1, b2, M, offset))
addCubic!(uc::UnitCell{D}, b1::Int64, b2::Int64, b3::Int64, M::Array{Float64,3},
          o2::NTuple{D,Int64}=nooffset(uc), o3::NTuple{D,Int64}=nooffset(uc)) where {D} =
    iszero(M) ? nothing : push!(uc.cubic, (b1, b2, b3, M, o2, o3))
addQuartic!(uc::UnitCell{D}, b1::Int64, b2::Int64, b3::Int64, b4::Int64, M::Array{Float64,4},
            o2::NTuple{D,Int64}=nooffset(uc), o3::NTuple{D,Int64}=nooffset(uc), o4::NTuple{D,Int64}=nooffset(uc)) where {D} =
    iszero(M) ? nothing : push!(uc.quartic, (b1, b2, b3, b4, M, o2, o3, o4))

Triangular() = UnitCell([1.0, 0.0], [cos(pi / 3), sin(pi / 3)])
Square() = UnitCell([1.0, 0.0], [0.0, 1.0])
FCC() = UnitCell(0.5 .* [0.0, 1, 1], 0.5 .* [1.0, 0, 1], 0.5 .* [1.0, 1, 0])
const _tetra = ([1.0, 1, 1], [1.0, -1, -1], [-1.0, 1, -1], [-1.0, -1, 1])
function Pyrochlore()
    uc = FCC()
    foreach(v -> addBasisSite!(uc, v ./ 8), _tetra)
    return uc
end
function BreathingPyrochlore(a::Float64=1.01)
    uc = FCC()
    foreach(v -> addBasisSite!(uc, a .* v ./ 8), _tetra)
    return uc
end
function Honeycomb()
    uc = Triangular()
    addBasisSite!(uc, [0.0, 0.0])
    addBasisSite!(uc, [0.0, 1.0] ./ sqrt(3))
    return uc
end

"Per-basis Zeeman vectors and on-site matrices, resolved with the rule of src/lattice.jl:117-140."
function resolve_site_terms(uc::UnitCell)
    nb = length(uc.basis)
    field = zeros(3, nb)
    onsite = zeros(9, nb)          # row-major 3x3 per column
    fidx = first.(uc.field)
    oidx = first.(uc.onsite)
    for i in 1:nb
        if i in fidx
            field[:, fidx[i]] .= uc.field[i][2]
        end
        if i in oidx
            onsite[:, oidx[i]] .= vec(permutedims(uc.onsite[i][2]))
        end
    end
    return field, onsite
end

"Packs a UnitCell + lattice description into the `csmc_model` struct; returns (model, buffers)."
function pack_model(uc::UnitCell{D}, shape::NTuple{D,Int64}, S::Real, bc::String) where {D}
    bc in ("periodic", "open") || error("Invalid boundary condition option")
    D <= 3 || error("at most 3 lattice dimensions are supported")
    field, onsite = resolve_site_terms(uc)
    i32(v) = Int32.(collect(v))
    bilB = isempty(uc.bilinear) ? Int32[0, 0] : i32(Iterators.flatten((t[1], t[2]) for t in uc.bilinear))
    bilO = isempty(uc.bilinear) ? zeros(Int32, D) : i32(Iterators.flatten(t[4] for t in uc.bilinear))
    bilJ = isempty(uc.bilinear) ? zeros(9) : collect(Iterators.flatten(vec(permutedims(t[3])) for t in uc.bilinear))
    cubB = isempty(uc.cubic) ? zeros(Int32, 3) : i32(Iterators.flatten((t[1], t[2], t[3]) for t in uc.cubic))
    cubO = isempty(uc.cubic) ? zeros(Int32, 2D) : i32(Iterators.flatten((t[5]..., t[6]...) for t in uc.cubic))
    cubT = isempty(uc.cubic) ? zeros(27) : collect(Iterators.flatten(vec(t[4]) for t in uc.cubic))
    quaB = isempty(uc.quartic) ? zeros(Int32, 4) : i32(Iterators.flatten((t[1], t[2], t[3], t[4]) for t in uc.quartic))
    quaO = isempty(uc.quartic) ? zeros(Int32, 3D) : i32(Iterators.flatten((t[6]..., t[7]..., t[8]...) for t in uc.quartic))
    quaT = isempty(uc.quartic) ? zeros(81) : collect(Iterators.flatten(vec(t[5]) for t in uc.quartic))
    buffers = (field, onsite, bilB, bilO, bilJ, cubB, cubO, cubT, quaB, quaO, quaT)
    shp = ntuple(d -> d <= D ? Int32(shape[d]) : Int32(1), 3)
    model = CsmcModel(D, shp, length(uc.basis), bc == "periodic" ? 1 : 0, Float64(S),
                      pointer(field), pointer(onsite),
                      length(uc.bilinear), pointer(bilB), pointer(bilO), pointer(bilJ),
                      length(uc.cubic), pointer(cubB), pointer(cubO), pointer(cubT),
                      length(uc.quartic), pointer(quaB), pointer(quaO), pointer(quaT))
    return model, buffers
end
