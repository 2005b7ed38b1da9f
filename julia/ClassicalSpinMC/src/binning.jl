# Observables with logarithmic binning and first-order error propagation: a restatement of the parts
# of BinningAnalysis.jl 0.6.1 that src/observables.jl:32-63 and src/hdf5.jl:220-227 use (kept local so
# the drop-in has no dependency on it; swap in `BinningAnalysis.ErrorPropagator` if preferred).

const NLEVELS = 32

mutable struct LogBinnedPairs
    sums::Matrix{Float64}            # 2 x levels
    prods::Array{Float64,3}          # 2 x 2 x levels
    count::Vector{Int64}
    held::Matrix{Float64}
    full::Vector{Bool}
    LogBinnedPairs() = new(zeros(2, NLEVELS), zeros(2, 2, NLEVELS), zeros(Int64, NLEVELS), zeros(2, NLEVELS), fill(false, NLEVELS))
end

function Base.push!(b::LogBinnedPairs, x1::Float64, x2::Float64)
    x = [x1, x2]
    for lvl in 1:NLEVELS
        b.sums[:, lvl] .+= x
        b.prods[:, :, lvl] .+= x * x'
        b.count[lvl] += 1
        if !b.full[lvl]
            b.held[:, lvl] .= x
            b.full[lvl] = true
            return b
        end
        b.full[lvl] = false
        x = 0.5 .* (b.held[:, lvl] .+ x)
    end
    return b
end

means(b::LogBinnedPairs, lvl=1) = b.sums[:, lvl] ./ max(b.count[lvl], 1)
reliable_level(b::LogBinnedPairs) = something(findlast(>=(32), b.count), 1)
function covariance(b::LogBinnedPairs, lvl)
    n = b.count[lvl]
    return (b.prods[:, :, lvl] .- b.sums[:, lvl] * b.sums[:, lvl]' ./ n) ./ (n - 1)
end
propagated_error(b::LogBinnedPairs, grad::Vector{Float64}, lvl=reliable_level(b)) =
    sqrt(abs(dot(grad, covariance(b, lvl) * grad) / b.count[lvl]))
std_error(b::LogBinnedPairs, i::Int, lvl=reliable_level(b)) = sqrt(abs(covariance(b, lvl)[i, i]) / b.count[lvl])

mutable struct Observables
    energy::LogBinnedPairs
    magnetization::LogBinnedPairs
    Observables() = new(LogBinnedPairs(), LogBinnedPairs())
end

function update_observables!(obs::Observables, energy::Float64, magnetization::Float64)
    push!(obs.energy, energy, energy^2)
    push!(obs.magnetization, magnetization, magnetization^2)
end

function specific_heat(obs::Observables, T, N)
    e = means(obs.energy)
    heat = (e[2] - e[1]^2) / (T^2 * N)
    return heat, propagated_error(obs.energy, [-2.0 * e[1] / (T^2 * N), 1 / (T^2 * N)])
end
function susceptibility(obs::Observables, T, N)
    m = means(obs.magnetization)
    chi = (m[2] - m[1]^2) / (T * N)
    return chi, propagated_error(obs.magnetization, [-2.0 * m[1] / (T * N), 1 / (T * N)])
end

# the reference's signatures (src/observables.jl:36-63): observables, temperature and size taken from the driver state
specific_heat(mc) = specific_heat(mc.observables, mc.T, mc.lattice.size)
susceptibility(mc) = susceptibility(mc.observables, mc.T, mc.lattice.size)
