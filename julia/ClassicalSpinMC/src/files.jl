# HDF5 files with the reference's layout (src/hdf5.jl:36-270; names pinned by util/load.py).
using HDF5

function dump_unit_cell!(fid, uc::UnitCell)
    g = create_group(fid, "unit_cell")
    g["lattice_vectors"] = reduce(hcat, uc.lattice_vectors)
    g["basis"] = reduce(vcat, transpose.(uc.basis))
    h = create_group(g, "field");    for (b, v) in uc.field;   h[string(b)] = v; end
    o = create_group(g, "onsite");   for (b, m) in uc.onsite;  o[string(b)] = m; end
    bl = create_group(g, "bilinear"); for (b1, b2, m, off) in uc.bilinear; bl[string("($b1,$b2),$off")] = m; end
    c = create_group(g, "cubic");    for (b1, b2, b3, m, o2, o3) in uc.cubic; c[string("($b1,$b2,$b3),$o2,$o3")] = m; end
    q = create_group(g, "quartic");  for (b1, b2, b3, b4, m, o2, o3, o4) in uc.quartic; q[string("($b1,$b2,$b3,$b4),$o2,$o3,$o4")] = m; end
end

function create_params_file(mc, filename)
    h5open(filename, "w") do f
        dump_unit_cell!(f, mc.lattice.unit_cell)
        l = create_group(f, "lattice")
        l["size"] = collect(mc.lattice.shape); l["S"] = mc.lattice.S; l["bc"] = mc.lattice.bc
        for name in fieldnames(SimulationParameters)
            write_attribute(f, String(name), getfield(mc.parameters, name))
        end
    end
    return filename
end

write_attributes(filename::String, d::Dict{String,<:Any}) = h5open(f -> foreach(kv -> write_attribute(f, kv[1], kv[2]), d), filename, "r+")

function read_unit_cell(fid)
    u = fid["unit_cell"]
    lv = read(u["lattice_vectors"])
    uc = UnitCell(eachcol(lv)...)
    foreach(r -> addBasisSite!(uc, collect(r)), eachrow(read(u["basis"])))
    for k in keys(u["field"]);  addZeemanCoupling!(uc, parse(Int64, k), read(u["field"][k])); end
    for k in keys(u["onsite"]); addOnSite!(uc, parse(Int64, k), read(u["onsite"][k])); end
    for k in keys(u["bilinear"]); (b, off) = eval(Meta.parse(k)); addBilinear!(uc, b[1], b[2], read(u["bilinear"][k]), off); end
    for k in keys(u["cubic"]); (b, o2, o3) = eval(Meta.parse(k)); addCubic!(uc, b[1], b[2], b[3], read(u["cubic"][k]), o2, o3); end
    for k in keys(u["quartic"]); (b, o2, o3, o4) = eval(Meta.parse(k)); addQuartic!(uc, b[1], b[2], b[3], b[4], read(u["quartic"][k]), o2, o3, o4); end
    return uc
end

function read_lattice(fid)::Lattice
    l = fid["lattice"]
    return Lattice(tuple(read(l["size"])...), read_unit_cell(fid), read(l["S"]), bc=read(l["bc"]))
end

function initialize_hdf5(path::String, T::Float64, paramsfile::String, spins, site_positions)
    h5open(path, "w") do f
        write_attribute(f, "T", T); write_attribute(f, "paramsfile", paramsfile)
        f["spins"] = spins; f["site_positions"] = site_positions
    end
end
initialize_hdf5(mc, paramsfile::String) = initialize_hdf5(mc.outpath, mc.T, paramsfile, mc.lattice.spins, mc.lattice.site_positions)

write_spins(path::String, spins) = h5open(f -> (f["spins"][:, :] = spins), path, "r+")
write_MC_checkpoint(mc) = write_spins(mc.outpath, mc.lattice.spins)

function write_initial_configuration(filename::String, T, configfile::String, spins)
    paramsfile = h5open(f -> read(attributes(f)["paramsfile"]), configfile, "r")
    h5open(filename, "w") do f
        write_attribute(f, "T", T); write_attribute(f, "paramsfile", paramsfile); f["spins"] = spins
    end
end

function write_observables(path::String, obs::Observables, T, N)
    heat, dheat = specific_heat(obs, T, N)
    chi, dchi = susceptibility(obs, T, N)
    h5open(path, "r+") do f
        haskey(f, "observables") && delete_object(f, "observables")
        g = create_group(f, "observables")
        g["specific_heat"] = heat;  g["specific_heat_err"] = abs(dheat / heat)
        g["susceptibility"] = chi;  g["susceptibility_err"] = abs(dchi / chi)
        g["magnetization"] = means(obs.magnetization)[1]; g["magnetization_err"] = std_error(obs.magnetization, 1)
        g["energy"] = means(obs.energy)[1];               g["energy_err"] = std_error(obs.energy, 1)
    end
end

function overwrite_keys!(fid, dict)
    for (k, v) in dict
        haskey(fid, k) && delete_object(fid, k)
        fid[k] = v
    end
end

read_spin_configuration!(lat::Lattice, filename::String) = h5open(f -> (lat.spins[:, :] = read(f["spins"])), filename, "r")
