# golden_from_reference.jl -- evaluate the REAL reference's soode_system! on the exported golden inputs and compare
# with the committed dv arrays (which came from the C restatement: "parity unpinned" until this has been run).
# NOT EXECUTED in this repository's CI (no Julia in the image).  Usage: see README.md in this directory.
using NBodySimulator, StaticArrays, SciMLBase, JSON

dir = ARGS[1]
manifest = JSON.parsefile(joinpath(dir, "manifest.json"))
readmat(name, field, n) = reshape(reinterpret(Float64, read(joinpath(dir, "$name.$field.f64"))), 3, :)
readvec(name, field) = collect(reinterpret(Float64, read(joinpath(dir, "$name.$field.f64"))))

function boundary(bc)
    bc[1] == "infinite" && return InfiniteBox()
    bc[1] == "cubic" && return CubicPeriodicBoundaryConditions(Float64(bc[2]))
    return PeriodicBoundaryConditions(Float64.(bc[2])...)
end

function thermostat(th)
    th === nothing && return NullThermostat()
    th["kind"] == "berendsen" && return BerendsenThermostat(th["T"], th["tau"])
    th["kind"] == "nosehoover" && return NoseHooverThermostat(th["T"], th["tau"])
    error("unsupported thermostat $(th["kind"])")
end

for (name, spec) in manifest
    startswith(name, "bench_") && continue
    get(spec, "water", false) && (println("$name: water systems are built from WaterSPCFw (positions of O only); skipped here"); continue)
    n = spec["n"]
    u = readmat(name, "u", n); v = readmat(name, "v", n); dv_gold = readmat(name, "dv", n); ms = readvec(name, "ms")
    bodies = if haskey(spec, "coulomb")
        qs = readvec(name, "qs")
        [ChargedParticle(SVector{3}(u[:, i]), SVector{3}(v[:, i]), ms[i], qs[i]) for i in 1:n]
    elseif haskey(spec, "dipole")
        mm = readmat(name, "mm", n)
        [MagneticParticle(SVector{3}(u[:, i]), SVector{3}(v[:, i]), ms[i], SVector{3}(mm[:, i])) for i in 1:n]
    else
        [MassBody(SVector{3}(u[:, i]), SVector{3}(v[:, i]), ms[i]) for i in 1:n]
    end
    pots = Dict{Symbol, NBodySimulator.PotentialParameters}()
    haskey(spec, "gravity") && (pots[:gravitational] = GravitationalParameters(spec["gravity"]["G"]))
    haskey(spec, "lj") && (pots[:lennard_jones] = LennardJonesParameters(spec["lj"]["eps"], spec["lj"]["sigma"], spec["lj"]["R"]))
    haskey(spec, "coulomb") && (pots[:electrostatic] = ElectrostaticParameters(spec["coulomb"]["k"], get(spec["coulomb"], "R", Inf)))
    haskey(spec, "dipole") && (pots[:magnetostatic] = MagnetostaticParameters(spec["dipole"]["mu_4pi"]))
    th = get(spec, "thermostat", nothing)
    kb = th === nothing ? 1.0 : th["kB"]
    sim = NBodySimulation(PotentialNBodySystem(bodies, pots), (0.0, 1.0), boundary(spec["bc"]), thermostat(th), kb)
    prob = SecondOrderODEProblem(sim)
    ncols = size(dv_gold, 2)
    uu = zeros(3, ncols); vv = zeros(3, ncols); dv = zeros(3, ncols)
    uu[:, 1:size(u, 2)] .= u; vv[:, 1:size(v, 2)] .= v
    prob.f.f1(dv, vv, uu, prob.p, 0.0)          # soode_system!(dv, v, u, p, t)  (src/nbody_to_ode.jl:474)
    err = maximum(i -> sqrt(sum(abs2, dv[:, i] .- dv_gold[:, i])) / max(sqrt(sum(abs2, dv_gold[:, i])), floatmin()), 1:n)
    println(rpad(name, 40), " max per-body relative difference reference vs golden = ", err, err == 0 ? "  (bit-identical)" : "")
end
