# bench_reference.jl -- time the REAL reference's CPU path on the benchmark workloads of bench.py.
# NOT EXECUTED in this repository's CI (no Julia in the image); bench.py --impl reference times the C restatement
# (oracle/nbody_oracle.c) instead and says so ("kind": "port").  Usage: see README.md; export with --bench.
using NBodySimulator, StaticArrays, SciMLBase, JSON, OrdinaryDiffEqSymplecticRK

dir = ARGS[1]
readmat(name, field) = reshape(reinterpret(Float64, read(joinpath(dir, "$name.$field.f64"))), 3, :)
readvec(name, field) = collect(reinterpret(Float64, read(joinpath(dir, "$name.$field.f64"))))
manifest = JSON.parsefile(joinpath(dir, "manifest.json"))

# (1) gravity, 262,144-body Plummer sphere: one RHS is 6.9e10 pair interactions (~5 min single-threaded), so time
#     gravitational_acceleration! (src/basic_potentials.jl:306-331) for a bounded sample of targets against all sources.
let name = "bench_gravity_262144"
    u = readmat(name, "u"); v = readmat(name, "v"); ms = readvec(name, "ms"); n = length(ms)
    bodies = [MassBody(SVector{3}(u[:, i]), SVector{3}(v[:, i]), ms[i]) for i in 1:n]
    sim = NBodySimulation(GravitationalSystem(bodies, 1.0), (0.0, 1.0))
    acc! = NBodySimulator.get_accelerating_function(GravitationalParameters(1.0), sim)
    targets = 1:256:n
    dv = zeros(3)
    acc!(dv, u, v, 0.0, 1)                        # compile
    t = @elapsed for i in targets
        fill!(dv, 0.0); acc!(dv, u, v, 0.0, i)
    end
    println("gravity: ", length(targets) * (n - 1) / t, " pair-interactions/s on 1 thread (", length(targets), " targets x ", n, " sources, ", t, " s)")
end

# (2) examples/liquid_argon.jl as shipped (216 atoms, VelocityVerlet): atom-steps/s of run_simulation
let name = "bench_argon_216", spec = manifest[name]
    u = readmat(name, "u"); v = readmat(name, "v"); ms = readvec(name, "ms"); n = length(ms)
    bodies = [MassBody(SVector{3}(u[:, i]), SVector{3}(v[:, i]), ms[i]) for i in 1:n]
    lj = LennardJonesParameters(spec["lj"]["eps"], spec["lj"]["sigma"], spec["lj"]["R"])
    sys = PotentialNBodySystem(bodies, Dict(:lennard_jones => lj))
    steps = 2000
    sim = NBodySimulation(sys, (0.0, steps * spec["dt"]), CubicPeriodicBoundaryConditions(spec["bc"][2]), 1.38e-23)
    run_simulation(sim, VelocityVerlet(), dt = spec["dt"])   # compile
    t = @elapsed run_simulation(sim, VelocityVerlet(), dt = spec["dt"])
    println("liquid argon 216: ", n * steps / t, " atom-steps/s on 1 thread (", steps, " steps, ", t, " s)")
end
