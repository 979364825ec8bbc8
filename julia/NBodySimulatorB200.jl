# NBodySimulatorB200.jl -- the reference-side binding of libnbody_b200.so (pure `ccall`, no CUDA.jl).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither this image nor the GPU box has Julia.  The same
# sequence of C-ABI calls is exercised from Python ctypes (nbodysimulator.jl_b200/_lib.py, tests/).
# Every ccall below matches one prototype of include/nbody_b200.h.
#
# Three adapters over the one C ABI (SURVEY.md section 8b):
#   (1) RHS drop-in   : B200Problem(sim) -> a normal SecondOrderODEProblem whose soode_system! is
#                       `ccall(:nbx_accel, ...)`; any DiffEq integrator / callback / accessor works.
#   (2) fused drop-in : run_simulation(sim, B200VelocityVerlet(); dt, saveat) runs the whole loop on
#                       the device (nbx_run_vv) and wraps the saved frames in a SimulationResult.
#   (3) plugin closure: GPUPotential(parameters) <: PotentialParameters whose get_accelerating_function
#                       returns the reference's own per-particle `acceleration!(dv, u, v, t, i)`
#                       (src/nbody_to_ode.jl:156-242): one device evaluation per sweep, served column by column.
# `device` is one CUDA ordinal or a collection of them (`device = 0:7`): several GPUs go through
# nbx_create_multi -- the same calls on one handle, fanned out inside the library (pair sharding, x-slabs or
# target blocks, exchanges over NVLink peer memory); nothing else changes on the Julia side.
module NBodySimulatorB200

using NBodySimulator
using NBodySimulator: NBodySimulation, PotentialNBodySystem, WaterSPCFw, get_masses,
                      gather_bodies_initial_coordinates, InfiniteBox, PeriodicBoundaryConditions,
                      CubicPeriodicBoundaryConditions, BerendsenThermostat, NoseHooverThermostat,
                      AndersenThermostat, LangevinThermostat, NullThermostat
using SciMLBase, RecursiveArrayTools

const LIB = get(ENV, "NBODY_B200_LIB", "libnbody_b200.so")

struct NbxError <: Exception
    code::Cint
    msg::String
end

mutable struct B200Context
    h::Ptr{Cvoid}
    n::Int
    ncols::Int
    function B200Context(device = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = if device isa Integer
            ccall((:nbx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
        else   # several GPUs of this process behind one handle
            devs = Cint[d for d in device]
            ccall((:nbx_create_multi, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), ref, length(devs), devs)
        end
        rc == 0 || throw(NbxError(rc, unsafe_string(ccall((:nbx_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        ctx = new(ref[], 0, 0)
        finalizer(c -> ccall((:nbx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

function check(ctx::B200Context, rc::Cint)
    rc == 0 && return nothing
    throw(NbxError(rc, unsafe_string(ccall((:nbx_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))))
end

# ---- lowering an NBodySimulation to the context (mirrors nbody_to_ode.jl:263-433) ----------------
function configure!(ctx::B200Context, s::NBodySimulation)
    sys = s.system
    ms = Vector{Float64}(get_masses(sys))
    n = length(ms)
    if sys isa WaterSPCFw
        qs = repeat([sys.qO, sys.qH, sys.qH], length(sys.bodies))
        check(ctx, ccall((:nbx_system, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                         ctx.h, n, ms, Vector{Float64}(qs), C_NULL, 1))
    else
        pots = sys.potentials
        qs = haskey(pots, :electrostatic) ? Float64[b.q for b in sys.bodies] : nothing
        mm = haskey(pots, :magnetostatic) ? Float64[b.mm[k] for k in 1:3, b in sys.bodies] : nothing
        check(ctx, ccall((:nbx_system, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                         ctx.h, n, ms, qs === nothing ? C_NULL : qs, mm === nothing ? C_NULL : mm, 0))
    end
    ctx.n = n

    bc = s.boundary_conditions
    if bc isa CubicPeriodicBoundaryConditions
        check(ctx, ccall((:nbx_boundary, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.h, 1, Float64[bc.L]))
    elseif bc isa PeriodicBoundaryConditions
        check(ctx, ccall((:nbx_boundary, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.h, 2, Float64[bc[i] for i in 1:6]))
    else
        check(ctx, ccall((:nbx_boundary, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.h, 0, C_NULL))
    end

    if sys isa WaterSPCFw
        lj, el, sp = sys.lj_parameters, sys.e_parameters, sys.scpfw_parameters
        check(ctx, ccall((:nbx_add_lj, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Float64), ctx.h, lj.ϵ, lj.σ, lj.R))
        check(ctx, ccall((:nbx_add_coulomb, LIB), Cint, (Ptr{Cvoid}, Float64, Float64), ctx.h, el.k, el.R))
        check(ctx, ccall((:nbx_add_spcfw, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Float64, Float64),
                         ctx.h, sp.rOH, sp.aHOH, sp.kb, sp.ka))
    else
        for (name, p) in sys.potentials
            if name == :lennard_jones
                check(ctx, ccall((:nbx_add_lj, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Float64), ctx.h, p.ϵ, p.σ, p.R))
            elseif name == :electrostatic
                check(ctx, ccall((:nbx_add_coulomb, LIB), Cint, (Ptr{Cvoid}, Float64, Float64), ctx.h, p.k, p.R))
            elseif name == :magnetostatic
                check(ctx, ccall((:nbx_add_dipole, LIB), Cint, (Ptr{Cvoid}, Float64), ctx.h, p.μ_4π))
            elseif name == :gravitational
                check(ctx, ccall((:nbx_add_gravity, LIB), Cint, (Ptr{Cvoid}, Float64), ctx.h, p.G))
            else
                error("potential $name has no B200 kernel; keep its closure on the host")
            end
        end
    end

    th = s.thermostat
    (N, Nc, _) = NBodySimulator.get_degrees_of_freedom(sys)
    kind, T0, par = th isa BerendsenThermostat ? (1, th.T, th.τ) :
                    th isa NoseHooverThermostat ? (2, th.T, th.τ) :
                    th isa AndersenThermostat ? (3, th.T, th.ν) :
                    th isa LangevinThermostat ? (4, th.T, th.γ) : (0, 0.0, 0.0)
    check(ctx, ccall((:nbx_thermostat, LIB), Cint, (Ptr{Cvoid}, Cint, Float64, Float64, Float64, Int64, Int64),
                     ctx.h, kind, T0, par, s.kb, N, Nc))
    ctx.ncols = n + (kind == 2 ? 1 : 0)
    return ctx
end

# ---- (1) RHS drop-in --------------------------------------------------------------------------------
"""
    B200Problem(simulation; device = 0)

Same object `SciMLBase.SecondOrderODEProblem(simulation)` returns (src/nbody_to_ode.jl:460-491), with
`soode_system!` evaluated on the GPU.  `u`, `v`, `dv` are the solver's own `Matrix{Float64}` (3 x ncols).
"""
function B200Problem(s::NBodySimulation; device = 0)
    ctx = configure!(B200Context(device), s)
    (u0, v0, n) = gather_bodies_initial_coordinates(s)
    function soode_system!(dv, v, u, p, t)
        check(ctx, ccall((:nbx_accel, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
                         ctx.h, u, v, t, dv))
        return nothing
    end
    return SecondOrderODEProblem(soode_system!, v0, u0, s.tspan)
end

# ---- (2) fused drop-in --------------------------------------------------------------------------------
struct B200VelocityVerlet end

function NBodySimulator.run_simulation(s::NBodySimulation, ::B200VelocityVerlet; dt, saveat::Integer = 1, device = 0)
    ctx = configure!(B200Context(device), s)
    (u0, v0, n) = gather_bodies_initial_coordinates(s)
    check(ctx, ccall((:nbx_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.h, u0, v0))
    nsteps = round(Int, (s.tspan[2] - s.tspan[1]) / dt)
    nfr = cld(nsteps, saveat)
    us = Array{Float64}(undef, 3, ctx.ncols, nfr)      # frame k = us[:, :, k]: consecutive 3 x ncols column-major arrays
    vs = Array{Float64}(undef, 3, ctx.ncols, nfr)
    got = Ref{Int64}(0)
    check(ctx, ccall((:nbx_run_vv, LIB), Cint, (Ptr{Cvoid}, Float64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Int64}),
                     ctx.h, dt, nsteps, saveat, us, vs, nfr, got))
    ts = [s.tspan[1]; [s.tspan[1] + min(k * saveat, nsteps) * dt for k in 1:got[]]]
    frames = [ArrayPartition(copy(v0), copy(u0)); [ArrayPartition(vs[:, :, k], us[:, :, k]) for k in 1:got[]]]
    # a DiffEq-shaped solution so that get_position / temperature / energies / rdf / msd work unchanged
    prob = B200Problem(s; device = device)
    sol = SciMLBase.build_solution(prob, B200VelocityVerlet(), ts, frames; retcode = SciMLBase.ReturnCode.Success)
    return NBodySimulator.SimulationResult(sol, s)
end

# Langevin thermostat: the SDE path of run_simulation (src/nbody_simulation_result.jl:488-492, calculate_simulation_sde) with the
# Euler-Maruyama steps fused on the device (nbx_step_em: atoms src/nbody_to_ode.jl:567-598, water :600-680).  Frames are laid
# out as the reference's SDEProblem state, hcat(u, v) (3 x 2 ncols).
struct B200EM end

function NBodySimulator.run_simulation(s::NBodySimulation, ::B200EM; dt, saveat::Integer = 1, seed::Integer = 1, device = 0)
    s.thermostat isa LangevinThermostat || error("B200EM needs a LangevinThermostat (the reference's SDEProblem)")
    ctx = configure!(B200Context(device), s)
    (u0, v0, n) = gather_bodies_initial_coordinates(s)
    check(ctx, ccall((:nbx_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx.h, u0, v0))
    nsteps = round(Int, (s.tspan[2] - s.tspan[1]) / dt)
    ts = [s.tspan[1]]
    frames = [hcat(u0, v0)]
    u = similar(u0); v = similar(v0)
    done = 0
    while done < nsteps
        k = min(saveat, nsteps - done)
        # the seed fixes the Philox key once; the library's step counter runs on across the calls
        check(ctx, ccall((:nbx_step_em, LIB), Cint, (Ptr{Cvoid}, Float64, Int64, UInt64), ctx.h, dt, k, done == 0 ? UInt64(seed) : UInt64(0)))
        check(ctx, ccall((:nbx_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ctx.h, u, v, C_NULL))
        done += k
        push!(ts, s.tspan[1] + done * dt)
        push!(frames, hcat(u, v))
    end
    prob = SciMLBase.SDEProblem(s)     # only its type and u0 shape are used by the result accessors
    sol = SciMLBase.build_solution(prob, B200EM(), ts, frames; retcode = SciMLBase.ReturnCode.Success)
    return NBodySimulator.SimulationResult(sol, s)
end

# ---- (3) the plugin interface itself: a PotentialParameters subtype served by the device ----------------------------
"""
    GPUPotential(parameters)

Wraps one of the reference's parameter structs (LennardJonesParameters, ElectrostaticParameters, MagnetostaticParameters,
GravitationalParameters).  `get_accelerating_function` then returns the reference's per-particle closure
`acceleration!(dv, u, v, t, i)` (src/nbody_to_ode.jl:156-242, src/basic_potentials.jl:10-19): the first index of a sweep
evaluates ALL accelerations of that potential on the device (one nbx_accel call on a context holding only this potential),
later indices add the cached column -- so a `PotentialNBodySystem(bodies, Dict(:custom => GPUPotential(p)))` runs through the
unmodified `soode_system!` loop (:474-488) of the reference.
"""
struct GPUPotential{P <: NBodySimulator.PotentialParameters} <: NBodySimulator.PotentialParameters
    parameters::P
    device::Any
end
GPUPotential(p) = GPUPotential(p, 0)

function NBodySimulator.get_accelerating_function(g::GPUPotential, simulation::NBodySimulation)
    p = g.parameters
    name = p isa NBodySimulator.LennardJonesParameters ? :lennard_jones :
           p isa NBodySimulator.ElectrostaticParameters ? :electrostatic :
           p isa NBodySimulator.MagnetostaticParameters ? :magnetostatic :
           p isa NBodySimulator.GravitationalParameters ? :gravitational : error("no B200 kernel for $(typeof(p))")
    only = NBodySimulation(PotentialNBodySystem(simulation.system.bodies, Dict(name => p)), simulation.tspan,
                           simulation.boundary_conditions, NullThermostat(), simulation.kb)
    ctx = configure!(B200Context(g.device), only)
    n = ctx.n
    cache = zeros(3, n)
    last_i = Ref(typemax(Int)); last_t = Ref(NaN); last_u = Ref{Ptr{Float64}}(C_NULL); last_sum = Ref(NaN)
    return function (dv, u, v, t, i)
        # a sweep visits i in ascending order (:475): a non-increasing index, another array, time or content starts a new one
        if i <= last_i[] || t != last_t[] || pointer(u) != last_u[] || sum(view(u, :, 1:n)) != last_sum[]
            uu = size(u, 2) == n ? u : u[:, 1:n]       # (Nose-Hoover states carry an extra column, :6-8)
            check(ctx, ccall((:nbx_accel, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
                             ctx.h, uu, C_NULL, t, cache))
            last_t[] = t; last_u[] = pointer(u); last_sum[] = sum(view(u, :, 1:n))
        end
        last_i[] = i
        dv .+= view(cache, :, i)
        return nothing
    end
end

# ---- (4) frame analysis on the device: rdf / msd (src/nbody_simulation_result.jl:664-783) ------------------------
# Same return values as NBodySimulator.rdf / msd; the O(frames x N^2) pair loop runs in rdf_kernel (integer histogram,
# identical counts), the normalisation is the reference's (:695-707).
function rdf_b200(sr::NBodySimulator.SimulationResult; device::Integer = 0)
    s = sr.simulation
    ctx = configure!(B200Context(device), s)
    L = s.boundary_conditions.L
    maxbin = 1000
    check(ctx, ccall((:nbx_rdf_reset, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, maxbin))
    for t in sr.solution.t
        cc = Matrix{Float64}(NBodySimulator.get_position(sr, t))
        check(ctx, ccall((:nbx_rdf_add, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, cc))
    end
    hist = zeros(Int64, maxbin)
    frames = Ref{Int64}(0)
    check(ctx, ccall((:nbx_rdf_get, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Int64}), ctx.h, hist, maxbin, frames))
    indlen = length(s.system.bodies)
    dr = L / maxbin
    c = 4 / 3 * π * indlen / L^3
    rs = [(bin - 1) * dr + dr / 2 for bin in 1:maxbin]
    gr = [(hist[bin] / (frames[] * indlen)) / (c * ((bin * dr)^3 - ((bin - 1) * dr)^3)) for bin in 1:maxbin]
    return (rs, gr)
end

function msd_b200(sr::NBodySimulator.SimulationResult; device::Integer = 0)
    s = sr.simulation
    ctx = configure!(B200Context(device), s)
    ts = sr.solution.t
    cc0 = Matrix{Float64}(NBodySimulator.get_position(sr, ts[1]))
    dr2 = zeros(length(ts))
    out = Ref{Float64}(0.0)
    for (k, t) in enumerate(ts)
        cc = Matrix{Float64}(NBodySimulator.get_position(sr, t))
        check(ctx, ccall((:nbx_msd, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}), ctx.h, cc0, cc, out))
        dr2[k] = out[]
    end
    return (ts, dr2)
end

export B200Context, B200Problem, B200VelocityVerlet, B200EM, GPUPotential, rdf_b200, msd_b200

end # module
