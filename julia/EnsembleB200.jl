# EnsembleB200.jl -- the Julia side of the drop-in boundary (BASELINE.json north_star: "The Julia API stays the
# drop-in surface ... a new `EnsembleB200(ngpus)` ensemble_algorithm calls a thin ccall C-ABI").
#
#     using NQCDynamics; include("julia/EnsembleB200.jl"); using .NQCDB200
#     run_dynamics(sim, tspan, distribution; trajectories, output, dt, saveat, reduction,
#                  ensemble_algorithm = EnsembleB200(ngpus))
#
# Seam: `SciMLBase.solve(ensemble_problem, algorithm, ensemble_algorithm; trajectories, kwargs...)`,
# reference src/Ensembles/run_dynamics.jl:91-97.  `EnsembleB200 <: SciMLBase.BasicEnsembleAlgorithm`, so SciMLBase's
# generic `__solve` keeps doing the batching (`batch_size`) and calls `prob.reduction(u, batch, I)`
# (src/Ensembles/reductions.jl) itself; this file only adds `SciMLBase.solve_batch` for the new algorithm.
#
# STATUS: Julia is not installed in the image this repository is built in (SURVEY.md 8c), so this file has NOT been
# executed.  It binds exactly the entry points of include/nqcb200.h; every call below has a tested twin in the Python
# mirror nqcdynamics.jl_b200/api.py (`run_dynamics`), which is driven through the same C ABI by tests/test_host_gpu.py.
# Names marked (EXTERNAL) belong to packages outside the reference tree (NQCModels field names, NQCCalculators
# accessors); they follow the call sites visible in the reference and must be confirmed against the installed versions.
#
# No CPU fallback: an unsupported (method, model, algorithm, output) combination is an `error`, never a silent
# `EnsembleThreads` run.
module NQCDB200

using SciMLBase
using Dictionaries: Dictionary
using LinearAlgebra: diag
using NQCDynamics
using NQCDynamics: DynamicsUtils, DynamicsMethods, Ensembles, RingPolymerSimulation, Simulation, masses
using NQCDynamics.DynamicsMethods: SurfaceHoppingMethods, EhrenfestMethods, ClassicalMethods, MappingVariableMethods,
                                   IntegrationAlgorithms
using NQCDynamics.Analysis.Postprocess: FakeSolution, FakeProblem
import NQCModels, NQCCalculators

export EnsembleB200, PositionOutside

# libnqcb200.so is built by `make -C nqcdynamics.jl_b200/csrc` (or `python -c "import __graft_entry__ as g; g.build()"`)
const LIB = get(ENV, "NQCB200_LIB", joinpath(@__DIR__, "..", "nqcdynamics.jl_b200", "csrc", "libnqcb200.so"))

"""
    EnsembleB200(ngpus = 1; device_ids = 0:ngpus-1)

`ensemble_algorithm` that steps every trajectory of a batch on B200 GPUs through `libnqcb200.so`.  Trajectories are
split into `ngpus` contiguous shards (one engine handle and one Julia task per shard; the Philox streams are keyed by
the global trajectory index, so results do not depend on `ngpus`).
"""
struct EnsembleB200 <: SciMLBase.BasicEnsembleAlgorithm
    ngpus::Int
    device_ids::Vector{Int}
end
EnsembleB200(ngpus::Integer = 1; device_ids = collect(0:ngpus-1)) = EnsembleB200(Int(ngpus), collect(Int, device_ids))

"""
    PositionOutside(lo, hi; dof = 1, outgoing = false, tcut = Inf)

Termination predicate the step kernels can evaluate (`callback = TerminatingCallback(PositionOutside(...))`,
src/DynamicsUtils/callbacks.jl:29): `r[dof] < lo || r[dof] > hi [&& moving outwards] || t > tcut`.  It is also an
ordinary `condition(u, t, integrator)`, so the same object works with the reference's CPU ensemble algorithms.
"""
struct PositionOutside
    lo::Float64
    hi::Float64
    dof::Int
    outgoing::Bool
    tcut::Float64
end
PositionOutside(lo, hi; dof = 1, outgoing = false, tcut = Inf) = PositionOutside(lo, hi, dof, outgoing, tcut)
function (f::PositionOutside)(u, t, integrator)
    r = DynamicsUtils.get_positions(u)[f.dof]
    v = DynamicsUtils.get_velocities(u)[f.dof]
    out = f.outgoing ? ((r < f.lo && v < 0) || (r > f.hi && v > 0)) : (r < f.lo || r > f.hi)
    return out || t > f.tcut
end

# ---------------------------------------------------------------------------------------------------------------
# mirror of `nqcb200_config` (include/nqcb200.h:167-201): field order and types must match exactly
# ---------------------------------------------------------------------------------------------------------------
Base.@kwdef mutable struct Config
    abi_version::Int32 = 1
    method::Int32 = 0
    model::Int32 = 0
    nstates::Int32 = 1
    ndofs::Int32 = 1
    nbeads::Int32 = 1
    nelectrons::Int32 = 0
    rescaling::Int32 = 0
    estimate_probability::Int32 = 1
    disable_hopping::Int32 = 0
    rng::Int32 = 0
    device::Int32 = 0
    save_every::Int32 = 1
    nsave::Int32 = 1
    per_trajectory::Int32 = 1
    diagnostics::Int32 = 0
    observables::UInt32 = 0
    reserved0::UInt32 = 0
    ntraj::Int64 = 0
    traj_offset::Int64 = 0
    seed::UInt64 = 0
    dt::Float64 = 1.0
    t0::Float64 = 0.0
    temperature::Float64 = 0.0
    nrpmd_gamma::Float64 = 0.5
    edc_C::Float64 = 0.0
    params::NTuple{32,Float64} = ntuple(_ -> 0.0, 32)
    masses::Ptr{Float64} = C_NULL
    bath_a::Ptr{Float64} = C_NULL
    bath_b::Ptr{Float64} = C_NULL
    nbath::Int32 = 0
    reserved1::Int32 = 0
end

const Handle = Ptr{Cvoid}

last_error(h::Handle) = unsafe_string(@ccall LIB.nqcb200_last_error(h::Handle)::Cstring)
check(rc::Integer, h::Handle = C_NULL) = rc == 0 ? nothing : error("EnsembleB200 [", rc, "]: ", last_error(h))

# ---------------------------------------------------------------------------------------------------------------
# enum tables (include/nqcb200.h:62-160)
# ---------------------------------------------------------------------------------------------------------------
const OBS_ADIABATIC_POP, OBS_DIABATIC_POP, OBS_POPCORR_DIABATIC, OBS_POPCORR_ADIABATIC = 0, 1, 2, 3
const OBS_KINETIC, OBS_POTENTIAL, OBS_TOTAL_ENERGY, OBS_POSITION, OBS_VELOCITY = 4, 5, 6, 7, 8
const OBS_DISCRETE_STATE, OBS_SCATTERING, OBS_SCATTERING_DIABATIC, OBS_SIGMA = 9, 10, 11, 12
const OBS_MAPPING_Q, OBS_MAPPING_P = 13, 14

method_id(::SurfaceHoppingMethods.FSSH) = 1
method_id(::EhrenfestMethods.Ehrenfest) = 2
method_id(::SurfaceHoppingMethods.AdiabaticIESH) = 3
method_id(::ClassicalMethods.Classical) = 4
method_id(::MappingVariableMethods.NRPMD) = 5
method_id(::EhrenfestMethods.EhrenfestNA) = 6
method_id(::ClassicalMethods.ThermalLangevin) = 7
method_id(m) = error("EnsembleB200: no device kernel for dynamics method $(typeof(m)) (no CPU fallback)")

# the engine implements the reference's DEFAULT algorithm of each method (IntegrationAlgorithms.jl:65-72,86-96,103)
function check_algorithm(sim, alg)
    expected = DynamicsMethods.select_algorithm(sim)
    typeof(alg) === typeof(expected) ||
        error("EnsembleB200 implements $(typeof(expected)) for $(typeof(sim.method)); got algorithm = $(typeof(alg))")
    return nothing
end

rescaling_id(s::Symbol) = s === :standard ? 0 : s === :vinversion ? 1 : s === :off ? 2 :
                          error("This mode of rescaling is not implemented: $s")      # surface_hopping.jl:90

# ---- model table: NQCModels type -> (enum nqcb200_model, params[], bath_a, bath_b, nelectrons) -----------------
# Field names are NQCModels' (EXTERNAL; recalled from its sources, the formulas are in the reference docs:
# docs/src/NQCModels/analyticmodels.md, systembathmodels.md:20-26,82-94,210-215, dynamicsmethods/iesh.md:71-105).
model_table(m::NQCModels.TullyModelOne)   = (1, (m.a, m.b, m.c, m.d), nothing, nothing, 0)
model_table(m::NQCModels.TullyModelTwo)   = (2, (m.a, m.b, m.c, m.d, m.e), nothing, nothing, 0)
model_table(m::NQCModels.TullyModelThree) = (3, (m.a, m.b, m.c), nothing, nothing, 0)
model_table(m::NQCModels.DoubleWell)      = (4, (m.mass, m.ω, m.γ, m.Δ), nothing, nothing, 0)
model_table(m::NQCModels.SpinBoson)       = (5, (m.ϵ, m.Δ), collect(Float64, m.ωⱼ), collect(Float64, m.cⱼ), 0)
function model_table(m::NQCModels.ThreeStateMorse)
    # params order of include/nqcb200.h:88-90: d1..3, alpha1..3, r1..3, c1..3, a12,a13,a23, alpha12,13,23, r12,13,23
    return (6, (m.d1, m.d2, m.d3, m.α1, m.α2, m.α3, m.r1, m.r2, m.r3, m.c1, m.c2, m.c3,
                m.a12, m.a13, m.a23, m.α12, m.α13, m.α23, m.r12, m.r13, m.r23), nothing, nothing, 0)
end
model_table(m::NQCModels.Harmonic) = (7, (m.m, m.ω, m.r₀), nothing, nothing, 0)
model_table(m::NQCModels.Free)     = (8, (), nothing, nothing, 0)
function model_table(m::NQCModels.AndersonHolstein)
    # H[1,1] = h(q) = U1 - U0, H[k+1,k+1] = eps_k, H[1,k+1] = V_k (iesh.md:71-76); the bath discretisation
    # (TrapezoidalRule / ShenviGaussLegendre / ...) is already evaluated in m.bath: only (eps_k, V_k) cross the ABI.
    eps = collect(Float64, m.bath.bathstates)
    V = collect(Float64, m.bath.bathcoupling) .* impurity_coupling(m.impurity_model)
    eps .-= m.fermi_level
    return (impurity_id(m.impurity_model), impurity_params(m.impurity_model), eps, V, Int(NQCModels.nelectrons(m)))
end
model_table(m) = error("EnsembleB200: no device implementation of model $(typeof(m)) (no CPU fallback)")

impurity_id(::NQCModels.MiaoSubotnik) = 9
impurity_params(m::NQCModels.MiaoSubotnik) = (m.m, m.ω, m.g, m.ΔG)
impurity_coupling(m::NQCModels.MiaoSubotnik) = sqrt(m.Γ / 2π)
impurity_id(::NQCModels.ErpenbeckThoss) = 10
# params order of include/nqcb200.h (ANDERSON_HOLSTEIN_ERPENBECK_THOSS): De, a, x0, c, D1, D2, a1, x01, Vinf, q, atilde, xtilde
#   U0 = De (exp(-a (x - x0)) - 1)^2 + c ; U1 = D1 exp(-2 a1 (x - x01)) - D2 exp(-a1 (x - x01)) + Vinf ;
#   V_k(x) = Vbar_k ((1 - q)/2 (1 - tanh((x - xtilde)/atilde)) + q)
impurity_params(m::NQCModels.ErpenbeckThoss) = (m.morse.Dₑ, m.morse.a, m.morse.x₀, m.c, m.D₁, m.D₂, m.a′, m.x₀′, m.V∞,
                                                m.q, m.ã, m.x̃)
impurity_coupling(m::NQCModels.ErpenbeckThoss) = sqrt(m.Γ / 2π)
impurity_id(m) = error("EnsembleB200: no device implementation of impurity model $(typeof(m))")

# ---- outputs -> device observables -----------------------------------------------------------------------------
# Outputs the device estimators reproduce directly (same value shapes as DynamicsOutputs.jl).  Anything else is
# evaluated on the host from streamed frames through Analysis.Postprocess.FakeSolution (src/Analysis/postprocess.jl).
struct DeviceOutput
    obs::Int
    kind::Symbol          # :series | :scatter | :popcorr
end
device_output(::typeof(OutputAdiabaticPopulation), sim) = DeviceOutput(OBS_ADIABATIC_POP, :series)
device_output(::typeof(OutputDiabaticPopulation), sim) = DeviceOutput(OBS_DIABATIC_POP, :series)
device_output(::typeof(OutputKineticEnergy), sim) = DeviceOutput(OBS_KINETIC, :scalar)
device_output(::typeof(OutputPotentialEnergy), sim) = DeviceOutput(OBS_POTENTIAL, :scalar)
device_output(::typeof(OutputTotalEnergy), sim) = DeviceOutput(OBS_TOTAL_ENERGY, :scalar)
device_output(::typeof(OutputCentroidPosition), sim) = DeviceOutput(OBS_POSITION, :nuclear)
device_output(::typeof(OutputCentroidVelocity), sim) = DeviceOutput(OBS_VELOCITY, :nuclear)
device_output(f::TimeCorrelationFunctions.PopulationCorrelationFunction, sim) =
    DeviceOutput(f.statetype isa Adiabatic ? OBS_POPCORR_ADIABATIC : OBS_POPCORR_DIABATIC, :popcorr)
device_output(f::OutputStateResolvedScattering1D, sim) =
    DeviceOutput(f.type === :adiabatic ? OBS_SCATTERING : OBS_SCATTERING_DIABATIC, :scatter)
device_output(f, sim) = nothing                                       # host evaluation from frames
# OutputPosition / OutputVelocity of a plain Simulation are the device streams themselves; for ring polymers they are
# bead-resolved (DynamicsOutputs.jl:39,66), which the engine does not stream (centroids only) -> host path refuses.
device_output(::typeof(OutputPosition), sim::Simulation) = DeviceOutput(OBS_POSITION, :nuclear)
device_output(::typeof(OutputVelocity), sim::Simulation) = DeviceOutput(OBS_VELOCITY, :nuclear)

has_sigma(sim) = method_id(sim.method) in (1, 2, 3, 6)
has_state(sim) = method_id(sim.method) in (1, 3)
frame_observables(sim) = (OBS_POSITION, OBS_VELOCITY, (has_sigma(sim) ? (OBS_SIGMA,) : ())...,
                          (has_state(sim) ? (OBS_DISCRETE_STATE,) : ())...,
                          (method_id(sim.method) == 5 ? (OBS_MAPPING_Q, OBS_MAPPING_P) : ())...)

"Bitmask of device observables needed for `functions`, and whether host-side frames are needed."
function obs_mask(sim, functions::Tuple)
    mask = UInt32(0)
    need_frames = false
    for f in functions
        d = device_output(f, sim)
        if d === nothing
            need_frames = true
        else
            mask |= UInt32(1) << d.obs
        end
    end
    if need_frames
        sim isa RingPolymerSimulation &&
            error("EnsembleB200: this output needs bead-resolved frames, which the device does not stream; use the " *
                  "Centroid / population / energy outputs with ring polymers")
        for o in frame_observables(sim)
            mask |= UInt32(1) << o
        end
    end
    return mask, need_frames
end

# ---------------------------------------------------------------------------------------------------------------
# one shard = one engine handle on one device
# ---------------------------------------------------------------------------------------------------------------
struct ShardPlan
    cfg::Config
    m::Vector{Float64}
    bath_a::Union{Nothing,Vector{Float64}}
    bath_b::Union{Nothing,Vector{Float64}}
end

edc_constant(method) = hasproperty(method, :decoherence) && method.decoherence isa SurfaceHoppingMethods.DecoherenceCorrectionEDC ?
                       Float64(method.decoherence.C) : 0.0

function make_plan(sim, tspan, dt, save_every, nsave, mask, II, device, seed)
    mid, params, ba, bb, ne = model_table(sim.cache.model)
    m = repeat(collect(Float64, masses(sim)); inner = size(sim)[1])         # one mass per nuclear dof, dof fastest
    B = sim isa RingPolymerSimulation ? length(sim.beads) : 1
    method = sim.method
    cfg = Config(; method = method_id(method), model = mid, nstates = NQCModels.nstates(sim), ndofs = length(m),
                 nbeads = B, nelectrons = ne,
                 rescaling = hasproperty(method, :rescaling) ? rescaling_id(method.rescaling) : 0,
                 estimate_probability = hasproperty(method, :estimate_probability) ? Int32(method.estimate_probability) : 1,
                 disable_hopping = hasproperty(method, :disable_hopping) ? Int32(method.disable_hopping) : 0,
                 device = device, save_every = save_every, nsave = nsave, per_trajectory = 1, observables = mask,
                 ntraj = length(II), traj_offset = first(II) - 1, seed = seed, dt = dt, t0 = tspan[1],
                 temperature = sim isa RingPolymerSimulation ? Float64(NQCDynamics.get_ring_polymer_temperature(sim)) / B : 0.0,
                 nrpmd_gamma = hasproperty(method, :γ) ? Float64(method.γ) : 0.5, edc_C = edc_constant(method),
                 params = ntuple(i -> i <= length(params) ? Float64(params[i]) : 0.0, 32))
    return ShardPlan(cfg, m, ba, bb)
end

function create_handle(plan::ShardPlan)
    h = Ref{Handle}(C_NULL)
    cfg = plan.cfg
    m, ba, bb = plan.m, plan.bath_a, plan.bath_b
    GC.@preserve m ba bb begin
        cfg.masses = pointer(m)
        if ba !== nothing
            cfg.bath_a = pointer(ba); cfg.bath_b = pointer(bb); cfg.nbath = length(ba)
        end
        rc = @ccall LIB.nqcb200_create(cfg::Ref{Config}, h::Ptr{Handle})::Cint      # the library copies what it needs
    end
    check(rc, C_NULL)
    return h[]
end

flat(x) = vec(collect(Float64, x))

"Pack the sampled `u0`s of one shard (trajectory-major == Julia column-major with the trajectory as last axis)."
function pack_state(sim, u0s)
    r = reduce(hcat, (flat(DynamicsUtils.get_positions(u)) for u in u0s))
    v = reduce(hcat, (flat(DynamicsUtils.get_velocities(u)) for u in u0s))
    σre = σim = nothing
    state = nothing
    if has_sigma(sim)
        σre = reduce(hcat, (flat(u.σreal) for u in u0s))
        σim = reduce(hcat, (flat(u.σimag) for u in u0s))
    end
    if has_state(sim)
        state = reduce(hcat, (round.(Int32, vec(collect(u.state))) for u in u0s))   # u.state is stored as Float64 (quirk Q7)
    end
    return r, v, σre, σim, state
end

ptr_or_null(x::Nothing) = Ptr{Float64}(C_NULL)
ptr_or_null(x::AbstractArray{Float64}) = pointer(x)
iptr_or_null(x::Nothing) = Ptr{Int32}(C_NULL)
iptr_or_null(x::AbstractArray{Int32}) = pointer(x)

function termination_of(kwargs)
    cb = get(kwargs, :callback, nothing)
    cb === nothing && return nothing
    cond = cb isa SciMLBase.DiscreteCallback ? cb.condition : nothing
    cond isa PositionOutside ||
        error("EnsembleB200: only TerminatingCallback(PositionOutside(...)) runs on the device (no host callbacks)")
    return cond
end

"Run one shard; returns Dict(obs id => Array(width, nsave, T)) and the termination steps."
function run_shard(sim, plan::ShardPlan, u0s, Zref, nsteps, term, obs_ids)
    h = create_handle(plan)
    try
        r, v, σre, σim, state = pack_state(sim, u0s)
        T = length(u0s)
        GC.@preserve r v σre σim state Zref begin
            if Zref !== nothing      # continue in LAPACK's eigenvector gauge (sim.cache.eigen.Z at each r0)
                check(@ccall(LIB.nqcb200_set_gauge_reference(h::Handle, Zref::Ptr{Float64}, 1::Int64)::Cint), h)
            end
            if term !== nothing
                check(@ccall(LIB.nqcb200_set_termination(h::Handle, (term.dof - 1)::Cint, term.lo::Cdouble, term.hi::Cdouble,
                                                         Cint(term.outgoing)::Cint, term.tcut::Cdouble)::Cint), h)
            end
            if method_id(sim.method) == 5       # NRPMD: r, v, then the mapping variables (nstates, nbeads) per trajectory
                check(@ccall(LIB.nqcb200_set_state(h::Handle, r::Ptr{Float64}, v::Ptr{Float64}, C_NULL::Ptr{Float64},
                                                   C_NULL::Ptr{Float64}, C_NULL::Ptr{Int32})::Cint), h)
                q = reduce(hcat, (flat(DynamicsUtils.get_mapping_positions(u)) for u in u0s))
                p = reduce(hcat, (flat(DynamicsUtils.get_mapping_momenta(u)) for u in u0s))
                check(@ccall(LIB.nqcb200_set_mapping(h::Handle, q::Ptr{Float64}, p::Ptr{Float64})::Cint), h)
                check(@threadcall((:nqcb200_run, LIB), Cint, (Handle, Int64), h, nsteps), h)
            else
                # one call per batch: set_state + run (launch-fused initialisation where the kernel family has one);
                # @threadcall keeps the Julia scheduler free during the multi-second blocking call
                check(@threadcall((:nqcb200_run_from_host, LIB), Cint,
                                  (Handle, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Cint, Int64),
                                  h, pointer(r), pointer(v), ptr_or_null(σre), ptr_or_null(σim), iptr_or_null(state),
                                  Ptr{Float64}(C_NULL), Cint(0), nsteps), h)
            end
        end
        out = Dict{Int,Array{Float64,3}}()
        nsave = Int(plan.cfg.nsave)
        for o in obs_ids
            w = Int(@ccall LIB.nqcb200_observable_width(h::Handle, o::Cint)::Cint)
            buf = Array{Float64,3}(undef, w, nsave, T)                 # out[(traj*nsave + isave)*width + k]
            check(@ccall(LIB.nqcb200_get_observable_per_trajectory(h::Handle, o::Cint, buf::Ptr{Float64},
                                                                    length(buf)::Int64)::Cint), h)
            out[o] = buf
        end
        tsteps = fill(Int64(-1), T)
        if term !== nothing
            check(@ccall(LIB.nqcb200_get_termination(h::Handle, tsteps::Ptr{Int64})::Cint), h)
        end
        return out, tsteps
    finally
        @ccall LIB.nqcb200_destroy(h::Handle)::Cint
    end
end

# ---------------------------------------------------------------------------------------------------------------
# engine arrays -> the values EnsembleSaver would have produced (Ensembles.jl:44-63, SURVEY.md A.1)
# ---------------------------------------------------------------------------------------------------------------
"Frames kept for a terminated trajectory: saveat points up to t_term, then the terminal state saved by the callback."
function frame_indices(term_step, save_every, nsave)
    term_step < 0 && return collect(1:nsave)
    kf, rem = divrem(term_step, save_every)
    return rem == 0 ? vcat(1:kf+1, kf + 1) : vcat(1:kf+1, kf + 2, kf + 2)
end

function shape_device_output(sim, d::DeviceOutput, a::AbstractMatrix{Float64}, idx)       # a: (width, nsave)
    n = NQCModels.nstates(sim)
    if d.kind === :scatter        # final frame only (DynamicsOutputs.jl:317-338)
        last = a[:, end]
        return ComponentArrays.ComponentVector(reflection = last[1:n], transmission = last[n+1:2n])
    elseif d.kind === :popcorr    # Vector{Matrix}: out[t][i, j] = P_i(0) P_j(t) (TimeCorrelationFunctions.jl:86-88)
        return [reshape(a[:, k], n, n) for k in idx]
    elseif d.kind === :scalar
        return [a[1, k] for k in idx]
    elseif d.kind === :nuclear    # (ndofs, natoms) per frame
        sz = size(sim)[1:2]
        return [reshape(a[:, k], sz) for k in idx]
    else
        return [a[:, k] for k in idx]
    end
end

"Rebuild DynamicsVariables frames from the streams (host evaluation of arbitrary output functions)."
function rebuild_frames(sim, template, streams, itraj, idx)
    us = Vector{typeof(template)}(undef, length(idx))
    for (j, k) in enumerate(idx)
        u = copy(template)
        vec(DynamicsUtils.get_positions(u)) .= @view streams[OBS_POSITION][:, k, itraj]
        vec(DynamicsUtils.get_velocities(u)) .= @view streams[OBS_VELOCITY][:, k, itraj]
        if has_sigma(sim)
            s = @view streams[OBS_SIGMA][:, k, itraj]
            half = length(s) ÷ 2
            vec(u.σreal) .= @view s[1:half]
            vec(u.σimag) .= @view s[half+1:end]
        end
        if has_state(sim)
            vec(u.state) .= @view streams[OBS_DISCRETE_STATE][:, k, itraj]
        end
        if method_id(sim.method) == 5
            vec(DynamicsUtils.get_mapping_positions(u)) .= @view streams[OBS_MAPPING_Q][:, k, itraj]
            vec(DynamicsUtils.get_mapping_momenta(u)) .= @view streams[OBS_MAPPING_P][:, k, itraj]
        end
        us[j] = u
    end
    return us
end

function trajectory_dictionary(sim, output_func, template, streams, itraj, iglobal, times, term_step, save_every, dt, t0)
    nsave = length(times)
    idx = frame_indices(term_step, save_every, nsave)
    t = term_step < 0 ? collect(times) : vcat(times[1:term_step ÷ save_every + 1],
                                              fill(t0 + dt * term_step, length(idx) - term_step ÷ save_every - 1))
    out = output_func.savetime ? Dictionary{Symbol,Any}([:Time], [t]) : Dictionary{Symbol,Any}()
    frames = nothing
    for f in output_func.functions
        name = f isa Function ? nameof(f) : nameof(typeof(f))               # Ensembles.jl:53-58
        d = device_output(f, sim)
        if d !== nothing
            insert!(out, name, shape_device_output(sim, d, @view(streams[d.obs][:, :, itraj]), idx))
        else
            if frames === nothing
                frames = FakeSolution(t, rebuild_frames(sim, template, streams, itraj, idx), FakeProblem(sim))
            end
            insert!(out, name, f(frames, iglobal))
        end
    end
    return out
end

# ---------------------------------------------------------------------------------------------------------------
# the hook: one batch of trajectories
# ---------------------------------------------------------------------------------------------------------------
function SciMLBase.solve_batch(prob, alg, e::EnsembleB200, II, pmap_batch_size; dt = 1.0, saveat = nothing, kwargs...)
    sim = prob.prob.p
    check_algorithm(sim, alg)
    prob.output_func isa Ensembles.EnsembleSaver ||
        error("EnsembleB200 expects the EnsembleSaver that run_dynamics builds as output_func")
    tspan = prob.prob.tspan
    nsteps = round(Int, (tspan[2] - tspan[1]) / dt)
    every = saveat === nothing ? dt : saveat isa Number ? saveat : (length(saveat) > 1 ? saveat[2] - saveat[1] : first(saveat))
    save_every = max(1, round(Int, every / dt))
    isapprox(save_every * dt, every; rtol = 1e-9) || error("EnsembleB200: saveat must be a multiple of dt (fixed-step integrators)")
    nsave = nsteps ÷ save_every + 1
    times = tspan[1] .+ dt .* save_every .* (0:nsave-1)
    mask, _ = obs_mask(sim, prob.output_func.functions)
    obs_ids = [o for o in 0:14 if (mask >> o) & 1 == 1]
    term = termination_of(kwargs)
    seed = rand(UInt64)

    # initial conditions exactly as the reference draws them (selections.jl:38-42,70-73): serial, because prob_func
    # mutates sim.cache (update_cache!) -- and that is what lets us read LAPACK's eigenvector gauge per trajectory
    II = collect(II)
    u0s = Vector{Any}(undef, length(II))
    n = NQCModels.nstates(sim)
    gauge = has_sigma(sim) && !(sim isa RingPolymerSimulation) && n > 1
    Zref = gauge ? Matrix{Float64}(undef, n * n, length(II)) : nothing
    for (k, i) in enumerate(II)
        u0s[k] = prob.prob_func(prob.prob, i, 1).u0
        if gauge
            r0 = DynamicsUtils.get_positions(u0s[k])
            NQCCalculators.update_cache!(sim.cache, r0)
            Zref[:, k] .= vec(NQCCalculators.get_eigen(sim.cache, r0).Z)            # (EXTERNAL) accessor, as bab_electronics.jl:84
        end
    end

    # shards: contiguous blocks, one task + one handle per device (SURVEY.md 8e)
    G = max(1, min(e.ngpus, length(II)))
    base, extra = divrem(length(II), G)
    bounds = cumsum(vcat(0, [base + (g <= extra ? 1 : 0) for g in 1:G]))
    tasks = map(1:G) do g
        rng = bounds[g]+1:bounds[g+1]
        plan = make_plan(sim, tspan, Float64(dt), save_every, nsave, mask, II[rng], e.device_ids[g], seed)
        Zg = Zref === nothing ? nothing : Zref[:, rng]
        Threads.@spawn run_shard(sim, plan, u0s[rng], Zg, nsteps, term, obs_ids)
    end
    results = fetch.(tasks)

    # one Dictionary per trajectory, in the order of II -- what EnsembleSaver returns (Ensembles.jl:44-51)
    out = Vector{Any}(undef, length(II))
    for g in 1:G
        streams, tsteps = results[g]
        for (j, k) in enumerate(bounds[g]+1:bounds[g+1])
            out[k] = trajectory_dictionary(sim, prob.output_func, u0s[k], streams, j, II[k], times, tsteps[j],
                                           save_every, Float64(dt), Float64(tspan[1]))
        end
    end
    return out
end

end # module
