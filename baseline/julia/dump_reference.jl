# dump_reference.jl -- per-step output of the REAL reference (NQCDynamics.jl) for the five BASELINE configs, so the
# CPU oracle (oracle/) and through it the CUDA engine can be pinned to the reference's own implementation
# (SURVEY.md 8c last row; north_star correctness level 1: w, d, sigma, forces within 1e-10; level 2: identical hop
# sequences with injected draws).
#
#     julia --project=<env with NQCDynamics, OrdinaryDiffEq, JSON> baseline/julia/dump_reference.jl [outdir]
#
# writes  tests/golden/julia_<cfg>.json  (default outdir = tests/golden).  `python -m pytest tests/test_julia_golden.py`
# then compares the oracle (and, with a GPU, the engine) with every julia_*.json present at 1e-10, feeding the dumped
# t0 eigenvectors through nqcb200_set_gauge_reference and the dumped uniform draws through nqcb200_set_draws.
#
# STATUS: Julia is not installed in the image this repository is built in, so this script has NOT been run there;
# it only uses functions whose call sites are visible in the reference tree (cited inline).
#
# How the draws are injected: the reference calls the global `rand()` once per trajectory per step
# (fssh.jl:112, iesh.jl:393).  The two methods below re-state exactly those functions with `rand()` replaced by the next
# entry of a recorded vector; nothing else of the reference is touched.
using NQCDynamics
using NQCDynamics: DynamicsMethods, DynamicsUtils
using NQCDynamics.DynamicsMethods: SurfaceHoppingMethods
using NQCDynamics.DynamicsUtils: get_positions, get_velocities, get_quantum_subsystem
using OrdinaryDiffEq
using LinearAlgebra
using Random
import NQCCalculators, NQCModels
import JSON

const DRAWS = Ref(Float64[])
const DRAW_POS = Ref(0)
next_draw() = (DRAW_POS[] += 1; DRAWS[][DRAW_POS[]])
use_draws!(xi) = (DRAWS[] = xi; DRAW_POS[] = 0)

@eval SurfaceHoppingMethods begin
    # fssh.jl:110-121, rand() -> injected
    function select_new_state(sim::AbstractSimulation{<:FSSH}, u)
        random_number = Main.next_draw()
        for (i, prob) in enumerate(sim.method.hopping_probability)
            if i != sim.method.state
                if prob > random_number
                    return i
                end
            end
        end
        return sim.method.state
    end
    # iesh.jl:390-397, rand() -> injected
    function iesh_check_hop!(u, t, integrator)::Bool
        sim = integrator.p
        ishoppingdisabled(sim.method) && return false
        random = Main.next_draw()
        evaluate_hopping_probability!(sim, u, OrdinaryDiffEq.get_proposed_dt(integrator), random)
        set_new_state!(sim.method, select_new_state(sim, u, random))
        return sim.method.new_state != sim.method.state
    end
end

flat(x) = vec(collect(Float64, x))

"Everything the parity test compares after a step (or at t0), from the integrator and the simulation cache."
function snapshot(sim, integrator)
    u = integrator.u
    r = get_positions(u)
    d = Dict{String,Any}("t" => integrator.t, "r" => flat(r), "v" => flat(get_velocities(u)))
    if hasproperty(u, :σreal)
        d["sigma_re"] = flat(u.σreal); d["sigma_im"] = flat(u.σimag)       # column-major (n, n) or (n, ne)
    end
    if hasproperty(u, :state)
        d["state"] = round.(Int, flat(u.state))
    end
    if hasproperty(u, :qmap)
        d["qmap"] = flat(u.qmap); d["pmap"] = flat(u.pmap)
    end
    if NQCModels.nstates(sim) > 1
        if sim isa RingPolymerSimulation
            eig = NQCCalculators.get_centroid_eigen(sim.cache, r)                         # bcb_electronics.jl:84
            nac = NQCCalculators.get_centroid_nonadiabatic_coupling(sim.cache, r)         # bcb_electronics.jl:83
            beads = NQCCalculators.get_eigen(sim.cache, r)                                # per bead: test/Core/calculators.jl:133
            d["Z_beads"] = reduce(vcat, (flat(beads[i].Z) for i in eachindex(beads)))     # [bead][n*n column-major]
        else
            eig = NQCCalculators.get_eigen(sim.cache, r)                                  # bab_electronics.jl:84
            nac = NQCCalculators.get_nonadiabatic_coupling(sim.cache, r)                  # bab_electronics.jl:83
        end
        d["w"] = flat(eig.w); d["Z"] = flat(eig.Z)                                        # Z column-major
        d["nac"] = reduce(vcat, (flat(nac[I]) for I in eachindex(nac)))                   # [dof][n*n column-major]
    end
    cache = integrator.cache
    if hasproperty(cache, :k)
        d["accel"] = flat(cache.k)              # the acceleration carried to the next half kick (quirk Q2)
    end
    return d
end

"Run one trajectory step by step with injected draws; returns the per-trajectory record."
function run_trajectory(sim, u0, tspan, dt, xi)
    use_draws!(xi)
    problem = DynamicsMethods.create_problem(u0, tspan, sim)                              # SurfaceHoppingMethods.jl:75-79
    integrator = init(problem, DynamicsMethods.select_algorithm(sim); dt = dt,
                      callback = DynamicsMethods.get_callbacks(sim), save_everystep = false)
    rec = Dict{String,Any}("t0" => snapshot(sim, integrator), "draws" => xi, "steps" => Any[])
    nsteps = round(Int, (tspan[2] - tspan[1]) / dt)
    for _ in 1:nsteps
        step!(integrator)
        push!(rec["steps"], snapshot(sim, integrator))
    end
    rec["draws_used"] = DRAW_POS[]
    return rec
end

function dump(outdir, name, sim, tspan, dt, u0s; meta = Dict{String,Any}())
    rng = MersenneTwister(20261017)
    nsteps = round(Int, (tspan[2] - tspan[1]) / dt)
    trajs = [run_trajectory(sim, u0, tspan, dt, rand(rng, nsteps)) for u0 in u0s]
    doc = merge(Dict{String,Any}("config" => name, "dt" => dt, "t0" => tspan[1], "nsteps" => nsteps,
                                 "masses" => flat(sim.atoms.masses), "nstates" => NQCModels.nstates(sim),
                                 "size" => collect(size(sim)), "method" => string(nameof(typeof(sim.method))),
                                 "model" => string(nameof(typeof(sim.cache.model))),
                                 "nqcdynamics_version" => string(pkgversion(NQCDynamics)),
                                 "trajectories" => trajs), meta)
    path = joinpath(outdir, "julia_$(name).json")
    open(io -> JSON.print(io, doc), path, "w")
    @info "wrote $path" trajectories = length(trajs) nsteps
end

function main(outdir)
    mkpath(outdir)
    Random.seed!(1)

    # ---- C1: TullyModelOne FSSH (docs/src/ensemble_simulations.md:39-55) ----------------------------------
    sim = Simulation{FSSH}(Atoms(2000), TullyModelOne())
    m = sim.cache.model
    u0s = [begin
               r = fill(-8.0 + 1.0 * randn(), 1, 1); v = fill(10.0 / 2000, 1, 1)
               NQCCalculators.update_cache!(sim.cache, r)
               DynamicsVariables(sim, v, r, PureState(2))
           end for _ in 1:8]
    dump(outdir, "C1_tully1_fssh", sim, (0.0, 3000.0), 1.0, u0s;
         meta = Dict("model_params" => Dict("a" => m.a, "b" => m.b, "c" => m.c, "d" => m.d), "rescaling" => "standard"))
    for rescaling in (:vinversion, :off)      # slow trajectories: frustrated hops (surface_hopping.jl:79-91,155-164)
        sim = Simulation{FSSH}(Atoms(2000), TullyModelOne(); rescaling)
        u0s = [begin
                   r = fill(-3.0 + 0.3 * randn(), 1, 1); v = fill(4.0 / 2000, 1, 1)
                   NQCCalculators.update_cache!(sim.cache, r)
                   DynamicsVariables(sim, v, r, PureState(1, Adiabatic()))
               end for _ in 1:8]
        dump(outdir, "C1_tully1_fssh_$(rescaling)", sim, (0.0, 2500.0), 1.0, u0s;
             meta = Dict("model_params" => Dict("a" => m.a, "b" => m.b, "c" => m.c, "d" => m.d), "rescaling" => string(rescaling)))
    end

    # ---- C1b: Ehrenfest on the same model ------------------------------------------------------------------
    sim = Simulation{Ehrenfest}(Atoms(2000), TullyModelOne())
    u0s = [begin
               r = fill(-5.0, 1, 1); v = fill((8.0 + 4k) / 2000, 1, 1)
               NQCCalculators.update_cache!(sim.cache, r)
               DynamicsVariables(sim, v, r, PureState(1))
           end for k in 1:4]
    dump(outdir, "C1_tully1_ehrenfest", sim, (0.0, 1500.0), 1.0, u0s;
         meta = Dict("model_params" => Dict("a" => m.a, "b" => m.b, "c" => m.c, "d" => m.d)))

    # ---- C2: SpinBoson, Debye bath, 100 modes (docs/src/examples/spinboson.md:22-28) ------------------------
    N = 100
    model = SpinBoson(DebyeSpectralDensity(0.25, 0.5), N, 0.0, 1.0)
    β = 5.0
    ω = model.ωⱼ
    σr = @. sqrt(1 / (2ω * tanh(β * ω / 2))); σv = @. sqrt(ω / (2 * tanh(β * ω / 2)))
    for (tag, M) in (("fssh", FSSH), ("ehrenfest", Ehrenfest))
        sim = Simulation{M}(Atoms(fill(1, N)), model)
        u0s = [begin
                   r = reshape(σr .* randn(N), 1, N); v = reshape(σv .* randn(N), 1, N)
                   NQCCalculators.update_cache!(sim.cache, r)
                   DynamicsVariables(sim, v, r, PureState(1))
               end for _ in 1:4]
        dump(outdir, "C2_spinboson_debye100_$(tag)", sim, (0.0, 20.0), 0.1, u0s;
             meta = Dict("model_params" => Dict("epsilon" => 0.0, "delta" => 1.0, "omega" => flat(model.ωⱼ), "c" => flat(model.cⱼ))))
    end

    # ---- C3: RPMD, 32 beads, Harmonic (test/Dynamics/algorithms/bcb.jl:14-17) -------------------------------
    T = 9.5e-4
    hm = Harmonic(m = 1837.0, ω = 0.005, r₀ = 0.1)
    sim = RingPolymerSimulation{Classical}(Atoms(1837.0), hm, 32; temperature = T)
    u0s = [DynamicsVariables(sim, sqrt(T * 32 / 1837.0) .* randn(1, 1, 32), 0.1 .+ 0.2 .* randn(1, 1, 32)) for _ in 1:4]
    dump(outdir, "C3_rpmd_harmonic32", sim, (0.0, 500.0), 2.5, u0s;
         meta = Dict("model_params" => Dict("m" => 1837.0, "omega" => 0.005, "r0" => 0.1), "nbeads" => 32, "temperature" => T))

    # ---- C4: AdiabaticIESH, MiaoSubotnik + TrapezoidalRule (test/Dynamics/iesh.jl:17-25 with the BASELINE model) ---
    Γ = 6.4e-3
    for M in (30, 100)
        W = 3Γ
        am = AndersonHolstein(MiaoSubotnik(; Γ), TrapezoidalRule(M, -W, W))
        sim = Simulation{AdiabaticIESH}(Atoms(2000), am)
        u0s = [begin
                   r = fill(8.0 + 10.0 * rand(), 1, 1); v = fill(-abs(randn()) * 2e-3, 1, 1)
                   NQCCalculators.update_cache!(sim.cache, r)
                   DynamicsVariables(sim, v, r)                                           # iesh.jl:89-97
               end for _ in 1:(M == 30 ? 4 : 2)]
        imp = am.impurity_model
        dump(outdir, "C4_iesh_miao_subotnik_m$(M)", sim, (0.0, M == 30 ? 200.0 : 50.0), M == 30 ? 5.0 : 1.0, u0s;
             meta = Dict("model_params" => Dict("m" => imp.m, "omega" => imp.ω, "g" => imp.g, "DeltaG" => imp.ΔG, "Gamma" => Γ,
                                                "eps" => flat(am.bath.bathstates), "V" => flat(am.bath.bathcoupling) .* sqrt(Γ / 2π)),
                         "nelectrons" => NQCModels.nelectrons(am)))
    end

    # ---- RPIESH / RP-EhrenfestNA with BCBWavefunction (test/Dynamics/rpiesh.jl:11-22, rp_ehrenfest_na.jl:10-31) ------------
    # These pin quirk Q5 (propagate_wavefunction!(.., vprev, rprev, ..), bcb_wavefunction.jl:67) and the bead-sum force.  The
    # replay takes no gauge reference for them: tests/julia_golden.py moves the dump into the identity-continuity gauge.
    # RP-EhrenfestNA contracts bead-basis derivatives with centroid-basis psi (rpehrenfest_na.jl:20-28), so its force depends
    # on the column signs LAPACK gives each bead at t0; a disagreement there is that dependence, not a propagation error.
    for (tag, Meth) in (("rpiesh", AdiabaticIESH), ("rp_ehrenfest_na", EhrenfestNA))
        M, B, T = 30, 4, 9.5e-4
        am = AndersonHolstein(MiaoSubotnik(; Γ), TrapezoidalRule(M, -3Γ, 3Γ))
        sim = RingPolymerSimulation{Meth}(Atoms(2000), am, B; temperature = T)
        u0s = [begin
                   r = (8.0 + 10.0 * rand()) .+ 0.5 .* randn(1, 1, B); v = -abs(randn()) * 2e-3 .+ 1e-4 .* randn(1, 1, B)
                   NQCCalculators.update_cache!(sim.cache, r)
                   DynamicsVariables(sim, v, r)
               end for _ in 1:4]
        imp = am.impurity_model
        dump(outdir, "RP_$(tag)_miao_subotnik_m$(M)_b$(B)", sim, (0.0, 200.0), 5.0, u0s;
             meta = Dict("model_params" => Dict("m" => imp.m, "omega" => imp.ω, "g" => imp.g, "DeltaG" => imp.ΔG, "Gamma" => Γ,
                                                "eps" => flat(am.bath.bathstates), "V" => flat(am.bath.bathcoupling) .* sqrt(Γ / 2π)),
                         "nelectrons" => NQCModels.nelectrons(am), "nbeads" => B, "temperature" => T))
    end

    # ---- C5: RPSH, 16 beads, ThreeStateMorse (docs/src/dynamicssimulations/dynamicsmethods/rpsh.md:40-68) -----------
    T = 9.5e-4
    tm = ThreeStateMorse()
    sim = RingPolymerSimulation{FSSH}(Atoms(20000), tm, 16; temperature = T)
    u0s = [begin
               r = 2.1 .+ (1 / sqrt(20000 * 0.005)) .* randn(1, 1, 16); v = sqrt(T * 16 / 20000) .* randn(1, 1, 16)
               NQCCalculators.update_cache!(sim.cache, r)
               DynamicsVariables(sim, v, r, PureState(1))
           end for _ in 1:4]
    dump(outdir, "C5_rpsh_morse3_16", sim, (0.0, 3000.0), 1.0, u0s;
         meta = Dict("model_params" => Dict(string(f) => getfield(tm, f) for f in fieldnames(typeof(tm))),
                     "nbeads" => 16, "temperature" => T))
end

main(length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "..", "tests", "golden"))
