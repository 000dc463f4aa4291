# bench_reference.jl -- the genuine CPU baseline: the five BASELINE.json configs through the UNMODIFIED reference with
# `ensemble_algorithm = EnsembleThreads()` (SURVEY.md 8d), timed by the `@timed` that run_dynamics already wraps around
# `SciMLBase.solve` (src/Ensembles/run_dynamics.jl:91).
#
#     julia -t $(nproc) --project=<env with NQCDynamics> baseline/julia/bench_reference.jl [scale]
#
# prints one JSON line per config in bench.py's format (`"impl": "reference-julia"`), `value` in trajectory-steps/s.
# `scale` (default 1.0) multiplies the trajectory counts, which are sized for roughly a minute per config on 16 cores.
#
# STATUS: Julia is not installed in the image this repository is built in; until this script has been run on the GPU
# box's host, BASELINE.md section 4's Julia row stays "not measured" and bench.py's reference arm reports the C++
# restatement (oracle/, `cpu_baseline.kind = "port"`).
using NQCDynamics
using Random
import JSON

const SCALE = length(ARGS) >= 1 ? parse(Float64, ARGS[1]) : 1.0
ntraj(n) = max(Threads.nthreads(), round(Int, n * SCALE))

function timed_run(name, sim, tspan, dist, dt, T; output, kwargs...)
    # run_dynamics' own precompile pass (run_dynamics.jl:100-120) warms the JIT; a second short call warms the threads
    run_dynamics(sim, (tspan[1], tspan[1] + 2dt), dist; output, dt, trajectories = Threads.nthreads(),
                 ensemble_algorithm = EnsembleThreads(), reduction = MeanReduction(), kwargs...)
    t = @elapsed run_dynamics(sim, tspan, dist; output, dt, trajectories = T, precompile_dynamics = false,
                              ensemble_algorithm = EnsembleThreads(), reduction = MeanReduction(), kwargs...)
    nsteps = round(Int, (tspan[2] - tspan[1]) / dt)
    line = Dict("impl" => "reference-julia", "metric" => "trajectory-steps/sec (FP64)", "value" => T * nsteps / t,
                "unit" => "trajectory-steps/s", "higher_is_better" => true, "dtype" => "f64",
                "config" => Dict("workload" => name, "trajectories" => T, "nuclear_steps" => nsteps, "dt" => dt),
                "cpu_baseline" => Dict("kind" => "reference", "cores" => Threads.nthreads(), "value" => T * nsteps / t,
                                       "sample" => "$T trajectories x $nsteps steps, EnsembleThreads, $(Sys.cpu_info()[1].model)"),
                "seconds" => t, "nqcdynamics_version" => string(pkgversion(NQCDynamics)))
    println(JSON.json(line))
    flush(stdout)
end

Random.seed!(20261017)

# C1  TullyModelOne FSSH (docs/src/ensemble_simulations.md:39-56)
let sim = Simulation{FSSH}(Atoms(2000), TullyModelOne())
    dist = DynamicalDistribution(10 / 2000, Normal(-8, 1), size(sim)) * PureState(2)
    timed_run("tully1_fssh", sim, (0.0, 3000.0), dist, 1.0, ntraj(2000);
              output = (OutputDiabaticPopulation, OutputStateResolvedScattering1D(sim, :adiabatic)), saveat = 10.0)
end

# C2  SpinBoson, Debye bath, 100 modes, FSSH and Ehrenfest (docs/src/examples/spinboson.md:22-66)
let N = 100, β = 5.0
    model = SpinBoson(DebyeSpectralDensity(0.25, 0.5), N, 0.0, 1.0)
    atoms = Atoms(fill(1, N))
    position = reshape([PositionHarmonicWigner(ω, β, 1) for ω in model.ωⱼ], 1, :)
    velocity = reshape([VelocityHarmonicWigner(ω, β, 1) for ω in model.ωⱼ], 1, :)
    for (tag, M) in (("fssh", FSSH), ("ehrenfest", Ehrenfest))
        sim = Simulation{M}(atoms, model)
        dist = DynamicalDistribution(velocity, position, size(sim)) * PureState(1)
        timed_run("spinboson_debye100_$(tag)", sim, (0.0, 20.0), dist, 0.1, ntraj(2000);
                  output = TimeCorrelationFunctions.PopulationCorrelationFunction(sim, Diabatic()), saveat = 0.1)
    end
end

# C3  RPMD, 32 beads, Harmonic, thermal normal-mode sample drawn on the host
let T = 9.5e-4, B = 32, m = 1837.0
    sim = RingPolymerSimulation{Classical}(Atoms(m), Harmonic(m = m, ω = 0.005, r₀ = 0.1), B; temperature = T)
    dist = DynamicalDistribution(Normal(0, sqrt(T * B / m)), Normal(0.1, 0.2), size(sim))
    timed_run("rpmd_harmonic32", sim, (0.0, 25000.0), dist, 2.5, ntraj(400);
              output = (OutputCentroidPosition, OutputKineticEnergy, OutputTotalEnergy), saveat = 250.0)
end

# C4  AdiabaticIESH, MiaoSubotnik + TrapezoidalRule, M = 100 and 200
for M in (100, 200)
    Γ = 6.4e-3; W = 3Γ; kT = 9.5e-4
    model = AndersonHolstein(MiaoSubotnik(; Γ), TrapezoidalRule(M, -W, W))
    sim = Simulation{AdiabaticIESH}(Atoms(2000), model)
    dist = DynamicalDistribution(Normal(0, sqrt(kT / 2000)), 21.0, size(sim))        # ground-state orbitals (iesh.jl:89-97)
    timed_run("iesh_anderson_holstein_m$(M)", sim, (0.0, M == 100 ? 1000.0 : 200.0), dist, 1.0, ntraj(M == 100 ? 32 : 16);
              output = (OutputAdiabaticPopulation, OutputKineticEnergy), saveat = 10.0)
end

# C5  RPSH and NRPMD, 16 beads, ThreeStateMorse (docs/src/dynamicssimulations/dynamicsmethods/rpsh.md:40-68)
let T = 9.5e-4, B = 16, m = 20000.0
    for (tag, M, kw) in (("rpsh", FSSH, (;)), ("nrpmd", NRPMD, (; γ = 0.5)))
        sim = RingPolymerSimulation{M}(Atoms(m), ThreeStateMorse(), B; temperature = T, kw...)
        dist = DynamicalDistribution(Normal(0, sqrt(T * B / m)), Normal(2.1, 1 / sqrt(m * 0.005)), size(sim)) * PureState(1)
        timed_run("$(tag)_morse3_16", sim, (0.0, 3000.0), dist, 1.0, ntraj(400);
                  output = TimeCorrelationFunctions.PopulationCorrelationFunction(sim, Diabatic()), saveat = 50.0)
    end
end
