"""nqcdynamics.jl_b200 -- B200-native ensemble-trajectory engine behind NQCDynamics.jl's
``run_dynamics(...; ensemble_algorithm=EnsembleB200(ngpus))`` seam.

Only the hot path lives here: ``csrc/`` (sm_100a CUDA kernels + the C ABI of ``include/nqcb200.h``) and a thin
Python mirror of the reference's user-facing interface for that path (models, Simulation types, distributions,
outputs, reductions, ``run_dynamics``).  There is no CPU fallback: without the CUDA library and a GPU every
compute call raises.
"""
from . import _abi
from ._abi import EngineError
from .models import *  # noqa: F401,F403
from .api import (Atoms, FSSH, Ehrenfest, EhrenfestNA, ThermalLangevin, Classical, AdiabaticIESH, NRPMD, Simulation, RingPolymerSimulation, Normal,
                  VelocityBoltzmann, Diabatic, Adiabatic, PureState, MixedState, FermiDiracState, DynamicalDistribution, ProductDistribution,
                  OutputDiabaticPopulation, OutputAdiabaticPopulation, OutputKineticEnergy, OutputPotentialEnergy,
                  OutputTotalEnergy, OutputPosition, OutputVelocity, OutputCentroidPosition, OutputCentroidVelocity,
                  OutputDiscreteState, OutputQuantumSubsystem, OutputSurfaceHops, OutputStateResolvedScattering1D,
                  OutputFinalKineticEnergy, OutputFirstPosition, OutputFirstVelocity, OutputFinalPosition,
                  OutputFinalVelocity, OutputTotalDiabaticPopulation, OutputTotalAdiabaticPopulation, OutputSpringEnergy,
                  OutputCentroidKineticEnergy, OutputFinalTime, OutputDynamicsVariables, OutputInitial, OutputFinal,
                  PopulationCorrelationFunction, SortByTrajectoryReduction, SortByOutputReduction, SumReduction, FileReduction,
                  MeanReduction, EnsembleB200, run_dynamics, TerminatingCallback, PositionOutside,
                  OutputSubsetKineticEnergy, OutputFinalSubsetKineticEnergy, OutputKineticTemperature,
                  OutputMappingPosition, OutputMappingMomentum, OutputOccupations)
