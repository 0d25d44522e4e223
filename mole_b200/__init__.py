"""mole_b200 — B200-native (sm_100a) walker-ensemble VMC/DMC hot path of Jvanrhijn/mole.

`from mole_b200 import *` plays the role of `use mole::prelude::*` (reference src/lib.rs:10-21).
All physics runs in libmole_b200.so (hand-written CUDA); this package is the thin host mirror of
the reference's trait surface.  No CPU fallback exists.
"""
from . import ffi
from .ffi import MoleError
from .api import (Context, default_context, comm_unique_id, derive_seed, WaveFunction, STO, GaussianWaveFunction, Hydrogen1sBasis, Orbital, SingleDeterminant,
                  SpinDeterminantProduct,
                  HeliumAtomWaveFunction, HydrogenMoleculeWaveFunction, H2WF, SlaterJastrow, LcaoSlaterJastrow, WaveFunctionMock,
                  LocalOperator, KineticEnergy, IonicPotential, ElectronicPotential, IonicHamiltonian,
                  ElectronicHamiltonian, HarmonicHamiltonian, ParameterGradient, WavefunctionValue, operators,
                  MetropolisBox, MetropolisDiffuse, Ensemble, acc_finalize, gram_finalize, rebalance_plan, series_block_sizes, Optimizer, SteepestDescent, MomentumDescent,
                  NesterovMomentum, OnlineLbfgs, StochasticReconfiguration, MonteCarloResult, Sampler, Runner,
                  VmcRunner, SRBrancher, SimpleBranching, DmcRunner)

__all__ = [n for n in dir() if not n.startswith("_")]
