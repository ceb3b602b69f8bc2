"""pdmp_b200 — B200-native ensemble engine for the hot path of
PiecewiseDeterministicMarkovProcesses.jl: `solve(problem, CHVGPU(model))` /
`solve(problem, RejectionGPU(model))` over many independent trajectories.

Host-side mirror (Python, ctypes) of the reference's Julia surface; the compute is in
csrc/ (CUDA, sm_100a) behind the C ABI of include/pdmp_b200.h.  Importing this package
does not load the CUDA library; the first solve does, and raises if it is missing.
"""
from . import _abi, _marshal, distributed, models
from .algorithms import (AbstractCHV, AbstractPDMPAlgorithm, AbstractRejection, AbstractRejectionExact, CHVGPU,
                         RejectionExactGPU, RejectionGPU)
from .problem import EnsembleResult, PDMPProblem, PDMPResult
from .rate import CompositeRate, ConstantRate, VariableRate
from .solve import Ensemble, solve
from ._lib import PDMPError, device_count, fp64_peak_tflops, list_models, model_info

__all__ = ["PDMPProblem", "PDMPResult", "EnsembleResult", "CHVGPU", "RejectionGPU", "RejectionExactGPU", "solve", "Ensemble",
           "models", "distributed", "PDMPError", "device_count", "fp64_peak_tflops", "list_models", "model_info",
           "AbstractPDMPAlgorithm", "AbstractCHV", "AbstractRejection", "VariableRate", "ConstantRate", "CompositeRate"]
