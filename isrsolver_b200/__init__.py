"""isrsolver_b200 -- B200-native (sm_100a, FP64) implementation of ISRSolver's data-parallel hot path.

The package holds the CUDA kernels + C ABI (csrc/ -> libisr_b200.so, declared in include/isr_b200.h),
the host-side mirror of the reference's PyISR interface (pyisr.py) and the one-process-per-GPU
sharding of toy campaigns (dist.py).  Importing the package does not load the shared library; the first
solver or kernel call does, and fails loudly if it is not built or no B200 is visible.
"""
from .pyisr import (DifferentModelCrossSectionSizes, Efficiency, EfficiencyDimensionException,  # noqa: F401
                    InterpRangeException, IsrGpuError, ISRSolverSLE, ISRSolverSLE2, ISRSolverTikhonov,
                    ISRSolverTSVD, IterISRInterpSolver, ConvolutionPlan, convolutionKuraevFadin, convolutionKuraevFadinBlurred,
                    draw_toys, draw_toys_device, kernelKuraevFadin)

__all__ = ["ISRSolverSLE", "ISRSolverSLE2", "ISRSolverTikhonov", "ISRSolverTSVD", "IterISRInterpSolver", "Efficiency",
           "kernelKuraevFadin", "convolutionKuraevFadin", "convolutionKuraevFadinBlurred", "ConvolutionPlan", "draw_toys", "draw_toys_device", "InterpRangeException", "EfficiencyDimensionException",
           "DifferentModelCrossSectionSizes", "IsrGpuError"]
