"""Algorithm types hung under the reference's type tree
(reference src/PiecewiseDeterministicMarkovProcesses.jl:10-15):

    AbstractPDMPAlgorithm <- AbstractCHV       <- CHVGPU
                          <- AbstractRejection <- RejectionGPU
                                               <- AbstractRejectionExact <- RejectionExactGPU

`CHVGPU(model, ode="dp5")` replaces `CHV(Tsit5())` (src/chvdiffeq.jl:2-4), and
`RejectionGPU(model, ode="dp5")` replaces `Rejection(Tsit5())` (src/rejectiondiffeq.jl:1-3),
for an ensemble of trajectories on the GPU.  `model` names the compiled-in CUDA device
functor for F/R/Delta.  There is no CPU fallback: solving without a GPU raises.
"""


class AbstractPDMPAlgorithm:
    pass


class AbstractCHV(AbstractPDMPAlgorithm):
    pass


class AbstractRejection(AbstractPDMPAlgorithm):
    pass


class AbstractRejectionExact(AbstractRejection):
    pass


class _GPUAlg:
    def __init__(self, model, ode="dp5", precision="fp64"):
        if ode not in ("dp5", "tsit5"):
            raise ValueError("ode must be 'dp5' (Dormand-Prince 5(4)) or 'tsit5' (Tsitouras 5(4))")
        self.model, self.ode, self.precision = str(model), ode, precision

    def __repr__(self):
        return f"{type(self).__name__}({self.model!r}, ode={self.ode!r}, precision={self.precision!r})"


class CHVGPU(_GPUAlg, AbstractCHV):
    algorithm = "chv"


class RejectionGPU(_GPUAlg, AbstractRejection):
    algorithm = "rejection"


class RejectionExactGPU(_GPUAlg, AbstractRejectionExact):
    """`RejectionExactGPU(model)` replaces `RejectionExact()` (reference src/rejection.jl:1, :3-91): thinning
    with the model's EXACT flow instead of an ODE solve, one warp per trajectory (csrc/exact_kernel.cuh)."""
    algorithm = "rejection_exact"

    def __init__(self, model):
        super().__init__(model, ode="dp5", precision="fp64")
