"""Rate wrappers of the reference (src/rate.jl).

    VariableRate(R)       src/rate.jl:8-14    the default: R is evaluated whenever the solver asks
    ConstantRate(R)       src/rate.jl:17-40   the total rate is constant between jumps (it depends on xd
                                              only): the solver caches it from one jump to the next
    CompositeRate(Rc, Rv) src/rate.jl:42-69   a constant part plus a variable part

With the GPU algorithms R names a compiled-in device functor, so the wrapper only selects how the
kernel uses it: ConstantRate makes the CHV field re-use the total rate computed at the post-jump
state (identical results — reference test/testRatesCst.jl:62-72 — for fewer rate evaluations).
It is accepted only for device functors whose rates depend on xd alone.
"""


class AbstractRate:
    def __init__(self, R):
        self.R = R


class VariableRate(AbstractRate):
    rate_kind = 0


class ConstantRate(AbstractRate):
    rate_kind = 1


class CompositeRate(AbstractRate):
    rate_kind = 2

    def __init__(self, Rc, Rv):
        self.Rcst = Rc if isinstance(Rc, AbstractRate) else ConstantRate(Rc)
        self.Rvar = Rv if isinstance(Rv, AbstractRate) else VariableRate(Rv)
        self.R = (self.Rcst, self.Rvar)
