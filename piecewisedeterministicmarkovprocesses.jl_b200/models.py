"""The shipped models: Python-side definitions mirroring the reference example scripts.

Each entry gives the data a user would pass to PDMPProblem (nu, xc0, xd0, parms, tspan,
n_jumps) exactly as the Julia script does, plus the name of the compiled-in CUDA device
functor that implements F / R / Delta (csrc/models.cuh).  `problem(name)` builds the
PDMPProblem of the example.
"""
import numpy as np

from .problem import PDMPProblem

# examples/ml.json:24-40 ("type II"), field order of examples/morris_lecar_variables.jl:1-17
ML_TYPE_II = dict(v_Na=3.7, g_Na=0.22, v_K=-0.9, g_K=0.4, v_L=-0.36, g_L=0.1, I_app=0.06,
                  gamma_na=1.22, k_na=-1.188, beta_na=0.1, gamma_k=0.8, k_k=-0.8, beta_k=0.04,
                  M=60.0, N=25.0)
ML_ORDER = ["v_Na", "g_Na", "v_K", "g_K", "v_L", "g_L", "I_app", "gamma_na", "k_na", "beta_na",
            "gamma_k", "k_k", "beta_k", "M", "N"]

EXAMPLES = {
    # examples/tcp.jl:47-53
    "tcp": dict(nu=[[1, 0], [0, -1]], xc0=[1.0], xd0=[0, 1], parms=[0.0], tspan=(0.0, 100000.0), n_jumps=10),
    # examples/tcp2d.jl:27-33
    "tcp2d": dict(nu=[[1, 0], [0, -1]], xc0=[0.05, 0.075], xd0=[0, 1], parms=[0.1], tspan=(0.0, 10000.0),
                  n_jumps=1000),
    # examples/morris_lecar.jl:31-48 (nu row 2 = [1 -1 0 1] is verbatim, SURVEY finding 0.7)
    "morris_lecar": dict(nu=[[-1, 1, 0, 0], [1, -1, 0, 1], [0, 0, -1, 1], [0, 0, 1, -1]], xc0=[-0.1],
                         xd0=[25, 0, 60, 0], parms=[ML_TYPE_II[k] for k in ML_ORDER], tspan=(0.0, 350.0),
                         n_jumps=2200),
    # examples/sir-rejection.jl:19-28
    "sir_rejection": dict(nu=[[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, 0, 1]], xc0=[0.0], xd0=[99, 10, 0, 0],
                          parms=[0.1 / 100.0, 0.01], tspan=(0.0, 150.0), n_jumps=1000),
    # examples/sir.jl:29-38
    "sir": dict(nu=[[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, 0, 1]], xc0=[0.0], xd0=[99, 10, 0, 0],
                parms=[0.1 / 100.0, 0.01], tspan=(0.0, 150.0), n_jumps=1000),
    # test/pdmpStiff.jl:92-99
    "pdmp_stiff": dict(nu=[[1, 0], [0, -1]], xc0=[1.0], xd0=[0, 0], parms=[0.0], tspan=(0.322156, 100000.0),
                       n_jumps=50),
    # test/pdmpStiff.jl:198-215: PDMPProblem(F!, R!, Delta!, 2, xc0, xd0, parms, (ti, tf)) — reaction_number form
    "pdmp_stiff_delta": dict(nu=2, xc0=[1.0], xd0=[0, 0], parms=[0.0], tspan=(0.322156, 100000.0), n_jumps=50),
    # test/testRatesCst.jl:5-33: rates 10 + xd[1] and parms[1], constant between jumps
    "rates_cst": dict(nu=[[1, 0], [0, -1]], xc0=[1.0], xd0=[0, 0], parms=[0.0], tspan=(0.0, 100000.0), n_jumps=14),
    # test/testRatesComposite.jl:49-55: rate 1 = 10 + xc[1] (variable), rate 2 = xd[2] (constant between jumps)
    "rates_composite": dict(nu=[[1, 0], [0, -1]], xc0=[1.0], xd0=[0, 0], parms=[0.0], tspan=(0.0, 100000.0), n_jumps=20),
    # examples/tcp_rejection.jl:67-73
    "tcp_rejection": dict(nu=[[1, 0], [0, -1]], xc0=[0.0], xd0=[0, 0], parms=[0.0], tspan=(0.0, 100000.0),
                          n_jumps=50),
    # examples/pdmpExplosion.jl:47-55
    "pdmp_explosion": dict(nu=[[1, 0], [0, -1]], xc0=[1.0], xd0=[0, 0], parms=[0.0], tspan=(0.332, 100000.0),
                           n_jumps=50),
    # examples/pdmp_example_eva.jl:47-52 (Delta resets xc on reaction 2)
    "neuron_eva": dict(nu=[[1, 0], [-1, 0], [0, 1]], xc0=[0.0], xd0=[0, 1], parms=[1.0], tspan=(0.0, 100.0),
                       n_jumps=200),
}


# examples/neuron_rejection_exact.jl:49-57: N = 100, xc0 = rand(N)*0.2 .+ 0.5 (Julia's seeded stream is not
# reproducible here: numpy PCG64 seed 1234 instead), xd0 = zeros, all-zero nu (the jumps are in Delta)
_N_MF = 100
EXAMPLES["neuron_mf"] = dict(nu=_N_MF, xc0=list(np.random.default_rng(1234).random(_N_MF) * 0.2 + 0.5),
                             xd0=[0] * _N_MF, parms=[0.1], tspan=(0.0, 10050.0), n_jumps=10000)


def problem(name, **override):
    """PDMPProblem of a reference example; F/R/Delta are the device functor's name."""
    e = dict(EXAMPLES[name])
    e.update(override)
    if isinstance(e["nu"], int):   # reaction_number constructor (src/problem.jl:190-193): the jumps live in Delta
        return PDMPProblem(name, name, name, e["nu"], e["xc0"], e["xd0"], e["parms"], e["tspan"])
    return PDMPProblem(name, name, np.array(e["nu"]), e["xc0"], e["xd0"], e["parms"], e["tspan"])
