"""Exact check that the device's phase separator represents the problem's cost function -- the reference's
`validate_circuit` (qaoa/utils/validation.py:16-79, Problem.validate_circuit base_problem.py:136-149).

The reference builds the 2^n x 2^n unitary of the phase-separator circuit and compares its diagonal with
exp(+i t cost(e)) up to a global phase (the docstring says exp(-j t cost), the code checks `np.exp(1j * t * costs)`, :43).
The engine has no circuit: its phase separator IS the diagonal exp(+i t c_dev(x)) of the cost diagonal it generated on the
device, so the same comparison is made with that diagonal read back (same arguments, same report keys, plus the largest
absolute difference of the costs themselves).  Brute force over 2^n strings: <= 16 qubits.
"""
import numpy as np


def _bitstring(i, n, flip=False):
    s = format(i, "0%db" % n)
    return s[::-1] if flip else s


def check_phase_separator_exact_qaoa(qaoa, *arg, **kwarg):
    return check_phase_separator_exact_problem(qaoa.problem, *arg, engine=qaoa._get_engine(), lowered=True, **kwarg)


def check_phase_separator_exact_problem(problem, t=1, flip=True, atol=1e-8, rtol=1e-8, engine=None, lowered=False):
    """Returns (ok, report).  ``engine``: an engine to lower the problem onto (default: a fresh complex128 engine on
    device 0); ``lowered``: the engine already carries this problem's cost."""
    n = problem.N_qubits
    if n > 16:
        raise ValueError("validate_circuit is a brute-force check for <= 16 qubits")
    if engine is None:
        from qaoa_b200.engine import Engine

        engine = Engine(n, dtype="complex128")
    if not lowered:
        problem.lower_cost(engine)
    dev = np.asarray(engine.read_cost_diagonal(), dtype=np.float64)
    costs = np.array([problem.cost(_bitstring(i, n, flip=flip)) for i in range(1 << n)], dtype=np.float64)
    expected = np.exp(1j * t * costs)
    diag = np.exp(1j * t * dev)
    g = diag[0] / expected[0]                       # global phase factor
    ratios = diag / (expected * g)
    mag_err = float(np.max(np.abs(np.abs(diag) - 1.0)))
    phase_err = float(np.max(np.abs(np.angle(ratios))))
    ok = (mag_err <= rtol) and (phase_err <= atol)
    report = {
        "n_qubits": n,
        "max_magnitude_error": mag_err,
        "max_phase_error_rad_after_global": phase_err,
        "global_phase_rad": float(np.angle(g)),
        "max_cost_error": float(np.max(np.abs(dev - costs))),
    }
    if not ok:
        worst = np.argsort(-np.abs(np.angle(ratios)))[:8]
        report["examples"] = [{"bitstring": list(_bitstring(int(k), n, flip=flip)), "diag_entry": complex(diag[k]),
                               "expected": complex(expected[k] * g), "phase_residual_rad": float(np.angle(ratios[k])),
                               "magnitude": float(np.abs(diag[k]))} for k in worst]
    return ok, report
