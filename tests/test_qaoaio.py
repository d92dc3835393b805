"""CPU: result files.  qaoa_b200.utils.qaoaIO reads and writes the reference's JSON schema
(qaoa/utils/qaoaIO.py:56-281): pinned by a result file the REFERENCE wrote (tests/golden/qaoaio_result_example.json,
extracted from examples/ExactCover/data/ by tests/golden/make_golden.py) with the assertions of the reference's own
unittests/test_qaoaIO_roundtrip.py, and by a live run through QAOAResult.from_qaoa."""
import json
import os

import numpy as np
import pytest

from qaoa_b200 import initialstates, mixers, problems
from qaoa_b200.utils import qaoaIO as qio
from test_host_api import make

EXAMPLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qaoaio_result_example.json")


def test_load_fields_of_a_reference_written_file():
    result = qio.QAOAResult.load(EXAMPLE)
    assert isinstance(result.problem, qio.ExactCoverProblemData)
    assert result.problem.problem_type == "ExactCover"
    assert result.problem.hamming_weight == 2
    assert isinstance(result.problem.columns, np.ndarray) and isinstance(result.problem.weights, np.ndarray)
    assert result.problem.columns.shape[1] == result.qaoa_params.N_qubits == 8
    assert result.qaoa_params.mixer_method == qio.MixerMethod.GROVER
    assert result.qaoa_params.init_method == qio.InitMethod.DICKE
    assert sorted(result.qaoa_params.depths) == [1, 2, 3]
    assert "timestamp" in result.metadata
    d1 = result.qaoa_params.depths[1]
    assert len(d1.optimal_angles) == 2 and sum(d1.histogram.values()) == 8192 and d1.opt_time > 0


def test_roundtrip_preserves_data(tmp_path):
    original = qio.QAOAResult.load(EXAMPLE)
    out = tmp_path / "again.json"
    original.save(str(out))
    reloaded = qio.QAOAResult.load(str(out))
    np.testing.assert_array_almost_equal(reloaded.problem.columns, original.problem.columns)
    np.testing.assert_array_almost_equal(reloaded.problem.weights, original.problem.weights)
    np.testing.assert_array_almost_equal(reloaded.problem.solution, original.problem.solution)
    assert reloaded.problem.hamming_weight == original.problem.hamming_weight
    qa, qb = reloaded.qaoa_params, original.qaoa_params
    assert (qa.cvar, qa.init_method, qa.mixer_method, qa.backend, qa.optimizer, qa.N_qubits) == \
           (qb.cvar, qb.init_method, qb.mixer_method, qb.backend, qb.optimizer, qb.N_qubits)
    assert set(qa.depths) == set(qb.depths)
    for d in qb.depths:
        np.testing.assert_array_almost_equal(qa.depths[d].optimal_angles, qb.depths[d].optimal_angles)
        assert qa.depths[d].histogram == qb.depths[d].histogram
        assert qa.depths[d].opt_time == pytest.approx(qb.depths[d].opt_time)
    assert reloaded.metadata == original.metadata
    # the file itself keeps the reference's layout: same keys at every level, depths keyed by strings
    a, b = json.load(open(EXAMPLE)), json.load(open(out))
    assert a.keys() == b.keys() and a["problem"].keys() == b["problem"].keys()
    assert a["qaoa_params"].keys() == b["qaoa_params"].keys()
    assert a["qaoa_params"]["depths"]["1"].keys() == b["qaoa_params"]["depths"]["1"].keys()
    assert b["qaoa_params"]["init_method"] == "DICKE" and b["qaoa_params"]["mixer_method"] == "GROVER"


def test_problem_instance_and_unknown_types():
    result = qio.QAOAResult.load(EXAMPLE)
    prob = result.get_problem_instance()
    assert isinstance(prob, problems.ExactCover) and prob.N_qubits == 8 and prob.hamming_weight == 2
    np.testing.assert_array_almost_equal(prob.columns, result.problem.columns)
    assert prob.isFeasible(prob.brute_force_solve())      # reference unittests/test_qaoaIO_roundtrip.py:96-108
    with pytest.raises(ValueError):
        qio.ProblemData.from_dict({"problem_type": "TravellingSalesman"})
    pf = qio.PortfolioOptimizationProblemData(risk=0.5, exp_returns=np.array([0.1, 0.2]),
                                              cov_matrix=np.eye(2), budget=1)
    back = qio.ProblemData.from_dict(pf.to_dict())
    assert isinstance(back, qio.PortfolioOptimizationProblemData) and back.cov_matrix.shape == (2, 2)
    inst = qio.QAOAResult(problem=back, qaoa_params=result.qaoa_params).get_problem_instance()
    assert isinstance(inst, problems.PortfolioOptimization)


@pytest.mark.parametrize("mixer_name", ["GROVER", "XYRING", "XYCHAIN"])
def test_from_qaoa_collects_a_live_run(tmp_path, mixer_name):
    loaded = qio.QAOAResult.load(EXAMPLE)
    prob = problems.ExactCover(columns=loaded.problem.columns.astype(int), weights=loaded.problem.weights,
                               hamming_weight=2)
    mixer = {"GROVER": lambda: mixers.Grover(initialstates.Dicke(2)), "XYRING": lambda: mixers.XY(case="ring"),
             "XYCHAIN": lambda: mixers.XY(case="chain")}[mixer_name]()
    q, fake = make(prob, mixer=mixer, init=initialstates.Dicke(2),
                   optimizer=[__import__("qaoa_b200").utils.COBYLA, {"maxiter": 15}], cvar=0.7)
    q.optimize(depth=2, angles={"gamma": [0, 2 * np.pi, 5], "beta": [0, 2 * np.pi, 5]})
    res = qio.QAOAResult.from_qaoa(q, hist_shots=256, solution=[1.0, 0, 0, 1.0, 0, 0, 0, 0])
    assert res.qaoa_params.mixer_method == qio.MixerMethod[mixer_name]
    assert res.qaoa_params.init_method == qio.InitMethod.DICKE
    assert res.qaoa_params.backend == q.backend.name and res.qaoa_params.optimizer == "COBYLA"
    assert sorted(res.qaoa_params.depths) == [1, 2]
    for d, r in res.qaoa_params.depths.items():
        assert len(r.optimal_angles) == 2 * d and sum(r.histogram.values()) == 256 and r.opt_time >= 0
        assert all(s.count("1") == 2 for s in r.histogram)      # Dicke + Grover / XY keep the Hamming weight
    out = tmp_path / "live.json"
    res.save(str(out))
    again = qio.QAOAResult.load(str(out))
    assert again.qaoa_params.depths[2].histogram == res.qaoa_params.depths[2].histogram
    assert again.qaoa_params.cvar == 0.7
