"""Result files of QAOA runs: the JSON schema of the reference's qaoa/utils/qaoaIO.py:56-281, so that files written
by either package load in the other.

  {"problem":     {"problem_type": "ExactCover" | "PortfolioOptimization", <problem fields>},
   "qaoa_params": {"cvar", "init_method": "PLUS" | "DICKE", "mixer_method": "X" | "GROVER" | "XYCHAIN" | "XYRING",
                   "backend", "optimizer", "N_qubits",
                   "depths": {"<depth>": {"optimal_angles": [...], "histogram": {bitstring: count}, "opt_time": s}}},
   "metadata":    {"timestamp", "system", "release", "python_version", "machine", "processor", ...}}

Same public names as the reference module (QAOAResult.save / load / from_qaoa / get_problem_instance, ProblemData and
its registry, DepthResult, QAOAParameters, InitMethod, MixerMethod).  The histograms of `from_qaoa` come from
QAOA.hist, i.e. from the on-device sampler."""
import json
import os
import platform
import subprocess
from dataclasses import dataclass, field, fields
from datetime import datetime
from enum import Enum
from typing import Dict, List, Type

import numpy as np


class InitMethod(Enum):
    PLUS = "PLUS"
    DICKE = "DICKE"


class MixerMethod(Enum):
    X = "X"
    GROVER = "GROVER"
    XYCHAIN = "XYCHAIN"
    XYRING = "XYRING"


@dataclass
class DepthResult:
    optimal_angles: List[float]
    histogram: Dict[str, int]
    opt_time: float  # seconds


@dataclass
class QAOAParameters:
    cvar: float
    init_method: InitMethod
    mixer_method: MixerMethod
    backend: str
    optimizer: str
    N_qubits: int
    depths: Dict[int, DepthResult] = field(default_factory=dict)


def _class_token(obj):
    """'Dicke' / 'Grover' / ... from an instance: the class name, which is also what the reference derives from str(obj)"""
    return type(obj).__name__.upper()


def run_metadata() -> dict:
    """when / where / on which commit a result was produced (the reference's metadata keys)"""
    here = os.path.dirname(os.path.abspath(__file__))
    try:
        commit = subprocess.check_output(["git", "rev-parse", "HEAD"], cwd=here, stderr=subprocess.DEVNULL).decode().strip()
    except Exception:
        commit = "unknown"
    return {
        "timestamp": datetime.now().isoformat(timespec="seconds"),
        "system": platform.system(), "release": platform.release(), "python_version": platform.python_version(),
        "machine": platform.machine(), "processor": platform.processor(),
        "conda_env": os.environ.get("CONDA_DEFAULT_ENV", "unknown"),
        "qaoa_repo_dir": here, "qaoa_git_commit": commit,
    }


def _plain(obj):
    """numpy arrays / scalars / enums -> what json.dump accepts"""
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, Enum):
        return obj.name
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


def _arrays(obj):
    """numeric (nested) lists of a loaded problem -> numpy arrays; anything ragged or non-numeric stays a list"""
    if isinstance(obj, dict):
        return {k: _arrays(v) for k, v in obj.items()}
    if isinstance(obj, list):
        if all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in obj):
            return np.array(obj)
        inner = [_arrays(v) for v in obj]
        if inner and all(isinstance(v, np.ndarray) for v in inner) and len({v.shape for v in inner}) == 1:
            return np.array(inner)
        return inner
    return obj


@dataclass
class ProblemData:
    """Problem fields stored next to the results; subclasses are looked up by `problem_type` on load."""
    problem_type: str = "Base"

    def to_dict(self):
        return _plain({f.name: getattr(self, f.name) for f in fields(self)})

    @classmethod
    def from_dict(cls, data: dict) -> "ProblemData":
        kind = data.get("problem_type", "Base")
        if kind not in problem_registry:
            raise ValueError(f"Unknown problem type for IO: {kind}")
        body = _arrays({k: v for k, v in data.items() if k != "problem_type"})
        return problem_registry[kind](**body)


@dataclass
class ExactCoverProblemData(ProblemData):
    columns: np.ndarray = None
    weights: np.ndarray = None
    solution: np.ndarray = None
    hamming_weight: int = None
    problem_type: str = "ExactCover"


@dataclass
class PortfolioOptimizationProblemData(ProblemData):
    risk: float = 0.0
    exp_returns: np.ndarray = None
    cov_matrix: np.ndarray = None
    budget: int = 0
    problem_type: str = "PortfolioOptimization"


problem_registry: Dict[str, Type[ProblemData]] = {
    "ExactCover": ExactCoverProblemData,
    "PortfolioOptimization": PortfolioOptimizationProblemData,
}



@dataclass
class QAOAResult:
    problem: ProblemData
    qaoa_params: QAOAParameters
    metadata: Dict[str, str] = field(default_factory=dict)

    def __post_init__(self):
        if not self.metadata:
            self.metadata = self._generate_metadata()

    # ---- file format ---------------------------------------------------------------------------------------
    def to_json_dict(self):
        qp = self.qaoa_params
        return {
            "problem": self.problem.to_dict(),
            "qaoa_params": {
                "cvar": _plain(qp.cvar), "init_method": qp.init_method.name, "mixer_method": qp.mixer_method.name,
                "backend": qp.backend, "optimizer": qp.optimizer, "N_qubits": int(qp.N_qubits),
                "depths": {int(d): {"optimal_angles": _plain(list(r.optimal_angles)), "histogram": _plain(dict(r.histogram)),
                                    "opt_time": float(r.opt_time)} for d, r in qp.depths.items()},
            },
            "metadata": self.metadata,
        }

    def save(self, filename: str):
        with open(filename, "w") as f:
            json.dump(self.to_json_dict(), f, indent=2)

    @classmethod
    def load(cls, filename: str) -> "QAOAResult":
        with open(filename, "r") as f:
            raw = json.load(f)
        qp = raw["qaoa_params"]
        params = QAOAParameters(
            cvar=qp["cvar"], init_method=InitMethod[qp["init_method"]], mixer_method=MixerMethod[qp["mixer_method"]],
            backend=qp["backend"], optimizer=qp["optimizer"], N_qubits=qp["N_qubits"],
            depths={int(d): DepthResult(**r) for d, r in qp["depths"].items()})
        return cls(problem=ProblemData.from_dict(raw["problem"]), qaoa_params=params, metadata=raw.get("metadata", {}))

    def _generate_metadata(self) -> dict:
        return run_metadata()

    # ---- from / to live objects ----------------------------------------------------------------------------
    @classmethod
    def from_qaoa(cls, qaoa, hist_shots: int = 1024, solution: str = None) -> "QAOAResult":
        """Collects the best angles, a sampled histogram and the optimisation time of every depth of an ExactCover run
        (reference qaoaIO.py:218-261; the histogram is drawn on the device, QAOA.hist)."""
        depths = {}
        for d, res in qaoa.optimization_results.items():
            best = res.get_best_angles()
            depths[d] = DepthResult(optimal_angles=best, histogram=qaoa.hist(best, hist_shots), opt_time=res.opt_time)
        problem = ExactCoverProblemData(columns=qaoa.problem.columns, weights=qaoa.problem.weights, solution=solution,
                                        hamming_weight=qaoa.problem.hamming_weight)
        init_method = InitMethod(_class_token(qaoa.initialstate))
        mixer = _class_token(qaoa.mixer)
        if mixer == "XY":
            mixer_method = {"ring": MixerMethod.XYRING, "chain": MixerMethod.XYCHAIN}.get(qaoa.mixer.case)
        else:
            mixer_method = MixerMethod(mixer)
        opt_name = getattr(qaoa.optimizer[0], "__name__", str(qaoa.optimizer))
        assert "COBYLA" in opt_name or "COBYLA" in str(qaoa.optimizer), f"unsupported optimizer {qaoa.optimizer} in qaoaIO"
        params = QAOAParameters(cvar=qaoa.cvar, init_method=init_method, mixer_method=mixer_method,
                                backend=qaoa.backend.name, optimizer="COBYLA", N_qubits=qaoa.problem.N_qubits, depths=depths)
        return cls(problem=problem, qaoa_params=params)

    def get_problem_instance(self):
        from qaoa_b200 import problems

        if isinstance(self.problem, ExactCoverProblemData):
            return problems.ExactCover(columns=self.problem.columns, weights=self.problem.weights,
                                       hamming_weight=self.problem.hamming_weight, scale_problem=True)
        if isinstance(self.problem, PortfolioOptimizationProblemData):
            # (the reference reads a field `exp_return` that its dataclass does not have, qaoaIO.py:92, 279)
            return problems.PortfolioOptimization(risk=self.problem.risk, budget=self.problem.budget,
                                                  cov_matrix=self.problem.cov_matrix, exp_return=self.problem.exp_returns)
        raise ValueError("no problem class for %s" % type(self.problem).__name__)
