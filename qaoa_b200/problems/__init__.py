from .base_problem import Problem
from .qubo_problem import QUBO
from .graph_problem import GraphProblem
from .exactcover_problem import ExactCover
from .portfolio_problem import PortfolioOptimization
from .maxkcut_binary_powertwo import MaxKCutBinaryPowerOfTwo
from .maxkcut_binary_fullH import MaxKCutBinaryFullH
from .maxcut_problem import MaxCut
