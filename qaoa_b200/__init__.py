"""B200-native QAOA statevector engine behind the OpenQuantumComputing/QAOA API.

    from qaoa_b200 import QAOA, problems, mixers, initialstates

mirrors ``from qaoa import QAOA, problems, mixers, initialstates`` of the reference.
"""

from . import initialstates, mixers, problems, utils
from .qaoa import QAOA, OptResult, B200Backend

__all__ = ["QAOA", "OptResult", "B200Backend", "problems", "mixers", "initialstates", "utils"]
__version__ = "0.1.0"
