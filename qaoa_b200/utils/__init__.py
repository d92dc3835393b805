from .statistic import Statistic, ExactStatistic
from .graphutils import GraphHandler
from .optimizers import COBYLA, NELDER_MEAD, L_BFGS_B, SPSA
from .flip import BitFlip
from .post import post_processing, post_process_all_depths
from . import validation
