from .base_mixer import Mixer
from .xy_mixer import XY
from .x_mixer import X
from .x_multiangle_mixer import XMultiAngle
from .grover_mixer import Grover
from .x_orbit_mixer import XOrbit
from .maxkcut_grover_mixer import MaxKCutGrover
from .xy_tensor import XYTensor
