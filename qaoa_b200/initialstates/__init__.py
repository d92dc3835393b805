from .base_initialstate import InitialState
from .plus_initialstate import Plus
from .dicke_initialstate import Dicke
from .statevector_initialstate import StateVector
from .plus_parameterized_initialstate import PlusParameterized
