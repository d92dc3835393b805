"""|+>^n (reference qaoa/initialstates/plus_initialstate.py:19-25)."""

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng


class Plus(InitialState):
    def __init__(self) -> None:
        super().__init__()

    def state_descriptor(self):
        return _eng.INIT_PLUS, 0, None
