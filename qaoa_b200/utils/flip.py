"""BitFlip: greedy random single-bit flips that raise the cost of a sampled bit-string -- the classical helper behind
the reference's `post` option (reference qaoa/utils/flip.py:5-88).  Strings are in the sampler's printing order (qubit 0
LAST), as in the reference: costs are taken of the reversed string.  Host-only; nothing here touches the device.

`create_circuit` keeps its signature but records the qubits to flip (`flip_qubits`, `flip_mask`) instead of building an
X-gate circuit: the engine executes no gates, `QAOA(flip=True)` hands the mask to the device, which applies the X gates
between two layers as an index permutation of the kept state (Engine.continue_from_kept_state)."""
import numpy as np


class BitFlip:
    def __init__(self, n):
        self.circuit = None
        self.N_qubits = n
        self.flip_qubits = []

    def boost_samples(self, problem, string, K=5):
        """K passes over the qubits in random order; a flip is kept whenever it increases problem.cost"""
        bits = [int(ch) for ch in string]
        best = problem.cost(string[::-1])
        for _ in range(int(K)):
            order = np.arange(self.N_qubits)
            np.random.shuffle(order)
            for i in order:
                trial = list(bits)
                trial[i] = 1 - trial[i]
                value = problem.cost("".join(map(str, trial))[::-1])
                if value > best:
                    best, bits = value, trial
        return "".join(map(str, bits))

    def xor(self, old_string, new_string):
        """position-wise XOR of two strings as a list of 0/1 flags (1 at position n-1-i: qubit i differs)"""
        return [int(a != b) for a, b in zip(old_string, new_string)]

    def create_circuit(self, xor) -> None:
        """qubits that get an X gate: the 1-flags of ``xor`` read from the right (position n-1-i <-> qubit i).  As in the
        reference (flip.py:86-88, `if np.any(indices_flip)`), a pattern that names qubit 0 ALONE flips nothing."""
        idx = np.where(np.asarray(xor)[::-1])[0]
        self.flip_qubits = [int(i) for i in idx] if np.any(idx) else []
        self.circuit = None

    @property
    def flip_mask(self):
        """bit mask of the qubits of the last create_circuit call (bit q = qubit q)"""
        return sum(1 << q for q in self.flip_qubits)
