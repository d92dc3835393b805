"""A few-qubit statevector calculator for the FIXED state-preparation circuits of the reference's k-cut initial
states (LessThanK, Dicke1_2: reference qaoa/initialstates/lessthank_initialstate.py:79-152,
dicke1_2_initialstate.py:24-41).  Those classes describe a state by a qiskit circuit on 2 or 3 qubits; the engine
needs the amplitudes, so the handful of gates they use is evaluated here, on the host, once per state (8 amplitudes
at most).  This is plan construction, not the hot path: nothing of a QAOA evaluation runs through it.

Conventions are qiskit's: qubit 0 is the least significant index bit; a two-qubit gate applied to (a, b) has its
matrix indexed by bit_a + 2 * bit_b."""
import numpy as np


class SmallState:
    def __init__(self, n):
        self.n = n
        self.v = np.zeros(1 << n, dtype=np.complex128)
        self.v[0] = 1.0

    def _apply(self, mat, qubits, ctrl=None, ctrl_state=1):
        """mat: 2^k x 2^k on `qubits` (first = least significant), optionally controlled on qubit `ctrl` == ctrl_state"""
        k = len(qubits)
        new = self.v.copy()
        for base in range(1 << self.n):
            if any((base >> q) & 1 for q in qubits):
                continue
            if ctrl is not None and ((base >> ctrl) & 1) != ctrl_state:
                continue
            idx = [base | sum(((j >> t) & 1) << qubits[t] for t in range(k)) for j in range(1 << k)]
            new[idx] = mat @ self.v[idx]
        self.v = new

    def x(self, q):
        self._apply(np.array([[0, 1], [1, 0]], dtype=complex), [q])

    def h(self, q, ctrl=None, ctrl_state=1):
        self._apply(np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2), [q], ctrl, ctrl_state)

    def ry(self, theta, q, ctrl=None, ctrl_state=1):
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        self._apply(np.array([[c, -s], [s, c]], dtype=complex), [q], ctrl, ctrl_state)

    def cx(self, c, t):
        self._apply(np.array([[0, 1], [1, 0]], dtype=complex), [t], c, 1)

    def cz(self, a, b):
        self._apply(np.diag([1, -1]).astype(complex), [b], a, 1)

    def xx_plus_yy(self, theta, beta, a, b):
        """XXPlusYYGate(theta, beta) on (a, b): rotation inside span{|01>, |10>}"""
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        m = np.eye(4, dtype=complex)
        m[1, 1] = c
        m[1, 2] = -1j * s * np.exp(-1j * beta)
        m[2, 1] = -1j * s * np.exp(1j * beta)
        m[2, 2] = c
        self._apply(m, [a, b])

    def evolve_yyy(self, time):
        """PauliEvolutionGate(Y^Y^Y, time) = exp(-i time Y(x)Y(x)Y) on three qubits"""
        Y = np.array([[0, -1j], [1j, 0]])
        P = np.kron(np.kron(Y, Y), Y)
        self._apply(np.cos(time) * np.eye(8) - 1j * np.sin(time) * P, [0, 1, 2])
