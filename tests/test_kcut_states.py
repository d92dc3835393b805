"""CPU: the k-cut "lego" initial states: LessThanK, Dicke1_2, Tensor, MaxKCutFeasible.

The reference describes them by fixed qiskit circuits; the host package evaluates those circuits' amplitudes
(qaoa_b200/utils/gatesim.py).  Pinned here by the reference's own checks -- unittests/test_maxkcut_feasible_initialstate.py:
the support avoids every infeasible colour string (binary encodings, :22-45), the one-hot node state is the W state
(:47-64) -- and by the documented meaning (equal superposition of the k feasible strings), which is what the oracle uses."""
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
from qaoa_b200 import initialstates as ist
from qaoa_b200.initialstates.tensor_initialstate import amplitudes_of


def _strings(vec):
    b = vec.size.bit_length() - 1
    return {"".join("1" if (x >> i) & 1 else "0" for i in range(b)) for x in range(vec.size) if abs(vec[x]) > 1e-12}


@pytest.mark.parametrize("k,enc", [(3, "LessThanK"), (5, "LessThanK"), (6, "LessThanK"), (7, "LessThanK"), (6, "Dicke1_2")])
def test_feasible_node_states_binary(k, enc):
    st = ist.MaxKCutFeasible(k, "binary", color_encoding=enc)
    kb = int(np.ceil(np.log2(k)))
    st.setNumQubits(kb)
    st.create_circuit()
    vec = st.state_vector()
    assert abs(np.linalg.norm(vec) - 1) < 1e-12
    assert not (_strings(vec) & set(st.infeasible))                  # the reference's assertion
    assert len(_strings(vec)) == k                                   # one string per colour ...
    assert np.allclose(vec, orc.feasible_node_state(k, enc), atol=1e-12)   # ... in equal, positive superposition


def test_feasible_node_state_onehot_is_the_w_state():
    for k in range(2, 9):
        st = ist.MaxKCutFeasible(k, "onehot")
        st.setNumQubits(k)
        vec = st.state_vector()
        expected = np.zeros(2 ** k)
        for i in range(k):
            expected[2 ** i] = 1 / np.sqrt(k)
        assert np.allclose(vec, expected)


def test_lessthank_powers_of_two_and_validation():
    for k in (2, 4, 8):
        v = ist.LessThanK(k).state_vector()
        assert np.allclose(v, np.full(k, 1 / np.sqrt(k)))
    with pytest.raises(ValueError):
        ist.LessThanK(9)
    with pytest.raises(ValueError):
        ist.MaxKCutFeasible(3, "unary")


def test_tensor_places_copy_i_on_its_own_qubits():
    one = ist.LessThanK(3)                      # {00, 10, 01} on qubits (0, 1)
    t = ist.Tensor(one, 3)
    assert t.N_qubits == 6
    vec = t.state_vector()
    assert np.allclose(vec, orc.tensor_state(orc.feasible_node_state(3), 3))
    for x in range(64):
        ok = all(((x >> (2 * v)) & 3) != 3 for v in range(3))
        assert (abs(vec[x]) > 1e-12) == ok
    mixed = ist.Tensor(ist.Dicke(1, 2), 2)      # sub-states described by kind: |D^2_1> (x) |D^2_1>
    v2 = mixed.state_vector()
    assert np.allclose(v2[[0b0101, 0b0110, 0b1001, 0b1010]], 0.5) and abs(np.linalg.norm(v2) - 1) < 1e-12
    assert np.allclose(amplitudes_of(ist.Plus(), 2), 0.5)


def test_node_grover_oracle_is_the_tensor_product_of_node_reflections():
    rng = np.random.default_rng(0)
    s = orc.feasible_node_state(5)
    psi = rng.normal(size=64) + 1j * rng.normal(size=64)
    beta = 0.7
    U1 = np.eye(8) + (np.exp(-1j * beta) - 1) * np.outer(s, s.conj())
    ref = np.kron(U1, U1) @ psi
    got = orc.apply_node_grover(psi.copy(), s, beta)
    assert np.allclose(got, ref)


def test_node_states_on_a_larger_register_leave_the_other_qubits_in_zero():
    """QAOA hands an initial state the problem's qubit count; the reference's LessThanK / Dicke1_2 circuits act on their
    first qubits only (lessthank_initialstate.py:79-152, dicke1_2_initialstate.py:28-41), powers of two put Hadamards
    on the whole register (:69-77)."""
    a = ist.LessThanK(3)
    a.setNumQubits(6)
    v = a.state_vector()
    assert v.size == 64 and np.allclose(v[:4], orc.feasible_node_state(3)) and not v[4:].any()
    b = ist.LessThanK(4)
    b.setNumQubits(5)
    assert np.allclose(b.state_vector(), 2.0 ** -2.5)
    c = ist.Dicke1_2()
    c.setNumQubits(5)
    v = c.state_vector()
    assert np.allclose(v[:8], orc.feasible_node_state(6, "Dicke1_2")) and not v[8:].any()
