"""CPU: the algebra behind the sharded symmetric register (csrc/tile_kernel.cuh "twisted
coordinates"), restated in numpy and checked against the oracle.

The engine stores phi~(x') = i^popc(x') psi(x', top = 0) * sqrt(2) for the n-1 low qubits x' only and splits it over
2^g ranks as (r, z): z = the EL = n-1-g low bits of x', r_j = x'[EL+j] XOR x'[0].  This file applies every operator in
exactly the form the kernels use -- real rotations in the S frame, the mirror pairing with sigma, qubit 0 as the
pairing (z_0 = 0, r) <-> (z_0 = 1, ~r), rank qubits with the sign twist -- on a [rank][z] array, so the formulas are
pinned independently of the CUDA code (which tests/test_gpu_sharded.py checks on the device)."""
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
import parity_helpers as H


def _popc(a):
    return np.array([bin(int(v)).count("1") for v in np.ravel(a)]).reshape(np.shape(a))


def twisted_energy(n, g, diag, angles, n_beta):
    el, G, mask = n - 1 - g, 1 << g, (1 << g) - 1
    z = np.arange(1 << el)
    xs = np.stack([z | ((np.where(z & 1, r ^ mask, r)) << el) for r in range(G)])      # x'(r, z), top qubit = 0
    pc = _popc(xs)
    cost = np.asarray(diag)[xs]
    phi = (np.sqrt(2.0) * 2.0 ** (-0.5 * n)) * (1j) ** (pc % 4)                          # |+>^n in the S frame
    sigma = np.real((1j) ** ((2 - n) % 4)) * (-1.0) ** pc                               # i^(2-n) (-1)^popc(x'), n even
    per = 1 + n_beta
    for d in range(len(angles) // per):
        gamma = angles[d * per]
        beta = lambda q: angles[d * per + 1 + (q if n_beta > 1 else 0)]
        phi = phi * np.exp(1j * gamma * cost)
        for q in range(1, el):                       # local qubits: a' = c a + s b, b' = c b - s a (a: bit = 0)
            c, s = np.cos(beta(q)), np.sin(beta(q))
            lo = z[(z >> q) & 1 == 0]
            a, b = phi[:, lo].copy(), phi[:, lo | (1 << q)].copy()
            phi[:, lo], phi[:, lo | (1 << q)] = c * a + s * b, c * b - s * a
        c, s = np.cos(beta(0)), np.sin(beta(0))      # qubit 0: (z_0 = 0, r) <-> (z_0 = 1, ~r)
        ev = z[z & 1 == 0]
        new = phi.copy()
        for r in range(G):
            a, b = phi[r, ev], phi[r ^ mask, ev | 1]
            new[r, ev], new[r ^ mask, ev | 1] = c * a + s * b, c * b - s * a
        phi = new
        for j in range(g):                           # rank qubits: the qubit is 0 where r_j == z_0
            c, s = np.cos(beta(el + j)), np.sin(beta(el + j))
            new = phi.copy()
            for r in range(G):
                if (r >> j) & 1:
                    continue
                r1 = r | (1 << j)
                for z0, (ra, rb) in ((0, (r, r1)), (1, (r1, r))):
                    zz = z[z & 1 == z0]
                    a, b = phi[ra, zz], phi[rb, zz]
                    new[ra, zz], new[rb, zz] = c * a + s * b, c * b - s * a
            phi = new
        c, s = np.cos(beta(n - 1)), np.sin(beta(n - 1))   # top qubit: the mirror image ~z lives on the SAME rank
        phi = c * phi + s * sigma * phi[:, z ^ ((1 << el) - 1)]
    prob = np.abs(phi) ** 2
    return float((prob * cost).sum()), float(prob.sum())


@pytest.mark.parametrize("n,g", [(8, 1), (8, 2), (10, 3), (10, 1), (12, 2)])
def test_twisted_shards_reproduce_the_oracle(n, g):
    edges, colors, table = H.regular_maxcut(n, seed=5)
    diag = H.maxcut_diag(n, edges, colors, table)
    assert np.array_equal(diag, diag[::-1]), "MaxCut is invariant under the global bit flip"
    rng = np.random.default_rng(n * 10 + g)
    for n_beta in (1, n):
        ang = rng.uniform(-1.5, 1.5, size=3 * (1 + n_beta))
        spec = orc.Spec(n, diag, mixer=("x",) if n_beta == 1 else ("xmulti",))
        ref = orc.energy(spec, ang)
        e, norm = twisted_energy(n, g, diag, ang, n_beta)
        assert abs(norm - 1) < 1e-12
        assert abs(e - ref[0]) < 1e-10 * max(1.0, abs(ref[0])), (n_beta, e, ref[0])
