"""Pin the oracle's gates against the reference's own golden vectors
(test/test_operator.py:31-110, test/test_inverserot.py:8-25)."""
import numpy as np

from oracle import operators as ops
from oracle.exact import all_states

# literals copied from the reference's test (test/test_operator.py:66-96); they are data, not code
CONNS_QUBIT = np.array([[[0.0, 1.0, 0.0], [1.0, 1.0, 0.0]], [[1.0, 1.0, 1.0], [0.0, 1.0, 1.0]]])
CONNS_SPIN = np.array([[[-1.0, 1.0, -1.0], [1.0, 1.0, -1.0]], [[1.0, 1.0, 1.0], [-1.0, 1.0, 1.0]]])
MELS_RX = np.array([[0.70710678 + 0.0j, 0.0 - 0.70710678j], [0.70710678 + 0.0j, 0.0 - 0.70710678j]])
MELS_RY = np.array([[0.70710678 + 0.0j, 0.70710678 + 0.0j], [0.70710678 + 0.0j, -0.70710678 + 0.0j]])
MELS_H = np.array([[0.70710678, 0.70710678], [-0.70710678, 0.70710678]])


def _states(local_states):
    st = all_states(3, local_states)
    return st[[2, 7]]


def test_numbers_to_states_ordering():
    # test/test_operator.py:38-41 + literals: state #2 is [0,1,0] / [-1,1,-1], #7 is all-up
    np.testing.assert_array_equal(_states((0.0, 1.0)), [[0, 1, 0], [1, 1, 1]])
    np.testing.assert_array_equal(_states((-1.0, 1.0)), [[-1, 1, -1], [1, 1, 1]])


def test_get_conns_and_mels_golden():
    for ls, conns_check in (((0.0, 1.0), CONNS_QUBIT), ((-1.0, 1.0), CONNS_SPIN)):
        sig = _states(ls)
        c, m = ops.conns_and_mels_rx(sig, 0, np.pi / 2, ls)
        np.testing.assert_allclose(c, conns_check)
        np.testing.assert_allclose(m, MELS_RX)
        assert m.dtype == np.complex128
        c, m = ops.conns_and_mels_ry(sig, 0, np.pi / 2, ls)
        np.testing.assert_allclose(c, conns_check)
        np.testing.assert_allclose(m, MELS_RY)
        c, m = ops.conns_and_mels_hadamard(sig, 0, ls)
        np.testing.assert_allclose(c, conns_check)
        np.testing.assert_allclose(m, MELS_H)
        assert m.dtype == np.float64


def test_inverse_rotation():
    # test/test_inverserot.py:8-25
    rx = ops.Rx(4, 0, 0.3)
    d, dh, dh_ = rx.to_dense(), rx.H.to_dense(), ops.Rx(4, 0, -0.3).to_dense()
    np.testing.assert_allclose(dh, dh_, atol=1e-12)
    np.testing.assert_allclose(dh @ d, np.eye(16), atol=1e-12)
    np.testing.assert_allclose(dh_ @ d, np.eye(16), atol=1e-12)


def _kron_site(N, idx, m):
    out = np.array([[1.0 + 0j]])
    for i in range(N):
        out = np.kron(out, m if i == idx else np.eye(2))
    return out


def test_dense_equals_local_operator():
    # mirrors test/test_operator.py:19-28 with explicit Pauli matrices.  For that reference test to
    # be green for all three gates, NetKet's sigma matrices must be the standard ones in LOCAL-INDEX
    # order (index 0 = local_states[0]); the conn kernels attach '+' to local_states[0]
    # (goldens test/test_operator.py:77-96).
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]], dtype=complex)
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    N = 3
    a = 0.23
    np.testing.assert_allclose(ops.Rx(N, 1, a).to_dense(),
                               np.cos(a / 2) * np.eye(8) - 1j * np.sin(a / 2) * _kron_site(N, 1, sx), atol=1e-14)
    a = 0.43
    np.testing.assert_allclose(ops.Ry(N, 1, a).to_dense(),
                               np.cos(a / 2) * np.eye(8) + 1j * np.sin(a / 2) * _kron_site(N, 1, sy), atol=1e-14)
    # singlequbit_gates.py:226-230: H = (sz + sx)/sqrt2
    np.testing.assert_allclose(ops.Hadamard(N, 0).to_dense(),
                               (_kron_site(N, 0, sz) + _kron_site(N, 0, sx)) / np.sqrt(2), atol=1e-14)
    h = ops.Hadamard(N, 0).to_dense()
    np.testing.assert_allclose(h @ h, np.eye(8), atol=1e-14)


def test_ising_dense():
    N = 4
    edges = ops.chain_edges(N, pbc=True)
    h, J = 0.7, 1.3
    H = ops.Ising(N, edges, h, J).to_dense()
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sz = np.array([[-1, 0], [0, 1]], dtype=complex)   # value of the spin (sz|x> = x|x>), zz is sign-blind
    ref = np.zeros((16, 16), dtype=complex)
    for i in range(N):
        ref -= h * _kron_site(N, i, sx)
    for i, j in edges:
        ref += J * _kron_site(N, i, sz) @ _kron_site(N, j, sz)
    np.testing.assert_allclose(H, ref, atol=1e-14)
    xp, mels = ops.Ising(N, edges, h, J).get_conn_padded(all_states(N)[5])
    assert xp.shape == (N + 1, N) and mels.shape == (N + 1,)
    U = ops.IsingPropagator(N, edges, h, J, 0.01).to_dense()
    np.testing.assert_allclose(U, np.eye(16) - 0.01j * ref, atol=1e-14)
