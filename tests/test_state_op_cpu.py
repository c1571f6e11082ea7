"""Ports of the reference's tests/pulser_simulation/test_qutip_state_op.py to ``B200State`` / ``B200Operator`` (the
``State`` / ``Operator`` protocol of seam S2; numpy arrays stand where the reference passes ``qutip.Qobj``)."""
import json
import re

import numpy as np
import pytest

from pulser_b200 import HAVE_PULSER

pytestmark = pytest.mark.skipif(not HAVE_PULSER, reason="pulser-core not importable here")


@pytest.fixture
def B():
    from pulser_b200 import backend

    return backend


def basis(d, k):
    v = np.zeros(d, dtype=complex)
    v[k] = 1.0
    return v


@pytest.fixture
def ket_r(B):
    return B.B200State(basis(2, 0), eigenstates=("r", "g"))


@pytest.fixture
def dm_g(B):
    return B.B200State(np.outer(basis(2, 1), basis(2, 1).conj()), eigenstates=("r", "g"))


@pytest.fixture
def ket_plus(B):
    return B.B200State.from_state_amplitudes(eigenstates=("r", "g"),
                                             amplitudes={"r": 1 / np.sqrt(2), "g": 1 / np.sqrt(2)})


class TestB200State:
    def test_init(self, B):
        """test_qutip_state_op.py:50-104"""
        with pytest.raises(ValueError, match="eigenstates must be represented by single characters"):
            B.B200State(basis(2, 0), eigenstates=["ground", "rydberg"])
        with pytest.raises(ValueError, match="can't contain repeated entries"):
            B.B200State(basis(2, 0), eigenstates=["r", "g", "r"])
        with pytest.raises(TypeError, match="must be a 'collections.Sequence'"):
            B.B200State(basis(2, 0), eigenstates={"r", "g"})
        with pytest.raises(TypeError, match="must be a ket"):
            B.B200State(np.arange(16).reshape(2, 8), eigenstates=["r", "g"])
        with pytest.raises(ValueError, match="incompatible with a system of 3-level qudits"):
            B.B200State(basis(2, 0), eigenstates=["r", "g", "h"])
        state = B.B200State(basis(3, 0).reshape(1, 3), eigenstates=["r", "g", "h"])  # a bra is accepted
        assert state.n_qudits == 1 and state.qudit_dim == 3 and state.eigenstates == ("r", "g", "h")
        np.testing.assert_array_equal(state.to_array(), basis(3, 0))
        with pytest.raises(RuntimeError, match="Failed to infer the 'one state'"):
            state.infer_one_state()
        three = np.kron(np.kron(basis(2, 1), basis(2, 1)), basis(2, 1))
        state = B.B200State(three, eigenstates=("r", "g"))
        assert state.n_qudits == 3 and state.qudit_dim == 2 and state.infer_one_state() == "r"
        two_qutrit_dm = np.outer(np.kron(basis(3, 0), basis(3, 0)), np.kron(basis(3, 0), basis(3, 0)))
        state = B.B200State(two_qutrit_dm, eigenstates=["r", "g", "h"])
        assert state.n_qudits == 2 and state.qudit_dim == 3 and not state.is_ket

    @pytest.mark.parametrize("eigenstates", [("g", "r"), ("g", "r", "x"), ("g", "h"), ("u", "d"), ("0", "1")])
    def test_infer_one_state(self, B, eigenstates):
        """:107-113"""
        assert B.B200State(basis(len(eigenstates), 0), eigenstates=eigenstates).infer_one_state() == eigenstates[1]

    def test_get_basis_state(self, B):
        """:115-131"""
        n = 3
        state = B.B200State.from_state_amplitudes(eigenstates=("r", "g", "h"), amplitudes={"g" * n: 1.0})
        for idx, s in ((0, "rrr"), (1, "rrg"), (2, "rrh"), (3, "rgr"), (4, "rgg"), (9, "grr"), (3**n - 1, "hhh")):
            assert state.get_basis_state_from_index(idx) == s
        with pytest.raises(ValueError, match="'index' must be a non-negative integer"):
            state.get_basis_state_from_index(-1)

    def test_overlap(self, B, ket_r, dm_g, ket_plus):
        """:133-172"""
        assert ket_r.overlap(ket_r) == 1.0
        assert dm_g.overlap(ket_r) == ket_r.overlap(dm_g) == 0.0
        assert ket_plus.overlap(ket_r) == ket_r.overlap(ket_plus)
        assert np.isclose(ket_plus.overlap(ket_r), 0.5)
        assert dm_g.overlap(ket_plus) == ket_plus.overlap(dm_g)
        assert np.isclose(dm_g.overlap(ket_plus), 0.5)
        with pytest.raises(TypeError, match="expects another 'B200State'"):
            dm_g.overlap(ket_r.to_array())
        with pytest.raises(ValueError, match="Can't calculate the overlap between a state with 1 "
                                             "2-dimensional qudits and another with 2 3-dimensional qudits"):
            ket_r.overlap(B.B200State.from_state_amplitudes(eigenstates=("r", "g", "h"), amplitudes={"rr": 1.0}))
        err_msg = "Can't calculate the overlap between states with eigenstates ('r', 'g') and {}."
        with pytest.raises(ValueError, match=re.escape(err_msg.format(("u", "d")))):
            ket_r.overlap(B.B200State(basis(2, 0), eigenstates=("u", "d")))
        with pytest.raises(NotImplementedError, match=re.escape(err_msg.format(("g", "r")))):
            ket_r.overlap(B.B200State(basis(2, 0), eigenstates=("g", "r")))

    def test_probabilities(self, B, ket_plus):
        """:174-207"""
        amps = {"rr": np.sqrt(0.5), "gg": 1j * np.sqrt(0.5 - 1e-12), "gr": 1e-6}
        state = B.B200State.from_state_amplitudes(eigenstates=("r", "g"), amplitudes=amps)
        probs = {k: np.abs(a) ** 2 for k, a in amps.items()}
        state_probs = state.probabilities(cutoff=9e-13)
        assert all(np.isclose(probs[k], state_probs[k]) for k in probs)
        probs.pop("gr")
        sum_ = sum(probs.values())
        probs = {k: v / sum_ for k, v in probs.items()}
        state_probs = state.probabilities()
        assert all(np.isclose(probs[k], state_probs[k]) for k in probs)
        assert state.infer_one_state() == "r"
        bp = state.bitstring_probabilities()
        assert set(bp) == {"11", "00"} and np.isclose(bp["11"], probs["rr"]) and np.isclose(bp["00"], probs["gg"])
        bp = state.bitstring_probabilities(one_state="g")
        assert np.isclose(bp["11"], probs["gg"]) and np.isclose(bp["00"], probs["rr"])
        v = ket_plus.to_array()
        dm_plus = B.B200State(np.outer(v, v.conj()), eigenstates=ket_plus.eigenstates)
        p = dm_plus.probabilities()
        assert np.isclose(p["r"], 0.5) and np.isclose(p["g"], 0.5)
        bp = dm_plus.bitstring_probabilities()
        assert np.isclose(bp["0"], 0.5) and np.isclose(bp["1"], 0.5)

    def test_sample(self, ket_r, dm_g):
        """:209-219"""
        shots = 2000
        assert ket_r.sample(num_shots=shots) == {"1": shots}
        assert ket_r.sample(num_shots=shots, one_state="g") == {"0": shots}
        assert ket_r.sample(num_shots=shots, p_false_pos=0.1) == {"1": shots}
        assert ket_r.sample(num_shots=shots, p_false_neg=0.1)["0"] > 0
        assert dm_g.sample(num_shots=shots) == {"0": shots}
        assert dm_g.sample(num_shots=shots, one_state="g") == {"1": shots}
        assert dm_g.sample(num_shots=shots, p_false_neg=0.1) == {"0": shots}
        assert dm_g.sample(num_shots=shots, p_false_pos=0.1)["1"] > 0

    @pytest.mark.parametrize("amplitudes", [{"rrh": 1.0}, {"rr": 0.5, "rgg": np.sqrt(0.75)}])
    def test_from_state_amplitudes_error(self, B, amplitudes):
        """:221-239"""
        with pytest.raises(ValueError, match=re.escape(
                "All basis states must be combinations of eigenstates with the same length. Expected combinations of "
                f"('r', 'g'), each with {len(list(amplitudes)[0])} elements.")):
            B.B200State.from_state_amplitudes(eigenstates=("r", "g"), amplitudes=amplitudes)

    def test_from_state_amplitudes(self, B):
        """:241-263"""
        mk = B.B200State.from_state_amplitudes
        np.testing.assert_array_equal(mk(eigenstates=("r", "g"), amplitudes={"g": 1.0}).to_array(), basis(2, 1))
        np.testing.assert_array_equal(mk(eigenstates=("g", "r"), amplitudes={"g": 1.0}).to_array(), basis(2, 0))
        np.testing.assert_array_equal(mk(eigenstates=("r", "g", "h"), amplitudes={"g": 1.0}).to_array(), basis(3, 1))
        r, g = basis(2, 0), basis(2, 1)
        got = mk(eigenstates=("r", "g"), amplitudes={"rr": -0.5j, "gr": 0.5, "rg": 0.5j, "gg": -0.5}).to_array()
        np.testing.assert_allclose(got, -0.5j * np.kron(r, r) + 0.5 * np.kron(g, r) + 0.5j * np.kron(r, g)
                                   - 0.5 * np.kron(g, g))

    def test_eq(self, B, ket_r, dm_g):
        """:273-280"""
        assert ket_r == B.B200State.from_state_amplitudes(eigenstates=("r", "g"), amplitudes={"r": 1.0})
        assert dm_g != B.B200State.from_state_amplitudes(eigenstates=("r", "g"), amplitudes={"g": 1.0})
        assert dm_g != np.outer(basis(2, 1), basis(2, 1))

    def test_abstract_repr(self, B, ket_r):
        """:282-305"""
        from pulser.exceptions.serialization import AbstractReprError
        from pulser.json.abstract_repr.serializer import AbstractReprEncoder

        kwargs = dict(eigenstates=("r", "g"), amplitudes={"g": 1.0})
        state = B.B200State.from_state_amplitudes(**kwargs)
        assert json.dumps(state, cls=AbstractReprEncoder) == json.dumps(kwargs)
        with pytest.raises(AbstractReprError, match=re.escape(
                "Failed to serialize state of type 'B200State' because it was not created via "
                "'B200State.from_state_amplitudes()'")):
            json.dumps(B.B200State(state.to_array(), eigenstates=state.eigenstates), cls=AbstractReprEncoder)
        state._state = ket_r._state
        with pytest.raises(AbstractReprError, match="modified in place after its creation"):
            json.dumps(state, cls=AbstractReprEncoder)


SX = np.array([[0, 1], [1, 0]], dtype=complex)
SY = np.array([[0, -1j], [1j, 0]], dtype=complex)
SZ = np.array([[1, 0], [0, -1]], dtype=complex)


def proj(d, k):
    return np.outer(basis(d, k), basis(d, k).conj())


class TestB200Operator:
    @pytest.fixture
    def pauli_i(self, B):
        return B.B200Operator(np.eye(2, dtype=complex), eigenstates=("r", "g"))

    @pytest.fixture
    def pauli_x(self, B):
        return B.B200Operator(SX, eigenstates=("r", "g"))

    @pytest.fixture
    def pauli_y(self, B):
        return B.B200Operator(SY, eigenstates=("r", "g"))

    @pytest.fixture
    def pauli_z(self, B):
        return B.B200Operator(SZ, eigenstates=("r", "g"))

    def test_init(self, B):
        """test_qutip_state_op.py:307-330"""
        with pytest.raises(ValueError, match="eigenstates must be represented by single characters"):
            B.B200Operator(SZ, eigenstates=["ground", "rydberg"])
        with pytest.raises(ValueError, match="can't contain repeated entries"):
            B.B200Operator(SZ, eigenstates=["r", "g", "r"])
        with pytest.raises(ValueError, match="incompatible with a system of 3-level qudits"):
            B.B200Operator(SZ, eigenstates=["r", "g", "h"])
        pauli_z = B.B200Operator(SZ, eigenstates=("r", "g"))
        assert pauli_z.eigenstates == ("r", "g")
        np.testing.assert_array_equal(pauli_z.to_array(), proj(2, 0) - proj(2, 1))

    @pytest.mark.parametrize("op_name", ["apply_to", "expect"])
    def test_errors_on_state(self, B, pauli_x, op_name):
        """:351-372"""
        op = getattr(pauli_x, op_name)
        with pytest.raises(TypeError, match=re.escape(f"'B200Operator.{op_name}()' expects a 'B200State' instance")):
            op(basis(2, 0))
        err_msg = (f"Can't apply B200Operator.{op_name}() between a B200Operator "
                   "with eigenstates ('r', 'g') and a B200State with {}")
        with pytest.raises(ValueError, match=re.escape(err_msg.format(("g", "h")))):
            op(B.B200State(basis(2, 0), eigenstates=("g", "h")))
        with pytest.raises(NotImplementedError, match=re.escape(err_msg.format(("g", "r")))):
            op(B.B200State(basis(2, 0), eigenstates=("g", "r")))

    @pytest.mark.parametrize("op_name", ["__add__", "__matmul__"])
    def test_errors_on_operator(self, B, pauli_x, ket_r, op_name):
        """:374-394"""
        op = getattr(pauli_x, op_name)
        with pytest.raises(TypeError, match=re.escape(f"'{op_name}' expects a 'B200Operator' instance")):
            op(ket_r)
        err_msg = (f"Can't apply {op_name} between a B200Operator with eigenstates "
                   "('r', 'g') and a B200Operator with {}")
        with pytest.raises(ValueError, match=re.escape(err_msg.format(("g", "h")))):
            op(B.B200Operator(proj(2, 0), eigenstates=("g", "h")))
        with pytest.raises(NotImplementedError, match=re.escape(err_msg.format(("g", "r")))):
            op(B.B200Operator(proj(2, 0), eigenstates=("g", "r")))

    def test_apply_to(self, B, ket_r, dm_g, pauli_x):
        """:396-402"""
        assert pauli_x.apply_to(ket_r) == B.B200State.from_state_amplitudes(eigenstates=("r", "g"),
                                                                            amplitudes={"g": 1.0})
        assert pauli_x.apply_to(dm_g) == B.B200State(proj(2, 0), eigenstates=dm_g.eigenstates)

    def test_expect(self, pauli_x, pauli_y, pauli_z, ket_r, dm_g, ket_plus):
        """:404-421"""
        assert pauli_x.expect(ket_r) == 0.0
        assert pauli_x.expect(dm_g) == 0.0
        assert np.isclose(pauli_x.expect(ket_plus), 1.0)
        ket_minus = pauli_y.apply_to(ket_plus)
        assert np.isclose(pauli_x.expect(ket_minus), -1.0)
        assert pauli_z.expect(ket_r) == 1.0
        assert pauli_z.expect(dm_g) == -1.0
        assert np.isclose(pauli_z.expect(ket_plus), 0.0)

    def test_add_rmul_matmul(self, B, pauli_i, pauli_x, pauli_y, pauli_z):
        """:423-453"""
        r, g = basis(2, 0), basis(2, 1)
        assert pauli_x + pauli_y == B.B200Operator((1 - 1j) * np.outer(r, g) + (1 + 1j) * np.outer(g, r),
                                                   eigenstates=pauli_x.eigenstates)
        assert pauli_i + pauli_z == B.B200Operator(2 * proj(2, 0), eigenstates=pauli_z.eigenstates)
        assert (1 - 2j) * pauli_i == B.B200Operator((1 - 2j) * np.eye(2), eigenstates=pauli_z.eigenstates)
        assert 0.5 * (pauli_i + pauli_z) == B.B200Operator(proj(2, 0), eigenstates=pauli_z.eigenstates)
        assert pauli_x @ pauli_x == pauli_y @ pauli_y == pauli_z @ pauli_z == pauli_i
        assert pauli_x @ pauli_z == -1j * pauli_y
        assert pauli_z @ pauli_x == 1j * pauli_y

    def test_from_operator_repr(self, B, pauli_i):
        """:455-562"""
        mk = B.B200Operator.from_operator_repr
        with pytest.raises(ValueError, match=re.escape(
                "Every QuditOp key must be made up of two eigenstates among ('r', 'g'); instead, got 'gggg'.")):
            mk(eigenstates=("r", "g"), n_qudits=2, operations=[(1.0, [({"gggg": 1.0, "rr": -1.0}, {0})])])
        with pytest.raises(ValueError, match=re.escape(
                "Every QuditOp key must be made up of two eigenstates among ('r', 'g'); instead, got 'hh'.")):
            mk(eigenstates=("r", "g"), n_qudits=2, operations=[(1.0, [({"hh": 1.0, "rr": -1.0}, {0})])])
        with pytest.raises(ValueError, match="Got invalid indices for a system with 2 qudits"):
            mk(eigenstates=("r", "g"), n_qudits=2, operations=[(1.0, [({"gg": 1.0, "rr": -1.0}, {3, 5, 9})])])
        with pytest.raises(ValueError, match=re.escape("only indices {1} were still available")):
            mk(eigenstates=("r", "g"), n_qudits=2,
               operations=[(1.0, [({"gg": 1.0, "rr": -1.0}, {0}), ({"rg": 1.0}, {0})])])
        got = mk(eigenstates=("r", "g", "h"), n_qudits=3,
                 operations=[(1.0, [({"rr": 1.0, "hh": -1.0}, {0}), ({"gr": -1j}, {2})])])
        want = np.kron(np.kron(proj(3, 0) - proj(3, 2), np.eye(3)), -1j * np.outer(basis(3, 1), basis(3, 0)))
        assert got == B.B200Operator(want, eigenstates=("r", "g", "h"))
        assert mk(eigenstates=("r", "g"), n_qudits=1, operations=[(1, [])]) == pauli_i
        got = mk(eigenstates=("r", "g"), n_qudits=2, operations=[(0.5, [({"rr": 1.0, "gg": -1.0}, {0})]), (0.5, [])])
        assert got == B.B200Operator(np.kron(proj(2, 0), np.eye(2)), eigenstates=("r", "g"))

    def test_eq(self, B, pauli_i, pauli_z, dm_g):
        """:573-578"""
        g_proj = 0.5 * (pauli_i + (-1) * pauli_z)
        assert g_proj == B.B200Operator(proj(2, 1), eigenstates=pauli_i.eigenstates)
        assert g_proj != dm_g

    def test_abstract_repr(self, B):
        """:580-605"""
        from pulser.exceptions.serialization import AbstractReprError
        from pulser.json.abstract_repr.serializer import AbstractReprEncoder

        kwargs = dict(eigenstates=("r", "g"), n_qudits=3,
                      operations=[(0.5, [({"rr": 1.0, "gg": 1.0j}, {0})]), (0.5, [])])
        op = B.B200Operator.from_operator_repr(**kwargs)
        ser_ops = [(0.5, [({"rr": 1.0, "gg": {"real": 0.0, "imag": 1.0}}, [0])]), (0.5, [])]
        assert json.dumps(op, cls=AbstractReprEncoder) == json.dumps({**kwargs, "operations": ser_ops})
        with pytest.raises(AbstractReprError, match=re.escape(
                "Failed to serialize state of type 'B200Operator' because it was not created via "
                "'B200Operator.from_operator_repr()'")):
            json.dumps(B.B200Operator(op.to_array(), eigenstates=op.eigenstates), cls=AbstractReprEncoder)
