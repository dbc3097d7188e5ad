"""Host-side operator helpers of the reference (src/utils.jl:1-40): the dense 2^n x 2^n matrix of a set of
non-overlapping localised gates.  Small-n algebra on the host in the reference too (it is what its own tests compare
`apply` against); kept so that the exported name `convert_gates_to_matrix` exists on this side of the boundary."""
from __future__ import annotations

import numpy as np

from .gates import Abstract1SpinGate, Abstract2SpinGate, IdentityGate, mat


def gatebits(gate):  # src/utils.jl:1-2
    return 2 if isinstance(gate, Abstract2SpinGate) else 1


def matrix_only_mat(gate):  # src/utils.jl:4-5 : 2-qubit gates as 4x4 with qubit K as the low index bit
    m = np.asarray(mat(gate))
    return m.reshape(4, 4, order="F") if isinstance(gate, Abstract2SpinGate) else m


def operation_tensor(mats):  # src/utils.jl:33-40 : kron(M_last, ..., M_first)
    a = np.asarray(mats[-1])
    for b in reversed(mats[:-1]):
        a = np.kron(a, np.asarray(b))
    return a


def convert_gates_to_matrix(nbits, gates):
    """src/utils.jl:10-26 : identity on untouched qubits, first qubit = least-significant Kronecker factor."""
    by_pos = {g.target_gate_dim: g for g in gates}
    out, i = [], 1
    while i <= nbits:
        if i in by_pos:
            out.append(matrix_only_mat(by_pos[i]))
            i += gatebits(by_pos[i])
        else:
            out.append(matrix_only_mat(IdentityGate()))
            i += 1
    return operation_tensor(out)
