"""GenericBrickworkCircuit and its geometry (src/circuits.jl:3-18,44-48)."""
from __future__ import annotations

import numpy as np

from . import _lib


def circuit_layer_starts(layer_number, nbits):
    """1-based first qubits of the gates of (1-based) half-layer `layer_number` (src/circuits.jl:9)."""
    return range(1 + (layer_number - 1) % 2, nbits, 2)


def brickwork_num_gates(nbits, nlayers):  # src/circuits.jl:10-12
    return _lib.lib().qcb_brickwork_num_gates(int(nbits), int(nlayers))


class GenericBrickworkCircuit:
    """nbits, nlayers, ngates and a 15 x ngates angle matrix (column g = gate g in application order)."""

    def __init__(self, nbits, nlayers, ngates=None, gate_angles=None):
        self.nbits, self.nlayers = int(nbits), int(nlayers)
        self.ngates = brickwork_num_gates(nbits, nlayers) if ngates is None else int(ngates)
        if gate_angles is None:
            gate_angles = np.zeros((15, self.ngates), dtype=np.float64, order="F")  # src/circuits.jl:16
        self.gate_angles = np.asfortranarray(gate_angles, dtype=np.float64)
        if self.gate_angles.shape != (15, self.ngates):
            raise ValueError(f"gate_angles must be 15 x {self.ngates}")

    def angles_abi(self):
        return np.ascontiguousarray(self.gate_angles.reshape(-1, order="F"))


def reconstruct(circuit, angle, angle_offset):
    """src/circuits.jl:44-48 : copy with one (0-based linear, column-major) angle shifted."""
    flat = circuit.gate_angles.reshape(-1, order="F").copy()
    flat[angle] += angle_offset
    return GenericBrickworkCircuit(circuit.nbits, circuit.nlayers, circuit.ngates, flat.reshape(15, -1, order="F"))
