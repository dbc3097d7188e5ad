"""HamiltonianLayer: the first external consumer of the boundary.

Mirrors examples/nns/circuit_layer.jl of the reference: a layer that maps a 15 x ngates angle array to the
energy <psi0| U(angles)^dagger H U(angles) |psi0>, with a reverse rule that calls
gradients!(cache..., H, psi0, circuit; calculate_energy=true) (circuit_layer.jl:32-58).  The reference wires it
into Flux through ChainRulesCore.rrule; here the same contract is a torch.autograd.Function, so the layer can sit
at the end of a torch network exactly like `hamiltonian_layer` does in examples/hpc/utils.jl:57-83.
Angles may arrive in float32 (Flux's default there); they are promoted to float64 for the gate construction.
"""
from __future__ import annotations

import numpy as np
import torch

from .circuits import GenericBrickworkCircuit, brickwork_num_gates
from .gradients import construct_grads_cache, gradients_
from .hamiltonians import measure
from .state import B200State


class _ApplyHamiltonian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate_angles, layer):
        a = gate_angles.detach().to("cpu", torch.float64).numpy()
        circuit = GenericBrickworkCircuit(layer.nbits, layer.nlayers, layer.ngates, np.asfortranarray(a))
        if gate_angles.requires_grad:
            E, grads = gradients_(*layer.cache, layer.H, layer.psi0, circuit, calculate_energy=True)  # circuit_layer.jl:33,47
            ctx.grads = torch.from_numpy(np.ascontiguousarray(grads))
        else:
            E = measure(layer.H, layer.psi0, circuit)  # circuit_layer.jl:21-27
            ctx.grads = None
        ctx.meta = (gate_angles.dtype, gate_angles.device)
        return torch.tensor(E, dtype=gate_angles.dtype, device=gate_angles.device)

    @staticmethod
    def backward(ctx, dE):
        dtype, device = ctx.meta
        return (dE * ctx.grads.to(device=device, dtype=dtype)), None  # pb(dE) = dE .* gradients  (circuit_layer.jl:34-42)


class HamiltonianLayer(torch.nn.Module):
    """HamiltonianLayer(nbits, nlayers, ngates, psi0, H, cache) -- examples/nns/circuit_layer.jl:8-15."""

    def __init__(self, nbits, nlayers, psi0, H, ngates=None, cache=None):
        super().__init__()
        self.nbits, self.nlayers = int(nbits), int(nlayers)
        self.ngates = brickwork_num_gates(nbits, nlayers) if ngates is None else int(ngates)
        self.psi0 = psi0 if isinstance(psi0, B200State) else B200State(psi0)
        self.H = H
        self.cache = construct_grads_cache(self.psi0) if cache is None else cache

    def forward(self, gate_angles):
        if tuple(gate_angles.shape) != (15, self.ngates):  # @boundscheck size(gate_angles) == (15, hl.ngates)
            raise ValueError(f"gate_angles must be 15 x {self.ngates}")
        return _ApplyHamiltonian.apply(gate_angles, self)


def apply_hamiltonian(hl: HamiltonianLayer, gate_angles):
    return hl(gate_angles)
