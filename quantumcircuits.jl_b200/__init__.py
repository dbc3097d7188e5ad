"""B200-native dense-statevector backend behind the QuantumCircuits.jl API.

The export list follows src/QuantumCircuits.jl:9-18 of the reference; Julia's `f!` is `f_` here.
The Julia-side ccall shim over the same C ABI is julia/B200Backend.jl.
"""
from ._lib import QCBError, GateOutsideQubitSpace, get_stat, init, set_option, sync, trim  # noqa: F401
from .gates import (AbstractGate, Abstract1SpinGate, Abstract2SpinGate, XGate, ZGate, NGate, HadamardGate,  # noqa: F401
                    IdentityGate, CNOTGate, ReverseCNOTGate, RZGate, RYGate, RXGate, RotationGate, Generic1SpinGate,
                    Generic2SpinGate, Localised1SpinGate, Localised2SpinAdjGate, mat, adjoint,
                    build_general_unitary_gate)
from .circuits import GenericBrickworkCircuit, circuit_layer_starts, brickwork_num_gates, reconstruct  # noqa: F401
from .state import B200State, zero_state_vec, zero_state_tensor  # noqa: F401
from .utils import convert_gates_to_matrix, operation_tensor  # noqa: F401
from .apply import B200ApplyCache, construct_apply_cache, apply, apply_, apply_circuit_host, right_apply_gate_  # noqa: F401
from .hamiltonians import (AbstractKernelHamiltonian, TFIMHamiltonian, EastModelHamiltonian, measure, measure_,  # noqa: F401
                           imaginary_energy_limit, build_sparse_tfim_hamiltonian, find_tfim_ground_state, east_model_matrix)
from . import dist  # noqa: F401
from .gradients import construct_grads_cache, gradients, gradients_, optimise_, propagate_forwards_  # noqa: F401
