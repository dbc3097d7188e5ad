# B200Backend.jl -- thin `ccall` shim that puts libqcb200.so behind the QuantumCircuits.jl API.
#
# Usage (from a session that has QuantumCircuits loaded):
#
#     ENV["QCB200_LIB"] = "/path/to/quantumcircuits.jl_b200/libqcb200.so"
#     include("B200Backend.jl"); using .B200Backend
#     psi = B200State(zero_state_tensor(30))            # where the reference's examples write CuArray(psi0)
#     E, grads = gradients(TFIMHamiltonian(1.0, 0.1, 0.5), psi, circuit; calculate_energy=true)
#
# No CUDA.jl, no KernelAbstractions dispatch, no CPU fallback: every method below is one ccall.
# NOTE: Julia is not installed in the build image, so this file has not been executed there; the same C ABI
# is exercised end to end by the Python mirror (quantumcircuits.jl_b200/*.py) and the tests.
module B200Backend

using LinearAlgebra
import QuantumCircuits
import SparseArrays
import QuantumCircuits: apply, apply!, measure, measure!, gradients, gradients!, optimise!,
    construct_apply_cache, construct_grads_cache, propagate_forwards!, right_apply_gate!, mat,
    GenericBrickworkCircuit, Localised1SpinGate, Localised2SpinAdjGate, AbstractGate,
    TFIMHamiltonian, EastModelHamiltonian, AbstractKernelHamiltonian, imaginary_energy_limit

export B200State, B200ApplyCache, B200ShardedState, dist_unique_id, dist_init, dist_finalize, map_peers!, unmap_peers!,
    upload_shard!, download_shard, local_qubits

const LIB = get(ENV, "QCB200_LIB", joinpath(@__DIR__, "..", "libqcb200.so"))

const QCB_C64 = Cint(0)
const QCB_C128 = Cint(1)

last_error() = unsafe_string(ccall((:qcb_last_error, LIB), Cstring, ()))

function check(rc::Cint)
    rc == 0 && return nothing
    msg = last_error()
    # the reference raises AssertionError for K outside the qubit space (src/apply.jl:27,45,112) and
    # ErrorException for a size mismatch (src/apply.jl:21,40,108)
    occursin("outside the qubit space", msg) && throw(AssertionError(msg))
    error(msg)
end

function __init__()
    dev = parse(Int, get(ENV, "QCB200_DEVICE", "0"))
    check(ccall((:qcb_init, LIB), Cint, (Cint,), dev))
end

dtype_code(::Type{Float32}) = QCB_C64
dtype_code(::Type{Float64}) = QCB_C128

# ------------------------------------------------------------------------------------------------
# device state: stands where the CuArray stands in the reference's GPU usage (examples/test_gpu.jl:21-24)
# ------------------------------------------------------------------------------------------------
mutable struct B200State{T<:AbstractFloat,N} <: AbstractArray{Complex{T},N}
    handle::Ptr{Cvoid}
    function B200State{T,N}() where {T,N}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qcb_state_create, LIB), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}), N, dtype_code(T), h))
        s = new{T,N}(h[])
        finalizer(x -> ccall((:qcb_state_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), s)
        return s
    end
end

function B200State(psi::Array{Complex{T},N}) where {T,N}
    all(==(2), size(psi)) || error("states are 2x2x...x2 tensors (src/utils.jl:61-66)")
    s = B200State{T,N}()
    check(ccall((:qcb_state_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s.handle, psi))
    return s
end

Base.size(::B200State{T,N}) where {T,N} = ntuple(_ -> 2, N)
Base.similar(::B200State{T,N}) where {T,N} = B200State{T,N}()
function Base.copy(s::B200State{T,N}) where {T,N}
    d = B200State{T,N}()
    check(ccall((:qcb_state_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), d.handle, s.handle))
    return d
end
function Base.copyto!(dst::B200State{T,N}, src::B200State{T,N}) where {T,N}
    check(ccall((:qcb_state_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dst.handle, src.handle))
    return dst
end
function Base.copyto!(dst::B200State{T,N}, src::Array{Complex{T},N}) where {T,N}
    check(ccall((:qcb_state_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dst.handle, src))
    return dst
end
function Base.Array(s::B200State{T,N}) where {T,N}
    out = Array{Complex{T},N}(undef, size(s))
    check(ccall((:qcb_state_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s.handle, out))
    return out
end
# scalar indexing exists only so that printing works; it downloads the whole state
Base.getindex(s::B200State, i::Int...) = Array(s)[i...]
# `psi .= psi0`, `psi'' .= psi` (src/gradients.jl:90,118,129,138) and `psi' .= 0` (src/apply.jl:135).  `dst .= src` lowers to
# materialize!(dst, broadcasted(identity, src)) -> copyto!(dst, ::Broadcasted): the documented extension point of the
# broadcast machinery, so only these three shapes are accepted and anything else is an error instead of a slow
# element-by-element loop over a device array.
Base.BroadcastStyle(::Type{<:B200State}) = Base.Broadcast.ArrayStyle{B200State}()
function Base.copyto!(dst::B200State, bc::Base.Broadcast.Broadcasted)
    bcf = Base.Broadcast.flatten(bc)
    if bcf.f === identity && length(bcf.args) == 1
        a = bcf.args[1]
        a isa Union{B200State,Array} && return copyto!(dst, a)
        if a isa Number
            iszero(a) || error("B200State: only `psi .= 0` is supported as a scalar fill")
            check(ccall((:qcb_state_fill_zero, LIB), Cint, (Ptr{Cvoid},), dst.handle))
            return dst
        end
    end
    error("B200State supports `dst .= src` and `dst .= 0` only; run arithmetic on the host copy `Array(psi)`")
end
LinearAlgebra.norm(s::B200State) = begin
    out = Ref{Cdouble}(0)
    check(ccall((:qcb_state_norm2, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
    sqrt(out[])
end

# ------------------------------------------------------------------------------------------------
# apply.jl seam: construct_apply_cache (src/apply.jl:95-104) and apply! (src/apply.jl:20,39,117,121)
# ------------------------------------------------------------------------------------------------
struct B200ApplyCache end
construct_apply_cache(::B200State) = B200ApplyCache()

gate_block(u::AbstractArray) = (m = zeros(ComplexF64, 16); m[1:length(u)] .= vec(ComplexF64.(u)); m)

function _check_pair(a::B200State, b::B200State)
    size(a) == size(b) && eltype(a) == eltype(b) || error("ψ′ must be the same size as ψ.")
    nothing
end

# (psi', psi) ping-pong of the callers: the gate pass reads psi and writes psi' (qcb_apply_gates_from: no copy of the
# state, no zero fill -- src/apply.jl:135 -- and one sweep instead of three)
function _apply_gate_from!(psi_out::B200State, psi::B200State, gate, arity::Integer)
    _check_pair(psi_out, psi)
    K = Cint[gate.target_gate_dim]; ar = Cint[arity]
    check(ccall((:qcb_apply_gates_from, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{ComplexF64}),
                psi_out.handle, psi.handle, 1, K, ar, gate_block(mat(gate))))
    return psi_out
end
apply!(psi_out::B200State, psi::B200State, gate::Localised2SpinAdjGate) = _apply_gate_from!(psi_out, psi, gate, 2)
apply!(psi_out::B200State, psi::B200State, gate::Localised1SpinGate) = _apply_gate_from!(psi_out, psi, gate, 1)
apply!(::B200ApplyCache, psi_out::B200State, psi::B200State, gate::Localised2SpinAdjGate) = apply!(psi_out, psi, gate)

# apply(psi, gates) (src/apply.jl:5-18): one fused call for the whole list
function apply(psi::B200State, gates::AbstractArray{<:AbstractGate})
    out = copy(psi)
    G = length(gates)
    K = Cint[g.target_gate_dim for g in gates]
    ar = Cint[g isa QuantumCircuits.Abstract2SpinGate ? 2 : 1 for g in gates]
    mats = zeros(ComplexF64, 16, G)
    for (i, g) in enumerate(gates)
        mats[:, i] .= gate_block(mat(g))
    end
    check(ccall((:qcb_apply_gates, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{ComplexF64}), out.handle, G, K, ar, mats))
    return out
end
apply(psi::B200State, gate::AbstractGate) = apply(psi, AbstractGate[gate])

# circuits.jl:19-42 (quirk Q3: the reference has no GPU method for this; the backend supplies one)
function _apply_circuit!(s::B200State, c::GenericBrickworkCircuit, first::Integer, last::Integer, adj::Bool)
    angles = Matrix{Float64}(c.gate_angles)
    check(ccall((:qcb_apply_brickwork_range, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Cint, Cint),
                s.handle, c.nbits, c.nlayers, angles, first, last, adj ? 1 : 0))
    return s
end
# psi' = circuit(psi) out of place: the first fused pass reads psi and writes psi' (no copy of the state)
function _apply_circuit_from!(dst::B200State, src::B200State, c::GenericBrickworkCircuit, first::Integer, last::Integer, adj::Bool)
    angles = Matrix{Float64}(c.gate_angles)
    check(ccall((:qcb_apply_brickwork_from, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Cint, Cint),
                dst.handle, src.handle, c.nbits, c.nlayers, angles, first, last, adj ? 1 : 0))
    return dst
end
function apply!(::B200ApplyCache, psi_out::B200State, psi::B200State, circuit::GenericBrickworkCircuit)
    _check_pair(psi_out, psi)
    return _apply_circuit_from!(psi_out, psi, circuit, 0, circuit.ngates, false)
end
apply(psi::B200State, circuit::GenericBrickworkCircuit) = _apply_circuit_from!(similar(psi), psi, circuit, 0, circuit.ngates, false)

# ------------------------------------------------------------------------------------------------
# hamiltonians.jl seam: measure! (src/hamiltonians/hamiltonians.jl:31-50)
# ------------------------------------------------------------------------------------------------
function _measure(H::TFIMHamiltonian, s::B200State)
    E = zeros(Cdouble, 2)
    check(ccall((:qcb_measure_tfim, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}), s.handle, H.J, H.h, H.g, E))
    return E
end
function _measure(H::EastModelHamiltonian, s::B200State)
    E = zeros(Cdouble, 2)
    check(ccall((:qcb_measure_east, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}), s.handle, H.c, H.A, H.B, E))
    return E
end
function measure!(::Union{B200State,Nothing}, H::AbstractKernelHamiltonian, psi::B200State)
    E = _measure(H, psi)
    if E[2] > imaginary_energy_limit(H)
        @warn "Imaginary energy of $(E[2]) -> above threshold!"
    end
    return E[1]
end
measure(H::AbstractKernelHamiltonian, psi::B200State) = measure!(nothing, H, psi)
measure(H::AbstractKernelHamiltonian, psi::B200State, circuit::GenericBrickworkCircuit) = measure!(nothing, H, apply(psi, circuit))

# matrix Hamiltonians (src/circuits.jl:49-62): dense or sparse H as a 0-based CSR triple, real(dot(psi, H, psi)) on the device
function _csr(H::AbstractMatrix)
    n = size(H, 1)
    rowptr = Vector{Int64}(undef, n + 1); rowptr[1] = 0
    cols = Int64[]; vals = ComplexF64[]
    Ht = H isa SparseArrays.AbstractSparseMatrix ? SparseArrays.sparse(transpose(H)) : nothing  # CSC of H^T = CSR of H
    for r in 1:n
        if Ht === nothing
            append!(cols, 0:(size(H, 2) - 1)); append!(vals, ComplexF64.(H[r, :]))
        else
            rng = SparseArrays.nzrange(Ht, r)
            append!(cols, SparseArrays.rowvals(Ht)[rng] .- 1); append!(vals, ComplexF64.(SparseArrays.nonzeros(Ht)[rng]))
        end
        rowptr[r + 1] = length(cols)
    end
    return rowptr, cols, vals
end
function measure(H::AbstractMatrix, psi::B200State)
    @assert ndims(H) == 2                                     # src/circuits.jl:50
    rowptr, cols, vals = _csr(H)
    E = zeros(Cdouble, 2)
    check(ccall((:qcb_measure_csr, LIB), Cint, (Ptr{Cvoid}, Clonglong, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Cdouble}),
                psi.handle, size(H, 1), rowptr, cols, vals, E))
    @assert E[2] < 1e-12                                      # :53
    return E[1]
end
measure!(::Union{B200State,Nothing}, H::AbstractMatrix, psi::B200State) = measure(H, psi)                       # :56-58
measure(H::AbstractMatrix, psi::B200State, circuit::GenericBrickworkCircuit) = measure(H, apply(psi, circuit))  # :59-62

# ------------------------------------------------------------------------------------------------
# gradients.jl seam (src/gradients.jl:1-17,54-152)
# ------------------------------------------------------------------------------------------------
ham_kind(::TFIMHamiltonian) = Cint(0)
ham_kind(::EastModelHamiltonian) = Cint(1)
ham_params(H::TFIMHamiltonian) = Cdouble[H.J, H.h, H.g]
ham_params(H::EastModelHamiltonian) = Cdouble[H.c, H.A, H.B]

# same 4-tuple as the reference (src/gradients.jl:73-81; splatted at examples/hpc/utils.jl:57 and
# examples/nns/circuit_layer.jl:33): (cache, psi', psi'', psi).  The adjoint sweep keeps its own work states inside the
# library, but code that touches cache[2:4] (the reference's propagate_forwards! callers) finds real states there.
construct_grads_cache(psi::B200State) = (construct_apply_cache(psi), similar(psi), similar(psi), similar(psi))

function gradients!(::B200ApplyCache, _psi_p, _psi_pp, _psi, H::AbstractKernelHamiltonian, psi0::B200State,
                    circuit::GenericBrickworkCircuit; calculate_energy::Bool=false)
    grads = similar(circuit.gate_angles, Float64)
    angles = Matrix{Float64}(circuit.gate_angles)
    E = Ref{Cdouble}(0)
    check(ccall((:qcb_gradients_brickwork, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                psi0.handle, circuit.nbits, circuit.nlayers, angles, ham_kind(H), ham_params(H), calculate_energy ? 1 : 0, E, grads))
    return calculate_energy ? (E[], grads) : grads
end
gradients(H::AbstractKernelHamiltonian, psi::B200State, circuit::GenericBrickworkCircuit; kwargs...) =
    gradients!(construct_grads_cache(psi)..., H, psi, circuit; kwargs...)

# the same sweep for an explicit (dense or sparse, Hermitian) Hamiltonian: examples/test_gradients.jl:25
function gradients!(::B200ApplyCache, _psi_p, _psi_pp, _psi, H::AbstractMatrix, psi0::B200State,
                    circuit::GenericBrickworkCircuit; calculate_energy::Bool=false)
    grads = similar(circuit.gate_angles, Float64)
    angles = Matrix{Float64}(circuit.gate_angles)
    rowptr, cols, vals = _csr(H)
    E = Ref{Cdouble}(0)
    check(ccall((:qcb_gradients_brickwork_csr, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Clonglong, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                psi0.handle, circuit.nbits, circuit.nlayers, angles, size(H, 1), rowptr, cols, vals, calculate_energy ? 1 : 0, E, grads))
    return calculate_energy ? (E[], grads) : grads
end
gradients(H::AbstractMatrix, psi::B200State, circuit::GenericBrickworkCircuit; kwargs...) =
    gradients!(construct_grads_cache(psi)..., H, psi, circuit; kwargs...)

function propagate_forwards!(::B200ApplyCache, psi_p::B200State, psi::B200State, circuit::GenericBrickworkCircuit,
                             start_layer::Int, layer_gate_idx::Int, gate_idx::Int)
    _apply_circuit_from!(psi_p, psi, circuit, gate_idx - 1, circuit.ngates, false)
    return psi_p, psi
end

# optimise! with quirk Q1 resolved to the intended semantics (SURVEY.md 8a): energy + full gradient per epoch
function optimise!(circuit::GenericBrickworkCircuit, H::AbstractKernelHamiltonian, psi0::B200State, epochs, lr; use_progress=true)
    energies = zeros(Float64, epochs + 1)
    angles = Matrix{Float64}(circuit.gate_angles)
    check(ccall((:qcb_optimise_brickwork, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}),
                psi0.handle, circuit.nbits, circuit.nlayers, angles, ham_kind(H), ham_params(H), epochs, lr, energies))
    circuit.gate_angles .= angles
    return energies
end

# ------------------------------------------------------------------------------------------------
# building blocks of the statevector polar optimiser (ext/MPSExt/polar_optimise.jl:83-102)
# ------------------------------------------------------------------------------------------------
"E[r, c] = sum_rest conj(lam[r, rest]) psi[c, rest] for the qubit pair (K, K+1), as a 4x4 ComplexF64 matrix."
function environment_2q(lam::B200State, psi::B200State, K::Integer)
    E = zeros(ComplexF64, 4, 4)
    check(ccall((:qcb_environment_2q, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{ComplexF64}), lam.handle, psi.handle, K, E))
    return E
end
function LinearAlgebra.dot(a::B200State, b::B200State)
    out = zeros(Cdouble, 2)
    check(ccall((:qcb_state_inner, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}), a.handle, b.handle, out))
    return complex(out[1], out[2])
end

# ------------------------------------------------------------------------------------------------
# right_apply_gate!(M', M, gate) (src/apply.jl:144-182): M' = M * U on a 2^n x 2^n operator.  Seen as a 2n-qubit state
# (column-major: row bits first) this is the gate v[i, j, a, b] = u[a, b, i, j] on qubits (K + n, K + n + 1).
# ------------------------------------------------------------------------------------------------
function right_apply_gate!(M_out::Matrix{Complex{T}}, M::Matrix{Complex{T}}, gate::Localised2SpinAdjGate) where {T}
    ndims(M) == 2 || error("Must be a matrix")
    size(M, 1) == size(M, 2) || error("Must be a square matrix")
    size(M_out) == size(M) || error("M′ must be the same size as M")
    nbits = round(Int, log2(size(M, 1)))
    @assert size(M, 1) == 2^nbits "The matrix must have a power of two edge"
    @assert 0 < gate.target_gate_dim < nbits "The gate cannot be applied as it sits outside the qubit space."
    u = ComplexF64.(mat(gate))
    v = permutedims(u, (3, 4, 1, 2))
    st = B200State(reshape(M, ntuple(_ -> 2, 2 * nbits)))
    check(ccall((:qcb_apply_2q, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint), st.handle, gate_block(v), gate.target_gate_dim + nbits))
    copyto!(M_out, reshape(Array(st), size(M)))
    return M_out
end

# ------------------------------------------------------------------------------------------------
# Sharded states: one Julia process per GPU (e.g. under MPI.jl or Distributed.jl), amplitudes sharded by the top
# log2(nranks) qubits (include/qcb200.h "multi-GPU").  No reference counterpart: the reference holds one array.
#
#     id = rank == 0 ? dist_unique_id() : nothing;  id = MPI.bcast(id, 0, comm)        # any host channel
#     dist_init(rank, nranks, id)
#     psi = B200ShardedState{Float64}(34)                    # 2^(34 - log2 nranks) amplitudes on this rank's GPU
#     map_peers!(psi, h -> MPI.Allgather(h, comm))           # optional: swaps fused into gate passes over NVLink (or, without
#                                                            # second buffers, in-place swaps over peer memory)
#     apply!(psi, circuit);  E = measure(H, psi);  E, g = gradients(H, psi, circuit; calculate_energy=true)
# ------------------------------------------------------------------------------------------------
dist_unique_id() = (id = zeros(UInt8, 128); check(ccall((:qcb_dist_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
dist_init(rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    (length(id) == 128 || error("the unique id is 128 bytes"); check(ccall((:qcb_dist_init, LIB), Cint, (Cint, Cint, Ptr{UInt8}), rank, nranks, id)))
dist_finalize() = check(ccall((:qcb_dist_finalize, LIB), Cint, ()))

mutable struct B200ShardedState{T<:AbstractFloat}
    handle::Ptr{Cvoid}
    nqubits::Int
    mapped::Bool
    function B200ShardedState{T}(nqubits::Integer) where {T}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qcb_state_create_sharded, LIB), Cint, (Cint, Cint, Ptr{Ptr{Cvoid}}), nqubits, dtype_code(T), h))
        s = new{T}(h[], nqubits, false)
        # peers must be unmapped on every rank (unmap_peers! + a host barrier) before any rank frees its shard
        finalizer(x -> (x.mapped || ccall((:qcb_state_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle)), s)
        return s
    end
end
Base.eltype(::B200ShardedState{T}) where {T} = Complex{T}
function local_qubits(s::B200ShardedState)
    out = Ref{Cint}(0)
    check(ccall((:qcb_state_local_qubits, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), s.handle, out))
    return Int(out[])
end
"`allgather(bytes::Vector{UInt8})` returns the concatenation over ranks, in rank order (MPI.Allgather, ...)."
function map_peers!(s::B200ShardedState, allgather)
    mine = zeros(UInt8, 128)
    check(ccall((:qcb_state_ipc_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), s.handle, mine))
    table = Vector{UInt8}(allgather(mine))
    check(ccall((:qcb_state_ipc_import, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), s.handle, table))
    s.mapped = true
    return s
end
"Unmap the peers' shards; synchronise the ranks (host barrier) between this call and the destruction of the state."
function unmap_peers!(s::B200ShardedState)
    check(ccall((:qcb_state_ipc_close, LIB), Cint, (Ptr{Cvoid},), s.handle))
    s.mapped = false
    finalizer(x -> ccall((:qcb_state_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), s)
    return s
end
set_zero_state!(s::B200ShardedState) = (check(ccall((:qcb_state_set_zero_state, LIB), Cint, (Ptr{Cvoid},), s.handle)); s)
"This rank's canonical shard: amplitudes [rank * 2^nloc, (rank + 1) * 2^nloc) of the full vector."
function upload_shard!(s::B200ShardedState{T}, shard::Vector{Complex{T}}) where {T}
    length(shard) == 1 << local_qubits(s) || error("ψ′ must be the same size as ψ.")
    check(ccall((:qcb_state_upload_shard, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s.handle, shard))
    return s
end
function download_shard(s::B200ShardedState{T}) where {T}
    out = Vector{Complex{T}}(undef, 1 << local_qubits(s))
    check(ccall((:qcb_state_download_shard, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s.handle, out))
    return out
end
LinearAlgebra.norm(s::B200ShardedState) = begin   # all-reduced over the ranks inside the library
    out = Ref{Cdouble}(0)
    check(ccall((:qcb_state_norm2, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
    sqrt(out[])
end
# in place: the state IS the distributed array, there is no second copy to ping-pong with
function apply!(s::B200ShardedState, c::GenericBrickworkCircuit)
    angles = Matrix{Float64}(c.gate_angles)
    check(ccall((:qcb_apply_brickwork_range, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Cint, Cint),
                s.handle, c.nbits, c.nlayers, angles, 0, c.ngates, 0))
    return s
end
function apply!(s::B200ShardedState, gates::AbstractArray{<:AbstractGate})
    G = length(gates)
    K = Cint[g.target_gate_dim for g in gates]
    ar = Cint[g isa QuantumCircuits.Abstract2SpinGate ? 2 : 1 for g in gates]
    mats = zeros(ComplexF64, 16, G)
    for (i, g) in enumerate(gates)
        mats[:, i] .= gate_block(mat(g))
    end
    check(ccall((:qcb_apply_gates, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{ComplexF64}), s.handle, G, K, ar, mats))
    return s
end
function measure(H::TFIMHamiltonian, s::B200ShardedState)
    E = zeros(Cdouble, 2)
    check(ccall((:qcb_measure_tfim, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}), s.handle, H.J, H.h, H.g, E))
    return E[1]
end
function measure(H::EastModelHamiltonian, s::B200ShardedState)
    E = zeros(Cdouble, 2)
    check(ccall((:qcb_measure_east, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}), s.handle, H.c, H.A, H.B, E))
    return E[1]
end
function gradients(H::AbstractKernelHamiltonian, psi0::B200ShardedState, circuit::GenericBrickworkCircuit; calculate_energy::Bool=false)
    grads = similar(circuit.gate_angles, Float64)
    angles = Matrix{Float64}(circuit.gate_angles)
    E = Ref{Cdouble}(0)
    check(ccall((:qcb_gradients_brickwork, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                psi0.handle, circuit.nbits, circuit.nlayers, angles, ham_kind(H), ham_params(H), calculate_energy ? 1 : 0, E, grads))
    return calculate_energy ? (E[], grads) : grads   # identical on every rank (all-reduced environments)
end
function optimise!(circuit::GenericBrickworkCircuit, H::AbstractKernelHamiltonian, psi0::B200ShardedState, epochs, lr; use_progress=true)
    energies = zeros(Float64, epochs + 1)
    angles = Matrix{Float64}(circuit.gate_angles)
    check(ccall((:qcb_optimise_brickwork, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}),
                psi0.handle, circuit.nbits, circuit.nlayers, angles, ham_kind(H), ham_params(H), epochs, lr, energies))
    circuit.gate_angles .= angles
    return energies
end

end # module
