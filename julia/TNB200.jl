# TNB200.jl -- Julia glue that re-points the ITensorsGPU.jl hot-path overrides at libtnb200.so.
#
# UNTESTED HERE: the build container has no Julia (see DESIGN.md).  Every method below replaces the
# BODY of a method the reference defines on CuDense storage (file:line cited per method, relative to
# the ITensorsGPU.jl tree); signatures and return values are the reference's, so ITensors.jl keeps
# dispatching to them unchanged.  `include` this file after `using ITensorsGPU`.
module TNB200

using CUDA, ITensors, ITensors.NDTensors, LinearAlgebra
import ITensorsGPU: CuDense, CuDenseTensor
import ITensors.NDTensors: ContractionProperties, Spectrum, Dense, Diag, Tensor, inds, dims, store, data, ind

const LIB = get(ENV, "TNB200_LIB", "libtnb200.so")
const HANDLE = Ref{Ptr{Cvoid}}(C_NULL)

struct BondDims
  chiL::Int64; chiR::Int64; d1::Int32; d2::Int32; wL::Int32; wM::Int32; wR::Int32
end

function handle()
  if HANDLE[] == C_NULL
    rc = ccall((:tnb_create, LIB), Cint, (Ref{Ptr{Cvoid}},), HANDLE)
    rc == 0 || error("tnb_create failed ($rc): a B200 (sm_100) device is required; there is no CPU fallback")
  end
  return HANDLE[]
end

function check(rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:tnb_last_error, LIB), Cstring, (Ptr{Cvoid},), handle()))
  rc == 2 && throw(DimensionMismatch(msg))      # src/cuitensor.jl:55,72,78
  rc == 1 && throw(ArgumentError(msg))          # src/mps/cumpo.jl:21
  error("libtnb200: $msg")                      # src/tensor/cudense.jl:165
end

dtype(::Type{Float64}) = Cint(0)
dtype(::Type{ComplexF64}) = Cint(1)
stream() = CUDA.stream().handle                  # task-local stream, as CUDA.jl library wrappers use
ptr(x::CuArray) = reinterpret(Ptr{Cvoid}, pointer(x))

# ---- _contract!  (src/tensor/cudense.jl:238-331): mode labels come straight from ContractionProperties
function NDTensors._contract!(CT::CuDenseTensor{El,NC}, AT::CuDenseTensor{El,NA}, BT::CuDenseTensor{El,NB},
                              props::ContractionProperties, α::Number=one(El), β::Number=zero(El)) where {El,NC,NA,NB}
  ea, eb, ec = Int64[dims(inds(AT))...], Int64[dims(inds(BT))...], Int64[dims(inds(CT))...]
  # NDTensors labels: negative = contracted, positive = free; shift to non-negative ints
  off = 1 - min(minimum(props.ai; init=0), minimum(props.bi; init=0))
  ma, mb, mc = Int32.(props.ai .+ off), Int32.(props.bi .+ off), Int32.(props.ci .+ off)
  a, b = Ref(El(α)), Ref(El(β))
  check(ccall((:tnb_contract, LIB), Cint,
              (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid},
               Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
              handle(), dtype(El), NA, ea, ma, ptr(data(store(AT))), NB, eb, mb, ptr(data(store(BT))),
              NC, ec, mc, ptr(data(store(CT))), a, b, 0, stream()))
  return data(store(CT))                          # the reference returns parent(Cdata): cudense.jl:330
end

# ---- workspace queries (host arithmetic; no GPU): reserve the arena once instead of growing it mid-sweep
bond_workspace_bytes(op::Integer, ::Type{El}, d::BondDims; ortho::Integer=0, noise::Bool=false, krylovdim::Integer=3) where {El} =
  ccall((:tnb_bond_workspace_bytes, LIB), Csize_t, (Cint, Cint, Ref{BondDims}, Cint, Cint, Cint, Ptr{Int32}),
        op, dtype(El), d, ortho, noise ? 1 : 0, krylovdim, C_NULL)
matrix_workspace_bytes(op::Integer, ::Type{El}, m::Integer, n::Integer) where {El} =
  ccall((:tnb_matrix_workspace_bytes, LIB), Csize_t, (Cint, Cint, Int64, Int64), op, dtype(El), m, n)
reserve!(nbytes::Integer) = check(ccall((:tnb_reserve, LIB), Cint, (Ptr{Cvoid}, Csize_t), handle(), nbytes))

# ---- planner dry run (no GPU): how a contraction would be matricised -- useful from the REPL when a layout is slow
struct PlanDesc
  M::Int64; N::Int64; K::Int64
  n_m::Int32; n_n::Int32; n_k::Int32; family::Int32
  ext_m::NTuple{12,Int64}; a_stride_m::NTuple{12,Int64}; c_stride_m::NTuple{12,Int64}
  ext_n::NTuple{12,Int64}; b_stride_n::NTuple{12,Int64}; c_stride_n::NTuple{12,Int64}
  ext_k::NTuple{12,Int64}; a_stride_k::NTuple{12,Int64}; b_stride_k::NTuple{12,Int64}
  a_k_major::Int32; b_k_major::Int32; a_vec::Int32; b_vec::Int32
  tile_m::Int32; tile_n::Int32; tile_k::Int32; herm_upper::Int32
  tiles::Int64; waves::Float64
end
function plan_describe(::Type{El}, ea::Vector{Int64}, ma::Vector{Int32}, eb::Vector{Int64}, mb::Vector{Int32},
                       ec::Vector{Int64}, mc::Vector{Int32}; flags::Integer=0, num_sms::Integer=148) where {El}
  out = Ref{PlanDesc}()
  err = zeros(UInt8, 512)
  rc = ccall((:tnb_plan_describe, LIB), Cint,
             (Cint, Cint, Ptr{Int64}, Ptr{Int32}, Cint, Ptr{Int64}, Ptr{Int32}, Cint, Ptr{Int64}, Ptr{Int32}, Cint, Cint,
              Ref{PlanDesc}, Ptr{UInt8}, Csize_t),
             dtype(El), length(ea), ea, ma, length(eb), eb, mb, length(ec), ec, mc, flags, num_sms, out, err, 512)
  rc == 0 || (rc == 2 ? throw(DimensionMismatch(unsafe_string(pointer(err)))) : throw(ArgumentError(unsafe_string(pointer(err)))))
  return out[]
end

"plan_describe for operands that are strided windows of larger tensors (element strides per mode)"
function plan_describe_strided(::Type{El}, ea::Vector{Int64}, ma::Vector{Int32}, sa::Vector{Int64}, eb::Vector{Int64},
                               mb::Vector{Int32}, sb::Vector{Int64}, ec::Vector{Int64}, mc::Vector{Int32}, sc::Vector{Int64};
                               flags::Integer=0, num_sms::Integer=148) where {El}
  out = Ref{PlanDesc}()
  err = zeros(UInt8, 512)
  rc = ccall((:tnb_plan_describe_strided, LIB), Cint,
             (Cint, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Int32},
              Ptr{Int64}, Cint, Cint, Ref{PlanDesc}, Ptr{UInt8}, Csize_t),
             dtype(El), length(ea), ea, ma, sa, length(eb), eb, mb, sb, length(ec), ec, mc, sc, flags, num_sms, out, err, 512)
  rc == 0 || (rc == 2 ? throw(DimensionMismatch(unsafe_string(pointer(err)))) : throw(ArgumentError(unsafe_string(pointer(err)))))
  return out[]
end

# ---- permute!  (src/tensor/cudense.jl:447-478)  and  + / -  (src/tensor/cudense.jl:333-445)
function permute_axpby!(B::CuDenseTensor{El}, A::CuDenseTensor{El}, α, β) where {El}
  n = length(inds(A))
  modeA = Int32.(1:n)
  modeB = Int32[findfirst(==(i), inds(A)) for i in inds(B)]
  a, b = Ref(El(α)), Ref(El(β))
  check(ccall((:tnb_permute_axpby, LIB), Cint,
              (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(El), n, Int64[dims(inds(A))...], modeA, ptr(data(store(A))), modeB, ptr(data(store(B))), a, b, stream()))
  return B
end
Base.permute!(B::CuDenseTensor, A::CuDenseTensor) = vec(data(store(permute_axpby!(B, A, 1, 0))))
Base.:+(B::CuDenseTensor, A::CuDenseTensor) = permute_axpby!(B, A, 1, 1)
Base.:-(B::CuDenseTensor, A::CuDenseTensor) = permute_axpby!(B, A, -1, 1)

# ---- norm  (src/tensor/cudense.jl:27)
function LinearAlgebra.norm(T::CuDenseTensor{El}) where {El}
  r = Ref{Float64}(0)
  check(ccall((:tnb_nrm2, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(El), length(data(store(T))), ptr(data(store(T))), C_NULL, r, stream()))
  return r[]
end

# ---- truncate!  (src/tensor/cutruncate.jl:1-93) -- CPU rule, one kernel + one 24-byte readback
function NDTensors.truncate!(P::CuVector{Float64}; kwargs...)
  maxdim = Int64(get(kwargs, :maxdim, length(P))); mindim = Int64(get(kwargs, :mindim, 1))
  cutoff = Float64(get(kwargs, :cutoff, 0.0))
  flags = Cint((get(kwargs, :absoluteCutoff, get(kwargs, :use_absolute_cutoff, false)) ? 1 : 0) |
               (get(kwargs, :doRelCutoff, get(kwargs, :use_relative_cutoff, true)) ? 0 : 2))
  n, err, docut = Ref{Int64}(0), Ref{Float64}(0), Ref{Float64}(0)
  check(ccall((:tnb_truncate, LIB), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Float64, Cint, Ref{Int64}, Ref{Float64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), ptr(P), length(P), maxdim, mindim, cutoff, flags, n, err, docut, stream()))
  return err[], docut[], P[1:n[]]
end

# ---- svd  (src/tensor/culinearalgebra.jl:33-72).  Follows the CPU convention: A = U*S*V with V already conjugated.
function LinearAlgebra.svd(T::CuDenseTensor{ElT,2,IndsT}; kwargs...) where {ElT,IndsT}
  m, n = dims(T)
  dotrunc = haskey(kwargs, :maxdim) || haskey(kwargs, :cutoff)
  maxdim = Int64(get(kwargs, :maxdim, min(m, n))); kmax = dotrunc ? min(m, n, maxdim) : min(m, n)
  A = copy(data(store(T)))
  U, S, V = CUDA.zeros(ElT, m * kmax), CUDA.zeros(Float64, kmax), CUDA.zeros(ElT, n * kmax)
  nk, err = Ref{Int64}(0), Ref{Float64}(0)
  check(ccall((:tnb_svd_trunc, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Float64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
               Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), m, n, ptr(A), maxdim, Int64(get(kwargs, :mindim, 1)), Float64(get(kwargs, :cutoff, 0.0)),
              0, dotrunc ? 1 : 0, ptr(U), ptr(S), ptr(V), nk, err, stream()))
  k = nk[]
  u = eltype(IndsT)(k); v = eltype(IndsT)(k)
  return Tensor(Dense(U[1:m*k]), IndsT((ind(T, 1), u))), Tensor(Diag(S[1:k]), IndsT((u, v))),
         Tensor(Dense(V[1:n*k]), IndsT((ind(T, 2), v))), Spectrum(S[1:k] .^ 2, err[])
end

# ---- eigen  (src/tensor/culinearalgebra.jl:74-108)
function LinearAlgebra.eigen(T::Hermitian{ElT,<:CuDenseTensor{ElT,2,IndsT}}; kwargs...) where {ElT,IndsT}
  n = dims(parent(T))[1]
  dotrunc = haskey(kwargs, :maxdim) || haskey(kwargs, :cutoff)
  maxdim = Int64(get(kwargs, :maxdim, n)); kmax = dotrunc ? min(n, maxdim) : n
  A = copy(data(store(parent(T))))
  D, U = CUDA.zeros(Float64, kmax), CUDA.zeros(ElT, n * kmax)
  nk, err = Ref{Int64}(0), Ref{Float64}(0)
  check(ccall((:tnb_eigh_trunc, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Int64, Int64, Float64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), n, ptr(A), maxdim, Int64(get(kwargs, :mindim, 1)), Float64(get(kwargs, :cutoff, 0.0)),
              0, dotrunc ? 1 : 0, ptr(D), ptr(U), nk, err, stream()))
  k = nk[]
  l = eltype(IndsT)(k); r = eltype(IndsT)(k)
  return Tensor(Diag(D[1:k]), IndsT((l, dag(r)))), Tensor(Dense(U[1:n*k]), IndsT((dag(ind(parent(T), 2)), dag(r)))), Spectrum(D[1:k], err[])
end

# ---- qr  (src/tensor/culinearalgebra.jl:110-121)
function LinearAlgebra.qr(T::CuDenseTensor{ElT,2,IndsT}; kwargs...) where {ElT,IndsT}
  m, n = dims(T); k = min(m, n)
  Q, R = CUDA.zeros(ElT, m * k), CUDA.zeros(ElT, k * n)
  check(ccall((:tnb_qr, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), m, n, ptr(data(store(T))), ptr(Q), ptr(R), stream()))
  q = eltype(IndsT)(k)
  return Tensor(Dense(Q), IndsT((ind(T, 1), q))), Tensor(Dense(R), IndsT((q, ind(T, 2))))
end

# ---- fused tier: product(::ProjMPO, ::ITensor) for the two-site case ([EXT] ITensors src/mps/projmpo.jl)
# The glue permutes L, R, W_b, W_{b+1} once per `position!` into the fixed layouts of include/tnb200.h and then
# calls tnb_heff_apply; kept out of this file's executable part because ProjMPO internals differ between
# ITensors 0.2.x patch releases -- see INTEGRATION.md for the call.

# ---- Diag x Dense contraction over ONE index of the Diag tensor (src/tensor/cudiag.jl:105-161 densifies instead)
# `shared` is the contracted label, `labelsC` the NDTensors output order with the Diag's other index already
# substituted for `shared` (the relabel is metadata only).
function diag_contract!(C::CuArray{ElT}, A::CuArray{ElT}, labelsA::Vector{Int32}, shared::Int32, d::CuVector,
                        labelsC::Vector{Int32}) where {ElT}
  check(ccall((:tnb_diag_contract, LIB), Cint,
              (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), ndims(A), Int64[size(A)...], labelsA, ptr(A), shared, ptr(d), dtype(eltype(d)), labelsC,
              ptr(C), stream()))
  return C
end

# ---- TEBD gate in B form (right-canonical tensors in the Schmidt bases + Schmidt values): the unit of work of an
# even/odd layer that can be spread over GPUs ([EXT] apply(gates, psi), examples/gate_evolution.jl:46)
function tebd_gate_bform!(G::CuArray, lamL::CuVector{Float64}, B1::CuArray{ElT,3}, B2::CuArray{ElT,3};
                          maxdim::Int=0, mindim::Int=1, cutoff::Float64=0.0) where {ElT}
  chiL, d1, chiM = size(B1); _, d2, chiR = size(B2)
  kmax = maxdim > 0 ? min(chiL * d1, d2 * chiR, maxdim) : min(chiL * d1, d2 * chiR)
  b1 = CUDA.zeros(ElT, max(length(B1), chiL * d1 * kmax)); copyto!(b1, vec(B1))
  b2 = CUDA.zeros(ElT, max(length(B2), kmax * d2 * chiR)); copyto!(b2, vec(B2))
  lam = CUDA.zeros(Float64, kmax); nk = Ref{Int64}(0); err = Ref{Float64}(0.0)
  check(ccall((:tnb_tebd_gate_bform, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
               Int64, Int64, Float64, Ptr{Cvoid}, Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), chiL, chiM, chiR, d1, d2, ptr(G), ptr(lamL), ptr(b1), ptr(b2),
              maxdim, mindim, cutoff, ptr(lam), nk, err, stream()))
  k = nk[]
  return reshape(b1[1:chiL*d1*k], chiL, d1, k), reshape(b2[1:k*d2*chiR], k, d2, chiR), lam[1:k], err[]
end

# ---- multi-GPU: peer-mapped buffers and the slab GEMM with the all-gather fused into its epilogue.
# One Julia process per GPU; exchange the 64-byte handles with MPI.jl / Distributed (host side only).
function peer_alloc(nbytes::Integer)
  p = Ref{Ptr{Cvoid}}(C_NULL); hd = zeros(UInt8, 64)
  check(ccall((:tnb_peer_alloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}, Ptr{UInt8}), handle(), nbytes, p, hd))
  return p[], hd
end
function peer_open(hd::Vector{UInt8})
  p = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:tnb_peer_open, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ref{Ptr{Cvoid}}), handle(), hd, p))
  return p[]
end
# out_peers / flag_peers: Vector{Ptr{Cvoid}} of length world (own buffer included); epoch increases by one per call
function heff_apply_shard_fused!(out_peers::Vector{Ptr{Cvoid}}, flag_peers::Vector{Ptr{Cvoid}}, epoch::Integer, rank::Integer,
                                 d::BondDims, Lslab::CuArray, W1::CuArray, W2::CuArray, R::CuArray, phi::CuArray)
  world = length(out_peers)
  check(ccall((:tnb_heff_apply_shard_fused, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
               Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, UInt64, Ptr{Cvoid}),
              handle(), dtype(eltype(phi)), Ref(d), rank, world, size(Lslab, 2), ptr(Lslab), ptr(W1), ptr(W2), ptr(R), ptr(phi),
              out_peers, flag_peers, UInt64(epoch), stream()))
end


# =====================================================================================================
# Tier 2 -- fused DMRG / TEBD entry points.  These replace the SEQUENCES of tier-1 calls that [EXT] ITensors.jl
# issues for `product(::ProjMPO, v)`, `makeL!/makeR!`, `KrylovKit.eigsolve`, `noiseterm`, `replacebond!` and one
# bond of `dmrg` (call sites: examples/dmrg.jl:25, test/dmrg.jl:27,75).  Fixed layouts (column-major CuArrays):
# phi[l,s1,s2,r], L[l,l',a], R[r,r',c], W[a,s,s',b] (s = ket), A[l,s,r].  `layout(T, inds...)` below is the one
# permute (tnb_permute_axpby) a maintainer needs when an ITensor's index order differs.
# =====================================================================================================
bond_dims(phi::CuArray{<:Any,4}, W1::CuArray{<:Any,4}, W2::CuArray{<:Any,4}) =
  Ref(BondDims(size(phi, 1), size(phi, 4), size(phi, 2), size(phi, 3), size(W1, 1), size(W1, 4), size(W2, 4)))

"ITensor -> CuArray in the index order `is` (a view when the order already matches, one fused permute otherwise)"
function layout(T::ITensor, is::Index...)
  Tp = inds(T) == IndexSet(is...) ? T : permute(T, is...)          # -> tnb_permute_axpby through permute! above
  return reshape(data(store(tensor(Tp))), dim.(is)...)
end

# ---- product(::ProjMPO, ::ITensor)  ([EXT] projmpo.jl; four cuTENSOR calls + three allocations in the reference)
function heff_apply!(out::CuArray{ElT,4}, L::CuArray{ElT,3}, W1::CuArray{ElT,4}, W2::CuArray{ElT,4}, R::CuArray{ElT,3},
                     phi::CuArray{ElT,4}) where {ElT}
  check(ccall((:tnb_heff_apply, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), bond_dims(phi, W1, W2), ptr(L), ptr(W1), ptr(W2), ptr(R), ptr(phi), ptr(out), stream()))
  return out
end

"drop-in body for `product(P::ProjMPO, v::ITensor)` in the two-site case (P.nsite == 2, both environments present)"
function product_two_site(P, v::ITensor, l::Index, s1::Index, s2::Index, r::Index)
  b = P.lpos + 1
  Lt, Rt = P.LR[P.lpos], P.LR[P.rpos]
  a, c = commonind(Lt, P.H[b]), commonind(Rt, P.H[b + 1])
  m = commonind(P.H[b], P.H[b + 1])
  phi = layout(v, l, s1, s2, r)
  out = similar(phi)
  heff_apply!(out, layout(Lt, l, l', a), layout(P.H[b], a, s1, s1', m), layout(P.H[b + 1], m, s2, s2', c),
              layout(Rt, r, r', c), phi)
  return itensor(out, l, s1, s2, r)                                 # == noprime(((((v*L)*W1)*W2)*R)
end

# ---- makeL! / makeR!  ([EXT] projmpo.jl `position!`)
function env_update_left!(Lnew::CuArray{ElT,3}, L::CuArray{ElT,3}, A::CuArray{ElT,3}, W::CuArray{ElT,4}) where {ElT}
  check(ccall((:tnb_env_update_left, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), size(A, 1), size(A, 3), size(A, 2), size(W, 1), size(W, 4), ptr(L), ptr(A), ptr(W), ptr(Lnew), stream()))
  return Lnew
end
function env_update_right!(Rnew::CuArray{ElT,3}, R::CuArray{ElT,3}, A::CuArray{ElT,3}, W::CuArray{ElT,4}) where {ElT}
  check(ccall((:tnb_env_update_right, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), size(A, 1), size(A, 3), size(A, 2), size(W, 1), size(W, 4), ptr(R), ptr(A), ptr(W), ptr(Rnew), stream()))
  return Rnew
end

# ---- KrylovKit.eigsolve(PH, phi, 1, :SR; ishermitian=true, tol=1e-14, krylovdim=3, maxiter=1)  ([EXT] dmrg.jl)
function eigsolve_lanczos!(phi::CuArray{ElT,4}, L, W1, W2, R; krylovdim::Int=3, maxiter::Int=1, tol::Float64=1e-14) where {ElT}
  e, nmv = Ref{Float64}(0.0), Ref{Cint}(0)
  check(ccall((:tnb_eigsolve_lanczos, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Float64,
               Ref{Float64}, Ref{Cint}, Ptr{Cvoid}),
              handle(), dtype(ElT), bond_dims(phi, W1, W2), ptr(L), ptr(W1), ptr(W2), ptr(R), ptr(phi), krylovdim, maxiter, tol,
              e, nmv, stream()))
  return e[], phi, Int(nmv[])                                       # phi overwritten with the normalised Ritz vector
end

# ---- noiseterm(::ProjMPO, phi, ortho)  ([EXT] projmpo.jl): rho (upper triangle) = noise * nt nt'
function noise_term!(rho::CuMatrix{ElT}, L, W1, W2, R, phi::CuArray{ElT,4}, ortho::String, noise::Float64;
                     accumulate::Bool=false) where {ElT}
  check(ccall((:tnb_noise_term, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Float64, Cint,
               Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), bond_dims(phi, W1, W2), ptr(L), ptr(W1), ptr(W2), ptr(R), ptr(phi), ortho == "left" ? 0 : 1,
              noise, accumulate ? 1 : 0, ptr(rho), stream()))
  return rho
end

"output capacity of the factorizing calls (include/tnb200.h, tnb_factorize_bond): eigen branch is bounded by the ortho side"
function kmax_for(m::Int, n::Int, ortho::String, maxdim::Int, eigen_branch::Bool)
  r = eigen_branch ? (ortho == "left" ? m : n) : min(m, n)
  return maxdim > 0 ? min(r, maxdim) : r
end

# ---- replacebond!(psi, b, phi; ortho, which_decomp, maxdim, mindim, cutoff, eigen_perturbation, normalize)
#      -> factorize  ([EXT] mps.jl / decomp.jl; svd / eigen bodies: src/tensor/culinearalgebra.jl:33-108)
function factorize_bond!(phi::CuArray{ElT,4}; ortho::String="left", which_decomp::Int=0, maxdim::Int=0, mindim::Int=1,
                         cutoff::Float64=0.0, rho_pert::Union{Nothing,CuMatrix{ElT}}=nothing, normalize::Bool=false) where {ElT}
  cl, d1, d2, cr = size(phi)
  eig = which_decomp == 2 || (which_decomp == 0 && (rho_pert !== nothing || cutoff > 1e-12))
  k = kmax_for(cl * d1, d2 * cr, ortho, maxdim, eig)
  A, B = CuArray{ElT}(undef, cl * d1 * k), CuArray{ElT}(undef, k * d2 * cr)
  nk, err = Ref{Int64}(0), Ref{Float64}(0.0)
  dims_ = Ref(BondDims(cl, cr, d1, d2, 1, 1, 1))
  check(ccall((:tnb_factorize_bond, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Float64, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid},
               Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), dims_, ptr(phi), ortho == "left" ? 0 : 1, which_decomp, maxdim, mindim, cutoff,
              rho_pert === nothing ? C_NULL : ptr(rho_pert), normalize ? 1 : 0, ptr(A), ptr(B), nk, err, stream()))
  n = Int(nk[])
  return reshape(view(A, 1:cl*d1*n), cl, d1, n), reshape(view(B, 1:n*d2*cr), n, d2, cr), Spectrum(nothing, err[])
end

# ---- one two-site bond of dmrg(): phi = A1*A2, eigsolve, noise term, replacebond!  (ONE call, one sync)
function dmrg_bond_step!(L, W1, W2, R, A1::CuArray{ElT,3}, A2::CuArray{ElT,3}; ortho::String="left", which_decomp::Int=0,
                         maxdim::Int, mindim::Int=1, cutoff::Float64=0.0, noise::Float64=0.0, krylovdim::Int=3,
                         maxiter::Int=1) where {ElT}
  cl, d1, cm = size(A1); _, d2, cr = size(A2)
  eig = which_decomp == 2 || (which_decomp == 0 && (noise > 0 || cutoff > 1e-12))
  k = kmax_for(cl * d1, d2 * cr, ortho, maxdim, eig)
  b1 = CuArray{ElT}(undef, max(length(A1), cl * d1 * k)); copyto!(b1, 1, vec(A1), 1, length(A1))
  b2 = CuArray{ElT}(undef, max(length(A2), k * d2 * cr)); copyto!(b2, 1, vec(A2), 1, length(A2))
  e, nk, err = Ref{Float64}(0.0), Ref{Int64}(0), Ref{Float64}(0.0)
  dims_ = Ref(BondDims(cl, cr, d1, d2, size(W1, 1), size(W1, 4), size(W2, 4)))
  check(ccall((:tnb_dmrg_bond_step, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint,
               Int64, Int64, Float64, Float64, Cint, Cint, Ref{Float64}, Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), dims_, cm, ptr(L), ptr(W1), ptr(W2), ptr(R), ptr(b1), ptr(b2), ortho == "left" ? 0 : 1,
              which_decomp, maxdim, mindim, cutoff, noise, krylovdim, maxiter, e, nk, err, stream()))
  n = Int(nk[])
  return e[], reshape(view(b1, 1:cl*d1*n), cl, d1, n), reshape(view(b2, 1:n*d2*cr), n, d2, cr), err[]
end

# ---- apply(gate, psi) on sites (n, n+1)  ([EXT] abstractmps.jl `product`; examples/gate_evolution.jl:46)
function tebd_apply_gate!(G::CuArray{ElT,4}, A1::CuArray{ElT,3}, A2::CuArray{ElT,3}; maxdim::Int=0, mindim::Int=1,
                          cutoff::Float64=0.0) where {ElT}
  cl, d1, cm = size(A1); _, d2, cr = size(A2)
  k = kmax_for(cl * d1, d2 * cr, "left", maxdim, cutoff > 1e-12)
  b1 = CuArray{ElT}(undef, max(length(A1), cl * d1 * k)); copyto!(b1, 1, vec(A1), 1, length(A1))
  b2 = CuArray{ElT}(undef, max(length(A2), k * d2 * cr)); copyto!(b2, 1, vec(A2), 1, length(A2))
  nk, err = Ref{Int64}(0), Ref{Float64}(0.0)
  check(ccall((:tnb_tebd_apply_gate, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Float64,
               Ref{Int64}, Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), cl, cm, cr, d1, d2, ptr(G), ptr(b1), ptr(b2), maxdim, mindim, cutoff, nk, err, stream()))
  n = Int(nk[])
  return reshape(view(b1, 1:cl*d1*n), cl, d1, n), reshape(view(b2, 1:n*d2*cr), n, d2, cr), err[]
end

# ---- plan cache / autotune (the `ContractionPlans` dictionary of src/ITensorsGPU.jl:54-55) and workspace limit
set_autotune(on::Bool) = check(ccall((:tnb_set_autotune, LIB), Cint, (Ptr{Cvoid}, Cint), handle(), on ? 1 : 0))
set_workspace_limit(bytes::Integer) = check(ccall((:tnb_set_workspace_limit, LIB), Cint, (Ptr{Cvoid}, Csize_t), handle(), bytes))

# ---- multi-GPU peer group (one Julia process per GPU, e.g. under MPI.jl: the 64-byte IPC handles travel by MPI.Allgather)
function comm_init(rank::Integer, flag_peers::Vector{Ptr{Cvoid}})
  check(ccall((:tnb_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Ptr{Cvoid}}), handle(), rank, length(flag_peers), flag_peers))
end
function dmrg_bond_step_shard!(Lslab, W1, W2, R, b1::CuVector{ElT}, b2::CuVector{ElT}, dims_::Ref{BondDims}, chiM::Integer,
                               out_a::Vector{Ptr{Cvoid}}, out_b::Vector{Ptr{Cvoid}}, stage::Vector{Ptr{Cvoid}}; ortho::String="left",
                               which_decomp::Int=0, maxdim::Int, mindim::Int=1, cutoff::Float64=0.0, noise::Float64=0.0,
                               krylovdim::Int=3, maxiter::Int=1) where {ElT}
  e, nk, err = Ref{Float64}(0.0), Ref{Int64}(0), Ref{Float64}(0.0)
  check(ccall((:tnb_dmrg_bond_step_shard, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint,
               Int64, Int64, Float64, Float64, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ref{Float64}, Ref{Int64},
               Ref{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), dims_, chiM, ptr(Lslab), ptr(W1), ptr(W2), ptr(R), ptr(b1), ptr(b2), ortho == "left" ? 0 : 1,
              which_decomp, maxdim, mindim, cutoff, noise, krylovdim, maxiter, out_a, out_b, stage, e, nk, err, stream()))
  return e[], Int(nk[]), err[]
end

# ---- one full sweep of dmrg(): 2(N-1) bond steps + environment updates, host loop inside the library
# A, W, env: Vector{CuArray} (site tensors / MPO tensors / one environment buffer per boundary 0..N), chi updated in place
function dmrg_sweep!(A::Vector{<:CuArray{ElT}}, W::Vector{<:CuArray{ElT}}, env::Vector{<:CuArray{ElT}}, chi::Vector{Int64},
                     d::Vector{Int32}, w::Vector{Int32}; build_right_envs::Bool, maxdim::Int, mindim::Int=1, cutoff::Float64=0.0,
                     noise::Float64=0.0, which_decomp::Int=0, krylovdim::Int=3, maxiter::Int=1) where {ElT}
  N = length(A)
  e, merr = Ref{Float64}(0.0), Ref{Float64}(0.0)
  be, bt = zeros(Float64, 2(N - 1)), zeros(Float64, 2(N - 1))
  check(ccall((:tnb_dmrg_sweep, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
               Ptr{Int64}, Cint, Int64, Int64, Float64, Float64, Cint, Cint, Cint, Ref{Float64}, Ref{Float64}, Ptr{Float64},
               Ptr{Float64}, Ptr{Cvoid}),
              handle(), dtype(ElT), N, chi, d, w, ptr.(A), Int64.(length.(A)), ptr.(W), ptr.(env), Int64.(length.(env)),
              build_right_envs ? 1 : 0, maxdim, mindim, cutoff, noise, which_decomp, krylovdim, maxiter, e, merr, be, bt, stream()))
  return e[], merr[], be, bt
end

# =====================================================================================================
# Remaining entry points of include/tnb200.h, so that every exported symbol has a Julia binding
# (tests/test_julia_glue.py checks names, argument counts and argument classes against the header).
# =====================================================================================================
# ---- scalar * and /  (src/tensor/cudense.jl:22,502): in place, no allocation
function scale!(x::CuArray{El}, α::Number) where {El}
  a = Ref(El(α))
  check(ccall((:tnb_scale, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(El), length(x), ptr(x), a, stream()))
  return x
end
# ---- dot = scalar(dag(A)*B)  (src/tensor/cudense.jl:25-26): device reduction + one 16-byte readback
function dot(x::CuArray{El}, y::CuArray{El}) where {El}
  r = Ref{ComplexF64}(0)
  check(ccall((:tnb_dot, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(El), length(x), ptr(x), ptr(y), C_NULL, r, stream()))
  return El <: Real ? real(r[]) : r[]
end

# ---- handle housekeeping / introspection
function destroy!()
  HANDLE[] == C_NULL && return
  check(ccall((:tnb_destroy, LIB), Cint, (Ptr{Cvoid},), HANDLE[])); HANDLE[] = C_NULL
end
version() = ccall((:tnb_version, LIB), Cint, ())
workspace_bytes() = ccall((:tnb_workspace_bytes, LIB), Csize_t, (Ptr{Cvoid},), handle())
launch_count() = ccall((:tnb_launch_count, LIB), UInt64, (Ptr{Cvoid},), handle())
workspace_limit() = ccall((:tnb_get_workspace_limit, LIB), Csize_t, (Ptr{Cvoid},), C_NULL)
function plan_cache_stats()               # the `ContractionPlans` analogue (src/ITensorsGPU.jl:54-55)
  v = [Ref{UInt64}(0) for _ in 1:4]
  check(ccall((:tnb_plan_cache_stats, LIB), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt64}, Ref{UInt64}, Ref{UInt64}),
              handle(), v[1], v[2], v[3], v[4]))
  return (entries=v[1][], hits=v[2][], misses=v[3][], autotuned=v[4][])
end
plan_cache_clear!() = check(ccall((:tnb_plan_cache_clear, LIB), Cint, (Ptr{Cvoid},), handle()))
function kernel_family_counts()
  v = zeros(UInt64, 4)
  check(ccall((:tnb_kernel_family_counts, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}), handle(), v))
  return (ldgsts=v[1], smallk=v[2], tma=v[3], tma_split_k=v[4])
end

# ---- host-buffer H_eff*phi: pinned host phi in, H*phi out, the PCIe copies pipelined behind steps 1 and 4
function heff_apply_host!(out_host::Vector{ElT}, L::CuArray{ElT,3}, W1::CuArray{ElT,4}, W2::CuArray{ElT,4}, R::CuArray{ElT,3},
                          phi_host::Vector{ElT}, d::BondDims) where {ElT}
  check(ccall((:tnb_heff_apply_host, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), Ref(d), ptr(L), ptr(W1), ptr(W2), ptr(R), pointer(phi_host), pointer(out_host), stream()))
  return out_host
end

# ---- multi-GPU (one process per GPU): slab form without the fused gather, peer group, sharded DMRG pieces
function heff_apply_shard!(out_slab::CuArray{ElT,4}, Lslab::CuArray{ElT,3}, W1::CuArray{ElT,4}, W2::CuArray{ElT,4},
                           R::CuArray{ElT,3}, phi::CuArray{ElT,4}) where {ElT}
  check(ccall((:tnb_heff_apply_shard, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), bond_dims(phi, W1, W2), size(Lslab, 2), ptr(Lslab), ptr(W1), ptr(W2), ptr(R), ptr(phi),
              ptr(out_slab), stream()))
  return out_slab
end
peer_close(p::Ptr{Cvoid}) = check(ccall((:tnb_peer_close, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle(), p))
peer_free(p::Ptr{Cvoid}) = check(ccall((:tnb_peer_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle(), p))
peer_status() = check(ccall((:tnb_peer_status, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle(), stream()))
comm_finalize() = check(ccall((:tnb_comm_finalize, LIB), Cint, (Ptr{Cvoid},), handle()))
comm_barrier() = check(ccall((:tnb_comm_barrier, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle(), stream()))
"every rank's region [off + rank*nbytes, +nbytes) of bufs[rank] lands in the same region of every peer buffer"
comm_allgather(bufs::Vector{Ptr{Cvoid}}, off::Integer, nbytes::Integer) =
  check(ccall((:tnb_comm_allgather, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Csize_t, Csize_t, Ptr{Cvoid}),
              handle(), bufs, off, nbytes, stream()))
shard_stage_bytes(::Type{El}, chi::Integer, d::Integer, w::Integer, world::Integer) where {El} =
  ccall((:tnb_shard_stage_bytes, LIB), Csize_t, (Cint, Int64, Int32, Int32, Cint), dtype(El), chi, d, w, world)
# environments stored as l' slabs per rank (never moved); stage_peers = peer-mapped staging buffers of every rank
function env_update_left_shard!(Lnew_slab::CuArray{ElT}, Lslab::CuArray{ElT}, A::CuArray{ElT,3}, W::CuArray{ElT,4},
                                stage_peers::Vector{Ptr{Cvoid}}) where {ElT}
  chiL, d, chiR = size(A)
  check(ccall((:tnb_env_update_left_shard, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), chiL, chiR, d, size(W, 1), size(W, 4), ptr(Lslab), ptr(A), ptr(W), stage_peers, ptr(Lnew_slab), stream()))
  return Lnew_slab
end
function env_update_right_shard!(Rnew::CuArray{ElT}, R::CuArray{ElT}, A::CuArray{ElT,3}, W::CuArray{ElT,4},
                                 stage_peers::Vector{Ptr{Cvoid}}) where {ElT}
  chiL, d, chiR = size(A)
  check(ccall((:tnb_env_update_right_shard, LIB), Cint,
              (Ptr{Cvoid}, Cint, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), chiL, chiR, d, size(W, 1), size(W, 4), ptr(R), ptr(A), ptr(W), stage_peers, ptr(Rnew), stream()))
  return Rnew
end
# Lanczos with the matvec sharded over l'; out_a / out_b: the two alternating peer-mapped full-vector buffers
function eigsolve_lanczos_shard!(phi::CuArray{ElT,4}, Lslab, W1, W2, R, out_a::Vector{Ptr{Cvoid}}, out_b::Vector{Ptr{Cvoid}};
                                 krylovdim::Int=3, maxiter::Int=1, tol::Float64=1e-14) where {ElT}
  e, nmv = Ref{Float64}(0), Ref{Cint}(0)
  check(ccall((:tnb_eigsolve_lanczos_shard, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
               Cint, Cint, Float64, Ref{Float64}, Ref{Cint}, Ptr{Cvoid}),
              handle(), dtype(ElT), bond_dims(phi, W1, W2), ptr(Lslab), ptr(W1), ptr(W2), ptr(R), ptr(phi), out_a, out_b,
              krylovdim, maxiter, tol, e, nmv, stream()))
  return e[], Int(nmv[])
end
# host-buffer form of the sharded matvec: this rank uploads 1/world of phi and downloads its own l' slab of H*phi
function heff_apply_shard_host!(out_host::Vector{ElT}, Lslab, W1, W2, R, phi_host::Vector{ElT}, d::BondDims,
                                phi_peers::Vector{Ptr{Cvoid}}, out_peers::Vector{Ptr{Cvoid}}) where {ElT}
  check(ccall((:tnb_heff_apply_shard_host, LIB), Cint,
              (Ptr{Cvoid}, Cint, Ref{BondDims}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
               Ptr{Cvoid}, Ptr{Cvoid}),
              handle(), dtype(ElT), Ref(d), ptr(Lslab), ptr(W1), ptr(W2), ptr(R), pointer(phi_host), phi_peers, out_peers,
              pointer(out_host), stream()))
  return out_host
end

end # module
