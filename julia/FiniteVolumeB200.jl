# FiniteVolumeB200.jl -- thin `ccall` layer that puts libfvb200.so (include/fvb200.h) behind the
# call surface of madsjulia/FiniteVolume.jl for the assemble -> solve hot path.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image or on the GPU
# box.  The file is kept deliberately thin (argument marshalling only; the integrator, the multi-GPU
# partitioning and every solve live behind the C ABI) so that what is tested -- the C ABI, through the
# Python ctypes harness and the plain-C demos that make exactly the same calls -- is what runs.
#
# Usage (drop-in for the functions of src/FiniteVolume.jl and src/transient.jl named below):
#     include("FiniteVolumeB200.jl"); import .FiniteVolumeB200 as FiniteVolume
#     head, ch, A, b, freenode = FiniteVolume.solvediffusion(neighbors, areasoverlengths,
#                                    conductivities, sources, dirichletnodes, dirichletheads)
module FiniteVolumeB200

import SparseArrays

const libfvb = get(ENV, "FVB200_LIB", joinpath(@__DIR__, "..", "finitevolume.jl_b200", "libfvb200.so"))

struct FVBError <: Exception
	status::Cint
	msg::String
end
Base.showerror(io::IO, e::FVBError) = print(io, e.msg)

function check(status::Cint)
	if status != 0
		msg = unsafe_string(ccall((:fvb_last_error, libfvb), Cstring, ()))
		# status 1 = the inputs the reference rejects with error(...) (src/FiniteVolume.jl:25-27)
		status == 1 ? error(msg) : throw(FVBError(status, msg))
	end
end

mutable struct System
	h::Ptr{Cvoid}
	nodelo::Int
	nodehi::Int
	function System(device::Integer=0)
		ref = Ref{Ptr{Cvoid}}(C_NULL)
		check(ccall((:fvb_create, libfvb), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref))
		s = new(ref[], 1, 0)
		finalizer(s->(ccall((:fvb_destroy, libfvb), Cint, (Ptr{Cvoid},), s.h); nothing), s)
		return s
	end
end

# `metaindex` is an arbitrary callable in the reference (src/FiniteVolume.jl:75); the C ABI takes its table.
metatable(metaindex, F) = metaindex === nothing ? nothing : Int64[metaindex(i) for i = 1:F]

function assemble!(s::System, neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector, metaindex=nothing, logtransformconductivity::Bool=false)
	F = length(neighbors)
	meta = metatable(metaindex, F)
	nb = reinterpret(Int64, neighbors)  # Pair{Int64,Int64} is isbits: 2F interleaved Int64, zero copy
	aol = convert(Vector{Float64}, areasoverlengths)
	cond = convert(Vector{Float64}, conductivities)
	src = convert(Vector{Float64}, sources)
	dh = convert(Vector{Float64}, dirichletheads)
	N = length(src)
	GC.@preserve nb aol cond src dh meta dirichletnodes begin
		check(ccall((:fvb_assemble, libfvb), Cint,
			(Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Cint, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}),
			s.h, N, 1, N, F, nb, aol, cond, length(cond), meta === nothing ? C_NULL : pointer(meta), logtransformconductivity, src, length(dirichletnodes), dirichletnodes, dh))
	end
	s.nodelo, s.nodehi = 1, N
	return s
end

function sizes(s::System)
	v = [Ref{Int64}(0) for i = 1:5]
	check(ccall((:fvb_sizes, libfvb), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), s.h, v...))
	return (nf=v[1][], nnz=v[2][], rowstart=v[3][], nfglobal=v[4][], nhalo=v[5][])
end

# A is symmetric, so the CSR arrays kept on the device ARE the colptr/rowval/nzval of the CSC matrix.
function getA(s::System)
	sz = sizes(s)
	colptr = Vector{Int64}(undef, sz.nf + 1); rowval = Vector{Int64}(undef, sz.nnz); nzval = Vector{Float64}(undef, sz.nnz)
	check(ccall((:fvb_get_csr, libfvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), s.h, colptr, rowval, nzval))
	return SparseArrays.SparseMatrixCSC(sz.nf, sz.nf, colptr, rowval, nzval)
end

function getb(s::System)
	b = Vector{Float64}(undef, sizes(s).nf)
	check(ccall((:fvb_get_b, libfvb), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, b))
	return b
end

function getfreenode(s::System)
	f = Vector{UInt8}(undef, s.nodehi - s.nodelo + 1)
	check(ccall((:fvb_get_freenode, libfvb), Cint, (Ptr{Cvoid}, Ptr{UInt8}), s.h, f))
	return f .!= 0
end

# what callers read from IterativeSolvers.ConvergenceHistory (ch.isconverged, ch.iters, ch.data[:resnorm])
struct ConvergenceHistory
	isconverged::Bool
	iters::Int
	data::Dict{Symbol, Any}
end

function solve!(s::System; maxiter=100_000, rtol=sqrt(eps(Float64)), x0=nothing)
	head = Vector{Float64}(undef, s.nodehi - s.nodelo + 1)
	hist = Vector{Float64}(undef, maxiter)
	iters = Ref{Int64}(0); conv = Ref{Cint}(0)
	check(ccall((:fvb_solve, libfvb), Cint,
		(Ptr{Cvoid}, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Cint}, Ptr{Float64}, Int64),
		s.h, rtol, maxiter, x0 === nothing ? C_NULL : pointer(x0), head, C_NULL, iters, conv, hist, maxiter))
	return head, ConvergenceHistory(conv[] != 0, iters[], Dict{Symbol, Any}(:resnorm=>hist[1:iters[]]))
end

# ---- solver controls (no counterpart in the reference: kernel / format / recurrence choice) -----------------
# fmt: 0 automatic, 1 always CSR, 2 diagonal copy with per-thread loads, 3 diagonal copy through the TMA pipeline
setspmvformat!(s::System, fmt::Integer) = (check(ccall((:fvb_set_spmv_format, libfvb), Cint, (Ptr{Cvoid}, Cint), s.h, fmt)); s)
# mode: 0 automatic (cold-started steady solves on the diagonal copy run CG on D^-1/2 A D^-1/2), 1 never
setpcgscaling!(s::System, mode::Integer) = (check(ccall((:fvb_set_pcg_scaling, libfvb), Cint, (Ptr{Cvoid}, Cint), s.h, mode)); s)
# kind: 0 Jacobi (default), 1 aggregation-multigrid V-cycle (box-structured matrices); zeros = default parameters
function setpreconditioner!(s::System, kind::Integer; nu::Integer=0, omega::Real=0.0, oc::Real=0.0)
	check(ccall((:fvb_set_preconditioner, libfvb), Cint, (Ptr{Cvoid}, Cint, Cint, Float64, Float64), s.h, kind, nu, omega, oc))
	return s
end
# values-only re-assembly on the retained structure (inverse loops, examples/box_model/ex.jl:53-64)
function updatevalues!(s::System, conductivities::Vector, logtransformconductivity::Bool=false)
	k = convert(Vector{Float64}, conductivities)
	check(ccall((:fvb_update_values, libfvb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Float64}),
		s.h, k, length(k), logtransformconductivity, C_NULL, C_NULL))
	return s
end

# ---- the reference's names ---------------------------------------------------------------------------
# src/FiniteVolume.jl:75
function assembleA(neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector, metaindex=nothing, logtransformconductivity::Bool=false)
	return getA(assemble!(System(), neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads, metaindex, logtransformconductivity))
end

# src/FiniteVolume.jl:110
function assembleb(neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector, metaindex=nothing, logtransformconductivity::Bool=false)
	return getb(assemble!(System(), neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads, metaindex, logtransformconductivity))
end

# src/FiniteVolume.jl:157 -- (head, ch, A, b, freenode); Jacobi-PCG replaces RS-AMG-PCG (north_star), so maxiter
# counts Jacobi iterations and defaults higher than the reference's 400.
function solvediffusion(neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector; maxiter=100_000, rtol=sqrt(eps(Float64)))
	s = assemble!(System(), neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads)
	head, ch = solve!(s; maxiter=maxiter, rtol=rtol)
	return head, ch, getA(s), getb(s), getfreenode(s)
end

# src/FiniteVolume.jl:32-44 -> (freenode::Vector{Bool}, nodei2freenodei::Vector{Int}); an assembly without faces
# builds exactly the node maps (the device scan that replaces the serial counter of :36-42).
function getfreenodes(n::Integer, dirichletnodes::Array{Int, 1})
	s = assemble!(System(), Pair{Int, Int}[], Float64[], Float64[], zeros(n), dirichletnodes, zeros(length(dirichletnodes)))
	map = Vector{Int64}(undef, n)
	check(ccall((:fvb_get_nodei2freenodei, libfvb), Cint, (Ptr{Cvoid}, Ptr{Int64}), s.h, map))
	return getfreenode(s), map
end

# src/FiniteVolume.jl:141-155 -> (head, freenode, nodei2freenodei)
function freenodes2nodes(result::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector)
	s = assemble!(System(), Pair{Int, Int}[], Float64[], Float64[], sources, dirichletnodes, dirichletheads)
	res = convert(Vector{Float64}, result)
	check(ccall((:fvb_vec_upload, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 0, res))
	head = Vector{Float64}(undef, length(sources))
	check(ccall((:fvb_vec_to_nodes, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 0, head))
	map = Vector{Int64}(undef, length(sources))
	check(ccall((:fvb_get_nodei2freenodei, libfvb), Cint, (Ptr{Cvoid}, Ptr{Int64}), s.h, map))
	return head, getfreenode(s), map
end

setstorage!(s::System, Ss::Number, volumes::Vector) = (check(ccall((:fvb_set_storage, libfvb), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), s.h, Ss, convert(Vector{Float64}, volumes))); s)

# ---- the `linearsolver(A, rhs, x0)` hook of backwardeulerintegrate (src/transient.jl:136, default :50-58) ----------
# The reference calls the hook with the row-scaled, diagonally shifted matrix At = D^-1 A + I/dt (:72) and
# rhs~ = D^-1 b + u_k/dt (:71).  The System holds the unscaled A and D = Ss*volumes (setstorage!), so the closure
#   * reads 1/dt off At's diagonal: 1/dt = At[i,i] - A[i,i]/D[i], taken at the row where A[i,i]/D[i] is smallest
#     (least cancellation; diag(A)./D is fetched once when the closure is made),
#   * solves the equivalent SPD system (A + D/dt) x = D .* rhs~ on the device, warm-started from x0 (fvb_solve_shifted).
# It can be passed as `linearsolver=` to the reference's own backwardeulerintegrate(u0, A, getb, dt0, t0, tfinal; ...).
function linearsolver(s::System, Ss::Number, volumes::Vector; rtol=sqrt(eps(Float64)), maxiter=100_000)
	setstorage!(s, Ss, volumes)
	nf = sizes(s).nf
	d = Vector{Float64}(undef, nf)
	check(ccall((:fvb_get_diag, libfvb), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, d))
	D = Ss .* convert(Vector{Float64}, volumes)[getfreenode(s)]
	scaleddiag = d ./ D
	imin = argmin(scaleddiag)
	return function (At, rhs::Vector{Float64}, x0::Vector{Float64})
		sigma = At[imin, imin] - scaleddiag[imin]
		sigma > 0 || error("time step must be positive")
		check(ccall((:fvb_vec_upload, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 0, D .* rhs))
		check(ccall((:fvb_vec_upload, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 1, x0))
		iters = Ref{Int64}(0); conv = Ref{Cint}(0)
		check(ccall((:fvb_solve_shifted, libfvb), Cint, (Ptr{Cvoid}, Cint, Cint, Float64, Cint, Float64, Int64, Ref{Int64}, Ref{Cint}),
			s.h, 0, 1, sigma, 2, rtol, maxiter, iters, conv))
		out = similar(x0)
		check(ccall((:fvb_vec_download, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 2, out))
		return out
	end
end

# ---- model-level transient entries (src/transient.jl:156-174, :188-205): the whole integrator behind one ccall -------
# fvb_integrate runs the reference's controller (backwardeulertwostep!, adaptivebackwardeulerstep! / fixedbackwardeulerstep!,
# the outer loop, :78-154) inside the library with device-resident states.
struct IntegrateOptions
	atol::Float64
	dt0::Float64
	fixed_step::Cint
	adjoint::Cint
	rtol::Float64
	maxiter::Int64
	getb::Ptr{Cvoid}
	getb_ctx::Ptr{Cvoid}
	callback::Ptr{Cvoid}
	callback_ctx::Ptr{Cvoid}
end

# C trampolines: ctx points at a Ref holding the Julia closure
function _getb_trampoline(t::Float64, out::Ptr{Float64}, ctx::Ptr{Cvoid})::Cvoid
	f, n = unsafe_pointer_to_objref(ctx)[]
	b = f(t)
	unsafe_copyto!(out, pointer(convert(Vector{Float64}, b)), n)
	return nothing
end
function _callback_trampoline(t::Float64, dt::Float64, ctx::Ptr{Cvoid})::Cvoid
	unsafe_pointer_to_objref(ctx)[](t, dt)
	return nothing
end

function integrate!(s::System, u0free::Vector{Float64}, t0, tfinal; dt0=1.0, atol=1e-4, fixed=false, adjoint=false, getb=nothing, callback=nothing, rtol=sqrt(eps(Float64)), maxiter=100_000, heads=true, maxstates=4096)
	nf = sizes(s).nf
	width = heads ? (s.nodehi - s.nodelo + 1) : nf
	getbref = Ref{Any}((getb, nf)); cbref = Ref{Any}(callback)
	while true
		ts = Vector{Float64}(undef, maxstates)
		out = Matrix{Float64}(undef, width, maxstates)   # column k = state k (column-major = the C layout [state][row])
		nstates = Ref{Int64}(0); nsolves = Ref{Int64}(0); nits = Ref{Int64}(0); natt = Ref{Int64}(0)
		status = GC.@preserve getbref cbref begin
			opt = Ref(IntegrateOptions(atol, dt0, fixed, adjoint, rtol, maxiter,
				getb === nothing ? C_NULL : @cfunction(_getb_trampoline, Cvoid, (Float64, Ptr{Float64}, Ptr{Cvoid})),
				getb === nothing ? C_NULL : pointer_from_objref(getbref),
				callback === nothing ? C_NULL : @cfunction(_callback_trampoline, Cvoid, (Float64, Float64, Ptr{Cvoid})),
				callback === nothing ? C_NULL : pointer_from_objref(cbref)))
			ccall((:fvb_integrate, libfvb), Cint,
				(Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Ref{IntegrateOptions}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
				s.h, u0free, t0, tfinal, opt, maxstates, ts, heads ? C_NULL : pointer(out), heads ? pointer(out) : C_NULL, nstates, nsolves, nits, natt)
		end
		if status == 5 && callback === nothing && occursin("max_states", unsafe_string(ccall((:fvb_last_error, libfvb), Cstring, ())))
			maxstates *= 4   # more accepted steps than room: run again with larger buffers
			continue
		end
		check(status)
		k = nstates[]
		return [out[:, i] for i = 1:k], ts[1:k]
	end
end

# src/transient.jl:156-163 (constant b) and :165-174 (caller's getb).  NOTE on getb: the reference's getb(t) returns
# the row-scaled D^-1 b(t); it is converted back to the unscaled right-hand side the library integrates with.
function backwardeulerintegrate(u0, tspan, Ss::Number, volumes::Vector, neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector, metaindex=nothing, logtransformconductivity=false; dt0=1.0, atol=1e-4, callback=nothing, fixed=false, kwargs...)
	s = assemble!(System(), neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads, metaindex, logtransformconductivity)
	setstorage!(s, Ss, volumes)
	freenode = getfreenode(s)
	return integrate!(s, convert(Vector{Float64}, u0)[freenode], tspan[1], tspan[2]; dt0=dt0, atol=atol, callback=callback, fixed=fixed, heads=true, kwargs...)
end
function backwardeulerintegrate(u0, tspan, getb::Function, Ss::Number, volumes::Vector, neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector, metaindex=nothing, logtransformconductivity=false; dt0=1.0, atol=1e-4, callback=nothing, fixed=false, kwargs...)
	s = assemble!(System(), neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads, metaindex, logtransformconductivity)
	setstorage!(s, Ss, volumes)
	freenode = getfreenode(s)
	D = Ss .* convert(Vector{Float64}, volumes)[freenode]
	return integrate!(s, convert(Vector{Float64}, u0)[freenode], tspan[1], tspan[2]; dt0=dt0, atol=atol, callback=callback, fixed=fixed, getb=t->D .* getb(t), heads=true, kwargs...)
end

# src/transient.jl:188-205 -> (lambdas, ts): gamma' = (D^-1 A)^T gamma + dg/du(T - t), gamma(0) = 0, reversed in time
function adjointintegrate(getdgdu::Function, tspan, Ss::Number, volumes::Vector, neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector, metaindex=nothing, logtransformconductivity=false; dt0=1.0, atol=1e-4, callback=nothing, fixed=false, kwargs...)
	s = assemble!(System(), neighbors, areasoverlengths, conductivities, sources, dirichletnodes, dirichletheads, metaindex, logtransformconductivity)
	setstorage!(s, Ss, volumes)
	nf = sizes(s).nf
	gammas, ts = integrate!(s, zeros(nf), tspan[1], tspan[2]; dt0=dt0, atol=atol, callback=callback, fixed=fixed, adjoint=true, getb=t->getdgdu(tspan[2] - t), heads=false, kwargs...)
	return reverse(gammas), reverse(map(t->tspan[2] - t, ts))
end

# ---- one Julia process, several GPUs (fvb_multi_*: the library partitions, connects and drives the devices) ---------
mutable struct MultiSystem
	m::Ptr{Cvoid}
	n::Int
	function MultiSystem(devices::Vector{<:Integer})
		ref = Ref{Ptr{Cvoid}}(C_NULL)
		ids = convert(Vector{Cint}, devices)
		check(ccall((:fvb_multi_create, libfvb), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(ids), ids, ref))
		ms = new(ref[], 0)
		finalizer(x->(ccall((:fvb_multi_destroy, libfvb), Cint, (Ptr{Cvoid},), x.m); nothing), ms)
		return ms
	end
end

# solvediffusion on `devices` (same arguments, same 5-tuple): src/FiniteVolume.jl:157
function solvediffusion(devices::Vector{<:Integer}, neighbors::Array{Pair{Int, Int}, 1}, areasoverlengths::Vector, conductivities::Vector, sources::Vector, dirichletnodes::Array{Int, 1}, dirichletheads::Vector; maxiter=100_000, rtol=sqrt(eps(Float64)), logtransformconductivity::Bool=false)
	ms = MultiSystem(devices)
	nb = reinterpret(Int64, neighbors)
	aol = convert(Vector{Float64}, areasoverlengths); cond = convert(Vector{Float64}, conductivities)
	src = convert(Vector{Float64}, sources); dh = convert(Vector{Float64}, dirichletheads)
	N, F = length(src), length(neighbors)
	GC.@preserve nb aol cond src dh dirichletnodes begin
		check(ccall((:fvb_multi_assemble, libfvb), Cint,
			(Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Cint, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}),
			ms.m, N, F, nb, aol, cond, length(cond), C_NULL, logtransformconductivity, src, length(dirichletnodes), dirichletnodes, dh))
	end
	nf = Ref{Int64}(0); nnz = Ref{Int64}(0); nd = Ref{Cint}(0)
	check(ccall((:fvb_multi_sizes, libfvb), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Cint}, Ptr{Int64}, Ptr{Int64}), ms.m, nf, nnz, nd, C_NULL, C_NULL))
	head = Vector{Float64}(undef, N); hist = Vector{Float64}(undef, maxiter)
	iters = Ref{Int64}(0); conv = Ref{Cint}(0)
	check(ccall((:fvb_multi_solve, libfvb), Cint, (Ptr{Cvoid}, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Cint}, Ptr{Float64}, Int64),
		ms.m, rtol, maxiter, head, C_NULL, iters, conv, hist, maxiter))
	colptr = Vector{Int64}(undef, nf[] + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{Float64}(undef, nnz[])
	check(ccall((:fvb_multi_get_csr, libfvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), ms.m, colptr, rowval, nzval))
	b = Vector{Float64}(undef, nf[])
	check(ccall((:fvb_multi_get_b, libfvb), Cint, (Ptr{Cvoid}, Ptr{Float64}), ms.m, b))
	fnode = Vector{UInt8}(undef, N)
	check(ccall((:fvb_multi_get_freenode, libfvb), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ms.m, fnode))
	ch = ConvergenceHistory(conv[] != 0, iters[], Dict{Symbol, Any}(:resnorm=>hist[1:iters[]]))
	return head, ch, SparseArrays.SparseMatrixCSC(nf[], nf[], colptr, rowval, nzval), b, fnode .!= 0
end

end # module
