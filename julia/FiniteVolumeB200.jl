# FiniteVolumeB200.jl -- thin `ccall` layer that puts libfvb200.so (include/fvb200.h) behind the
# call surface of madsjulia/FiniteVolume.jl for the assemble -> solve hot path.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image or on the GPU
# box.  The file is kept deliberately thin (argument marshalling only) so that what is tested --
# the C ABI, through the Python ctypes harness that makes exactly the same calls -- is what runs.
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

# The `linearsolver(A, rhs, x0)` hook of backwardeulerintegrate (src/transient.jl:136): a closure over a System
# whose storage term D = Ss*volumes has been set.  One call = one device-resident backward-Euler solve
# (fvb_step); `A` is ignored because the System already holds the unshifted matrix and applies 1/dt itself.
function linearsolver(s::System, Ss::Number, volumes::Vector; rtol=sqrt(eps(Float64)), maxiter=100_000)
	vol = convert(Vector{Float64}, volumes)
	check(ccall((:fvb_set_storage, libfvb), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), s.h, Ss, vol))
	return function (dt, b_unscaled::Vector{Float64}, u::Vector{Float64})
		check(ccall((:fvb_vec_upload, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 0, b_unscaled))
		check(ccall((:fvb_vec_upload, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 1, u))
		iters = Ref{Int64}(0); conv = Ref{Cint}(0)
		check(ccall((:fvb_step, libfvb), Cint, (Ptr{Cvoid}, Cint, Cint, Float64, Cint, Cint, Float64, Int64, Ref{Int64}, Ref{Cint}),
			s.h, 0, 1, dt, 2, 0, rtol, maxiter, iters, conv))
		out = similar(u)
		check(ccall((:fvb_vec_download, libfvb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.h, 2, out))
		return out
	end
end

end # module
