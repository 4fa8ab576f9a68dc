# nuPGCMB200Ext — package extension that puts nuPGCM's GPU() architecture on libnupgcm_b200.so.
#
# It replaces ext/nuPGCMCUDAExt.jl (reference ext/nuPGCMCUDAExt.jl:24-33: CuArray / CuSparseMatrixCSR)
# and the Krylov.krylov_solve! call of src/iterative_solvers.jl:58.  nuPGCM's own source is not
# touched: InversionToolkit / EvolutionToolkit / invert! / evolve! / run! run as they are, because
#   * on_architecture(GPU(), ::Vector / ::SparseMatrixCSC) return B200Vector / B200CSR,
#   * B200Vector is an AbstractVector{Float64} with exactly what those functions and the
#     Krylov.jl workspace CONSTRUCTORS use (S(undef, n), similar, fill!, `.=` of linear combinations,
#     x[::Vector{Int}], B*x, Diagonal(x)),
#   * iterative_solve!(::IterativeSolverToolkit{<:B200CSR}) is one ccall per solve.
# Load it instead of CUDA.jl:   ENV["NUPGCM_B200_LIB"] = "/path/to/libnupgcm_b200.so"; using nuPGCM, Libdl
# with, in nuPGCM's Project.toml,   [weakdeps] Libdl = "8f399da3-3557-5675-b5ff-fb832c97cbdb"
#                                   [extensions] nuPGCMB200Ext = "Libdl"
# (next to nuPGCMCUDAExt = "CUDA", Project.toml:22-26; load only one of the two: both define GPU()).
#
# STATUS: written against include/nupgcm_b200.h; Julia is not installed in the build environment of
# this repository, so this file has never been executed.  The identical call sequence is executed by
# the ctypes binding (nupgcm_b200/lib.py) and by the C test tests/capi/test_abi.c.
module nuPGCMB200Ext

using nuPGCM, SparseArrays, LinearAlgebra, Libdl
import Krylov
import nuPGCM: on_architecture, architecture, vector_type, print_memory_status,
               iterative_solve!, IterativeSolverToolkit, GPU, CPU

const LIB = Ref{String}("")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

lasterr() = unsafe_string(ccall((:nupgcm_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[]))
check(rc) = rc == 0 || error("libnupgcm_b200 error $rc: " * lasterr())

# mirror of nupgcm_solve_stats (include/nupgcm_b200.h): 80 bytes, checked against the library at load
struct SolveStats
    niter::Int64
    solved::Int32
    inconsistent::Int32
    breakdown::Int32
    reserved::Int32
    rnorm::Float64
    rnorm0::Float64
    device_ms::Float32
    launches::Int32
    hist_len::Int64
    phase_frac::NTuple{4,Float32}
    sm_mhz::Float32
    reserved2::Float32
end

function __init__()
    LIB[] = get(ENV, "NUPGCM_B200_LIB", joinpath(@__DIR__, "..", "deps", "libnupgcm_b200.so"))
    Libdl.dlopen(LIB[])                                          # fails loudly when the library is missing
    nbytes = ccall((:nupgcm_solve_stats_size, LIB[]), Int64, ())
    nbytes == sizeof(SolveStats) || error("ABI mismatch: nupgcm_solve_stats is $nbytes bytes in the library, " *
                                          "$(sizeof(SolveStats)) in nuPGCMB200Ext")
    ctx = Ref{Ptr{Cvoid}}()
    rc = ccall((:nupgcm_create, LIB[]), Int32, (Int32, Ptr{Ptr{Cvoid}}), parse(Int32, get(ENV, "NUPGCM_B200_DEVICE", "0")), ctx)
    rc == 0 || error("nupgcm_create failed (no CPU fallback): " *
                     unsafe_string(ccall((:nupgcm_last_error, LIB[]), Cstring, (Ptr{Cvoid},), C_NULL)))
    CTX[] = ctx[]
    name = Vector{UInt8}(undef, 64); sm = Ref{Int32}(); maj = Ref{Int32}(); mnr = Ref{Int32}()
    check(ccall((:nupgcm_device_info, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}),
                CTX[], sm, maj, mnr, name))
    @info "B200 device: $(unsafe_string(pointer(name))) ($(sm[]) SMs, sm_$(maj[])$(mnr[]))"    # nuPGCMCUDAExt.jl:8-16
end

# ---- device vector ---------------------------------------------------------------------------------
mutable struct B200Vector <: AbstractVector{Float64}
    h::Ptr{Cvoid}
    n::Int
    uniform::Union{Nothing,Float64}    # all entries equal this value (set at upload; Diagonal((1/h^dim) ones), inversion.jl:54)
    function B200Vector(::UndefInitializer, n::Integer)          # what Krylov workspaces call: S(undef, n); zero-filled
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:nupgcm_vec_create, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), CTX[], n, h))
        v = new(h[], n, nothing)
        finalizer(x -> ccall((:nupgcm_vec_destroy, LIB[]), Int32, (Ptr{Cvoid},), x.h), v)
        return v
    end
end
function B200Vector(a::AbstractVector{<:Real})
    v = B200Vector(undef, length(a))
    host = convert(Vector{Float64}, a)
    check(ccall((:nupgcm_vec_upload, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), v.h, host, length(host)))
    v.uniform = (!isempty(host) && all(==(host[1]), host)) ? host[1] : nothing
    return v
end
Base.size(v::B200Vector) = (v.n,)
Base.IndexStyle(::Type{B200Vector}) = IndexLinear()
Base.similar(v::B200Vector) = B200Vector(undef, v.n)
Base.similar(v::B200Vector, ::Type{Float64}, dims::Dims{1}) = B200Vector(undef, dims[1])
function Base.Array(v::B200Vector)
    a = Vector{Float64}(undef, v.n)
    check(ccall((:nupgcm_vec_download, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), v.h, a, v.n))
    return a
end
Base.Vector(v::B200Vector) = Array(v)
Base.getindex(v::B200Vector, i::Int) = Array(v)[i]               # scalar indexing: show() and debugging only (slow)
function Base.fill!(v::B200Vector, x::Real)
    check(ccall((:nupgcm_vec_fill, LIB[]), Int32, (Ptr{Cvoid}, Float64), v.h, x)); v.uniform = Float64(x); v
end
function Base.copyto!(dst::B200Vector, src::B200Vector)
    check(ccall((:nupgcm_vec_copy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), dst.h, src.h)); dst.uniform = src.uniform; dst
end
function Base.copyto!(dst::B200Vector, src::Vector{Float64})
    check(ccall((:nupgcm_vec_upload, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), dst.h, src, length(src))); dst.uniform = nothing; dst
end
axpby!(α::Real, x::B200Vector, β::Real, y::B200Vector) =        # y = αx + βy
    (check(ccall((:nupgcm_vec_axpby, LIB[]), Int32, (Ptr{Cvoid}, Float64, Ptr{Cvoid}, Float64), y.h, α, x.h, β)); y.uniform = nothing; y)
function LinearAlgebra.dot(x::B200Vector, y::B200Vector)
    r = Ref{Float64}(); check(ccall((:nupgcm_vec_dot, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), x.h, y.h, r)); r[]
end
function LinearAlgebra.norm(x::B200Vector)
    r = Ref{Float64}(); check(ccall((:nupgcm_vec_norm2, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), x.h, r)); r[]
end
function Base.maximum(::typeof(abs), x::B200Vector)              # blow-up check, model.jl:149-150
    m = Ref{Float64}(); nan = Ref{Int32}()
    check(ccall((:nupgcm_vec_maxabs, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int32}), x.h, 0, m, nan))
    nan[] != 0 ? NaN : m[]
end

# x[inv_perm] (model.jl:282,312): gather on the device; the index vectors are uploaded once and cached
const INDEX_CACHE = IdDict{Vector{Int},Ptr{Cvoid}}()
function Base.getindex(v::B200Vector, idx::Vector{Int})
    hidx = get!(INDEX_CACHE, idx) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:nupgcm_index_create, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Ptr{Ptr{Cvoid}}),
                    CTX[], idx, length(idx), 1, h))               # index_base = 1
        h[]
    end
    out = B200Vector(undef, length(idx))
    check(ccall((:nupgcm_vec_gather, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), out.h, v.h, hidx))
    return out
end

# Broadcasting: everything nuPGCM broadcasts over device vectors is a LINEAR COMBINATION —
#   workspace.x .= zero(T)                                         inversion.jl:85, evolution.jl:121
#   solver.y .= B*b .+ b₀                                          inversion.jl:104
#   model.evolution.rhsᵥ .= rhsᵥ                                   model.jl:243-244
#   @. solver.y = rhs_adv + θ*rhs_diff + Δt*rhs_flux - (rhsₘ + θ*(rhsₕ + rhsᵥ))      model.jl:278
# so a Broadcasted tree of +, -, scalar * over B200Vectors and numbers is flattened into
# Σ cᵢ vᵢ + c₀ and evaluated with nupgcm_vec_fill / nupgcm_vec_axpby.  Anything else is refused.
struct B200Style <: Broadcast.AbstractArrayStyle{1} end
B200Style(::Val{1}) = B200Style()
B200Style(::Val{0}) = B200Style()
Base.BroadcastStyle(::Type{B200Vector}) = B200Style()
Base.similar(bc::Broadcast.Broadcasted{B200Style}, ::Type{Float64}) = B200Vector(undef, length(axes(bc)[1]))

function lincomb!(terms::Vector{Tuple{Float64,B200Vector}}, c0::Ref{Float64}, c::Float64, x)
    if x isa B200Vector
        push!(terms, (c, x))
    elseif x isa Number
        c0[] += c * x
    elseif x isa Base.RefValue
        lincomb!(terms, c0, c, x[])
    elseif x isa Broadcast.Broadcasted
        f, a = x.f, x.args
        if f === identity && length(a) == 1
            lincomb!(terms, c0, c, a[1])
        elseif f === (+)
            foreach(t -> lincomb!(terms, c0, c, t), a)
        elseif f === (-) && length(a) == 1
            lincomb!(terms, c0, -c, a[1])
        elseif f === (-) && length(a) == 2
            lincomb!(terms, c0, c, a[1]); lincomb!(terms, c0, -c, a[2])
        elseif f === (*) && length(a) == 2 && a[1] isa Number
            lincomb!(terms, c0, c * a[1], a[2])
        elseif f === (*) && length(a) == 2 && a[2] isa Number
            lincomb!(terms, c0, c * a[2], a[1])
        elseif f === (/) && length(a) == 2 && a[2] isa Number
            lincomb!(terms, c0, c / a[2], a[1])
        else
            error("nuPGCMB200Ext: only linear combinations of device vectors can be broadcast (got $f)")
        end
    else
        error("nuPGCMB200Ext: cannot broadcast over $(typeof(x)) with device vectors")
    end
end
function Base.copyto!(dest::B200Vector, bc::Broadcast.Broadcasted{B200Style})
    terms = Tuple{Float64,B200Vector}[]; c0 = Ref(0.0)
    lincomb!(terms, c0, 1.0, bc)
    self = findfirst(t -> t[2] === dest, terms)                   # dest on the right-hand side: keep its coefficient
    β = self === nothing ? 0.0 : terms[self][1]
    self === nothing || deleteat!(terms, self)
    any(t -> t[2] === dest, terms) && error("nuPGCMB200Ext: destination appears twice in a broadcast")
    if c0[] != 0.0
        β == 0.0 || error("nuPGCMB200Ext: constant term together with an in-place update")
        fill!(dest, c0[]); β = 1.0
    elseif β == 0.0 && isempty(terms)
        return fill!(dest, 0.0)
    end
    first_term = true
    for (c, v) in terms
        axpby!(c, v, (first_term && β == 0.0 && c0[] == 0.0) ? 0.0 : 1.0, dest); first_term = false
    end
    isempty(terms) && β != 1.0 && axpby!(0.0, dest, β, dest)
    return dest
end
Base.copyto!(dest::B200Vector, bc::Broadcast.Broadcasted{<:Broadcast.AbstractArrayStyle{0}}) = fill!(dest, bc.f(bc.args...))

# ---- device CSR matrix -----------------------------------------------------------------------------
mutable struct B200CSR
    h::Ptr{Cvoid}
    m::Int
    n::Int
    function B200CSR(A::SparseMatrixCSC{Float64,Int}; drop_zeros::Bool=false)
        At = SparseMatrixCSC(transpose(A))            # CSC of Aᵀ == CSR of A, 1-based Int64 indices
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:nupgcm_csr_create, LIB[]), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Int32, Ptr{Ptr{Cvoid}}),
                    CTX[], size(A, 1), size(A, 2), nnz(At), At.colptr, At.rowval, At.nzval, 1, drop_zeros, h))
        M = new(h[], size(A)...)
        finalizer(x -> ccall((:nupgcm_csr_destroy, LIB[]), Int32, (Ptr{Cvoid},), x.h), M)
        return M
    end
end
Base.size(A::B200CSR) = (A.m, A.n)
Base.size(A::B200CSR, d::Integer) = d == 1 ? A.m : d == 2 ? A.n : 1
Base.eltype(::B200CSR) = Float64
Base.summary(A::B200CSR) = "$(A.m)×$(A.n) B200CSR"
function LinearAlgebra.mul!(y::B200Vector, A::B200CSR, x::B200Vector, α::Number=1.0, β::Number=0.0)
    check(ccall((:nupgcm_spmv, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64), A.h, x.h, y.h, α, β))
    y.uniform = nothing
    return y
end
Base.:*(A::B200CSR, x::B200Vector) = mul!(B200Vector(undef, A.m), A, x)       # B*b, inversion.jl:104

# ---- the four functions of src/architectures.jl (ext/nuPGCMCUDAExt.jl:24-33) -----------------------
on_architecture(::GPU, a::Vector{<:Real}) = B200Vector(a)
on_architecture(::CPU, a::B200Vector) = Array(a)
on_architecture(::GPU, a::B200Vector) = a
on_architecture(::GPU, a::SparseMatrixCSC) = B200CSR(SparseMatrixCSC{Float64,Int}(a))
on_architecture(::GPU, a::B200CSR) = a
architecture(::B200Vector) = GPU()
architecture(::B200CSR) = GPU()
vector_type(::GPU, ::Type{Float64}) = B200Vector
vector_type(::GPU, T) = error("libnupgcm_b200 is FP64 only (got $T)")
function print_memory_status(::GPU)
    f = Ref{Csize_t}(); t = Ref{Csize_t}()
    check(ccall((:nupgcm_mem_status, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Csize_t}, Ptr{Csize_t}), CTX[], f, t))
    println("GPU memory usage: ", round((t[] - f[]) / 2^30, digits=3), " GiB of ", round(t[] / 2^30, digits=1))
end

# ---- src/iterative_solvers.jl:31-68 — one ccall per solve instead of Krylov.krylov_solve! -----------
# P is what the reference builds: Diagonal(on_architecture(arch, …)) — (1/h^dim) ones for the inversion
# (inversion.jl:54), 1 ./ diag(A) for the evolution (evolution.jl:149,167; model.jl:256).
function preconditioner_args(P)
    P isa Diagonal{Float64,B200Vector} || error("nuPGCMB200Ext: preconditioner must be Diagonal(::B200Vector), got $(typeof(P))")
    d = P.diag
    return d.uniform === nothing ? (d.h, 1.0) : (C_NULL, d.uniform)        # a multiple of I is passed as a scalar
end

function iterative_solve!(tk::IterativeSolverToolkit{<:B200CSR})
    kw, ws = tk.kwargs, tk.workspace
    tk.x === ws.x || error("nuPGCMB200Ext: solver.x must alias workspace.x (iterative_solvers.jl:26-29)")
    dinv, pscale = preconditioner_args(tk.P)
    history = get(kw, :history, false)
    cap = history ? 1 + (get(kw, :itmax, 0) == 0 ? 2 * size(tk.A, 1) : kw[:itmax]) : 0
    hist = Vector{Float64}(undef, max(cap, 1))
    st = Ref{SolveStats}()
    atol, rtol, itmax = Float64(get(kw, :atol, 1e-6)), Float64(get(kw, :rtol, 1e-6)), Int64(get(kw, :itmax, 0))
    if ws isa Krylov.GmresWorkspace                              # inversion.jl:84
        get(kw, :restart, true) || error("nuPGCMB200Ext: restart=false GMRES is not provided by libnupgcm_b200")
        memory = length(ws.V)                                    # GmresWorkspace(N, N, VT; memory)
        orth = parse(Int32, get(ENV, "NUPGCM_B200_ORTH", "0"))   # 0 = modified Gram-Schmidt, what Krylov.jl does
        check(ccall((:nupgcm_gmres_solve, LIB[]), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Int64, Int32, Int32,
                     Ptr{Float64}, Int64, Ptr{SolveStats}),
                    tk.A.h, dinv, pscale, tk.y.h, tk.x.h, atol, rtol, itmax, memory, orth, hist, cap, st))
    elseif ws isa Krylov.CgWorkspace                             # evolution.jl:120
        check(ccall((:nupgcm_cg_solve, LIB[]), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Int64,
                     Ptr{Float64}, Int64, Ptr{SolveStats}),
                    tk.A.h, dinv, pscale, tk.y.h, tk.x.h, atol, rtol, itmax, hist, cap, st))
    else
        error("nuPGCMB200Ext: unsupported Krylov workspace $(typeof(ws))")
    end
    tk.x.uniform = nothing
    s = st[]
    ws.stats.niter = s.niter                                     # what the reference reads (iterative_solvers.jl:61-63)
    ws.stats.solved = s.solved != 0
    ws.stats.inconsistent = s.inconsistent != 0
    ws.stats.timer = s.device_ms / 1e3
    history && (empty!(ws.stats.residuals); append!(ws.stats.residuals, view(hist, 1:s.hist_len)))
    @debug "$(tk.label) iterative solve: solved=$(s.solved != 0), niter=$(s.niter), time=$(s.device_ms / 1e3)"
    return tk                                                    # non-convergence is not an error (:58-67)
end

end # module
