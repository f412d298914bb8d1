# BandedMatricesB200.jl -- Julia glue for libbmb200 (include/bmb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia.  It is the literal binding a maintainer adds next to
# BandedMatrices.jl; every method shadows one call site of the reference (cited) with a `ccall` into the C ABI.
# The same ABI is exercised end to end by the Python host mirror (bandedmatrices.jl_b200/) and tests/.
module BandedMatricesB200

using LinearAlgebra, BandedMatrices, ArrayLayouts
import LinearAlgebra: BlasInt, LAPACK
import BandedMatrices: banded_gbmv!, _gbmm!, bandeddata, bandwidth, BandedMatrix, _BandedMatrix

const libbmb200 = get(ENV, "LIBBMB200", "libbmb200.so")
const Handle = Ptr{Cvoid}
const HANDLE = Ref{Handle}(C_NULL)

function handle()
    if HANDLE[] == C_NULL
        rc = ccall((:bmb200_create, libbmb200), Cint, (Ref{Handle}, Cint, Ptr{Cvoid}), HANDLE, 0, C_NULL)
        rc == 0 || error("bmb200_create failed ($rc): no sm_100a GPU; there is no CPU fallback")
    end
    HANDLE[]
end
chk(rc, what) = rc == 0 ? nothing :
    rc < 0 && rc > -100 ? throw(ArgumentError("invalid argument #$(-rc) to $what")) :
    error("$what failed: " * unsafe_string(ccall((:bmb200_last_error, libbmb200), Cstring, (Handle,), handle())))

# ---- device array: the container type that routes BandedMatrix to the BLAS layouts (BandedMatrix.jl:37-41) ----
mutable struct B200Array{T,N} <: DenseArray{T,N}
    ptr::Ptr{T}
    dims::NTuple{N,Int}
    function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        p = Ref{Ptr{Cvoid}}()
        chk(ccall((:bmb200_malloc, libbmb200), Cint, (Handle, Ref{Ptr{Cvoid}}, Csize_t), handle(), p, sizeof(T) * prod(dims)), "malloc")
        a = new{T,N}(Ptr{T}(p[]), dims)
        finalizer(x -> ccall((:bmb200_free, libbmb200), Cint, (Handle, Ptr{Cvoid}), handle(), x.ptr), a)
    end
end
Base.size(a::B200Array) = a.dims
Base.strides(a::B200Array{T,2}) where {T} = (1, a.dims[1])
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where {T} = a.ptr
Base.similar(a::B200Array{T}, ::Type{T}, dims::Dims{N}) where {T,N} = B200Array{T,N}(undef, dims)
ArrayLayouts.MemoryLayout(::Type{<:B200Array}) = DenseColumnMajor()
function B200Array(h::Array{T,N}) where {T,N}
    d = B200Array{T,N}(undef, size(h))
    chk(ccall((:bmb200_memcpy_h2d, libbmb200), Cint, (Handle, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), handle(), d.ptr, h, sizeof(h)), "h2d")
    d
end
function Base.Array(d::B200Array{T,N}) where {T,N}
    h = Array{T,N}(undef, d.dims)
    chk(ccall((:bmb200_memcpy_d2h, libbmb200), Cint, (Handle, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), handle(), h, d.ptr, sizeof(h)), "d2h")
    h
end
const DVec = B200Array{Float64,1}
const DMat = B200Array{Float64,2}
const DBanded = BandedMatrix{Float64,DMat}

# ---- y <- alpha*op(A)*x + beta*y : shadows banded_gbmv! (src/generic/matmul.jl:21-23) ----
function banded_gbmv!(tA, α, A::DBanded, x::DVec, β, y::DVec)
    D = bandeddata(A)
    chk(ccall((:bmb200_dgbmv, libbmb200), Cint,
              (Handle, UInt8, Int64, Int64, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64),
              handle(), tA, size(A, 1), size(D, 2), bandwidth(A, 1), bandwidth(A, 2), α, D, stride(D, 2), x, stride(x, 1), β, y, stride(y, 1)), "dgbmv")
    y
end

# ---- C <- alpha*A*B + beta*C, banded x banded : shadows _gbmm! (src/banded/gbmm.jl:296-340) ----
function _gbmm!(α::Float64, A_data::DMat, B_data::DMat, β, C_data::Union{DMat,SubArray{Float64,2,DMat}}, (n, ν, m), (Al, Au), (Bl, Bu), (Cl, Cu), Czero)
    chk(ccall((:bmb200_dgbmm_bb, libbmb200), Cint,
              (Handle, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64),
              handle(), n, ν, m, Al, Au, Bl, Bu, Cl, Cu, α, A_data, stride(A_data, 2), B_data, stride(B_data, 2), β, pointer(C_data), stride(C_data, 2)), "dgbmm_bb")
    C_data
end

# ---- banded x dense : shadows materialize!(MatMulMatAdd{<:BandedColumns,...}) (src/generic/matmul.jl:243-256) ----
function ArrayLayouts.materialize!(M::ArrayLayouts.MatMulMatAdd{<:BandedMatrices.BandedColumns,<:Any,<:Any,Float64,<:DBanded,<:DMat,<:DMat})
    A, B, C = M.A, M.B, M.C
    D = bandeddata(A)
    chk(ccall((:bmb200_dgbmm_bd, libbmb200), Cint,
              (Handle, UInt8, Int64, Int64, Int64, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64),
              handle(), 'N', size(A, 1), size(A, 2), bandwidth(A, 1), bandwidth(A, 2), size(B, 2), M.α, D, stride(D, 2), B, stride(B, 2), M.β, C, stride(C, 2)), "dgbmm_bd")
    C
end

# ---- _fill_lmul! / zero! on device blocks (src/generic/utils.jl:29-31) ----
function LinearAlgebra.lmul!(β::Number, C::DMat)
    chk(ccall((:bmb200_dfill_lmul, libbmb200), Cint, (Handle, Float64, Ptr{Float64}, Int64, Int64, Int64, Int64),
              handle(), β, C, size(C, 1), size(C, 2), stride(C, 2), 1), "dfill_lmul")
    C
end
ArrayLayouts.zero!(C::DMat) = lmul!(0.0, C)

# ---- lu: widening copy (src/banded/BandedLU.jl:108-111) + gbtrf! (BandedLU.jl:98) ----
function BandedMatrix{Float64}(A::DBanded, (l, u2)::NTuple{2,Integer})   # only the (l, l+u) widening used by _lu
    l0, u0 = bandwidths(A)
    @assert l == l0 && u2 == l0 + u0
    W = _BandedMatrix(DMat(undef, (2l0 + u0 + 1, size(A, 2))), size(A, 1), l0, l0 + u0)
    chk(ccall((:bmb200_dband_widen, libbmb200), Cint, (Handle, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
              handle(), size(A, 2), l0, u0, bandeddata(A), stride(bandeddata(A), 2), bandeddata(W), stride(bandeddata(W), 2)), "dband_widen")
    W
end

const DEVICE_IPIV = IdDict{Any,B200Array{Int64,1}}()   # device mirror of BandedLU.ipiv, keyed by the host vector

function LAPACK.gbtrf!(kl::Integer, ku::Integer, m::Integer, AB::DMat)
    n = size(AB, 2)
    dip = B200Array{Int64,1}(undef, (min(m, n),))
    info = Ref{Cint}(0)
    chk(ccall((:bmb200_dgbtrf, libbmb200), Cint, (Handle, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Ref{Cint}),
              handle(), m, n, kl, ku, AB, stride(AB, 2), dip, info), "dgbtrf")
    LAPACK.chklapackerror(BlasInt(info[]))           # info > 0 -> LAPACKException, as the stdlib wrapper does
    ipiv = Array(dip)                                 # BandedLU.ipiv is a host Vector{Int64} (BandedLU.jl:12)
    DEVICE_IPIV[ipiv] = dip
    AB, ipiv
end

# ---- lu(A): widening copy + factorisation in one call (src/banded/BandedLU.jl:106-111) ----
function BandedMatrices._lu(::BandedMatrices.BandedColumns, axes, A::DBanded, pivot = Val(true); check::Bool = true)
    m, n = size(A);  l, u = bandwidths(A)
    W = BandedMatrix{Float64}(undef, (m, n), (l, l + u))          # B200Array container (similar(A) in the real glue)
    AB = bandeddata(W);  D = bandeddata(A)
    dip = B200Array{Int64,1}(undef, (min(m, n),));  info = Ref{Cint}(0)
    chk(ccall((:bmb200_dgbtrf_from, libbmb200), Cint,
              (Handle, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Ref{Cint}),
              handle(), m, n, l, u, D, stride(D, 2), AB, stride(AB, 2), dip, info), "dgbtrf_from")
    check && LAPACK.chklapackerror(BlasInt(info[]))
    BandedMatrices.BandedLU{Float64,typeof(W)}(W, Array(dip), BlasInt(info[]))
end

# ---- dense x banded (src/generic/matmul.jl:258-271): one launch instead of one strided gbmv per row of C ----
function ArrayLayouts.materialize!(M::ArrayLayouts.MatMulMatAdd{<:Any,<:BandedMatrices.BandedColumns,<:Any,Float64,<:DMat,<:DBanded,<:DMat})
    α, A, B, β, C = M.α, M.A, M.B, M.β, M.C
    P = bandeddata(B)
    chk(ccall((:bmb200_dgbmm_db, libbmb200), Cint,
              (Handle, UInt8, Int64, Int64, Int64, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64),
              handle(), 'N', size(A, 1), size(B, 1), size(B, 2), bandwidth(B, 1), bandwidth(B, 2), α, A, stride(A, 2), P, stride(P, 2),
              β, C, stride(C, 2)), "dgbmm_db")
    C
end

function LAPACK.gbtrs!(trans::AbstractChar, kl::Integer, ku::Integer, m::Integer, AB::DMat, ipiv::Vector{BlasInt}, B::Union{DVec,DMat})
    dip = get!(() -> B200Array(ipiv), DEVICE_IPIV, ipiv)
    chk(ccall((:bmb200_dgbtrs, libbmb200), Cint, (Handle, UInt8, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64),
              handle(), trans, size(AB, 2), kl, ku, size(B, 2), AB, stride(AB, 2), dip, B, max(1, stride(B, 2))), "dgbtrs")
    B
end


# ---- triangular band solve / multiply: shadow tbsv! / tbmv! (src/blas.jl:121-141, :83-101), reached from ldiv! / lmul! of
# ---- UpperTriangular / LowerTriangular{<:BandedMatrix} (src/tribanded.jl:47-84); `A` is bandeddata of the triangular view,
# ---- i.e. a row-range view of the parent's device data array (stride(A,2) = l+u+1).  trans 'N' and 'T'/'C'.
const DBandData = Union{DMat,SubArray{Float64,2,<:DMat}}
for (jl, sym) in ((:tbsv!, :bmb200_dtbsv), (:tbmv!, :bmb200_dtbmv))
    @eval function BandedMatrices.$jl(uplo::AbstractChar, trans::AbstractChar, diag::AbstractChar, m::Int, k::Int, A::DBandData, x::DVec)
        n = size(A, 2)
        size(A, 1) ≥ k + 1 || throw(ArgumentError("triangular banded data missing"))
        n == m || throw(DimensionMismatch("matrix is not square: dimensions are $n, $m"))
        n == length(x) || throw(DimensionMismatch("size of A is $n != length(x) = $(length(x))"))
        n == 0 && return x
        chk(ccall(($(QuoteNode(sym)), libbmb200), Cint, (Handle, UInt8, UInt8, UInt8, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
                  handle(), uplo, trans, diag, m, k, pointer(A), max(1, stride(A, 2)), pointer(x), stride(x, 1)), $(string(jl)))
        x
    end
end

# ---- symmetric band matvec: shadows banded_sbmv! (src/symbanded/symbanded.jl:72-73) for device-resident data ----
function BandedMatrices.banded_sbmv!(uplo, α::Float64, A::Symmetric{Float64,<:DBanded}, x::DVec, β::Float64, y::DVec)
    D = BandedMatrices.symbandeddata(A)   # row-range view of the parent's device data array
    chk(ccall((:bmb200_dsbmv, libbmb200), Cint, (Handle, UInt8, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64),
              handle(), uplo, size(D, 2), BandedMatrices.bandwidth(A), α, pointer(D), max(1, stride(D, 2)), pointer(x), stride(x, 1), β, pointer(y), stride(y, 1)), "dsbmv")
    y
end

# ---- band-aligned elementwise operations between different bandwidths: shadow banded_axpy! (src/banded/BandedMatrix.jl:1006-1015)
# ---- and the identity broadcast copyto! (src/generic/broadcast.jl:175-230) for device-resident data ----
function BandedMatrices.banded_axpy!(a::Number, X::DBanded, Y::DBanded)
    size(X) == size(Y) || throw(DimensionMismatch("X has size $(size(X)) but Y has size $(size(Y))"))
    (xl, xu), (yl, yu) = bandwidths(X), bandwidths(Y)
    out = Ref{Int64}(0)
    chk(ccall((:bmb200_dband_axpy, libbmb200), Cint, (Handle, Int64, Int64, Float64, Int64, Int64, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ref{Int64}),
              handle(), size(X, 1), size(X, 2), Float64(a), xl, xu, X.data, stride(X.data, 2), yl, yu, Y.data, stride(Y.data, 2), out), "dband_axpy")
    out[] == 0 || throw(BandError(Y, xl > yl ? xl : -xu))
    Y
end
function BandedMatrices._banded_broadcast!(dest::DBanded, ::typeof(identity), src::DBanded, ::BandedMatrices.BandedColumns, ::BandedMatrices.BandedColumns)
    size(dest) == size(src) || throw(DimensionMismatch())
    (sl, su), (dl, du) = bandwidths(src), bandwidths(dest)
    out = Ref{Int64}(0)
    chk(ccall((:bmb200_dband_copy, libbmb200), Cint, (Handle, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ref{Int64}),
              handle(), size(src, 1), size(src, 2), sl, su, src.data, stride(src.data, 2), dl, du, dest.data, stride(dest.data, 2), out), "dband_copy")
    out[] == 0 || throw(BandError(dest, sl > dl ? sl : -su))
    dest
end

# ---- banded Cholesky: shadows pbtrf! / pbtrs! (src/lapack.jl:268-332), reached from banded_chol! (src/symbanded/BandedCholesky.jl:2-13)
# ---- and ldiv!(::Cholesky{T,<:BandedMatrix}, B) (BandedCholesky.jl:72-80), for device-resident band data ----
function BandedMatrices.pbtrf!(uplo::Char, m::Int, kd::Int, A::DBandData)
    LinearAlgebra.chkuplo(uplo)
    n = size(A, 2)
    n ≠ m && throw(ArgumentError("Matrix must be square"))
    size(A, 1) < kd + 1 && throw(ArgumentError("Not enough bands"))
    info = Ref{Cint}(0)
    chk(ccall((:bmb200_dpbtrf, libbmb200), Cint, (Handle, UInt8, Int64, Int64, Ptr{Float64}, Int64, Ref{Cint}),
              handle(), uplo, n, kd, pointer(A), max(1, stride(A, 2)), info), "dpbtrf")
    A, BlasInt(info[])     # info > 0 -> PosDefException in cholesky! (checkpositivedefinite), exactly as today
end
function BandedMatrices.pbtrs!(uplo::Char, m::Int, kd::Int, A::DBandData, B::Union{DVec,DMat})
    LinearAlgebra.chkuplo(uplo)
    n = size(A, 2)
    (m != n || m != size(B, 1)) && throw(DimensionMismatch("matrix A has dimensions $(size(A)), but right hand side matrix B has dimensions $(size(B))"))
    size(A, 1) < kd + 1 && throw(ArgumentError("Not enough bands"))
    chk(ccall((:bmb200_dpbtrs, libbmb200), Cint, (Handle, UInt8, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
              handle(), uplo, n, kd, size(B, 2), pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2))), "dpbtrs")
    B
end

# ---- the other three BLAS element types (src/blas.jl:4-7; LAPACK.gbtrf! / gbtrs! incl. the conjugate-transpose solve, linalg.jl:57-63):
# ---- same argument lists, typed device pointers, alpha / beta by reference as in the Fortran interface ----
for (p, T) in ((:s, Float32), (:c, ComplexF32), (:z, ComplexF64))
    gbmv, gbtrf, gbtrs = Symbol(:bmb200_, p, :gbmv), Symbol(:bmb200_, p, :gbtrf), Symbol(:bmb200_, p, :gbtrs)
    hbmv = p === :s ? :bmb200_ssbmv : Symbol(:bmb200_, p, :hbmv)
    @eval begin
        function BandedMatrices.gbmv!(trans::Char, m::Int, kl::Int, ku::Int, α::$T, A::B200Array{$T,2}, x::B200Array{$T,1}, β::$T, y::B200Array{$T,1})
            chk(ccall(($(QuoteNode(gbmv)), libbmb200), Cint,
                      (Handle, UInt8, Int64, Int64, Int64, Int64, Ref{$T}, Ptr{$T}, Int64, Ptr{$T}, Int64, Ref{$T}, Ptr{$T}, Int64),
                      handle(), trans, m, size(A, 2), kl, ku, α, A, max(1, stride(A, 2)), x, stride(x, 1), β, y, stride(y, 1)), $(string(gbmv)))
            y
        end
        function BandedMatrices.$(p === :s ? :sbmv! : :hbmv!)(uplo::Char, k::Int, α::$T, A::B200Array{$T,2}, x::B200Array{$T,1}, β::$T, y::B200Array{$T,1})
            chk(ccall(($(QuoteNode(hbmv)), libbmb200), Cint,
                      (Handle, UInt8, Int64, Int64, Ref{$T}, Ptr{$T}, Int64, Ptr{$T}, Int64, Ref{$T}, Ptr{$T}, Int64),
                      handle(), uplo, size(A, 2), k, α, A, max(1, stride(A, 2)), x, 1, β, y, 1), $(string(hbmv)))
            y
        end
        function LAPACK.gbtrf!(kl::Integer, ku::Integer, m::Integer, AB::B200Array{$T,2})
            dip = B200Array{Int64,1}(undef, (min(m, size(AB, 2)),)); info = Ref{Cint}(0)
            chk(ccall(($(QuoteNode(gbtrf)), libbmb200), Cint, (Handle, Int64, Int64, Int64, Int64, Ptr{$T}, Int64, Ptr{Int64}, Ref{Cint}),
                      handle(), m, size(AB, 2), kl, ku, AB, stride(AB, 2), dip, info), $(string(gbtrf)))
            LAPACK.chklapackerror(BlasInt(info[]))
            AB, Array(dip)
        end
        function LAPACK.gbtrs!(trans::AbstractChar, kl::Integer, ku::Integer, m::Integer, AB::B200Array{$T,2}, ipiv::Vector{BlasInt}, B::Union{B200Array{$T,1},B200Array{$T,2}})
            dip = B200Array(ipiv)
            chk(ccall(($(QuoteNode(gbtrs)), libbmb200), Cint, (Handle, UInt8, Int64, Int64, Int64, Int64, Ptr{$T}, Int64, Ptr{Int64}, Ptr{$T}, Int64),
                      handle(), trans, size(AB, 2), kl, ku, size(B, 2), AB, stride(AB, 2), dip, B, max(1, stride(B, 2))), $(string(gbtrs)))
            B
        end
    end
end

end # module
