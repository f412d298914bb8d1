"""bandedmatrices.jl_b200 -- B200 (sm_100a) hot path of BandedMatrices.jl.

Only the path named by BASELINE.json's north_star lives here: band-storage matvec / matmul
(``gbmv!``, ``gbmm!``) and banded LU factor / solve (``gbtrf!`` / ``gbtrs!``) behind the
reference's ``mul!`` / ``*`` / ``lu`` / ``ldiv!`` / ``\\`` API, on a device-resident
``BandedMatrix``.  ``csrc/`` holds the hand-written CUDA kernels and the C ABI
(``include/bmb200.h`` -> ``libbmb200.so``); the modules here are the host-side mirror of the
reference's driver logic over that ABI (Julia, the reference's host language, is not available
in this image -- the Julia glue a maintainer would add is in ``julia/`` and INTEGRATION.md).

The directory name contains a dot, so import it as ``import bandedmatrices_b200`` (a shim at the
repository root loads this directory under that name).
"""
from ._lib import BMB200Error, Handle, LIB_PATH, PROTOTYPES, handle, load
from .banded import (BandedMatrix, BandError, DimensionMismatch, LAPACKException, Transposed, bandeddata, bandwidth,
                     bandwidths, brand, colmajor, to_colmajor)
from .linalg import (gbmv_, hbmv_, gbtrf_, gbtrs_, gbmm_typed_, gbmm_bd_typed_, BandedCholesky, PosDefException, cholesky, cholesky_, ldiv_chol_, pbtrf_, pbtrs_, BandedLU, TransposeFact, axpby_, axpy_, badd, bscale, bsub, copyto_, similar, factorize, gbmm_, gbmv_host, ldiv_, ldiv_tri_, lmul_tri_, lu, lu_, matmul, mul_,
                     mul_sym_, sbmv_, solve, tbmv_, tbsv_)

__all__ = [
    "BandedMatrix", "Transposed", "BandedLU", "TransposeFact", "BandError", "DimensionMismatch", "LAPACKException",
    "BMB200Error", "Handle", "handle", "load", "LIB_PATH", "PROTOTYPES", "bandeddata", "bandwidth", "bandwidths",
    "brand", "colmajor", "to_colmajor", "mul_", "matmul", "gbmm_", "lu", "lu_", "ldiv_", "solve", "factorize",
    "gbmv_host", "tbsv_", "tbmv_", "ldiv_tri_", "lmul_tri_", "sbmv_", "mul_sym_", "axpy_", "copyto_", "axpby_", "badd", "bsub",
    "bscale", "similar", "gbmv_", "hbmv_", "gbtrf_", "gbtrs_", "gbmm_typed_", "gbmm_bd_typed_", "BandedCholesky", "PosDefException", "cholesky", "cholesky_", "ldiv_chol_", "pbtrf_", "pbtrs_",
]
