"""CPU: the host-side data model (storage rule, views, transposes, band counting) on CPU tensors.
Kernels are never called here."""
import numpy as np
import pytest
import torch

import bandedmatrices_b200 as bm
from bandedmatrices_b200.linalg import _num_zeroband_l, _num_zeroband_u, materialize_transpose


def _rand_banded(rng, m, n, l, u):
    D = rng.standard_normal((m, n))
    return np.triu(np.tril(D, u), -l) if -l <= u else np.zeros((m, n))


@pytest.mark.parametrize("shape", [(10, 12, 2, 3), (12, 10, 3, 2), (7, 7, 0, 0), (5, 9, 1, 6), (9, 5, 6, 1), (10, 10, -1, 2),
                                   (10, 10, 2, -1), (1, 10, 0, 9), (6, 6, 8, 8)])
def test_storage_rule_roundtrip(rng, shape):
    m, n, l, u = shape
    D = _rand_banded(rng, m, n, l, u)
    A = bm.BandedMatrix.from_dense(D, (l, u), device="cpu")
    assert A.data.shape == (n, max(0, l + u + 1))
    # data[u+k-j, j] = A[k,j]  (src/banded/BandedMatrix.jl:414-419)
    d = A.banddata_host()
    for j in range(n):
        for k in range(max(0, j - u), min(m - 1, j + l) + 1):
            assert d[u + k - j, j] == D[k, j]
    assert np.array_equal(A.to_dense(), D)
    assert bm.bandwidths(A) == (l, u) and bm.bandwidths(A.T) == (u, l)
    assert bm.bandeddata(A).shape == (max(0, l + u + 1), n)


def test_views_shift_bandwidths(rng):
    m, n, l, u = 11, 13, 3, 2
    D = _rand_banded(rng, m, n, l, u)
    A = bm.BandedMatrix.from_dense(D, (l, u), device="cpu")
    for s in (0, 1, 2, 4):
        V = A.view_cols(s)
        assert (V.l, V.u) == (l + s, u - s) and V.shape == (m, n - s)
        assert np.array_equal(V.to_dense(), D[:, s:])
        W = A.view_rows(s)
        assert (W.l, W.u) == (l - s, u + s) and W.shape == (m - s, n)
        assert np.array_equal(W.to_dense(), D[s:, :])


def test_band_bookkeeping_refuses_cpu_tensors(rng):
    """The transposed copy and the zero-band counts of the gbmm! driver are device kernels now (no eager tensor ops on the
    product path): CPU operands are refused, not silently computed.  Their arithmetic is checked in tests/test_gpu_ewise.py."""
    D = _rand_banded(rng, 8, 11, 2, 3)
    A = bm.BandedMatrix.from_dense(D, (2, 3), device="cpu")
    with pytest.raises(TypeError):
        materialize_transpose(A.T)
    with pytest.raises(TypeError):
        _num_zeroband_u(A)
    with pytest.raises(TypeError):
        _num_zeroband_l(A)


def test_constructor_checks():
    with pytest.raises(ValueError):  # BandedMatrix.jl:22-24
        bm.BandedMatrix(torch.zeros((5, 4), dtype=torch.float64), 5, 1, 1)
    with pytest.raises(TypeError):
        bm.BandedMatrix(torch.zeros((5, 3), dtype=torch.float32), 5, 1, 1)


def test_dimension_checks_before_any_kernel():
    A = bm.BandedMatrix(torch.zeros((6, 3), dtype=torch.float64), 6, 1, 1)
    with pytest.raises(bm.DimensionMismatch):  # test/test_linalg.jl:333-338
        bm.mul_(torch.zeros(6, dtype=torch.float64), A, torch.zeros(5, dtype=torch.float64))
    R = bm.BandedMatrix(torch.zeros((7, 3), dtype=torch.float64), 6, 1, 1)
    with pytest.raises(bm.DimensionMismatch):  # non-square solve, test/test_bandedlu.jl:79-87
        bm.solve(R, torch.zeros(6, dtype=torch.float64))


def test_new_wrappers_validate_before_touching_the_device():
    """tbsv_/tbmv_/sbmv_/axpy_/copyto_ raise the reference's argument errors (src/blas.jl:86-93,124-131;
    src/generic/broadcast.jl:979-983) from the host checks alone, and refuse CPU tensors instead of computing on them."""
    A = bm.BandedMatrix(torch.zeros((6, 4), dtype=torch.float64), 6, 1, 2)   # 6 x 6, (l,u) = (1,2), CPU tensors
    x = torch.zeros(6, dtype=torch.float64)
    with pytest.raises(bm.DimensionMismatch):
        bm.tbsv_("U", "N", "N", 7, 2, A.data[:, :3], x)                      # matrix is not square: dimensions are 6, 7
    with pytest.raises(bm.DimensionMismatch):
        bm.tbmv_("U", "N", "N", 6, 2, A.data[:, :3], x[:5])                  # size of A != length(x)
    with pytest.raises(ValueError):
        bm.tbsv_("U", "N", "N", 6, 3, A.data[:, :3], x)                      # triangular banded data missing
    with pytest.raises(bm.DimensionMismatch):
        bm.sbmv_("U", 2, 1.0, A.data[:, :3], x[:5], 0.0, x)
    with pytest.raises(bm.DimensionMismatch):
        bm.axpy_(1.0, A, bm.BandedMatrix(torch.zeros((5, 4), dtype=torch.float64), 6, 1, 2))
    with pytest.raises(bm.DimensionMismatch):
        bm.mul_sym_(x, "U", bm.BandedMatrix(torch.zeros((5, 4), dtype=torch.float64), 6, 1, 2), x)
    with pytest.raises(TypeError):                                           # valid arguments, CPU tensors: no CPU fallback
        bm.tbsv_("U", "N", "N", 6, 2, A.data[:, :3], x)
    with pytest.raises(TypeError):
        bm.axpy_(1.0, A, A)


def test_cholesky_and_typed_wrappers_validate_before_touching_the_device():
    """pbtrf_/pbtrs_ raise the reference's argument errors (src/lapack.jl:278-282, 308-314), the typed wrappers refuse element
    types outside the four BLAS floats and mixed types, and none of them computes on CPU tensors."""
    d = torch.zeros((6, 3), dtype=torch.float64)  # 6 columns, 3 band rows, CPU
    with pytest.raises(ValueError):
        bm.pbtrf_("X", 6, 2, d)               # chkuplo
    with pytest.raises(ValueError):
        bm.pbtrf_("U", 7, 2, d)               # Matrix must be square
    with pytest.raises(ValueError):
        bm.pbtrf_("U", 6, 3, d)               # Not enough bands
    with pytest.raises(bm.DimensionMismatch):
        bm.pbtrs_("U", 6, 2, d, torch.zeros(5, dtype=torch.float64))
    with pytest.raises(TypeError):            # valid arguments, CPU tensors: no CPU fallback
        bm.pbtrf_("U", 6, 2, d)
    with pytest.raises(TypeError):
        bm.cholesky(bm.BandedMatrix(torch.zeros((6, 3), dtype=torch.float64), 6, 1, 1))
    h = torch.zeros((6, 3), dtype=torch.float16)
    with pytest.raises(TypeError):            # not a BLAS float
        bm.gbmv_("N", 6, 1, 1, 1.0, h, torch.zeros(6, dtype=torch.float16), 0.0, torch.zeros(6, dtype=torch.float16))
    s = torch.zeros((6, 3), dtype=torch.float32)
    with pytest.raises(TypeError):            # element types differ
        bm.gbmv_("N", 6, 1, 1, 1.0, s, torch.zeros(6, dtype=torch.float64), 0.0, torch.zeros(6, dtype=torch.float32))
    with pytest.raises(bm.DimensionMismatch):
        bm.gbmv_("N", 6, 1, 1, 1.0, s, torch.zeros(5, dtype=torch.float32), 0.0, torch.zeros(6, dtype=torch.float32))
    with pytest.raises(TypeError):            # valid arguments, CPU tensors
        bm.gbmv_("C", 6, 1, 1, 1.0, torch.zeros((6, 3), dtype=torch.complex64), torch.zeros(6, dtype=torch.complex64), 0.0,
                 torch.zeros(6, dtype=torch.complex64))
    with pytest.raises(TypeError):
        bm.gbtrf_(6, 1, 1, torch.zeros((6, 4), dtype=torch.complex128))


def test_sharded_cholesky_solve_bounds():
    from bandedmatrices_b200.sharded import rhs_bounds

    for nrhs, world in ((7, 2), (256, 8), (3, 4)):
        owned = [rhs_bounds(nrhs, r, world) for r in range(world)]
        assert owned[0][0] == 0 and owned[-1][1] == nrhs and all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))
