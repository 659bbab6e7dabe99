"""Eigen 3.4.0's operation order for `LLT<Matrix<float,N,N,RowMajor>, Lower>` and its solve --
TEST INFRASTRUCTURE (restatement of a third-party dependency that is absent from /root/reference:
Eigen 3.4.0, vcpkg port `eigen3`; call sites /root/reference/cpp/pnp/lev_marq.h:102-106,299-314).

What is restated (x86-64 build without -march flags, i.e. SSE2 packets of 4 floats and no FMA --
the reference's CMake adds no architecture flags, cpp/CMakeLists.txt:10-16):
  * llt_inplace<float, Lower>::unblocked (size < 32): per column k
      x = A(k,k) - A10.squaredNorm(); A(k,k) = sqrt(x); A21 -= A20 * A10^T; A21 /= x
    with `squaredNorm` / dots evaluated by Eigen's linear-vectorised reduction: packets of four
    products added lane-wise, the packet reduced as (p0+p2)+(p1+p3), remaining products added
    one by one; the row-major GEMV kernel forms each row's dot the same way.
  * triangular_solve_vector, forward (row-major lower, panels of 8) and backward (the adjoint seen
    as column-major upper, panels of 8, axpy updates inside the panel).
The packet boundaries of `squaredNorm` depend on the 16-byte alignment of the row start, i.e. on
where the 9x9 matrix sits in memory (36-byte rows); `base` selects that offset (0 = 16-byte aligned).

PARITY NOTE: the reference's one numeric fixture for this code,
cpp/examples/levmarq_ill_conditioned_float32_issue.cpp:16-63 (residual 0.0028946274, expected cost
change +0.000244110823), was searched for with this restatement over 50 build variants (packet
size 4 / 8, FMA contraction in packets and / or scalars, every alignment offset, no
vectorisation); none reproduces the recorded digits (the system has condition number 4.4e10, so
they depend on the exact binary that produced them, which is unknown).  tests/test_oracle_solvers.py
asserts what every variant agrees on."""
from __future__ import annotations

import numpy as np

F = np.float32


def _predux4(c):
    return F(F(c[0] + c[2]) + F(c[1] + c[3]))


def _redux_dot(a, b, aligned_start=0):
    """sum_i a_i * b_i as Eigen's redux_impl<LinearVectorizedTraversal, NoUnrolling> evaluates it."""
    n = len(a)
    prod = (np.asarray(a, F) * np.asarray(b, F)).astype(F)
    s0 = min(aligned_start, n)
    a1 = ((n - s0) // 4) * 4
    a2 = ((n - s0) // 8) * 8
    if a1 == 0:
        res = prod[0]
        for i in range(1, n):
            res = F(res + prod[i])
        return res
    p0 = prod[s0:s0 + 4].copy()
    if a1 > 4:
        p1 = prod[s0 + 4:s0 + 8].copy()
        i = s0 + 8
        while i < s0 + a2:
            p0 = (p0 + prod[i:i + 4]).astype(F)
            p1 = (p1 + prod[i + 4:i + 8]).astype(F)
            i += 8
        p0 = (p0 + p1).astype(F)
        if a1 > a2:
            p0 = (p0 + prod[s0 + a2:s0 + a2 + 4]).astype(F)
    res = _predux4(p0)
    for i in range(s0):
        res = F(res + prod[i])
    for i in range(s0 + a1, n):
        res = F(res + prod[i])
    return res


def _gemv_row_dot(a, b):
    """One row of general_matrix_vector_product<RowMajor>: packet accumulators over full packets,
    (c0+c2)+(c1+c3), then the remaining columns one by one."""
    n = len(a)
    a, b = np.asarray(a, F), np.asarray(b, F)
    c = np.zeros(4, F)
    j = 0
    while j + 4 <= n:
        c = ((a[j:j + 4] * b[j:j + 4]).astype(F) + c).astype(F)
        j += 4
    cc = _predux4(c)
    while j < n:
        cc = F(F(a[j] * b[j]) + cc)
        j += 1
    return cc


def _first_aligned(byte_offset, n):
    return min(((16 - byte_offset % 16) % 16) // 4, n)


def llt_lower(A, base=0):
    """LLT<RowMajor, Lower>::compute on the lower triangle of A.  Returns (L, ok)."""
    n = A.shape[0]
    L = np.tril(np.array(A, F, copy=True))
    for k in range(n):
        rs = n - k - 1
        x = L[k, k]
        if k > 0:
            x = F(x - _redux_dot(L[k, :k], L[k, :k], _first_aligned(base + 4 * n * k, k)))
        if not (x > 0):
            return L, False
        x = F(np.sqrt(x))
        L[k, k] = x
        if k > 0 and rs > 0:
            for i in range(k + 1, n):
                # a single remaining row falls back to .dot() (GeneralProduct.h, scaleAndAddTo)
                d = _redux_dot(L[i, :k], L[k, :k]) if rs == 1 else _gemv_row_dot(L[i, :k], L[k, :k])
                L[i, k] = F(L[i, k] - d)
        if rs > 0:
            L[k + 1:, k] = (L[k + 1:, k] / x).astype(F)
    return L, True


def llt_solve(L, b):
    """LLT::solve: matrixL().solveInPlace, then matrixU().solveInPlace (panel width 8)."""
    n = len(b)
    rhs = np.array(b, F, copy=True)
    pi = 0
    while pi < n:                                       # forward, row-major lower
        apw = min(n - pi, 8)
        if pi > 0:
            for i in range(pi, pi + apw):
                rhs[i] = F(rhs[i] - _gemv_row_dot(L[i, :pi], rhs[:pi]))
        for k in range(apw):
            i = pi + k
            if k > 0:
                rhs[i] = F(rhs[i] - _redux_dot(L[i, pi:i], rhs[pi:i]))
            if rhs[i] != 0:
                rhs[i] = F(rhs[i] / L[i, i])
        pi += 8
    pi = n
    while pi > 0:                                       # backward, L^T as column-major upper
        apw = min(pi, 8)
        start = pi - apw
        for k in range(apw):
            i = pi - k - 1
            if rhs[i] != 0:
                rhs[i] = F(rhs[i] / L[i, i])
                for j in range(start, i):
                    rhs[j] = F(rhs[j] - F(rhs[i] * L[i, j]))
        for i in range(start):                          # column-major GEMV, sequential over the panel's columns
            c = F(0)
            for j in range(start, pi):
                c = F(F(L[j, i] * rhs[j]) + c)
            rhs[i] = F(rhs[i] - c)
        pi -= 8
    return rhs


def selfadjoint_lower_times(A, v):
    """selfadjointView<Lower>() * v for a row-major matrix of size <= 9 (the scalar loop of
    selfadjoint_matrix_vector_product: the two-column vector body only starts at size 10)."""
    n = len(v)
    assert n <= 9
    v = np.asarray(v, F)
    res = np.zeros(n, F)
    for j in range(n):
        t1 = v[j]
        t2 = F(0)
        res[j] = F(res[j] + F(A[j, j] * t1))
        for i in range(j):
            res[i] = F(res[i] + F(A[j, i] * t1))
            t2 = F(t2 + F(A[j, i] * v[i]))
        res[j] = F(res[j] + t2)
    return res


def dot_fixed(a, b):
    """Inner product / squaredNorm of fixed-size vectors (<= 9 entries): two packets, then the tail."""
    return _redux_dot(a, b, 0)
