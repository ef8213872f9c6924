"""MPS input: thin binding of the native reader in libblp.so (``blp_mps_*``, include/blp.h).

The reference loads its random test models with ``MILPInstance(file_name=...)``
(test_simple_mip_solver/helpers.py:42), which goes through CLP's MPS reader; the fixtures under
``scale_1_models`` / ``example_models`` are CLP-written. The parsing is done in C++
(csrc/blp_mps.cpp); this module only wraps the arrays. A pure-Python
restatement of the same dialect lives in ``oracle/mps_py.py`` and is used by the tests as the
checker, never here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np
import scipy.sparse as sp


@dataclass
class MpsModel:
    name: str = ''
    row_names: List[str] = field(default_factory=list)
    row_senses: List[str] = field(default_factory=list)      # 'L', 'G', 'E' per constraint row
    col_names: List[str] = field(default_factory=list)
    A: sp.csr_matrix = None
    rhs: np.ndarray = None
    c: np.ndarray = None
    obj_offset: float = 0.0
    l: np.ndarray = None
    u: np.ndarray = None
    integer_indices: List[int] = field(default_factory=list)


def read_mps(path: str) -> MpsModel:
    """Parse an MPS file with the native reader. Raises ``BlpError`` if the file cannot be read
    or libblp.so has not been built (there is no Python fallback)."""
    from ..engine import BlpError, load_library
    lib = load_library()
    h = C.c_void_p()
    if lib.blp_mps_read(str(path).encode(), C.byref(h)) != 0:
        raise BlpError(lib.blp_mps_last_error().decode())
    try:
        m, n, n_int, nnz = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        lib.blp_mps_dims(h, C.byref(m), C.byref(n), C.byref(nnz), C.byref(n_int))
        m, n, nnz, n_int = m.value, n.value, nnz.value, n_int.value
        rowptr = np.zeros(m + 1, dtype=np.int32)
        colidx = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        rhs, sense = np.zeros(m), np.zeros(m, dtype='S1')
        c, l, u = np.zeros(n), np.zeros(n), np.zeros(n)
        ints = np.zeros(n_int, dtype=np.int32)
        off = C.c_double(0.0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.blp_mps_copy(h, p(rowptr), p(colidx), p(val), p(rhs), p(sense), p(c),
                         C.cast(C.byref(off), C.c_void_p), p(l), p(u), p(ints))
        mdl = MpsModel(name=lib.blp_mps_name(h).decode(),
                       row_names=[lib.blp_mps_row_name(h, i).decode() for i in range(m)],
                       row_senses=[s.decode() for s in sense],
                       col_names=[lib.blp_mps_col_name(h, j).decode() for j in range(n)],
                       A=sp.csr_matrix((val, colidx, rowptr), shape=(m, n)), rhs=rhs, c=c,
                       obj_offset=off.value, l=l, u=u, integer_indices=[int(j) for j in ints])
    finally:
        lib.blp_mps_free(h)
    return mdl
