"""TEST INFRASTRUCTURE: pure-Python restatement of the MPS dialect the native reader accepts.

The product reads MPS files with the C++ reader in libblp.so (csrc/blp_mps.cpp, bound by
simple_mip_solver_b200/compat/mps.py); this module is the checker the tests compare it with and
is never imported by the product.

The reference loads its random test models with ``MILPInstance(file_name=...)``
(test_simple_mip_solver/helpers.py:42), which goes through CLP's MPS reader. The fixtures under
``scale_1_models`` / ``example_models`` are CLP-written: whitespace separated fields, sections
ROWS / COLUMNS / RHS / BOUNDS / ENDATA, integer columns flagged by ``UI`` bounds (or MARKER
lines). This reader accepts that dialect (plus RANGES-free general MPS) and returns raw arrays;
it tolerates empty rows and empty columns, which several fixtures contain.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np
import scipy.sparse as sp

_INF = float('inf')


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
    mdl = MpsModel()
    obj_row = None
    row_idx: Dict[str, int] = {}
    col_idx: Dict[str, int] = {}
    entries = []                      # (row, col, value)
    obj_coefs: Dict[int, float] = {}
    rhs_vals: Dict[int, float] = {}
    lower: Dict[int, float] = {}
    upper: Dict[int, float] = {}
    integer = set()
    in_int_marker = False
    section = None

    def col_of(name):
        j = col_idx.get(name)
        if j is None:
            j = len(mdl.col_names)
            col_idx[name] = j
            mdl.col_names.append(name)
            if in_int_marker:
                integer.add(j)
        return j

    with open(path) as fh:
        for raw in fh:
            if not raw.strip() or raw.lstrip().startswith('*'):
                continue
            tok = raw.split()
            if not raw[0].isspace():                      # section header
                section = tok[0].upper()
                if section == 'NAME' and len(tok) > 1:
                    mdl.name = tok[1]
                if section == 'ENDATA':
                    break
                continue
            if section == 'ROWS':
                sense, name = tok[0].upper(), tok[1]
                if sense == 'N':
                    if obj_row is None:
                        obj_row = name
                else:
                    row_idx[name] = len(mdl.row_names)
                    mdl.row_names.append(name)
                    mdl.row_senses.append(sense)
            elif section == 'COLUMNS':
                if len(tok) >= 3 and tok[1].upper() == "'MARKER'":
                    in_int_marker = tok[2].upper() == "'INTORG'"
                    continue
                j = col_of(tok[0])
                for k in range(1, len(tok) - 1, 2):
                    rname, val = tok[k], float(tok[k + 1])
                    if rname == obj_row:
                        obj_coefs[j] = val
                    elif rname in row_idx:
                        entries.append((row_idx[rname], j, val))
            elif section == 'RHS':
                start = 1 if len(tok) % 2 == 1 else 0     # optional set name
                for k in range(start, len(tok) - 1, 2):
                    rname, val = tok[k], float(tok[k + 1])
                    if rname == obj_row:
                        mdl.obj_offset = -val
                    elif rname in row_idx:
                        rhs_vals[row_idx[rname]] = val
            elif section == 'BOUNDS':
                kind = tok[0].upper()
                # "UI BOUND x_0 100." (set name present) or "UI x_0 100."
                if kind in ('FR', 'MI', 'PL', 'BV'):
                    cname = tok[2] if len(tok) >= 3 else tok[1]
                    val = None
                else:
                    cname = tok[2] if len(tok) >= 4 else tok[1]
                    val = float(tok[-1])
                j = col_of(cname)
                if kind == 'UP':
                    upper[j] = val
                    if val < 0 and j not in lower:
                        lower[j] = -_INF
                elif kind == 'UI':
                    upper[j] = val
                    integer.add(j)
                elif kind == 'LO':
                    lower[j] = val
                elif kind == 'LI':
                    lower[j] = val
                    integer.add(j)
                elif kind == 'FX':
                    lower[j] = upper[j] = val
                elif kind == 'FR':
                    lower[j], upper[j] = -_INF, _INF
                elif kind == 'MI':
                    lower[j] = -_INF
                elif kind == 'PL':
                    upper[j] = _INF
                elif kind == 'BV':
                    lower[j], upper[j] = 0.0, 1.0
                    integer.add(j)
            # RANGES and anything else: not produced by the reference's writer; ignored.

    n, m = len(mdl.col_names), len(mdl.row_names)
    if entries:
        r, cidx, v = zip(*entries)
    else:
        r, cidx, v = (), (), ()
    mdl.A = sp.csr_matrix((np.asarray(v, dtype=float), (np.asarray(r, dtype=int),
                                                        np.asarray(cidx, dtype=int))), shape=(m, n))
    mdl.rhs = np.array([rhs_vals.get(i, 0.0) for i in range(m)])
    mdl.c = np.array([obj_coefs.get(j, 0.0) for j in range(n)])
    mdl.l = np.array([lower.get(j, 0.0) for j in range(n)])
    mdl.u = np.array([upper.get(j, _INF) for j in range(n)])
    mdl.integer_indices = sorted(integer)
    return mdl
