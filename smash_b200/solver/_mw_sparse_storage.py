"""Drop-in for ``smash.solver._mw_sparse_storage`` (routine/mw_sparse_storage.f90:12-260): the sparse index of the active
cells and the matrix <-> vector conversions in that order (the order of ``mesh.path`` restricted to active cells)."""
from __future__ import annotations

import numpy as np

from ._derived_types import compute_rowcol_to_ind_sparse  # noqa: F401  (mw_sparse_storage.f90:12-49)


def _path_cells(mesh):
    """(row, col, active) of the cells ``path`` lists, 0-based (the Python side of ``path`` is 0-based,
    _f90wrap_decorator.py:72-106; unused entries are negative)."""
    rows, cols = np.asarray(mesh.path[0]), np.asarray(mesh.path[1])
    ok = (rows >= 0) & (cols >= 0)
    r, c = rows[ok], cols[ok]
    return r, c, np.asarray(mesh.active_cell)[r, c] == 1


def _matrix_to_vector(mesh, matrix, vector):
    r, c, act = _path_cells(mesh)
    n = int(act.sum())
    vector[:n] = np.asarray(matrix)[r[act], c[act]]                       # :73-93


def _vector_to_matrix(mesh, vector, matrix, na_value):
    r, c, act = _path_cells(mesh)
    matrix[r[act], c[act]] = np.asarray(vector)[: int(act.sum())]         # :176-186
    matrix[r[~act], c[~act]] = na_value                                   # :188-198


def sparse_matrix_to_vector_r(mesh, matrix, vector):
    _matrix_to_vector(mesh, matrix, vector)


def sparse_matrix_to_vector_i(mesh, matrix, vector):
    _matrix_to_vector(mesh, matrix, vector)


def sparse_vector_to_matrix_r(mesh, vector, matrix, na_value=None):
    _vector_to_matrix(mesh, vector, matrix, np.float32(-99.0) if na_value is None else na_value)


def sparse_vector_to_matrix_i(mesh, vector, matrix, na_value=None):
    _vector_to_matrix(mesh, vector, matrix, -99 if na_value is None else na_value)
