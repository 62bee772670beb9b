"""``smash.solver._mw_sparse_storage`` (routine/mw_sparse_storage.f90:12-49)."""
from ._derived_types import compute_rowcol_to_ind_sparse  # noqa: F401
