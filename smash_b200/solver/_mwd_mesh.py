"""``smash.solver._mwd_mesh`` (derived_type/mwd_mesh.f90)."""
from ._derived_types import MeshDT  # noqa: F401
