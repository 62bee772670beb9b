"""``smash.solver._mwd_setup`` (derived_type/mwd_setup.f90)."""
from ._derived_types import Optimize_SetupDT, SetupDT  # noqa: F401
