"""``smash.solver._mwd_states`` (derived_type/mwd_states.f90)."""
from ._derived_types import Hyper_StatesDT, StatesDT  # noqa: F401
