"""Mirror of the reference's generated ``smash.solver`` package for the forward/adjoint hot path."""
