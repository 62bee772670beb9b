"""TEST INFRASTRUCTURE: the CPU oracle behind the call signatures of ``smash_b200.solver._mw_forward`` /
``_mw_multiple_run``, so that the host-side drivers (optimisers, L-curve, ensemble callers) can be checked against
the reference's golden values on a machine without a GPU.  Never imported by the package."""
import numpy as np

import oracle


def forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output, cost=0.0):
    return np.float32(oracle.forward(setup, mesh, input_data, parameters, parameters_bgd, states, states_bgd, output))


def forward_b(setup, mesh, input_data, parameters, parameters_b, parameters_bgd, parameters_bgd_b, states, states_b,
              states_bgd, states_bgd_b, output, output_b, cost=0.0, cost_b=1.0):
    return np.float32(oracle.forward_b(setup, mesh, input_data, parameters, parameters_b, parameters_bgd, states, states_b,
                                       states_bgd, output))


def hyper_forward(setup, mesh, input_data, parameters, hyper_parameters, hyper_parameters_bgd, states, hyper_states,
                  hyper_states_bgd, output, cost=0.0):
    return np.float32(oracle.hyper_forward(setup, mesh, input_data, parameters, hyper_parameters, states, hyper_states,
                                           output))


def hyper_forward_b(setup, mesh, input_data, parameters, parameters_b, hyper_parameters, hyper_parameters_b,
                    hyper_parameters_bgd, hyper_parameters_bgd_b, states, states_b, hyper_states, hyper_states_b,
                    hyper_states_bgd, hyper_states_bgd_b, output, output_b, cost=0.0, cost_b=1.0):
    return np.float32(oracle.hyper_forward_b(setup, mesh, input_data, parameters, hyper_parameters, hyper_parameters_b,
                                             states, hyper_states, hyper_states_b, output))


def compute_multiple_run(setup, mesh, input_data, parameters, states, output, sample, ind_parameters_states, res_cost,
                         res_qsim):
    oracle.compute_multiple_run(setup, mesh, input_data, parameters, states, output, sample, ind_parameters_states,
                                res_cost, res_qsim, nthreads=4)
