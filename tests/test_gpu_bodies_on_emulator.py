"""CPU run of the bodies of GPU parity tests whose device work is only statevector simulation + Pauli
expectation values: the same assertions, with the emulator engine (tests/emu_engine.py) in place of the CUDA
engine.  Covers the Python host glue (QNN value / gradient assembly, reference goldens) without a GPU."""

import squlearn_b200 as sq
from tests import test_gpu_parity as gpu_tests
from tests.emu_engine import EmuExecutor


def test_qnn_goldens_and_gradients_on_emulator(goldens, regression_data):
    gpu_tests.test_qnn_goldens_and_gradients(sq, EmuExecutor(), goldens, regression_data)


def test_optree_gradient_golden_on_emulator(goldens):
    gpu_tests.test_optree_gradient_golden_through_the_product(sq, EmuExecutor(), goldens)


def test_input_derivative_on_emulator(goldens):
    gpu_tests.test_input_derivative_golden_and_finite_differences(sq, EmuExecutor(), goldens)


def test_qrc_and_highdim_goldens_on_emulator(goldens, regression_data):
    gpu_tests.test_qrc_and_highdim_goldens(sq, EmuExecutor(), goldens, regression_data)


def test_closed_form_kernel_goldens_on_emulator(goldens):
    gpu_tests.test_closed_form_goldens_through_the_product(sq, EmuExecutor(), goldens)


def test_fidelity_kernel_c1_and_duplicate_policies_on_emulator():
    gpu_tests.test_fidelity_kernel_config_c1_and_duplicate_policies(sq, EmuExecutor())


def test_projected_quantum_kernel_c2_on_emulator():
    gpu_tests.test_projected_quantum_kernel_config_c2(sq, EmuExecutor())


def test_fidelity_kernel_regularization_mitigation_and_none_policy_on_emulator():
    gpu_tests.test_fidelity_kernel_regularization_mitigation_and_none_policy(sq, EmuExecutor())


def test_qnn_regressor_reference_golden_and_fit_on_emulator(goldens, regression_data):
    gpu_tests.test_qnn_regressor_reference_golden_and_fit(sq, EmuExecutor(), goldens, regression_data)


def test_second_order_derivatives_on_emulator():
    gpu_tests.test_second_order_derivatives_against_finite_differences(sq, EmuExecutor())
