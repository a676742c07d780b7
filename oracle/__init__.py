"""CPU oracle for the sQUlearn statevector hot path (TEST INFRASTRUCTURE ONLY).

This package is a NumPy complex128 restatement of what the reference
(sQUlearn v0.11.0) computes on its statevector path: gate list -> batched
statevectors -> fidelity Gram matrix / Pauli expectation values / parameter
gradients.  It exists to CHECK the CUDA engine in ``squlearn_b200``; it is not a
product path and is not a fallback.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``squlearn_b200/`` imports it (``tests/test_no_oracle_in_product.py`` enforces this).

Parity status: PINNED.  The reference itself cannot be imported in this image
(pennylane / qiskit / qulacs are not installed and there is no network), and the
arithmetic lives in those un-vendored third-party engines (pyproject.toml:31-48
pins only lower bounds: pennylane>=0.34.0, qiskit>=1.0, qulacs>=0.6.0).  Parity is
anchored instead on the reference's own known-answer tests, transcribed into
``tests/golden/reference_goldens.json`` (each with its reference file:line) and
reproduced by this oracle in ``tests/test_oracle_goldens.py``.

Conventions:
  * qubit order is little-endian (Qiskit / Qulacs): qubit 0 is the least
    significant bit of the amplitude index;  ``to_pennylane_order`` converts to the
    bit-reversed order ``Executor.pennylane_execute`` returns
    (src/squlearn/util/pennylane/pennylane_circuit.py:397-400, :633-634).
  * a batch of states is an ``(N, 2**n)`` complex128 array, one row per data point.
"""

from .gates import gate_matrix, GATE_ARITY, DIAGONAL_GATES
from .statevector import (
    Gate,
    simulate,
    simulate_reference_style,
    to_pennylane_order,
)
from .circuits import (
    chebyshev_pqc,
    chebyshev_pqc_num_parameters,
    chebyshev_pqc_initial_parameters,
    hubregtsen,
    hubregtsen_num_parameters,
    hubregtsen_initial_parameters,
    chebyshev_rx,
    chebyshev_rx_num_parameters,
    chebyshev_rx_initial_parameters,
    layered_ry_rx,
    highdim,
)
from .observables import (
    pauli_expectation,
    single_pauli_string,
    summed_paulis,
    summed_paulis_squared,
    xyz_measurement_strings,
    random_pauli_list,
)
from .kernels import (
    fidelity_gram_pairloop,
    fidelity_gram_zgemm,
    fidelity_kernel,
    projected_kernel_features,
    gaussian_outer_kernel,
    projected_quantum_kernel,
)
from .qnn import qnn_evaluate, parameter_shift_variants
