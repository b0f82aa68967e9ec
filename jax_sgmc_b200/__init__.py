"""jax_sgmc_b200 -- B200-native sampling hot path of jax-sgmc.

The operator API of jax-sgmc for its data-parallel sampling step
(``potential``, ``integrator``, ``adaption``, ``solver``, ``alias`` and the thin
``scheduler`` / ``data`` / ``io`` glue) on top of hand-written sm_100a CUDA
kernels exposed through the C ABI in ``include/sgmc_b200.h``.
"""
__version__ = "0.1.0"
