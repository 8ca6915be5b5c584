"""magpy_rv_b200 -- B200-native (sm_100a CUDA) implementation of the MAGPy-RV likelihood hot path.

Like the reference package (``src/magpy_rv/__init__.py`` is empty) the sub-modules are imported
explicitly: ``parameters``, ``kernels``, ``models``, ``gp_likelihood``, ``mcmc``, ``mcmc_aux``,
``auxiliary`` keep the reference's names and signatures; ``engine`` / ``sharding`` / ``_native`` are
the batched boundary to the CUDA library (``include/magpy_rv_b200.h``).
"""
__version__ = "0.1.0"
