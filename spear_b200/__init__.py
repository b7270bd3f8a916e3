"""spear_b200 -- B200-native replacement for Spear's CPU particle hot path.

The product is the CUDA shared library spear_b200/libspear_particles.so behind the
C ABI of include/spear_particles.h. This package is only its ctypes binding
(spear_b200.particles) plus the ctypes mirror of the POD descriptors
(spear_b200.abi). There is no CPU fallback: importing spear_b200.particles
without the built library, or creating a context without a B200, raises.
"""
from . import abi  # noqa: F401
