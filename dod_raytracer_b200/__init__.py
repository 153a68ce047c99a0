"""dod_raytracer_b200 -- B200-native (sm_100a) ray-query path for AVassilev98/dod_raytracer.

The product is the C-ABI library ``lib/libdodrt_cuda.so`` (sources in ``csrc/``, interface in
``include/dodrt.h``).  ``capi`` is its ctypes binding; ``host`` mirrors the reference's host side
(scene construction, kd-tree build, ray tables) above the C ABI.  Nothing in this package imports
``oracle/`` and there is no CPU fallback for any query.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
