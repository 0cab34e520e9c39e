"""mflbm-b200: B200-native (sm_100a) implementation of the lanl/MF-LBM time-step hot path.

The product is the CUDA shared library behind include/mflbm.h (csrc/); this package holds its
ctypes binding and the host-side mirror of the reference driver used by the tests and bench.py.
The directory name contains a hyphen, so import it through the root-level shim ``mflbm_b200``.
"""
from .binding import (Config, Arrays, Context, MflbmError, EXPORTS, SOLID_DTYPE, FLUID_DTYPE,  # noqa: F401
                      SOLVER_MULTIPHASE, SOLVER_SINGLEPHASE, build, load, nccl_unique_id, field_shape, geometry_preprocess, GeometryConfig)
from .driver import Driver, write_control_file, write_wall_array, build_host, load_host  # noqa: F401,E402
