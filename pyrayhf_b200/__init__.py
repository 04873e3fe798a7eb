"""pyrayhf_b200 -- B200-native (sm_100a) vertical forward operator for PyRayHF.

Drop-in for ``PyRayHF.library.vertical_forward_operator`` (PyRayHF/library.py:459-509)
plus a batched ``[n_profiles x n_alt]`` form.  The compute path is hand-written CUDA
behind a C ABI (``include/pyrayhf_b200.h``); there is no CPU fallback.
"""
import logging

# Same logger name as the reference package (PyRayHF/__init__.py:4-6) so that
# applications capturing 'PyRayHF_logger' keep seeing the shape-mismatch message.
logger = logging.getLogger('PyRayHF_logger')

__version__ = "0.1.0"

from pyrayhf_b200 import library  # noqa: E402,F401
from pyrayhf_b200.library import (  # noqa: E402,F401
    vertical_forward_operator,
    vertical_forward_operator_batched,
    vertical_forward_operator_streamed,
    pinned_empty,
    find_mu_mup,
    residual_VH_batched,
    install,
    uninstall,
)
from pyrayhf_b200 import inversion  # noqa: E402,F401
from pyrayhf_b200.inversion import (  # noqa: E402,F401
    brute_force_fit,
    brute_force_search,
    brute_grid,
    chapman_profile_builder,
    freq2den,
    minimize_parameters,
    nmf2_from_max_frequency,
)
from pyrayhf_b200 import stages  # noqa: E402,F401
from pyrayhf_b200.stages import (  # noqa: E402,F401
    constants,
    den2freq,
    find_X,
    find_Y,
    smooth_nonuniform_grid,
    regrid_to_nonuniform_grid,
    find_vh,
)
from pyrayhf_b200 import snell  # noqa: E402,F401
from pyrayhf_b200.snell import (  # noqa: E402,F401
    trace_ray_cartesian_snells,
    trace_ray_spherical_snells,
    trace_rays_snells_batched,
)
from pyrayhf_b200 import sharding  # noqa: E402,F401
