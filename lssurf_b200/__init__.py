"""lssurf_b200: B200-native smooth_fit hot path behind the LSsurf Python API.

Public names mirror ``LSsurf/__init__.py`` for the hot path: ``fd_grid``, ``lin_op``
(+ ``fd_op``), ``smooth_fit``, ``iterate_fit``, ``RDE``, ``calc_sigma_extra``, the grid / constraint /
bias set-up functions, plus ``point_data`` containers standing in for pointCollection.  Everything
numerical runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/lssurf_b200.h``;
importing the package does not touch the GPU, calling into it without one raises.
"""
from . import _cabi, parallel, point_data, synthetic      # noqa: F401
from .RDE import RDE                                      # noqa: F401
from .bias_functions import assign_bias_ID, setup_bias_fit, parse_biases       # noqa: F401
from .calc_sigma_extra import calc_sigma_extra            # noqa: F401
from .constraint_functions import setup_smoothness_constraints, build_reference_epoch_matrix   # noqa: F401
from .fd_grid import fd_grid                              # noqa: F401
from .calc_sigma_extra import calc_sigma_extra_on_grid    # noqa: F401
from .data_slope_bias import data_slope_bias              # noqa: F401
from .grid_functions import setup_grids, sum_cell_area, calc_cell_area, setup_averaging_ops, setup_z0_avg, setup_avg_mask_ops, sym_range, \
    setup_mask, validate_by_dz_mask   # noqa: F401
from .lin_op import lin_op, fd_op, DeviceCSR              # noqa: F401
from .match_priors import make_prior_op, match_prior_dz, match_tile_edges   # noqa: F401
from .setup_grid_bias import setup_grid_bias              # noqa: F401
from .setup_sensor_grid_bias import setup_sensor_grid_bias, parse_sensor_bias_grids      # noqa: F401
from .smooth_fit import smooth_fit, iterate_fit           # noqa: F401
from .solve import solve                                  # noqa: F401
from .unique_by_rows import unique_by_rows                # noqa: F401


def matlab_to_year(t):
    """Approximate decimal year of a MATLAB serial date, with 730486 = 1 January 2000 (LSsurf/matlab_to_year.py:9-13)."""
    return 2000. + (t - 730486.) / 365.25


__version__ = '0.2.0'
