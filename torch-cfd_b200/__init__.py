"""torch-cfd_b200: B200-native (sm_100a) implementation of torch-cfd's spectral hot path."""
__version__ = "0.1.0"

from .grids import Grid  # noqa: F401
from .forcings import ForcingFn, KolmogorovForcing  # noqa: F401
from .spectral import (brick_wall_filter_2d, fft_mesh_2d, spectral_curl_2d, spectral_div_2d,  # noqa: F401
                       spectral_grad_2d, spectral_laplacian_2d, spectral_rot_2d, vorticity_to_velocity)
from .equations import (IMEXStepper, ImplicitExplicitODE, NavierStokes2DSpectral,  # noqa: F401
                        RK4CrankNicolsonStepper, stable_time_step)
from .solvers import (get_trajectory_imex, get_trajectory_imex_crank_nicolson, get_trajectory_imex_sharded,  # noqa: F401
                      imex_crank_nicolson_step, postprocess_trajectory, update_residual)
from . import fft  # noqa: F401
from . import fno  # noqa: F401
from . import initial_conditions  # noqa: F401
from .initial_conditions import GRF2d, vorticity_field  # noqa: F401
