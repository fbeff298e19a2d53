"""FNO / SFNO spectral-convolution layers backed by the pruned-FFT sm_100a kernels of libtcfd
(reference: fno/fno3d.py, fno/sfno.py, fno/base.py)."""
from .spectral_conv import SpectralConv3d, SpectralConvS, SpectralConvT, spectral_conv3d  # noqa: F401
from .fno3d import FNO3d, MLP  # noqa: F401
from .sfno import (SFNO, FNOBase, HelmholtzProjection, LayerNormnd, LiftingOperator, OutConv, PointwiseFFN,  # noqa: F401
                   SpaceTimePositionalEncoding)
