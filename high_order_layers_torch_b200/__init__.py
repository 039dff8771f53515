"""B200-native (sm_100a) piecewise Lagrange-polynomial layers.

Drop-in for the hot path of jloveric/high-order-layers-torch: the same constructors
(`high_order_fc_layers`, `HighOrderMLP`, `high_order_convolution_layers`, `PiecewisePolynomial`,
`PiecewiseDiscontinuousPolynomial`, `PiecewisePolynomialConvolution2d`), parameter names and
shapes; forward/backward run hand-written CUDA kernels through a C-ABI library
(`csrc/libhol_b200.so`, `include/hol_b200.h`).  CUDA only -- there is no CPU fallback.
"""
from . import layers  # noqa: F401
from .layers import (  # noqa: F401
    L2Normalization,
    MaxAbsNormalization,
    MaxAbsNormalizationLast,
    MaxAbsNormalizationND,
    MaxCenterNormalization,
    MaxCenterNormalizationND,
    normalization_layers,
    convolutional_layers,
    convolutional_transpose_layers,
    fc_layers,
    high_order_convolution_layers,
    high_order_convolution_transpose_layers,
    high_order_fc_layers,
)
from .FunctionalConvolutionTranspose import PiecewisePolynomialConvolutionTranspose2d  # noqa: F401
from .PolynomialLayers import PiecewiseDiscontinuousPolynomial, PiecewisePolynomial, Polynomial  # noqa: F401
from .FunctionalConvolution import (  # noqa: F401
    PiecewiseDiscontinuousPolynomialConvolution1d,
    PiecewiseDiscontinuousPolynomialConvolution2d,
    PiecewisePolynomialConvolution1d,
    PiecewisePolynomialConvolution2d,
)
from .networks import HighOrderMLP  # noqa: F401
from .refinement import (  # noqa: F401
    hp_refine_high_order_mlp,
    initialize_network_polynomial_layers,
    initialize_polynomial_layer,
    interpolate_high_order_mlp,
    interpolate_polynomial_layer,
    refine_discontinuous_polynomial_layer,
    refine_polynomial_layer,
    smooth_discontinuous_layer,
    smooth_discontinuous_network,
)

__version__ = "0.1.0"
