"""Drop-in replacement for the reference's `builders` package (same module and class names)."""
from .build_network_from_config import NetworkFromConfig, get_activation_module  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .encoder import Encoder  # noqa: F401
from .resblocks import BasicBlockD, BottleneckD, DropPath, SqueezeExcite, StackedResidualBlocks  # noqa: F401
from .simple_conv_blocks import ConvDropoutNormReLU, StackedConvBlocks  # noqa: F401
from .utils import (get_matching_convtransp, get_matching_instancenorm, get_matching_pool_op,  # noqa: F401
                    get_n_blocks_per_stage, get_pool_and_conv_props, maybe_convert_scalar_to_list, pad_shape)
