"""ORACLE SCAFFOLDING.  resblocks.py:10 imports ConvDropoutNormReLU from DNA; the
reference's vendored twin (builders/simple_conv_blocks.py:13) has identical attribute
names, so state_dict keys match."""
from builders.simple_conv_blocks import ConvDropoutNormReLU  # noqa: F401
