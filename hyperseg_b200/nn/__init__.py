"""Host-side mirror of the reference's nn.Module surface (hyperseg/models/*)."""
from .meta_sequential import MetaSequential
from .meta_conv import MetaConv2d, make_meta_conv2d_block
from .meta_patch import MetaPatch, MetaPatchConv2d, make_meta_patch_conv2d_block
