from .substituton import Substitution, LAYER, build_layer
from .simple_conv import SimpleConv
from .low_rank_conv import LowRankExpConvV1, LowRankExpConvV2, SeparableConv
from .depth_seperable_conv import ParallelConv, CascadeConv, FixPaddingBias
from .drop import DropPath
from .msca import MSCA, MSCAProfile
from .mscan_attention import MSCANAttention
from .mscan_block import MSCANFFN, MSCANBlock
from .dummy import DummyLayer
