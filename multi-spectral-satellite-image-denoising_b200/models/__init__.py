"""Mirror of the reference's ``src.models`` exports (src/models/__init__.py:7-18)."""
from .blocks import ConvBlock, ResidualBlock, UpsampleBlock
from .decoder import LightweightDecoder
from .encoder import SatMAEEncoder, convert_block_keys
from .export_graph import export_mode
from .satmae_restoration import SatMAERestoration

__all__ = ["SatMAERestoration", "SatMAEEncoder", "LightweightDecoder", "ConvBlock", "ResidualBlock", "UpsampleBlock",
           "convert_block_keys", "export_mode"]
