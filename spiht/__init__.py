"""`spiht`: the reference's package name (spiht/__init__.py:1-2), backed by spiht_b200.

    from spiht import encode_image, decode_image, EncodingResult, SpihtSettings, encode, decode
    from spiht.spiht_wrapper import SpihtSettings, get_slices_and_h_w
    from spiht.utils import imload

so that the reference's scripts (encode_decode.py, demonstrate.py, make_gif.py) and tests import unchanged.
Every submodule is the spiht_b200 module of the same name (one module object, not a copy).
"""
import sys as _sys

import spiht_b200 as _impl  # noqa: F401
from spiht_b200 import color_models, spiht, spiht_wrapper, utils  # noqa: F401
from spiht_b200.spiht_wrapper import *  # noqa: F401,F403  (the reference's `from .spiht_wrapper import *`)
from spiht_b200.spiht_wrapper import (ENCODER_DECODER_VERSION, EncodingResult, SpihtSettings, decode_image,  # noqa: F401
                                      decode_images, encode_image, encode_images)
from spiht_b200.spiht import decode, decode_with_metadata, encode  # noqa: F401

for _name in ("color_models", "spiht", "spiht_wrapper", "utils"):
    _sys.modules[__name__ + "." + _name] = _sys.modules["spiht_b200." + _name]
del _name
