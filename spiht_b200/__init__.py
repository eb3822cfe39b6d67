"""spiht_b200: a B200-native (sm_100a CUDA) SPIHT image codec behind the Python
API of theAdamColton/spiht (spiht/__init__.py:1-2 exports the same names).

    import spiht_b200 as spiht
    enc = spiht.encode_image(img, spiht.SpihtSettings(), max_bits=...)
    rec = spiht.decode_image(enc, spiht.SpihtSettings())
"""
from .spiht_wrapper import (encode_image, decode_image, encode_images, decode_images, decode_image_prefixes,
                            EncodingResult, SpihtSettings, ENCODER_DECODER_VERSION)
from .spiht import encode, decode

__all__ = ["encode_image", "decode_image", "encode_images", "decode_images", "decode_image_prefixes", "EncodingResult",
           "SpihtSettings",
           "ENCODER_DECODER_VERSION", "encode", "decode"]
