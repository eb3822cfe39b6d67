"""spiht/color_models.py of the reference.  Only RGB <-> IPT is supported: it is
the one colour model on the accelerated path, where it is fused with the
transform kernels (spihtb_forward / spihtb_inverse).  `convert` keeps the
reference's signature and error behaviour for everything else.
"""
SUPPORTED_MODELS = {"RGB", "IPT"}


def normalise(name):
    """None, or the canonical name of a supported model (case-insensitive, as colour.convert is)."""
    if name is None:
        return None
    up = str(name).upper()
    if up not in SUPPORTED_MODELS:
        raise ValueError(f'{name} is not a supported color model. Supported models are {SUPPORTED_MODELS}')
    return up
