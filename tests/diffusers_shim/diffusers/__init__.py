"""Stand-in for diffusers==0.25.1 (test infrastructure, see ../README.md)."""
__version__ = "0.25.1+shim"


class AutoencoderKLTemporalDecoder:  # imported (never used) by svd/temporal_controlnet.py:25
    def __init__(self, *a, **k):
        raise RuntimeError("diffusers shim: AutoencoderKLTemporalDecoder is a placeholder")
