class Transformer2DModel:
    def __init__(self, *a, **k):
        raise RuntimeError("diffusers shim: Transformer2DModel is a placeholder (dead 3D/Motion blocks only)")
