class DualTransformer2DModel:
    def __init__(self, *a, **k):
        raise RuntimeError("diffusers shim: DualTransformer2DModel is a placeholder (dead 3D/Motion blocks only)")
