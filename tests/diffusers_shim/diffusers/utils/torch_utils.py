def apply_freeu(*a, **k):  # only reached by the dead 3D / Motion blocks
    raise RuntimeError("diffusers shim: apply_freeu is a placeholder")


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: CPU generator -> draw on CPU then move."""
    import torch
    rand_device = device
    if generator is not None:
        gdev = generator.device.type if not isinstance(generator, list) else generator[0].device.type
        if gdev != (device.type if isinstance(device, torch.device) else str(device)) and gdev == "cpu":
            rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
