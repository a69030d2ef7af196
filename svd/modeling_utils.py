"""Minimal stand-ins for the diffusers ModelMixin / ConfigMixin surface the reference's callers touch
(test_code/inference.py:331-378, svd/pipeline_stable_video_diffusion_controlnet.py:236-241,493-496,560,588):
`.config.<field>`, `from_pretrained(path, subfolder=, ...)`, `save_pretrained`, `.dtype`, `.device`.
diffusers itself is not a dependency of this repository.
"""
from __future__ import annotations

import inspect
import json
import os
from types import SimpleNamespace
from typing import Any, Dict

import torch
from torch import nn


class FrozenConfig(SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)

    def to_dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)


def register_to_config(init):
    """Records the constructor arguments on `self.config` (diffusers' decorator of the same name)."""
    sig = inspect.signature(init)

    def wrapper(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        init(self, *args, **kwargs)
        cfg["_class_name"] = type(self).__name__
        self._internal_config = FrozenConfig(**cfg)

    wrapper.__wrapped__ = init
    return wrapper


class ModelBase(nn.Module):
    config_name = "config.json"
    weights_name = "diffusion_pytorch_model.safetensors"

    @property
    def config(self) -> FrozenConfig:
        return self._internal_config

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    # ---- persistence in the diffusers directory layout (<dir>/[subfolder]/{config.json, *.safetensors})
    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **_unused) -> None:
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.to_dict().items()}
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(cfg, f, indent=2)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, self.weights_name))
        else:
            torch.save(sd, os.path.join(save_directory, "diffusion_pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: str = None, torch_dtype=None,
                        variant: str = None, low_cpu_mem_usage: bool = True, **_unused):
        root = pretrained_model_name_or_path
        if subfolder:
            root = os.path.join(root, subfolder)
        cfg_path = os.path.join(root, cls.config_name)
        if not os.path.isfile(cfg_path):
            raise EnvironmentError(
                f"{cfg_path} not found: this build has no network access, pass a local diffusers-format directory")
        with open(cfg_path) as f:
            cfg = json.load(f)
        params = inspect.signature(cls.__init__.__wrapped__ if hasattr(cls.__init__, "__wrapped__")
                                   else cls.__init__).parameters
        kwargs = {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if k in params}
        model = cls(**kwargs)
        names = [cls.weights_name]
        if variant:
            names.insert(0, cls.weights_name.replace(".safetensors", f".{variant}.safetensors"))
        sd = None
        for n in names:
            p = os.path.join(root, n)
            if os.path.isfile(p):
                from safetensors.torch import load_file
                sd = load_file(p)
                break
        if sd is None:
            p = os.path.join(root, "diffusion_pytorch_model.bin")
            if os.path.isfile(p):
                sd = torch.load(p, map_location="cpu")
        if sd is None:
            raise EnvironmentError(f"no weights found under {root}")
        model.load_state_dict(sd, strict=True)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        model.eval()
        return model

    # ---- diffusers API kept as no-ops: the CUDA engine has one attention / feed-forward implementation
    def enable_gradient_checkpointing(self):
        raise NotImplementedError("training (backward) is out of scope for the sm_100a inference engine")

    def set_attn_processor(self, processor) -> None:
        return None

    def set_default_attn_processor(self) -> None:
        return None

    def enable_forward_chunking(self, chunk_size=None, dim: int = 0) -> None:
        if dim not in [0, 1]:
            raise ValueError(f"Make sure to set `dim` to either 0 or 1, not {dim}")

    def enable_xformers_memory_efficient_attention(self, *a, **k) -> None:
        return None
