"""Drop-ins for the two `transformers` modules the reference's conditioning builder calls
(encode_clip, svd/pipeline_stable_video_diffusion_controlnet.py:130-188):

  * CLIPVisionModelWithProjection — `self.image_encoder(image).image_embeds` (:155); loaded by the reference at
    test_code/inference.py:325-327 (`from_pretrained(path, subfolder="image_encoder", revision=None, variant="fp16")`)
  * CLIPTextModel — `text_encoder(prompt)[0]` (:166) with `prompt` = token ids [B, 77]; loaded at :347-348

Same class names, `from_pretrained(path, subfolder=, variant=)` on the HF directory layout (`config.json` +
`model[.variant].safetensors`, HF parameter names), `.config`, `.dtype`, `.parameters()`, `.to()`, `.requires_grad_()`
and call results (`.image_embeds`, `[0]` / `.last_hidden_state`). The module tree only OWNS the parameters; the
arithmetic runs on the sm_100a kernels through this_and_that_vdm_b200.clip_engine. A live `transformers` module can be
wrapped with `from_hf(module)`. There is no CPU / eager implementation: calling a tower that is not on a CUDA sm_100
device raises.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch
from torch import nn


def _vision_shapes(c: dict) -> Dict[str, tuple]:
    C, ps = c["hidden_size"], c["patch_size"]
    n_pos = (c["image_size"] // ps) ** 2 + 1
    s = {"vision_model.embeddings.class_embedding": (C,),
         "vision_model.embeddings.patch_embedding.weight": (C, c.get("num_channels", 3), ps, ps),
         "vision_model.embeddings.position_embedding.weight": (n_pos, C),
         "vision_model.pre_layrnorm.weight": (C,), "vision_model.pre_layrnorm.bias": (C,),
         "vision_model.post_layernorm.weight": (C,), "vision_model.post_layernorm.bias": (C,),
         "visual_projection.weight": (c["projection_dim"], C)}
    s.update(_layer_shapes("vision_model", c))
    return s


def _text_shapes(c: dict) -> Dict[str, tuple]:
    C = c["hidden_size"]
    s = {"text_model.embeddings.token_embedding.weight": (c["vocab_size"], C),
         "text_model.embeddings.position_embedding.weight": (c["max_position_embeddings"], C),
         "text_model.final_layer_norm.weight": (C,), "text_model.final_layer_norm.bias": (C,)}
    s.update(_layer_shapes("text_model", c))
    return s


def _layer_shapes(prefix: str, c: dict) -> Dict[str, tuple]:
    C, inter = c["hidden_size"], c["intermediate_size"]
    s = {}
    for i in range(c["num_hidden_layers"]):
        p = f"{prefix}.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[f"{p}.self_attn.{n}.weight"], s[f"{p}.self_attn.{n}.bias"] = (C, C), (C,)
        for n in ("layer_norm1", "layer_norm2"):
            s[f"{p}.{n}.weight"], s[f"{p}.{n}.bias"] = (C,), (C,)
        s[f"{p}.mlp.fc1.weight"], s[f"{p}.mlp.fc1.bias"] = (inter, C), (inter,)
        s[f"{p}.mlp.fc2.weight"], s[f"{p}.mlp.fc2.bias"] = (C, inter), (C,)
    return s


class _Node(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard rail
        raise RuntimeError("parameter container of a CLIP tower; call the tower itself on a CUDA (sm_100) device")


class _TowerBase(nn.Module):
    _kind = ""
    _defaults: dict = {}
    # replay each tower as one CUDA graph per input shape (TTVDM_CLIP_GRAPH=0 launches kernel by kernel)
    use_cuda_graph = os.environ.get("TTVDM_CLIP_GRAPH", "1") != "0"

    def __init__(self, config=None, **kwargs):
        super().__init__()
        cfg = dict(self._defaults)
        if config is not None:
            cfg.update(config if isinstance(config, dict) else
                       (config.to_dict() if hasattr(config, "to_dict") else vars(config)))
        cfg.update(kwargs)
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        shapes = _vision_shapes(cfg) if self._kind == "vision" else _text_shapes(cfg)
        for key, shape in shapes.items():
            node = self
            parts = key.split(".")
            for part in parts[:-1]:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            node.register_parameter(parts[-1], nn.Parameter(torch.empty(*shape).normal_(0.0, 0.02), requires_grad=False))
        self._engine = None

    # ---- HF-style persistence (local directories only; there is no hub access)
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, revision=None,
                        variant: Optional[str] = None, torch_dtype=None, **_unused):
        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        cfg_path = os.path.join(root, "config.json")
        if not os.path.isfile(cfg_path):
            raise EnvironmentError(f"{cfg_path} not found: pass a local HF-format directory (no network access)")
        with open(cfg_path) as f:
            cfg = json.load(f)
        sub = "vision_config" if cls._kind == "vision" else "text_config"
        if sub in cfg and "hidden_size" not in cfg:  # a full CLIPConfig: take the tower's part
            proj = cfg.get("projection_dim")
            cfg = dict(cfg[sub])
            cfg.setdefault("projection_dim", proj)
        model = cls(cfg)
        names = ["model.safetensors"]
        if variant:
            names.insert(0, f"model.{variant}.safetensors")
        sd = None
        for n in names:
            if os.path.isfile(os.path.join(root, n)):
                from safetensors.torch import load_file
                sd = load_file(os.path.join(root, n))
                break
        if sd is None and os.path.isfile(os.path.join(root, "pytorch_model.bin")):
            sd = torch.load(os.path.join(root, "pytorch_model.bin"), map_location="cpu")
        if sd is None:
            raise EnvironmentError(f"no weights found under {root}")
        model.load_state_dict(sd)
        return model.to(torch_dtype) if torch_dtype is not None else model

    def save_pretrained(self, save_directory: str, **_unused) -> None:
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(self._cfg, f, indent=2)
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, "model.safetensors"))

    @classmethod
    def from_hf(cls, module):
        """Wrap a live `transformers` CLIP tower (copies its config and weights)."""
        model = cls(module.config)
        model.load_state_dict(module.state_dict())
        p = next(module.parameters())
        return model.to(device=p.device, dtype=p.dtype)

    def load_state_dict(self, state_dict, strict: bool = True, **k):
        self._engine = None
        # buffers of the HF modules (position_ids) are not parameters of the computation
        sd = {key: v for key, v in state_dict.items() if not key.endswith("position_ids")}
        return super().load_state_dict(sd, strict=strict, **k)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _get_engine(self):
        from this_and_that_vdm_b200.clip_engine import ClipTowerEngine
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} runs only on a CUDA sm_100 device (hand-written kernels in libttvdm_sm100.so); "
                "move it with .to('cuda') — there is no CPU / eager fallback")
        if self._engine is None or self._engine.device != dev:
            self._engine = ClipTowerEngine(self.state_dict(), self._cfg, self._kind, dev)
        return self._engine


class _Output(dict):
    """Attribute + integer indexing like transformers' ModelOutput."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return super().__getitem__(k)


class CLIPVisionModelWithProjection(_TowerBase):
    _kind = "vision"
    _defaults = dict(hidden_size=768, intermediate_size=3072, projection_dim=512, num_hidden_layers=12,
                     num_attention_heads=12, num_channels=3, image_size=224, patch_size=32, hidden_act="quick_gelu",
                     layer_norm_eps=1e-5)

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor = None, **_unused):
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        emb = self._get_engine().image_embeds(pixel_values, use_graph=self.use_cuda_graph)
        return _Output(image_embeds=emb.to(pixel_values.dtype if pixel_values.is_floating_point() else self.dtype))


class CLIPTextModel(_TowerBase):
    _kind = "text"
    _defaults = dict(vocab_size=49408, hidden_size=512, intermediate_size=2048, num_hidden_layers=12,
                     num_attention_heads=8, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor = None, **_unused):
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        hs = self._get_engine().last_hidden_state(input_ids, use_graph=self.use_cuda_graph)
        return _Output(last_hidden_state=hs.to(self.dtype))
