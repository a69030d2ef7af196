"""Generates tests/golden/clip_golden.pt: outputs of the LIBRARY the reference calls for its conditioning —
`transformers` CLIPVisionModelWithProjection(...).image_embeds and CLIPTextModel(...)[0]
(svd/pipeline_stable_video_diffusion_controlnet.py:155, :166) — on seeded tiny configs, plus outputs of the reference's
own `_resize_with_antialiasing` (its source text, :741-845, is executed unmodified; the module itself cannot be imported
because diffusers is absent).

Run here (transformers 5.5 is in the image; /root/reference is mounted):  python -m tests.golden.make_clip_golden
Weights and inputs are regenerated from seeds by tests/common.py; only the expected outputs are stored."""
import ast
import json
from pathlib import Path

import torch

from tests.common import (TINY_CLIP_TEXT, TINY_CLIP_VISION, TINY_CLIP_VISION_D80, clip_inputs, clip_text_sd,
                          clip_vision_sd)

HERE = Path(__file__).parent
REF = Path("/root/reference/svd/pipeline_stable_video_diffusion_controlnet.py")


def reference_resize():
    src = REF.read_text()
    tree = ast.parse(src)
    want = {"_resize_with_antialiasing", "_compute_padding", "_filter2d", "_gaussian", "_gaussian_blur2d"}
    code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body
                       if isinstance(n, ast.FunctionDef) and n.name in want)
    ns = {"torch": torch}
    exec(compile(code, str(REF), "exec"), ns)
    return ns["_resize_with_antialiasing"]


def compute():
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPVisionConfig, CLIPVisionModelWithProjection
    out = {}
    for name, cfg in (("vision", TINY_CLIP_VISION), ("vision_d80", TINY_CLIP_VISION_D80)):
        m = CLIPVisionModelWithProjection(CLIPVisionConfig(**cfg)).eval()
        m.load_state_dict(clip_vision_sd(cfg), strict=True)
        px, _ = clip_inputs(cfg, TINY_CLIP_TEXT, n=2)
        with torch.no_grad():
            out[name] = m(px).image_embeds
    t = CLIPTextModel(CLIPTextConfig(**TINY_CLIP_TEXT)).eval()
    missing = t.load_state_dict(clip_text_sd(TINY_CLIP_TEXT), strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    _, ids = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT, n=2)
    with torch.no_grad():
        out["text"] = t(ids)[0]
    resize = reference_resize()
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 3, 96, 160, generator=g) * 2 - 1
    out["resize_96x160_to_56"] = resize(img, (56, 56))
    out["resize_40x40_to_56"] = resize(img[..., :40, :40], (56, 56))
    return out


if __name__ == "__main__":
    out = compute()
    torch.save({k: v.to(torch.float32) for k, v in out.items()}, HERE / "clip_golden.pt")
    import transformers
    (HERE / "clip_golden.json").write_text(json.dumps(
        {"tensors": {k: list(v.shape) for k, v in out.items()}, "transformers": transformers.__version__,
         "generator": "tests/golden/make_clip_golden.py (transformers CLIP modules; reference _resize_with_antialiasing "
                      "source executed unmodified)", "seeds": {"vision": 77, "text": 78, "inputs": 9, "resize": 3}},
        indent=1))
    print({k: tuple(v.shape) for k, v in out.items()})
