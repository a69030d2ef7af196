"""Shared machinery of the two drop-in pipelines (VL: svd/pipeline_stable_video_diffusion.py, VGL:
svd/pipeline_stable_video_diffusion_controlnet.py of the reference): component registry, the pre/post-processing
steps that live OUTSIDE the hot path (they call the caller-supplied CLIP / VAE modules exactly like the reference),
and the denoising loop, which runs on this_and_that_vdm_b200.sampler.FusedDenoiser.

Benchmark / latent mode: with `vae=None` and `image_encoder=None` the caller passes precomputed conditioning through
the extra keyword arguments `encoder_hidden_states`, `image_latents` and `controlnet_cond_latents` together with
`output_type="latent"`. These are the only additions to the reference's `__call__` signatures.
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
from torch import nn

from .scheduler import EulerDiscreteScheduler

try:  # PIL is only needed for pil inputs / outputs
    import PIL.Image
except Exception:  # pragma: no cover
    PIL = None


@dataclass
class StableVideoDiffusionPipelineOutput:
    frames: Union[List[List["PIL.Image.Image"]], np.ndarray, torch.Tensor]


def _append_dims(x, target_dims):
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: sample on the generator's device, then move."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if isinstance(generator, list):
        shape1 = (1,) + tuple(shape[1:])
        return torch.cat([randn_tensor(shape1, g, device, dtype) for g in generator], dim=0)
    gdev = generator.device if generator is not None else device
    return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)


def _gaussian_blur_resize(image: torch.Tensor, size, interpolation="bicubic", align_corners=True):
    """_resize_with_antialiasing of the reference pipelines (Gaussian pre-blur + bicubic), CLIP pre-processing."""
    import torch.nn.functional as F
    h, w = image.shape[-2:]
    factors = (h / size[0], w / size[1])
    sigmas = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
    ks = int(max(2.0 * 2.0 * sigmas[0], 3)), int(max(2.0 * 2.0 * sigmas[1], 3))
    ks = (ks[0] + 1 if ks[0] % 2 == 0 else ks[0], ks[1] + 1 if ks[1] % 2 == 0 else ks[1])

    def k1d(n, sigma):
        x = torch.arange(n, dtype=image.dtype, device=image.device) - n // 2
        if n % 2 == 0:
            x = x + 0.5
        g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
        return g / g.sum()

    ky, kx = k1d(ks[0], sigmas[0]), k1d(ks[1], sigmas[1])
    c = image.shape[1]
    pad = [ks[1] // 2, ks[1] // 2, ks[0] // 2, ks[0] // 2]
    x = F.pad(image, pad, mode="reflect")
    x = F.conv2d(x, kx.view(1, 1, 1, -1).expand(c, 1, 1, -1), groups=c)
    x = F.conv2d(x, ky.view(1, 1, -1, 1).expand(c, 1, -1, 1), groups=c)
    return F.interpolate(x, size=size, mode=interpolation, align_corners=align_corners)


class SVDPipelineBase:
    model_cpu_offload_seq = "image_encoder->unet->vae"
    _callback_tensor_inputs = ["latents"]

    def __init__(self, vae=None, image_encoder=None, unet=None, scheduler=None, feature_extractor=None):
        self.vae, self.image_encoder, self.unet = vae, image_encoder, unet
        self.scheduler = scheduler if scheduler is not None else EulerDiscreteScheduler()
        self.feature_extractor = feature_extractor
        if vae is not None:
            self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1)
        else:
            self.vae_scale_factor = 8
        self._progress = {}
        self._device = None
        self._guidance_scale = None
        self._num_timesteps = None
        self._denoiser = None

    # ---- diffusers DiffusionPipeline surface used by test_code/inference.py:171-180 and app.py
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, vae=None, image_encoder=None, unet=None,
                        scheduler=None, feature_extractor=None, revision=None, torch_dtype=None, **_unused):
        """Components are passed in by the callers (as the reference's callers do for vae / image_encoder / unet);
        the scheduler defaults to the SVD EulerDiscreteScheduler config. No network access."""
        if unet is None:
            raise EnvironmentError("pass `unet=` (this build has no network access to download components)")
        return cls(vae=vae, image_encoder=image_encoder, unet=unet, scheduler=scheduler,
                   feature_extractor=feature_extractor)

    def to(self, device=None, dtype=None):
        for m in (self.vae, self.image_encoder, self.unet):
            if isinstance(m, nn.Module):
                m.to(device=device) if dtype is None else m.to(device=device, dtype=dtype)
        self._device = torch.device(device) if device is not None else self._device
        return self

    def set_progress_bar_config(self, **kwargs):
        self._progress = kwargs

    def maybe_free_model_hooks(self):
        return None

    @property
    def _execution_device(self):
        if self._device is not None:
            return self._device
        return self.unet.device

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    # ---- pre / post processing (NOT the hot path; mirrors the reference call for call)
    def check_inputs(self, image, height, width):
        ok = isinstance(image, (torch.Tensor, list)) or (PIL is not None and isinstance(image, PIL.Image.Image))
        if not ok:
            raise ValueError(
                "`image` has to be of type `torch.FloatTensor` or `PIL.Image.Image` or `List[PIL.Image.Image]` but is"
                f" {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    @staticmethod
    def _pil_to_pt(image) -> torch.Tensor:
        imgs = image if isinstance(image, list) else [image]
        arr = np.stack([np.array(i.convert("RGB")).astype(np.float32) / 255.0 for i in imgs], axis=0)
        return torch.from_numpy(arr.transpose(0, 3, 1, 2))

    def _preprocess_image(self, image, height, width) -> torch.Tensor:
        """diffusers VaeImageProcessor.preprocess(image, height, width) as the reference calls it
        (svd/pipeline_stable_video_diffusion_controlnet.py:541): PIL images are converted to RGB, Lanczos-resized and
        mapped [0,1] -> [-1,1]; tensors (or lists of tensors) are concatenated / stacked, returned untouched when they
        are 4-channel latents, otherwise resized with F.interpolate (nearest) and mapped [0,1] -> [-1,1] unless they
        already contain negative values (diffusers warns and skips the normalisation in that case)."""
        if isinstance(image, torch.Tensor):
            image = [image]
        if isinstance(image, list) and len(image) > 0 and isinstance(image[0], torch.Tensor):
            t = torch.cat(image, dim=0) if image[0].ndim == 4 else torch.stack(image, dim=0)
            if t.shape[1] == 4:
                return t
            if tuple(t.shape[-2:]) != (height, width):
                t = torch.nn.functional.interpolate(t, size=(height, width))
            if t.min() < 0:
                return t
            return 2.0 * t - 1.0
        imgs = image if isinstance(image, list) else [image]
        imgs = [i.convert("RGB").resize((width, height), resample=PIL.Image.LANCZOS) for i in imgs]
        return 2.0 * self._pil_to_pt(imgs) - 1.0

    def encode_clip(self, image, prompt, use_text, text_encoder, device, num_videos_per_prompt,
                    do_classifier_free_guidance, use_instructpix2pix=False):
        dtype = next(self.image_encoder.parameters()).dtype
        if not isinstance(image, torch.Tensor):
            image = self._pil_to_pt(image) * 2.0 - 1.0
            image = _gaussian_blur_resize(image, (224, 224))
            image = (image + 1.0) / 2.0
            if self.feature_extractor is not None:
                image = self.feature_extractor(images=image, do_normalize=True, do_center_crop=False, do_resize=False,
                                               do_rescale=False, return_tensors="pt").pixel_values
            else:  # CLIPImageProcessor defaults (the SVD repo's feature_extractor/preprocessor_config.json)
                mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
                std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
                image = (image - mean) / std
        image = image.to(device=device, dtype=dtype)
        emb = self.image_encoder(image).image_embeds
        text = text_encoder(prompt)[0] if use_text else None
        # tail of encode_clip (:156-186): [text | image] concat, fresh LayerNorm((78, 1024)), CFG zero stack — on the
        # sm_100a kernels (ttvdm_layernorm_flat), fp32 result cast to the towers' dtype like the reference's
        from this_and_that_vdm_b200.clip_engine import assemble_conditioning
        return assemble_conditioning(emb, text, do_classifier_free_guidance, num_videos_per_prompt,
                                     use_instructpix2pix).to(dtype)

    def _encode_vae_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance,
                          use_instructpix2pix=False):
        image = image.to(device=device)
        lat = self.vae.encode(image).latent_dist.mode()
        if do_classifier_free_guidance:
            neg = torch.zeros_like(lat)
            lat = torch.cat([lat, lat, neg]) if use_instructpix2pix else torch.cat([neg, lat])
        return lat.repeat(num_videos_per_prompt, 1, 1, 1)

    def _get_add_time_ids(self, fps, motion_bucket_id, noise_aug_strength, dtype, batch_size, num_videos_per_prompt,
                          do_classifier_free_guidance, guess_mode=False, use_instructpix2pix=False):
        ids = [fps, motion_bucket_id, noise_aug_strength]
        passed = self.unet.config.addition_time_embed_dim * len(ids)
        expected = self.unet.add_embedding.linear_1.in_features
        if expected != passed:
            raise ValueError(
                f"Model expects an added time embedding vector of length {expected}, but a vector of {passed} was "
                "created. The model has an incorrect config. Please check `unet.config.time_embedding_type` and "
                "`text_encoder_2.config.projection_dim`.")
        t = torch.tensor([ids], dtype=dtype).repeat(batch_size * num_videos_per_prompt, 1)
        if do_classifier_free_guidance:
            t = torch.cat([t, t, t]) if use_instructpix2pix else torch.cat([t, t])
        return t

    def decode_latents(self, latents, num_frames, decode_chunk_size=14):
        latents = latents.flatten(0, 1)
        latents = 1 / self.vae.config.scaling_factor * latents
        accepts = "num_frames" in set(inspect.signature(self.vae.forward).parameters.keys())
        frames = []
        for i in range(0, latents.shape[0], decode_chunk_size):
            chunk = latents[i:i + decode_chunk_size]
            kw = {"num_frames": chunk.shape[0]} if accepts else {}
            frames.append(self.vae.decode(chunk, **kw).sample)
        frames = torch.cat(frames, dim=0)
        return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_frames, num_channels_latents // 2, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(
                f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def prepare_condition_image(self, condition_img, width, height, batch_size, num_videos_per_prompt, device, dtype,
                                do_classifier_free_guidance=False, guess_mode=False):
        t = torch.from_numpy(condition_img) if isinstance(condition_img, np.ndarray) else condition_img
        return t.to(torch.float16).to(device)

    @staticmethod
    def _tensor2vid(video: torch.Tensor, output_type="pil"):
        """tensor2vid of the reference (:53-65): [B, C, F, H, W] in [-1, 1] -> per batch element
        VaeImageProcessor.postprocess(frames [F, C, H, W], output_type): "pt" tensor in [0, 1], "np" [F, H, W, C],
        "pil" list of images."""
        outs = []
        for vid in video:
            frames = (vid.permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
            if output_type == "pt":
                outs.append(frames)
                continue
            arr = frames.cpu().permute(0, 2, 3, 1).float().numpy()
            if output_type == "pil":
                arr = [PIL.Image.fromarray((f * 255).round().astype("uint8")) for f in arr]
            outs.append(arr)
        return outs

    # ---- the hot loop
    def _denoise(self, latents, image_latents_b, encoder_hidden_states, added_time_ids, guidance_vec, num_frames,
                 timesteps, controlnet=None, controlnet_cond=None, cond_scales=None, callback_on_step_end=None,
                 callback_on_step_end_tensor_inputs=("latents",)):
        """latents [N, F, 4, h, w] (already x init_noise_sigma). Each video n is an independent CFG pair
        (encoder_hidden_states / image_latents rows [n] = uncond half, [N + n] = cond half).

        Two things to know for N > 1 (the reference's callers always run N = 1; the VGL pipeline rejects N > 1 like the
        reference does):
          * the reference runs ONE batch of 2N sequences, so its temporal cross-attention quirk
            (svd/diffusion_arch/transformer_temporal.py:310-319) indexes the context by (b*S + s) mod 2N and mixes the
            contexts of DIFFERENT videos; here every video is its own pair (mod 2), i.e. the N = 1 behaviour of the
            reference for each video — results differ from the reference's batched call for N > 1;
          * with num_videos_per_prompt > 1 the reference lays image_latents out as repeat([neg, lat]) = [neg, lat, neg, lat]
            and still reads rows [n] / [N + n]; _encode_vae_image reproduces that layout, so the same rows are paired."""
        from this_and_that_vdm_b200.sampler import FusedDenoiser
        dev = latents.device
        N = latents.shape[0]
        do_cfg = encoder_hidden_states.shape[0] == 2 * N
        h, w = latents.shape[-2:]
        cn_engine = controlnet._get_engine() if controlnet is not None else None
        den = FusedDenoiser.cached(self.unet._get_engine(), cn_engine)
        out_dtype = latents.dtype
        result = []
        for n in range(N):
            idx = [n, N + n] if do_cfg else [n]
            state = latents[n].to(torch.float32).contiguous().clone()
            den.prepare(encoder_hidden_states[idx], image_latents_b[idx], added_time_ids[idx], self.scheduler.sigmas,
                        timesteps, guidance_vec, num_frames=num_frames, height=h, width=w,
                        controlnet_cond=controlnet_cond, conditioning_scale=1.0)
            for i, t in enumerate(timesteps):
                if cond_scales is not None:
                    den.cond_scale = float(cond_scales[i])
                eps = den.predict(i, state)
                rows = num_frames * h * w
                den.euler_update(i, state, eps[:rows], eps[rows:] if do_cfg else eps[:rows])
                if callback_on_step_end is not None:
                    cb = {"latents": state[None].to(out_dtype)}
                    outp = callback_on_step_end(self, i, t, {k: cb[k] for k in callback_on_step_end_tensor_inputs})
                    new = outp.pop("latents", None) if isinstance(outp, dict) else None
                    if new is not None:
                        state = new[0].to(torch.float32).contiguous().clone()
            result.append(state.to(out_dtype))
        self.scheduler._step_index = len(timesteps)
        return torch.stack(result, 0)
