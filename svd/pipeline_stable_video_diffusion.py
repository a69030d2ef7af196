"""StableVideoDiffusionPipeline (VL) — drop-in for svd/pipeline_stable_video_diffusion.py of the reference.

`__call__` keeps the reference's keyword arguments and defaults (:324-347); the denoising loop (:527-562) runs on
the fused sm_100a path (no ControlNet). With several input images each video is denoised as an independent CFG pair.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Union

import torch

from .pipeline_common import (PIL, StableVideoDiffusionPipelineOutput, SVDPipelineBase, _append_dims, randn_tensor)


class StableVideoDiffusionPipeline(SVDPipelineBase):
    @torch.no_grad()
    def __call__(
        self,
        image=None,
        prompt=None,
        use_text: bool = False,
        text_encoder=None,
        height: int = 576,
        width: int = 1024,
        num_frames: Optional[int] = None,
        num_inference_steps: int = 25,
        min_guidance_scale: float = 1.0,
        max_guidance_scale: float = 3.0,
        fps: int = 7,
        motion_bucket_id: int = 127,
        noise_aug_strength: float = 0.02,
        decode_chunk_size: Optional[int] = None,
        num_videos_per_prompt: Optional[int] = 1,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.FloatTensor] = None,
        output_type: Optional[str] = "pil",
        callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
        callback_on_step_end_tensor_inputs: List[str] = ["latents"],
        return_dict: bool = True,
        use_instructpix2pix: bool = False,
        image_guidance_scale: float = 7.5,
        # --- additions for latent / benchmark mode
        encoder_hidden_states: Optional[torch.Tensor] = None,
        image_latents: Optional[torch.Tensor] = None,
    ):
        if use_instructpix2pix:
            raise NotImplementedError("use_instructpix2pix=True (3-way CFG) is not implemented")
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        decode_chunk_size = decode_chunk_size if decode_chunk_size is not None else num_frames
        device = self._execution_device
        do_cfg = max_guidance_scale > 1.0

        if encoder_hidden_states is not None:
            if image_latents is None:
                raise ValueError("latent mode needs encoder_hidden_states and image_latents")
            if height % 8 != 0 or width % 8 != 0:
                raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
            ehs = encoder_hidden_states.to(device)
            batch_size = ehs.shape[0] // (2 if do_cfg else 1)
            img_lat = image_latents.to(device)
        else:
            self.check_inputs(image, height, width)
            if PIL is not None and isinstance(image, PIL.Image.Image):
                batch_size = 1
            elif isinstance(image, list):
                batch_size = len(image)
            else:
                batch_size = image.shape[0]
            ehs = self.encode_clip(image, prompt, use_text, text_encoder, device, num_videos_per_prompt, do_cfg)
            # the reference draws the augmentation noise where the preprocessed image lives (the host for PIL / CPU
            # tensors) and only then moves it: same RNG stream as the reference under torch.manual_seed / a CPU generator
            image_t = self._preprocess_image(image, height, width)
            noise = randn_tensor(image_t.shape, generator=generator, device=image_t.device, dtype=image_t.dtype)
            image_t = (image_t + noise_aug_strength * noise).to(device)
            needs_upcasting = (self.vae.dtype == torch.float16 and self.vae.config.force_upcast
                               and not getattr(self.vae, "_ttvdm_native", False))
            if needs_upcasting:
                self.vae.to(dtype=torch.float32)
            img_lat = self._encode_vae_image(image_t, device, num_videos_per_prompt, do_cfg).to(ehs.dtype)
            if needs_upcasting:
                self.vae.to(dtype=torch.float16)

        fps = fps - 1
        added_time_ids = self._get_add_time_ids(fps, motion_bucket_id, noise_aug_strength, ehs.dtype, batch_size,
                                                num_videos_per_prompt, do_cfg).to(device)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = self.scheduler.timesteps
        latents = self.prepare_latents(batch_size * num_videos_per_prompt, num_frames, self.unet.config.in_channels,
                                       height, width, ehs.dtype, device, generator, latents)
        guidance_vec = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames)
        gs = guidance_vec.unsqueeze(0).to(device, latents.dtype).repeat(batch_size * num_videos_per_prompt, 1)
        self._guidance_scale = _append_dims(gs, latents.ndim)
        self._num_timesteps = len(timesteps)

        latents = self._denoise(latents, img_lat, ehs, added_time_ids, guidance_vec, num_frames, timesteps,
                                callback_on_step_end=callback_on_step_end,
                                callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs)

        if not output_type == "latent":
            frames = self.decode_latents(latents.to(self.vae.dtype), num_frames, decode_chunk_size)
            frames = self._tensor2vid(frames, output_type=output_type)
        else:
            frames = latents
        self.maybe_free_model_hooks()
        if not return_dict:
            return frames
        return StableVideoDiffusionPipelineOutput(frames=frames)
