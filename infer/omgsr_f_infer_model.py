"""OMGSR-F one-step SR inference module — B200-native drop-in for the reference's
infer/omgsr_f_infer_model.py (same module-level helpers, class name, constructor and method signatures).

Dataflow (reference lines in brackets): FLUX-VAE encode, posterior sample, (z - shift) * scale [:15-18,:325] ->
2x2 pack to tokens [:21-26] -> one FLUX DiT forward at sigma = schedule[-(244+1)] = 0.5051 [:191-200] -> Euler
step to sigma 0 [:202] -> unpack, un-scale [:204-210] -> VAE decode [:211]; latents above tile_size^2 use the
overlapping latent-tile path [:214-320].
"""
import math
import os
import time
from typing import Callable

import torch

from omgsr_b200 import capi
from omgsr_b200 import dist as D
from omgsr_b200 import weights as W
from .vaehook import VAEHook


def encode_images(pixels: torch.Tensor, vae: torch.nn.Module, weight_dtype):
    fused = getattr(vae, "encode_scaled", None)
    if fused is not None:  # B200AutoencoderKL: shift / scale applied inside the posterior-sample kernel
        return fused(pixels.to(vae.dtype)).to(weight_dtype)
    pixel_latents = vae.encode(pixels.to(vae.dtype)).latent_dist.sample()
    pixel_latents = (pixel_latents - vae.config.shift_factor) * vae.config.scaling_factor
    return pixel_latents.to(weight_dtype)


def _pack_latents(latents, batch_size, num_channels_latents, height, width):
    """[B,C,h,w] -> [B,(h/2)(w/2),4C], channel order (c, dy, dx). CUDA tensors use the pack kernel."""
    if latents.is_cuda and latents.dtype == torch.bfloat16:
        return capi.pack_latents(latents)
    x = latents.reshape(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
    return x.permute(0, 2, 4, 1, 3, 5).reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)


def _unpack_latents(latents, height, width, vae_scale_factor):
    """inverse of _pack_latents; height/width are PIXEL sizes (latent size = 2 * (pixels // (2 * vae_scale_factor)))."""
    batch_size, _, channels = latents.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    if latents.is_cuda and latents.dtype == torch.bfloat16:
        return capi.unpack_latents(latents, channels // 4, h, w)
    x = latents.reshape(batch_size, h // 2, w // 2, channels // 4, 2, 2)
    return x.permute(0, 3, 1, 4, 2, 5).reshape(batch_size, channels // 4, h, w)


def time_shift(mu: float, sigma: float, t: torch.Tensor):
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_lin_function(x1: float = 256, y1: float = 0.5, x2: float = 4096, y2: float = 1.15) -> Callable[[float], float]:
    slope = (y2 - y1) / (x2 - x1)
    icpt = y1 - slope * x1
    return lambda x: slope * x + icpt


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                 shift: bool = True) -> list:
    # num_steps + 1 points from 1 to 0, resolution-dependent shift towards high sigma
    ts = torch.linspace(1, 0, num_steps + 1)
    if shift:
        mu = get_lin_function(y1=base_shift, y2=max_shift)(image_seq_len)
        ts = time_shift(mu, 1.0, ts)
    return ts.tolist()


def get_flux_setting_timesteps(n=999):
    return get_schedule(n, (1024 // 8) * (1024 // 8) // 4, shift=True)


def _parse_spec(spec):
    body = spec[len("random://"):]
    name, _, q = body.partition("?")
    kv = dict(p.split("=") for p in q.split("&") if p)
    return name, {k: int(v) for k, v in kv.items()}


def build_f_state(flux_path, lora_path, weight_dtype=torch.bfloat16, device="cpu"):
    """-> (vae_sd, flux_sd, vae_cfg, flux_cfg) with the LoRA deltas merged. The reference casts the base
    weights to bf16 BEFORE merging for OMGSR-F (infer/omgsr_f_infer_model.py:111-118); `weight_dtype` follows it."""
    if isinstance(flux_path, dict):
        st = flux_path
        return st["vae"], st["flux"], st["vae_cfg"], st["flux_cfg"]
    if flux_path.startswith("random://"):
        name, kv = _parse_spec(flux_path)
        seed = kv.get("seed", 0)
        vae_cfg, flux_cfg = (W.FLUX_VAE, W.FLUX_DEV) if name == "flux_dev" else (W.TINY_FLUX_VAE, W.TINY_FLUX)
        vae_sd = W.random_state_dict(W.vae_shapes(vae_cfg), seed)
        flux_sd = W.random_state_dict(W.flux_shapes(flux_cfg), seed + 1, dtype=weight_dtype, device=device)
    else:
        from omgsr_b200.checkpoint import load_flux_checkpoint
        vae_sd, flux_sd, vae_cfg, flux_cfg = load_flux_checkpoint(flux_path)
    if lora_path:
        if lora_path.startswith("random://"):
            _, kv = _parse_spec(lora_path)
            seed = kv.get("seed", 1)
            rank = kv.get("rank", 64)  # configs/omgsr_f_1024.yml:8-9
            vae_ad = W.random_lora(W.vae_shapes(vae_cfg), W.LORA_TARGETS_SD, rank, seed, prefix="encoder.")
            flux_ad = W.random_lora(W.flux_shapes(flux_cfg), W.LORA_TARGETS_FLUX, rank, seed + 1, device=device)
        else:
            from omgsr_b200.checkpoint import load_peft_adapter
            vae_ad = load_peft_adapter(os.path.join(lora_path, "vae_encoder_adapter"), prefix="encoder.")
            flux_ad = load_peft_adapter(os.path.join(lora_path, "flux_adapter"))
        vae_sd = W.merge_lora(vae_sd, vae_ad)
        flux_sd = W.merge_lora(flux_sd, flux_ad)
    return vae_sd, flux_sd, vae_cfg, flux_cfg


class OMGSR_F_Infer(torch.nn.Module):
    def __init__(self, flux_path, lora_path, device, weight_dtype=torch.bfloat16, mid_timestep=244, guidance_scale=1.0):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("OMGSR_F_Infer (omgsr_b200) needs a CUDA device: the hot path has no CPU fallback")
        from omgsr_b200.engine_flux import B200FluxTransformer
        from omgsr_b200.engine_vae import B200AutoencoderKL
        from .omgsr_s_infer_model import _select_device
        _select_device(device)
        print("Loading adapers...")
        vae_sd, flux_sd, vae_cfg, flux_cfg = build_f_state(flux_path, lora_path, weight_dtype, device)
        self.vae_scale_factor = 2 ** (len(vae_cfg.block_out_channels) - 1)
        self.guidance_scale = guidance_scale
        self.mid_timestep = mid_timestep
        self.weight_dtype = weight_dtype
        flux_timesteps = get_flux_setting_timesteps()
        self.t_curr = flux_timesteps[-(self.mid_timestep + 1)]
        self.t_prev = flux_timesteps[-1]  # 0.0
        print(f"Current One mid-timestep settings: {mid_timestep}")
        self.vae = B200AutoencoderKL(vae_sd, vae_cfg, device=device, dtype=weight_dtype)
        self.flux_transformer = B200FluxTransformer(flux_sd, flux_cfg, device=device, dtype=weight_dtype)
        self.vae.eval()
        self.flux_transformer.eval()
        self.device = device
        self._latent_channels = vae_cfg.latent_channels
        self._wts = {}
        # self._init_tiled_vae(encoder_tile_size=1024, decoder_tile_size=224)   (off, as in the reference :133)

    def _init_tiled_vae(
            self,
            encoder_tile_size=256,
            decoder_tile_size=256,
            fast_decoder=False,
            fast_encoder=False,
            color_fix=False,
            vae_to_gpu=True):
        if not hasattr(self.vae.encoder, 'original_forward'):
            setattr(self.vae.encoder, 'original_forward', self.vae.encoder.forward)
        if not hasattr(self.vae.decoder, 'original_forward'):
            setattr(self.vae.decoder, 'original_forward', self.vae.decoder.forward)
        self.vae.encoder.forward = VAEHook(
            self.vae.encoder, encoder_tile_size, is_decoder=False, fast_decoder=fast_decoder,
            fast_encoder=fast_encoder, color_fix=color_fix, to_gpu=vae_to_gpu)
        self.vae.decoder.forward = VAEHook(
            self.vae.decoder, decoder_tile_size, is_decoder=True, fast_decoder=fast_decoder,
            fast_encoder=fast_encoder, color_fix=color_fix, to_gpu=vae_to_gpu)

    def _gaussian_weights(self, tile_width, tile_height, nbatches):
        """same asymmetric Gaussian as OMGSR-S, tiled over the 16 latent channels (reference :157-172)."""
        var = 0.01
        mid = (tile_width - 1) / 2
        xs = [math.exp(-(x - mid) * (x - mid) / (tile_width * tile_width) / (2 * var)) / math.sqrt(2 * math.pi * var)
              for x in range(tile_width)]
        mid = tile_height / 2
        ys = [math.exp(-(y - mid) * (y - mid) / (tile_height * tile_height) / (2 * var)) / math.sqrt(2 * math.pi * var)
              for y in range(tile_height)]
        wts = torch.outer(torch.tensor(ys, dtype=torch.float64), torch.tensor(xs, dtype=torch.float64))
        return torch.tile(wts.to(self.device), (nbatches, 16, 1, 1))

    def _dit(self, packed, prompt_embeds, pooled_prompt_embeds, text_ids, latent_image_ids, seq_parallel=None):
        bsz = packed.shape[0]
        guidance_vec = torch.full((bsz,), self.guidance_scale, dtype=torch.float32)  # host: constant per run
        return self.flux_transformer(
            hidden_states=packed,
            timestep=torch.tensor([self.t_curr]),  # One-Step (host scalar: no device sync to read it back)
            guidance=guidance_vec,
            pooled_projections=pooled_prompt_embeds,
            encoder_hidden_states=prompt_embeds,
            txt_ids=text_ids,
            img_ids=latent_image_ids,
            return_dict=False,
            **({"seq_parallel": seq_parallel} if seq_parallel is not None else {}),
        )[0]

    def _decode_latent(self, lq_latent_f32_or_bf16):
        z = lq_latent_f32_or_bf16.to(self.vae.dtype)
        return self.vae.decode(z, return_dict=False)[0]

    def _forward_no_tile(self, lq_latent, prompt_embeds, pooled_prompt_embeds, text_ids, latent_image_ids):
        bsz, c, h, w = lq_latent.shape
        packed = _pack_latents(lq_latent.contiguous(), bsz, c, h, w)
        model_pred = self._dit(packed, prompt_embeds, pooled_prompt_embeds, text_ids, latent_image_ids)
        # Euler step to sigma = 0 (:202), unpack (:204-209), un-scale (:210)
        packed = capi.latent_axpby(packed.contiguous(), model_pred.contiguous(), 1.0, self.t_prev - self.t_curr)
        lat = _unpack_latents(packed, h * self.vae_scale_factor, w * self.vae_scale_factor, self.vae_scale_factor)
        lat = capi.latent_axpby(lat, None, 1.0 / self.vae.config.scaling_factor, 0.0, self.vae.config.shift_factor)
        return self._decode_latent(lat)

    def _forward_tile(self, lq_latent, prompt_embeds, pooled_prompt_embeds, text_ids, latent_image_ids, tile_size,
                      tile_overlap):
        from .omgsr_s_infer_model import OMGSR_S_Infer
        B, C, h, w = lq_latent.shape
        tile_size = min(tile_size, min(h, w))
        key = (tile_size, str(lq_latent.device))
        if key not in self._wts:
            self._wts[key] = self._gaussian_weights(tile_size, tile_size, 1)[0, 0].float().contiguous()
        wts = self._wts[key]
        offs = OMGSR_S_Infer._tile_offsets(h, w, tile_size, tile_overlap)  # identical grid walk (:220-248)
        nt = len(offs)
        # the reference runs the DiT once per tile at batch 1 (:255-289); the tiles are independent and share
        # prompt and position ids, so they are stacked on the batch axis
        tiles = torch.stack([lq_latent[:, :, oy:oy + tile_size, ox:ox + tile_size] for ox, oy in offs], 0)
        tiles = tiles.reshape(nt * B, C, tile_size, tile_size).contiguous()
        pe, pp = prompt_embeds, pooled_prompt_embeds
        if pe.shape[0] not in (1, nt * B):
            pe, pp = pe.repeat(nt, 1, 1), pp.repeat(nt, 1)
        packed = _pack_latents(tiles, nt * B, C, tile_size, tile_size)
        if D.tile_parallel():  # single large image on several GPUs: each rank runs the DiT on its share of the tiles
            if pe.shape[0] != 1:
                raise RuntimeError("tile-parallel mode expects one shared prompt")
            # tiles that fill a wave go one per rank; the left-over ones (9 tiles on 8 GPUs: the ninth) are computed by
            # all ranks together through the engine's sequence-parallel forward instead of costing another wave
            import torch.distributed as tdist
            world, rank = tdist.get_world_size(), tdist.get_rank()
            n_tok = packed.shape[1] + pe.shape[1]
            fn_all = None
            if n_tok % world == 0 and (nt * B) % world:
                fn_all = lambda t: self._dit(t, pe, pp, text_ids, latent_image_ids, seq_parallel=(rank, world))
            pred = D.tile_parallel_forward(lambda t: self._dit(t, pe, pp, text_ids, latent_image_ids), packed,
                                           fn_all=fn_all)
        else:
            pred = self._dit(packed, pe, pp, text_ids, latent_image_ids)  # [nt*B, (T/2)^2, 4C] bf16
        T2 = tile_size // 2
        # logical [nt, B, C, T, T] view of the packed prediction: channel c, pixel (y,x) -> token (y/2)*T2 + x/2,
        # feature c*4 + (y%2)*2 + (x%2); not expressible with one stride per axis, so unpack first (kernel)
        pred_t = _unpack_latents(pred.contiguous(), tile_size * self.vae_scale_factor,
                                 tile_size * self.vae_scale_factor, self.vae_scale_factor)
        pred_t = pred_t.view(nt, B, C, tile_size, tile_size)
        blended = torch.empty((B, C, h, w), device=lq_latent.device, dtype=torch.float32)
        ofs = torch.tensor(offs, dtype=torch.int32, device=lq_latent.device)
        capi.tile_blend(pred_t, wts, ofs, blended, tile_size)
        # x <- x + (t_prev - t_curr) * v ; / scale + shift (:316-317), in fp32 like the reference's promoted math
        lat = capi.latent_axpby(lq_latent.float().contiguous(), blended, 1.0, self.t_prev - self.t_curr)
        lat = capi.latent_axpby(lat, None, 1.0 / self.vae.config.scaling_factor, 0.0, self.vae.config.shift_factor)
        return self._decode_latent(lat)

    def forward(self, lq_img, prompt_embeds, pooled_prompt_embeds, text_ids, latent_image_ids, tile_size,
                tile_overlap):
        torch.cuda.synchronize()
        start_time = time.time()
        lq_latent = encode_images(lq_img.to(self.vae.dtype), self.vae, self.weight_dtype)
        _, _, h, w = lq_latent.shape
        if h * w <= tile_size * tile_size:
            print(f"[Tiled Latent]: the input size is tiny and unnecessary to tile.")
            pred_img = self._forward_no_tile(lq_latent, prompt_embeds, pooled_prompt_embeds, text_ids,
                                             latent_image_ids)
        else:
            print(f"[Tiled Latent]: the input size is {lq_img.shape[-2]}x{lq_img.shape[-1]}, need to tiled")
            pred_img = self._forward_tile(lq_latent, prompt_embeds, pooled_prompt_embeds, text_ids,
                                          latent_image_ids, tile_size, tile_overlap)
        torch.cuda.synchronize()
        t = time.time() - start_time
        print(f"Inference time per image: {t}s")
        return pred_img, t
