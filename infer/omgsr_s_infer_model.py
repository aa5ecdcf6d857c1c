"""OMGSR-S one-step SR inference module — B200-native drop-in for the reference's
infer/omgsr_s_infer_model.py (same class name, constructor and method signatures, return contract).

Dataflow (reference lines in brackets): VAE-encode the LQ image and sample the posterior, times the
scaling factor [:173] -> one UNet forward at the fixed mid-timestep [:75-79] -> epsilon-to-x0 step
[:80-82] -> VAE-decode and clamp [:83-85]; latents larger than tile_size^2 go through the overlapping
latent-tile path with Gaussian blending [:88-168]. Everything between the NCHW tensors at this boundary
runs in hand-written sm_100a kernels (omgsr_b200/csrc) through the C ABI; there is no PyTorch/CPU fallback.

`sd_path` / `lora_path` accept, besides checkpoint directories (omgsr_b200.checkpoint), the synthetic specs
"random://sd21?seed=0" / "random://tiny?seed=0" and "random://lora?seed=1" because neither HF weights nor
diffusers/peft exist offline (additive constructor path).
"""
import math
import os
import time

import torch

from omgsr_b200 import capi
from omgsr_b200 import dist as D
from omgsr_b200 import weights as W
from omgsr_b200.engine_unet import B200UNet
from omgsr_b200.engine_vae import B200AutoencoderKL
from .vaehook import VAEHook


def _select_device(device):
    """make `device` the current CUDA device of this process: the C-ABI kernels launch on the current device / stream
    (the reference's drivers pass --device cuda:N, infer/infer_omgsr_s.py:116)."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is not None:
        torch.cuda.set_device(d)


def _parse_spec(spec):
    body = spec[len("random://"):]
    name, _, q = body.partition("?")
    kv = dict(p.split("=") for p in q.split("&") if p)
    return name, {k: int(v) for k, v in kv.items()}


def build_s_state(sd_path, lora_path):
    """-> (vae_sd, unet_sd, vae_cfg, unet_cfg) with the LoRA deltas merged in fp32
    (vae.encoder + unet only; the decoder is frozen — train/train_omgsr_s.py:215-222)."""
    if isinstance(sd_path, dict):  # pre-built state: {"vae":..., "unet":..., "vae_cfg":..., "unet_cfg":...}
        st = sd_path
        return st["vae"], st["unet"], st["vae_cfg"], st["unet_cfg"]
    if sd_path.startswith("random://"):
        name, kv = _parse_spec(sd_path)
        seed = kv.get("seed", 0)
        vae_cfg, unet_cfg = (W.SD21_VAE, W.SD21_UNET) if name == "sd21" else (W.TINY_VAE, W.TINY_UNET)
        vae_sd = W.random_state_dict(W.vae_shapes(vae_cfg), seed)
        unet_sd = W.random_state_dict(W.unet_shapes(unet_cfg), seed + 1)
    else:
        from omgsr_b200.checkpoint import load_sd_checkpoint
        vae_sd, unet_sd, vae_cfg, unet_cfg = load_sd_checkpoint(sd_path)
    if lora_path:
        if lora_path.startswith("random://"):
            _, kv = _parse_spec(lora_path)
            seed = kv.get("seed", 1)
            r_vae, r_unet = kv.get("vae_rank", 16), kv.get("unet_rank", 32)  # configs/omgsr_s_512.yml:8-9
            vae_ad = W.random_lora(W.vae_shapes(vae_cfg), W.LORA_TARGETS_SD, r_vae, seed, prefix="encoder.")
            unet_ad = W.random_lora(W.unet_shapes(unet_cfg), W.LORA_TARGETS_SD, r_unet, seed + 1)
        else:
            from omgsr_b200.checkpoint import load_peft_adapter
            vae_ad = load_peft_adapter(os.path.join(lora_path, "vae_encoder_lora_adapter"), prefix="encoder.")
            unet_ad = load_peft_adapter(os.path.join(lora_path, "unet_lora_adapter"))
        vae_sd = W.merge_lora(vae_sd, vae_ad)
        unet_sd = W.merge_lora(unet_sd, unet_ad)
    return vae_sd, unet_sd, vae_cfg, unet_cfg


class _Scheduler:
    """the only thing the reference reads from DDPMScheduler is alphas_cumprod (:13-14)."""

    def __init__(self):
        self.alphas_cumprod = W.ddpm_alphas_cumprod()


class OMGSR_S_Infer(torch.nn.Module):
    def __init__(self, sd_path, lora_path, mid_timestep, device, weight_dtype):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("OMGSR_S_Infer (omgsr_b200) needs a CUDA device: the hot path has no CPU fallback")
        _select_device(device)
        vae_sd, unet_sd, vae_cfg, unet_cfg = build_s_state(sd_path, lora_path)
        self.mid_timestep = mid_timestep
        self.scheduler = _Scheduler()
        self.alpha_t = self.scheduler.alphas_cumprod[mid_timestep]
        print(f"Current One mid-timestep settings: {mid_timestep}")
        self.vae = B200AutoencoderKL(vae_sd, vae_cfg, device=device, dtype=weight_dtype)
        self.unet = B200UNet(unet_sd, unet_cfg, mid_timestep, device=device, dtype=weight_dtype)
        self.vae.eval()
        self.unet.eval()
        self.device = device
        self._s1 = float((1 - self.alpha_t).sqrt())
        self._s2 = float(self.alpha_t.sqrt())
        self._wts = {}
        pq = self.vae.post_quant_conv
        self._pq = (pq.wf, pq.bf) if pq is not None else (None, None)
        # self._init_tiled_vae(encoder_tile_size=1024, decoder_tile_size=224)   (off by default, as in the reference :32)

    def _init_tiled_vae(
            self,
            encoder_tile_size=256,
            decoder_tile_size=256,
            fast_decoder=False,
            fast_encoder=False,
            color_fix=False,
            vae_to_gpu=True):
        # save original forward (only once)
        if not hasattr(self.vae.encoder, 'original_forward'):
            setattr(self.vae.encoder, 'original_forward', self.vae.encoder.forward)
        if not hasattr(self.vae.decoder, 'original_forward'):
            setattr(self.vae.decoder, 'original_forward', self.vae.decoder.forward)
        encoder = self.vae.encoder
        decoder = self.vae.decoder
        self.vae.encoder.forward = VAEHook(
            encoder, encoder_tile_size, is_decoder=False, fast_decoder=fast_decoder, fast_encoder=fast_encoder,
            color_fix=color_fix, to_gpu=vae_to_gpu)
        self.vae.decoder.forward = VAEHook(
            decoder, decoder_tile_size, is_decoder=True, fast_decoder=fast_decoder, fast_encoder=fast_encoder,
            color_fix=color_fix, to_gpu=vae_to_gpu)

    def _gaussian_weights(self, tile_width, tile_height, nbatches):
        """Gaussian tile mask, float64, tiled to (nbatches, in_channels, H, W) like the reference (:56-71)."""
        var = 0.01
        midpoint = (tile_width - 1) / 2
        x_probs = [math.exp(-(x - midpoint) * (x - midpoint) / (tile_width * tile_width) / (2 * var)) /
                   math.sqrt(2 * math.pi * var) for x in range(tile_width)]
        midpoint = tile_height / 2  # (sic) the reference is asymmetric between x and y
        y_probs = [math.exp(-(y - midpoint) * (y - midpoint) / (tile_height * tile_height) / (2 * var)) /
                   math.sqrt(2 * math.pi * var) for y in range(tile_height)]
        wts = torch.outer(torch.tensor(y_probs, dtype=torch.float64), torch.tensor(x_probs, dtype=torch.float64))
        return torch.tile(wts.to(self.device), (nbatches, self.unet.config.in_channels, 1, 1))

    # -- tails -------------------------------------------------------------------------------------
    def _decode_from_eps(self, lq_latent, eps_hat_nhwc):
        """x0 = (z - sqrt(1-a) eps)/sqrt(a) (:80-82), / scaling_factor (:84), decode, clamp (:85)."""
        sf = self.vae.config.scaling_factor
        patched = "forward" in self.vae.decoder.__dict__
        if patched:  # tiled decoder (VAEHook): go through vae.decode so post_quant_conv + hook run
            z0 = capi.x0_unscale(lq_latent.contiguous(), eps_hat_nhwc, self._s1, self._s2, sf)
            return self.vae.decode(z0.to(self.vae.dtype)).sample.clamp(-1, 1).to(self.vae.dtype)
        z0 = capi.x0_unscale(lq_latent.contiguous(), eps_hat_nhwc, self._s1, self._s2, sf, 0.0, *self._pq)
        return self.vae.decoder.forward(z0, clamp=True)

    def _forward_no_tile(self, lq_latent, prompt_embeds):
        eps_hat = self.unet.eps_hat_nhwc(lq_latent, prompt_embeds)
        return self._decode_from_eps(lq_latent, eps_hat)

    @staticmethod
    def _tile_offsets(h, w, tile_size, tile_overlap):
        """(ofs_x, ofs_y) per tile in the reference's visiting order (:94-118)."""
        grid_rows = 0
        cur_x = 0
        while cur_x < w:
            cur_x = max(grid_rows * tile_size - tile_overlap * grid_rows, 0) + tile_size
            grid_rows += 1
        grid_cols = 0
        cur_y = 0
        while cur_y < h:
            cur_y = max(grid_cols * tile_size - tile_overlap * grid_cols, 0) + tile_size
            grid_cols += 1
        offs = []
        ofs_x = ofs_y = 0
        for row in range(grid_rows):
            for col in range(grid_cols):
                if col < grid_cols - 1 or row < grid_rows - 1:
                    ofs_x = max(row * tile_size - tile_overlap * row, 0)
                    ofs_y = max(col * tile_size - tile_overlap * col, 0)
                if row == grid_rows - 1:
                    ofs_x = w - tile_size
                if col == grid_cols - 1:
                    ofs_y = h - tile_size
                offs.append((ofs_x, ofs_y))
        return offs

    def _forward_tile(self, lq_latent, prompt_embeds, tile_size, tile_overlap):
        B, C, h, w = lq_latent.shape
        tile_size = min(tile_size, min(h, w))
        key = (tile_size, str(lq_latent.device))
        if key not in self._wts:
            self._wts[key] = self._gaussian_weights(tile_size, tile_size, 1)[0, 0].float().contiguous()
        wts = self._wts[key]
        offs = self._tile_offsets(h, w, tile_size, tile_overlap)
        nt = len(offs)
        # the reference runs the tiles one by one at batch 1 (:129-133); they are independent, so they are
        # stacked on the batch axis and run as ONE UNet forward of batch nt*B.
        tiles = torch.stack([lq_latent[:, :, oy:oy + tile_size, ox:ox + tile_size] for ox, oy in offs], 0)
        tiles = tiles.reshape(nt * B, C, tile_size, tile_size).contiguous()
        if prompt_embeds.shape[0] not in (1, nt * B):  # per-image prompts: repeat for every tile of that image
            prompt_embeds = prompt_embeds.repeat(nt, 1, 1)
        if D.tile_parallel():  # single large image on several GPUs: each rank denoises its share of the tiles
            if prompt_embeds.shape[0] != 1:
                raise RuntimeError("tile-parallel mode expects one shared prompt")
            eps_t = D.tile_parallel_forward(lambda t: self.unet.eps_hat_nhwc(t, prompt_embeds)[..., :C], tiles)
        else:
            eps_t = self.unet.eps_hat_nhwc(tiles, prompt_embeds)  # fp32 NHWC view [nt*B,T,T,C] (ld 8)
        ld = eps_t.stride(2)
        et = torch.as_strided(eps_t, (nt, B, C, tile_size, tile_size),
                              (B * tile_size * tile_size * ld, tile_size * tile_size * ld, 1, tile_size * ld, ld),
                              eps_t.storage_offset())
        blended = torch.empty((B, h, w, 8), device=lq_latent.device, dtype=torch.float32)
        okey = ("ofs", h, w, tile_size, tile_overlap, str(lq_latent.device))
        if okey not in self._wts:  # cached: no host->device copy inside a CUDA-graph capture
            self._wts[okey] = torch.tensor(offs, dtype=torch.int32, device=lq_latent.device)
        capi.tile_blend(et, wts, self._wts[okey], blended[..., :C].permute(0, 3, 1, 2), tile_size)
        return self._decode_from_eps(lq_latent, blended[..., :C])

    def _encode(self, lq_img):
        """vae.encode(x).latent_dist.sample() * scaling_factor (:173), fused: the posterior kernel applies
        the scale; quant_conv is folded into conv_out."""
        vae = self.vae
        if "forward" in vae.encoder.__dict__:  # tiled encoder installed
            return vae.encode(lq_img.to(vae.dtype)).latent_dist.sample() * vae.config.scaling_factor
        x = lq_img.to(vae.dtype)
        m = vae.encoder.moments_nhwc(x, *(vae._folded or (None, None)))
        B, hh, ww, _ = m.shape
        cz = vae.cfg.latent_channels
        eps = vae.posterior_eps
        if eps is None:
            eps = D.broadcast_(torch.randn((B, cz, hh, ww), device=m.device, dtype=torch.float32))
        return capi.posterior_sample(m, eps.to(device=m.device, dtype=torch.float32), cz, 0.0,
                                     vae.config.scaling_factor, vae.dtype)

    # -- CUDA graph replay (additive) ---------------------------------------------------------------------
    def enable_cuda_graph(self, on=True):
        """Capture the whole launch sequence (~660 kernels for SD-2.1 shapes) of one (shape, tile setting) into a CUDA
        graph on first use and replay it afterwards. The reference's use case is batch 1 per image, where the
        Python/ctypes launch overhead (not the GPU) bounds the latency; every C-ABI entry only enqueues on the
        current stream, so the sequence is capturable as is. The posterior noise is drawn outside the graph."""
        self._graphs = {} if on else None

    def _run(self, lq_img, prompt_embeds, tile_size, tile_overlap):
        lq_latent = self._encode(lq_img)
        _, _, h, w = lq_latent.size()
        if h * w <= tile_size * tile_size:
            return self._forward_no_tile(lq_latent, prompt_embeds), False
        return self._forward_tile(lq_latent, prompt_embeds, tile_size, tile_overlap), True

    def _run_graphed(self, lq_img, prompt_embeds, tile_size, tile_overlap):
        if "forward" in self.vae.encoder.__dict__ or D.tile_parallel():
            raise RuntimeError("CUDA-graph replay supports the plain (untiled-VAE, single-GPU) path only")
        key = (tuple(lq_img.shape), lq_img.dtype, tuple(prompt_embeds.shape), tile_size, tile_overlap)
        cz = self.vae.cfg.latent_channels
        B, _, H, W = lq_img.shape
        ent = self._graphs.get(key)
        if ent is None:
            st = {"x": torch.empty_like(lq_img), "ctx": prompt_embeds.to(torch.bfloat16).clone(),
                  "eps": torch.empty((B, cz, H // 8, W // 8), device=lq_img.device, dtype=torch.float32)}
            st["x"].copy_(lq_img)
            st["eps"].normal_()
            user_eps = self.vae.posterior_eps
            self.vae.posterior_eps = st["eps"]
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):  # warm-up: lazy attribute setting / caches happen outside the capture
                    self.unet.cache_kv = False
                    self._run(st["x"], st["ctx"], tile_size, tile_overlap)
                torch.cuda.current_stream().wait_stream(side)
                self.unet.cache_kv = True
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out, tiled = self._run(st["x"], st["ctx"], tile_size, tile_overlap)
            finally:
                self.unet.cache_kv = True
                self.vae.posterior_eps = user_eps
            ent = self._graphs[key] = (graph, st, out, tiled)
        graph, st, out, tiled = ent
        st["x"].copy_(lq_img)
        st["ctx"].copy_(prompt_embeds)
        if self.vae.posterior_eps is not None:
            st["eps"].copy_(self.vae.posterior_eps)
        else:
            st["eps"].normal_()
        graph.replay()
        return out.clone(), tiled

    def forward(self, lq_img, prompt_embeds, tile_size, tile_overlap):
        torch.cuda.synchronize()
        start_time = time.time()
        if getattr(self, "_graphs", None) is not None:
            pred_img, tiled = self._run_graphed(lq_img, prompt_embeds, tile_size, tile_overlap)
        else:
            pred_img, tiled = self._run(lq_img, prompt_embeds, tile_size, tile_overlap)
        if not tiled:
            print(f"[Tiled Latent]: the input size is tiny and unnecessary to tile.")
        else:
            print(f"[Tiled Latent]: the input size is {lq_img.shape[-2]}x{lq_img.shape[-1]}, need to tiled")
        torch.cuda.synchronize()
        t = time.time() - start_time
        print(f"Inference time per image: {t}s")
        return pred_img, t
