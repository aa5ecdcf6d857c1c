"""Synthetic model specifications shared by the product path, the oracle and bench.py's reference arm.

Pure PyTorch, no native code: architecture configs, diffusers-0.34.0-keyed shape tables (SURVEY.md §8c /
Appendix C), seeded random-init state dicts and random PEFT-format LoRA adapters of exactly those shapes
(neither the HF checkpoints nor diffusers/peft exist offline; BASELINE.json asks for random-init weights), and the
SD-2.1 DDPM alphas. Lives outside `omgsr_b200/` so that `bench.py --impl reference` and `oracle/` can build the very
same weights without importing the product package (which dlopen()s the CUDA library).
LoRA targets: train/train_omgsr_s.py:61-72,89-100 and train/train_omgsr_f.py:155-169 of the reference.
"""
import dataclasses
import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch


# ------------------------------------------------------------------------------------------------
# configs
# ------------------------------------------------------------------------------------------------
@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    shift_factor: float = 0.0
    use_quant_conv: bool = True
    eps: float = 1e-6


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    layers_per_block: int = 2
    num_heads: Tuple[int, ...] = (5, 10, 20, 20)  # "attention_head_dim" used as head count => d = 64
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5

    @property
    def time_embed_dim(self):
        return 4 * self.block_out_channels[0]


@dataclass
class FluxConfig:
    in_channels: int = 64
    num_layers: int = 19
    num_single_layers: int = 38
    num_heads: int = 24
    head_dim: int = 128
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    axes_dims_rope: Tuple[int, ...] = (16, 56, 56)
    guidance_embeds: bool = True
    mlp_ratio: int = 4

    @property
    def dim(self):
        return self.num_heads * self.head_dim


SD21_VAE = VAEConfig()
FLUX_VAE = VAEConfig(latent_channels=16, scaling_factor=0.3611, shift_factor=0.1159, use_quant_conv=False)
SD21_UNET = UNetConfig()
FLUX_DEV = FluxConfig()
# reduced-width / reduced-depth variants with the same topology, used by the CPU-sized parity tests
TINY_VAE = VAEConfig(block_out_channels=(64, 64, 128, 128))
TINY_FLUX_VAE = VAEConfig(latent_channels=16, scaling_factor=0.3611, shift_factor=0.1159, use_quant_conv=False,
                          block_out_channels=(64, 64, 128, 128))
TINY_UNET = UNetConfig(block_out_channels=(64, 128, 256, 256), num_heads=(1, 2, 4, 4), cross_attention_dim=128)
TINY_FLUX = FluxConfig(num_layers=2, num_single_layers=3, num_heads=4, joint_attention_dim=256,
                       pooled_projection_dim=64, axes_dims_rope=(16, 56, 56))


# ------------------------------------------------------------------------------------------------
# shape tables
# ------------------------------------------------------------------------------------------------
def _conv(sd, name, cin, cout, k):
    sd[name + ".weight"] = (cout, cin, k, k)
    sd[name + ".bias"] = (cout,)


def _lin(sd, name, cin, cout, bias=True):
    sd[name + ".weight"] = (cout, cin)
    if bias:
        sd[name + ".bias"] = (cout,)


def _norm(sd, name, c):
    sd[name + ".weight"] = (c,)
    sd[name + ".bias"] = (c,)


def _resnet(sd, name, cin, cout, temb=None):
    _norm(sd, name + ".norm1", cin)
    _conv(sd, name + ".conv1", cin, cout, 3)
    if temb:
        _lin(sd, name + ".time_emb_proj", temb, cout)
    _norm(sd, name + ".norm2", cout)
    _conv(sd, name + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(sd, name + ".conv_shortcut", cin, cout, 1)


def _vae_attn(sd, name, c):
    _norm(sd, name + ".group_norm", c)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(sd, f"{name}.{p}", c, c)


def vae_shapes(cfg: VAEConfig) -> Dict[str, tuple]:
    sd = {}
    ch = cfg.block_out_channels
    # encoder
    _conv(sd, "encoder.conv_in", cfg.in_channels, ch[0], 3)
    cin = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", cin, co)
            cin = co
        if i < len(ch) - 1:
            _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
    _resnet(sd, "encoder.mid_block.resnets.0", ch[-1], ch[-1])
    _vae_attn(sd, "encoder.mid_block.attentions.0", ch[-1])
    _resnet(sd, "encoder.mid_block.resnets.1", ch[-1], ch[-1])
    _norm(sd, "encoder.conv_norm_out", ch[-1])
    _conv(sd, "encoder.conv_out", ch[-1], 2 * cfg.latent_channels, 3)
    # decoder
    rev = list(reversed(ch))
    _conv(sd, "decoder.conv_in", cfg.latent_channels, rev[0], 3)
    _resnet(sd, "decoder.mid_block.resnets.0", rev[0], rev[0])
    _vae_attn(sd, "decoder.mid_block.attentions.0", rev[0])
    _resnet(sd, "decoder.mid_block.resnets.1", rev[0], rev[0])
    cin = rev[0]
    for i, co in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", cin, co)
            cin = co
        if i < len(ch) - 1:
            _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
    _norm(sd, "decoder.conv_norm_out", rev[-1])
    _conv(sd, "decoder.conv_out", rev[-1], cfg.out_channels, 3)
    if cfg.use_quant_conv:
        _conv(sd, "quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        _conv(sd, "post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    return sd


def _transformer2d(sd, name, c, ctx):
    _norm(sd, name + ".norm", c)
    _lin(sd, name + ".proj_in", c, c)
    tb = name + ".transformer_blocks.0"
    _norm(sd, tb + ".norm1", c)
    for p in ("to_q", "to_k", "to_v"):
        _lin(sd, f"{tb}.attn1.{p}", c, c, bias=False)
    _lin(sd, tb + ".attn1.to_out.0", c, c)
    _norm(sd, tb + ".norm2", c)
    _lin(sd, tb + ".attn2.to_q", c, c, bias=False)
    _lin(sd, tb + ".attn2.to_k", ctx, c, bias=False)
    _lin(sd, tb + ".attn2.to_v", ctx, c, bias=False)
    _lin(sd, tb + ".attn2.to_out.0", c, c)
    _norm(sd, tb + ".norm3", c)
    _lin(sd, tb + ".ff.net.0.proj", c, 8 * c)
    _lin(sd, tb + ".ff.net.2", 4 * c, c)
    _lin(sd, name + ".proj_out", c, c)


def unet_shapes(cfg: UNetConfig) -> Dict[str, tuple]:
    sd = {}
    ch = cfg.block_out_channels
    te = cfg.time_embed_dim
    _lin(sd, "time_embedding.linear_1", ch[0], te)
    _lin(sd, "time_embedding.linear_2", te, te)
    _conv(sd, "conv_in", cfg.in_channels, ch[0], 3)
    skips = [ch[0]]
    cin = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet(sd, f"down_blocks.{i}.resnets.{j}", cin, co, te)
            if cfg.down_has_attn[i]:
                _transformer2d(sd, f"down_blocks.{i}.attentions.{j}", co, cfg.cross_attention_dim)
            cin = co
            skips.append(co)
        if i < len(ch) - 1:
            _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
            skips.append(co)
    _resnet(sd, "mid_block.resnets.0", ch[-1], ch[-1], te)
    _transformer2d(sd, "mid_block.attentions.0", ch[-1], cfg.cross_attention_dim)
    _resnet(sd, "mid_block.resnets.1", ch[-1], ch[-1], te)
    rev = list(reversed(ch))
    rev_attn = list(reversed(cfg.down_has_attn))
    for i, co in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            skip = skips.pop()
            _resnet(sd, f"up_blocks.{i}.resnets.{j}", cin + skip, co, te)
            if rev_attn[i]:
                _transformer2d(sd, f"up_blocks.{i}.attentions.{j}", co, cfg.cross_attention_dim)
            cin = co
        if i < len(ch) - 1:
            _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
    _norm(sd, "conv_norm_out", ch[0])
    _conv(sd, "conv_out", ch[0], cfg.out_channels, 3)
    return sd


def flux_shapes(cfg: FluxConfig) -> Dict[str, tuple]:
    sd = {}
    d = cfg.dim
    hd = cfg.head_dim
    _lin(sd, "x_embedder", cfg.in_channels, d)
    _lin(sd, "context_embedder", cfg.joint_attention_dim, d)
    for e, cin in (("timestep_embedder", 256), ("guidance_embedder", 256), ("text_embedder", cfg.pooled_projection_dim)):
        if e == "guidance_embedder" and not cfg.guidance_embeds:
            continue
        _lin(sd, f"time_text_embed.{e}.linear_1", cin, d)
        _lin(sd, f"time_text_embed.{e}.linear_2", d, d)
    for i in range(cfg.num_layers):
        b = f"transformer_blocks.{i}"
        _lin(sd, b + ".norm1.linear", d, 6 * d)
        _lin(sd, b + ".norm1_context.linear", d, 6 * d)
        for p in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            _lin(sd, f"{b}.attn.{p}", d, d)
        for p in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            sd[f"{b}.attn.{p}.weight"] = (hd,)
        for ff in ("ff", "ff_context"):
            _lin(sd, f"{b}.{ff}.net.0.proj", d, cfg.mlp_ratio * d)
            _lin(sd, f"{b}.{ff}.net.2", cfg.mlp_ratio * d, d)
    for i in range(cfg.num_single_layers):
        b = f"single_transformer_blocks.{i}"
        _lin(sd, b + ".norm.linear", d, 3 * d)
        _lin(sd, b + ".proj_mlp", d, cfg.mlp_ratio * d)
        _lin(sd, b + ".proj_out", d + cfg.mlp_ratio * d, d)
        for p in ("to_q", "to_k", "to_v"):
            _lin(sd, f"{b}.attn.{p}", d, d)
        for p in ("norm_q", "norm_k"):
            sd[f"{b}.attn.{p}.weight"] = (hd,)
    _lin(sd, "norm_out.linear", d, 2 * d)
    _lin(sd, "proj_out", d, cfg.in_channels)
    return sd


# ------------------------------------------------------------------------------------------------
# random init (synthetic weights: there is no network access for checkpoints)
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class ClipTextConfig:
    """Text tower of CLIP as transformers' CLIPTextModel builds it (key names `text_model.*`)."""
    vocab_size: int = 49408
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_layers: int = 23
    num_heads: int = 16
    max_positions: int = 77
    hidden_act: str = "gelu"       # "gelu" (erf) | "quick_gelu" (x * sigmoid(1.702 x))
    layer_norm_eps: float = 1e-5
    eos_token_id: int = 49407
    pad_token_id: int = 0          # SD-2.x tokenizer pads with "!" (id 0); the OpenAI CLIP-L tokenizer with EOS


# stabilityai/stable-diffusion-2-1-base/text_encoder (OpenCLIP ViT-H/14 text tower without its last layer): the
# `CLIPTextModel` of infer/infer_omgsr_s.py:19-45
SD21_CLIP_H = ClipTextConfig()
# openai/clip-vit-large-patch14 text tower: `text_encoder` of FluxPipeline.encode_prompt (infer_omgsr_f.py:36-43)
FLUX_CLIP_L = ClipTextConfig(hidden_size=768, intermediate_size=3072, num_layers=12, num_heads=12,
                             hidden_act="quick_gelu", pad_token_id=49407)
TINY_CLIP = ClipTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2,
                           eos_token_id=999, pad_token_id=0)
# sizes that match TINY_FLUX (pooled projection 64 wide, joint attention 256 wide)
TINY_FLUX_CLIP = ClipTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=128, num_layers=2, num_heads=2,
                                hidden_act="quick_gelu", eos_token_id=999, pad_token_id=999)
TINY_CLIP_QUICK = ClipTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2,
                                 hidden_act="quick_gelu", eos_token_id=999, pad_token_id=999)


def clip_text_shapes(cfg: ClipTextConfig) -> Dict[str, tuple]:
    sd = {}
    C = cfg.hidden_size
    sd["text_model.embeddings.token_embedding.weight"] = (cfg.vocab_size, C)
    sd["text_model.embeddings.position_embedding.weight"] = (cfg.max_positions, C)
    for i in range(cfg.num_layers):
        b = f"text_model.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _lin(sd, f"{b}.self_attn.{n}", C, C)
        _norm(sd, f"{b}.layer_norm1", C)
        _lin(sd, f"{b}.mlp.fc1", C, cfg.intermediate_size)
        _lin(sd, f"{b}.mlp.fc2", cfg.intermediate_size, C)
        _norm(sd, f"{b}.layer_norm2", C)
    _norm(sd, "text_model.final_layer_norm", C)
    return sd


@dataclasses.dataclass
class T5EncoderConfig:
    """Encoder stack of T5 v1.1 as transformers' T5EncoderModel builds it (gated GELU-tanh feed-forward, RMS norms,
    relative position bias shared by all layers, no biases)."""
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_heads: int = 64
    num_buckets: int = 32
    max_distance: int = 128
    layer_norm_eps: float = 1e-6
    eos_token_id: int = 1
    pad_token_id: int = 0


# google/t5-v1_1-xxl encoder: `text_encoder_2` of FluxPipeline.encode_prompt (infer_omgsr_f.py:36-43), 512 tokens
FLUX_T5_XXL = T5EncoderConfig()
TINY_T5 = T5EncoderConfig(vocab_size=1000, d_model=128, d_kv=32, d_ff=256, num_layers=2, num_heads=4)
TINY_FLUX_T5 = T5EncoderConfig(vocab_size=1000, d_model=256, d_kv=32, d_ff=512, num_layers=2, num_heads=4)
SMALL_T5 = T5EncoderConfig(vocab_size=1000, d_model=512, d_kv=64, d_ff=1024, num_layers=3, num_heads=8)


def t5_encoder_shapes(cfg: T5EncoderConfig) -> Dict[str, tuple]:
    sd = {"shared.weight": (cfg.vocab_size, cfg.d_model)}
    inner = cfg.num_heads * cfg.d_kv
    for i in range(cfg.num_layers):
        b = f"encoder.block.{i}.layer"
        for n in ("q", "k", "v"):
            sd[f"{b}.0.SelfAttention.{n}.weight"] = (inner, cfg.d_model)
        sd[f"{b}.0.SelfAttention.o.weight"] = (cfg.d_model, inner)
        if i == 0:
            sd[f"{b}.0.SelfAttention.relative_attention_bias.weight"] = (cfg.num_buckets, cfg.num_heads)
        sd[f"{b}.0.layer_norm.weight"] = (cfg.d_model,)
        sd[f"{b}.1.DenseReluDense.wi_0.weight"] = (cfg.d_ff, cfg.d_model)
        sd[f"{b}.1.DenseReluDense.wi_1.weight"] = (cfg.d_ff, cfg.d_model)
        sd[f"{b}.1.DenseReluDense.wo.weight"] = (cfg.d_model, cfg.d_ff)
        sd[f"{b}.1.layer_norm.weight"] = (cfg.d_model,)
    sd["encoder.final_layer_norm.weight"] = (cfg.d_model,)
    return sd


def empty_prompt_ids_t5(cfg: T5EncoderConfig, length=512):
    """T5 tokenisation of the empty prompt "" at max_length padding: [EOS, pad, pad, ...] (sentencepiece adds only the
    end-of-sequence token)."""
    return torch.tensor([[cfg.eos_token_id] + [cfg.pad_token_id] * (length - 1)], dtype=torch.int64)


def empty_prompt_ids(cfg: ClipTextConfig):
    """Token ids of the empty prompt "" at max_length padding - the reference's default prompt
    (infer_omgsr_s.py `--prompt ""`): [BOS, EOS, pad, pad, ...]. No tokenizer files exist offline; any other prompt
    has to arrive tokenized."""
    return torch.tensor([[49406 if cfg.vocab_size >= 49408 else cfg.vocab_size - 2, cfg.eos_token_id] +
                         [cfg.pad_token_id] * (cfg.max_positions - 2)], dtype=torch.int64)


def random_state_dict(shapes: Dict[str, tuple], seed: int, dtype=torch.float32, out_gain: Dict[str, float] = None,
                      device="cpu"):
    """Variance-preserving init: conv/linear weights ~ N(0, 1/fan_in), biases ~ N(0, 0.02^2), norm affine
    ~ (1 + N(0, 0.1^2), N(0, 0.1^2)). `out_gain` rescales selected layers (substring match) so that the
    end-to-end outputs have O(1) range and the PSNR / relative-error bars are meaningful."""
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda shp: torch.randn(shp, generator=g, device=device)
    sd = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            t = rn(shp) / math.sqrt(fan_in)
        elif name.endswith(".weight"):
            t = 1.0 + 0.1 * rn(shp)
        else:
            is_norm = name.replace(".bias", ".weight") in shapes and len(shapes[name.replace(".bias", ".weight")]) == 1
            t = (0.1 if is_norm else 0.02) * rn(shp)
        if out_gain:
            for key, gain in out_gain.items():
                if key in name and name.endswith(".weight") and len(shp) >= 2:
                    t = t * gain
        sd[name] = t.to(dtype)
    return sd


# LoRA target modules by suffix (train/train_omgsr_s.py:61-72,89-100; train/train_omgsr_f.py:155-169)
LORA_TARGETS_SD = ("conv1", "conv2", "conv_in", "conv_shortcut", "conv", "conv_out", "to_q", "to_k", "to_v", "to_out.0")
LORA_TARGETS_FLUX = ("x_embedder", "attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "attn.add_q_proj",
                     "attn.add_k_proj", "attn.add_v_proj", "attn.to_add_out", "ff.net.0.proj", "ff.net.2",
                     "ff_context.net.0.proj", "ff_context.net.2")


def lora_target_names(shapes: Dict[str, tuple], targets, prefix=""):
    out = []
    for name in shapes:
        if not name.endswith(".weight") or len(shapes[name]) < 2 or not name.startswith(prefix):
            continue
        mod = name[: -len(".weight")]
        if any(mod == t or mod.endswith("." + t) for t in targets):
            out.append(mod)
    return out


def random_lora(shapes, targets, rank, seed, prefix="", b_std=0.02, device="cpu"):
    """PEFT-format adapter tensors {module: (A, B)}: Linear A [r,in], B [out,r]; Conv2d A [r,Cin,k,k]
    (same kernel as the base conv), B [Cout,r,1,1]. A ~ N(0, 1/r) ('gaussian' init, train_omgsr_s.py:78);
    PEFT zero-inits B, here B ~ N(0, b_std^2) so that the merge is actually exercised."""
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda shp: torch.randn(shp, generator=g, device=device)
    ad = {}
    for mod in lora_target_names(shapes, targets, prefix):
        shp = shapes[mod + ".weight"]
        if len(shp) == 4:
            a = rn((rank, shp[1], shp[2], shp[3])) / rank
            b = rn((shp[0], rank, 1, 1)) * b_std
        else:
            a = rn((rank, shp[1])) / rank
            b = rn((shp[0], rank)) * b_std
        ad[mod] = (a, b)
    return ad


def ddpm_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """SD-2.1-base DDPMScheduler: scaled-linear betas (the reference only reads alphas_cumprod[273],
    infer/omgsr_s_infer_model.py:13-14)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)
