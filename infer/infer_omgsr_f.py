"""OMGSR-F inference driver — same flags as the reference's infer/infer_omgsr_f.py, plus `--synthetic N` and
`--prompt_embeds file.pt` (a dict with prompt_embeds / pooled_prompt_embeds / text_ids) for boxes without the
FLUX text encoders."""
import argparse
import os
import sys

import torch

SCRIPT_DIR = os.path.dirname(os.path.realpath(__file__))
sys.path.append(os.path.dirname(SCRIPT_DIR))
from infer import _driver_common as dc  # noqa: E402


def _prepare_latent_image_ids(height, width, device, dtype):
    """[height*width, 3] position ids (0, row, col) for the packed latent grid (reference :17-28)."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] += torch.arange(height)[:, None]
    ids[..., 2] += torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)


def encode_prompt_flux(args):
    """`FluxPipeline.encode_prompt(prompt, prompt_2=None)` (reference :36-48) through the sm_100a text engines:
    pooled_prompt_embeds = CLIP-L(ids_77).pooler_output, prompt_embeds = T5-XXL encoder(ids_512)[0], text_ids = zeros.
    Tokenisation is host string processing and stays with the checkpoint's own tokenizers; `random://` weights carry none,
    so only the empty prompt (the reference's default) can be encoded there: CLIP [BOS, EOS, EOS...], T5 [EOS, pad...]."""
    import omgsr_synth as syn
    from omgsr_b200.engine_clip import build_clip_text
    from omgsr_b200.engine_t5 import build_t5_encoder
    if args.flux_path.startswith("random://"):
        if args.prompt:
            raise ValueError("random:// checkpoints carry no tokenizer: only the empty prompt \"\" can be encoded; "
                             "pass --prompt_embeds for anything else")
        full = "flux_dev" in args.flux_path
        clip = build_clip_text("random://clip_l?seed=1235" if full else "random://clip_tiny_flux?seed=1235", args.device)
        ids_clip = syn.empty_prompt_ids(clip.cfg)
        t5 = build_t5_encoder("random://t5_xxl?seed=1236" if full else "random://t5_tiny_flux?seed=1236", args.device)
        ids_t5 = syn.empty_prompt_ids_t5(t5.cfg, 512 if full else 32)
    else:
        from transformers import AutoTokenizer
        tok1 = AutoTokenizer.from_pretrained(args.flux_path, subfolder="tokenizer")
        tok2 = AutoTokenizer.from_pretrained(args.flux_path, subfolder="tokenizer_2")
        clip = build_clip_text(os.path.join(args.flux_path, "text_encoder"), args.device)
        t5 = build_t5_encoder(os.path.join(args.flux_path, "text_encoder_2"), args.device)
        ids_clip = tok1(args.prompt, padding="max_length", max_length=tok1.model_max_length, truncation=True,
                        return_tensors="pt").input_ids
        ids_t5 = tok2(args.prompt, padding="max_length", max_length=512, truncation=True, return_tensors="pt").input_ids
    _, pooled = clip(ids_clip)
    del clip
    prompt_embeds = t5(ids_t5)
    del t5
    torch.cuda.empty_cache()
    return prompt_embeds, pooled, torch.zeros(prompt_embeds.shape[1], 3)


def main(args):
    from infer.omgsr_f_infer_model import OMGSR_F_Infer
    omgsr = OMGSR_F_Infer(args.flux_path, args.lora_path, device=args.device, guidance_scale=args.guidance_scale,
                          mid_timestep=args.mid_timestep)
    if args.prompt_embeds:
        pe = torch.load(args.prompt_embeds)
        prompt_embeds, pooled, text_ids = pe["prompt_embeds"], pe["pooled_prompt_embeds"], pe["text_ids"]
    else:
        prompt_embeds, pooled, text_ids = encode_prompt_flux(args)
    to = dict(device=args.device, dtype=args.weight_dtype)
    prompt_embeds, pooled, text_ids = prompt_embeds.to(**to), pooled.to(**to), text_ids.to(**to)
    grid = (args.process_size // 8) // 2
    latent_image_ids = _prepare_latent_image_ids(grid, grid, args.device, args.weight_dtype)
    if args.synthetic:
        items = dc.synthetic_images(args.synthetic, args.process_size // args.upscale)
    else:
        from PIL import Image
        items = [(os.path.basename(p).split('.')[0] + ".png", Image.open(p).convert('RGB'))
                 for p in dc.list_inputs(args.input_image)]
    os.makedirs(args.output_dir, exist_ok=True)
    print(f"There are {len(items)} images.")
    total_time = 0
    if args.batch_size > 1 and not args.host_preprocess:
        ts, to = args.process_size // 8, args.process_size // 16
        total_time, n = dc.run_batched(
            items, lambda x: omgsr(x, prompt_embeds, pooled, text_ids, latent_image_ids, ts, to), args, args.device,
            args.weight_dtype)
        print(f"Average inference time: {total_time / max(n, 1)}s")
        return
    for bname, img in items:
        if args.host_preprocess:
            lq, resize_flag, orig = dc.preprocess(img, args.process_size, args.upscale)
            lq_img = dc.pil_to_model_input(lq, args.device, args.weight_dtype)
        else:  # same chain on the device, bit-identical to PIL (omgsr_resize_u8)
            lq, lq_img, resize_flag, orig = dc.preprocess_gpu(img, args.process_size, args.upscale, args.device,
                                                             args.weight_dtype)
        tile_size = args.process_size // 8
        tile_overlap = tile_size // 2
        with torch.no_grad():
            output_image, t = omgsr(lq_img, prompt_embeds, pooled, text_ids, latent_image_ids, tile_size, tile_overlap)
            total_time += t
        (dc.finish if args.host_preprocess else dc.finish_gpu)(output_image, lq, args.align_method, resize_flag, orig,
                                                               args.upscale, os.path.join(args.output_dir, bname))
    print(f"Average inference time: {total_time / max(len(items), 1)}s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="OMGSR-F Inference Script")
    ap.add_argument("--input_image", type=str, default="")
    ap.add_argument("--output_dir", type=str, required=True)
    ap.add_argument("--flux_path", type=str, required=True)
    ap.add_argument("--lora_path", type=str, required=True)
    ap.add_argument("--device", type=str, default="cuda:0")
    ap.add_argument("--process_size", type=int, default=1024)
    ap.add_argument("--upscale", type=int, default=4)
    ap.add_argument("--align_method", type=str, default="adain", choices=["wavelet", "adain", "nofix"])
    ap.add_argument("--weight_dtype", type=str, default="bf16", choices=["bf16"])
    ap.add_argument("--prompt", type=str, default="")
    ap.add_argument("--mid_timestep", type=int, default=244)
    ap.add_argument("--guidance_scale", type=float, default=1.0)
    ap.add_argument("--batch_size", type=int, default=1, help="stack up to N same-size images per forward")
    ap.add_argument("--host_preprocess", action="store_true",
                    help="resize / colour-align through PIL on the host like the reference (default: on the GPU)")
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--prompt_embeds", type=str, default="")
    a = ap.parse_args()
    a.weight_dtype = torch.bfloat16
    main(a)
