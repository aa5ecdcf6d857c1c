"""OMGSR-S inference driver — same flags as the reference's infer/infer_omgsr_s.py, plus `--synthetic N` and
`--prompt_embeds file.pt` for boxes without checkpoints/text encoders (sd_path/lora_path may be the
`random://...` specs documented in infer/omgsr_s_infer_model.py)."""
import argparse
import os
import sys

import torch

SCRIPT_DIR = os.path.dirname(os.path.realpath(__file__))
sys.path.append(os.path.dirname(SCRIPT_DIR))
from infer import _driver_common as dc  # noqa: E402
from infer.omgsr_s_infer_model import OMGSR_S_Infer  # noqa: E402


def encode_prompt_clip(args):
    """CLIP-H text encoding of the prompt (reference :19-45): `CLIPTextModel(ids)[0]` through the sm_100a text engine
    (omgsr_b200/engine_clip.py). Tokenisation is host string processing and stays with the checkpoint's own tokenizer;
    without one (random:// weights, no tokenizer files offline) only the empty prompt can be tokenised here:
    [BOS, EOS, pad...]."""
    import omgsr_synth as syn
    from omgsr_b200.engine_clip import build_clip_text
    prompts = args.prompt if isinstance(args.prompt, list) else [args.prompt]
    print(f"Current prompt: {prompts}")
    if args.sd_path.startswith("random://"):
        enc = build_clip_text("random://clip_h?seed=1235" if "sd21" in args.sd_path else "random://clip_tiny?seed=1235",
                              args.device)
        if any(p for p in prompts):
            raise ValueError("random:// checkpoints carry no tokenizer: only the empty prompt \"\" (the reference's "
                             "default) can be encoded; pass --prompt_embeds for anything else")
        ids = syn.empty_prompt_ids(enc.cfg).repeat(len(prompts), 1)
    else:
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(args.sd_path, subfolder="tokenizer")
        enc = build_clip_text(os.path.join(args.sd_path, "text_encoder"), args.device)
        ids = torch.cat([tok(p, max_length=tok.model_max_length, padding="max_length", truncation=True,
                             return_tensors="pt").input_ids for p in prompts])
    embs, _ = enc(ids)
    del enc
    torch.cuda.empty_cache()
    return embs.to(dtype=args.weight_dtype)


def main(args):
    if args.prompt_embeds:
        prompt_embeds = torch.load(args.prompt_embeds).to(args.device, dtype=args.weight_dtype)
    else:
        prompt_embeds = encode_prompt_clip(args)
    net_sr = OMGSR_S_Infer(sd_path=args.sd_path, lora_path=args.lora_path, mid_timestep=args.mid_timestep,
                           device=args.device, weight_dtype=args.weight_dtype)
    if args.synthetic:
        items = dc.synthetic_images(args.synthetic, args.process_size // args.upscale)
    else:
        from PIL import Image
        items = [(os.path.basename(p).split('.')[0] + ".png", Image.open(p).convert('RGB'))
                 for p in dc.list_inputs(args.input_image)]
    os.makedirs(args.output_dir, exist_ok=True)
    print(f"There are {len(items)} images.")
    if args.cuda_graph:
        net_sr.enable_cuda_graph()  # replay the captured launch sequence per (shape, tiling) - batch-1 latency
    import time
    wall0 = time.perf_counter()
    total_time = 0
    if args.batch_size > 1 and not args.host_preprocess:
        ts, to = args.process_size // 8, args.process_size // 16
        total_time, n = dc.run_batched(items, lambda x: net_sr(x, prompt_embeds, ts, to), args, args.device,
                                       args.weight_dtype)
        print(f"Average inference time: {total_time / max(n, 1)}s")
        print(f"Driver wall time (pre-processing, model, colour fix, PNG): {time.perf_counter() - wall0:.3f}s for {n} images")
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
            output_image, t = net_sr(lq_img, prompt_embeds, tile_size, tile_overlap)
            total_time += t
        (dc.finish if args.host_preprocess else dc.finish_gpu)(output_image, lq, args.align_method, resize_flag, orig,
                                                               args.upscale, os.path.join(args.output_dir, bname))
    print(f"Average inference time: {total_time / max(len(items), 1)}s")
    print(f"Driver wall time (pre-processing, model, colour fix, PNG): {time.perf_counter() - wall0:.3f}s for {len(items)} images")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="OMGSR-S Inference Script")
    ap.add_argument("--input_image", type=str, default="", help="image, directory, or .txt list")
    ap.add_argument("--output_dir", type=str, required=True)
    ap.add_argument("--sd_path", type=str, required=True)
    ap.add_argument("--lora_path", type=str, required=True)
    ap.add_argument("--device", type=str, default="cuda:0")
    ap.add_argument("--process_size", type=int, default=512)
    ap.add_argument("--upscale", type=int, default=4)
    ap.add_argument("--align_method", type=str, default="adain", choices=["wavelet", "adain", "nofix"])
    ap.add_argument("--weight_dtype", type=str, default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--prompt", type=str, nargs="+", default="")
    ap.add_argument("--mid_timestep", type=int, default=273)
    ap.add_argument("--batch_size", type=int, default=1, help="stack up to N same-size images per forward")
    ap.add_argument("--cuda_graph", action="store_true", help="capture / replay the launch sequence (untiled VAE only)")
    ap.add_argument("--host_preprocess", action="store_true",
                    help="resize / colour-align through PIL on the host like the reference (default: on the GPU)")
    ap.add_argument("--synthetic", type=int, default=0, help="run on N seeded synthetic LQ images")
    ap.add_argument("--prompt_embeds", type=str, default="", help="precomputed [1,77,D] prompt embeddings (.pt)")
    a = ap.parse_args()
    a.weight_dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[a.weight_dtype]
    main(a)
