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
    """CLIP-H text encoding of the prompt (reference :19-45); needs the SD checkpoint's tokenizer/text_encoder."""
    from transformers import AutoTokenizer, CLIPTextModel
    tok = AutoTokenizer.from_pretrained(args.sd_path, subfolder="tokenizer")
    enc = CLIPTextModel.from_pretrained(args.sd_path, subfolder="text_encoder").to(args.device, dtype=args.weight_dtype)
    prompts = args.prompt if isinstance(args.prompt, list) else [args.prompt]
    print(f"Current prompt: {prompts}")
    with torch.no_grad():
        embs = [enc(tok(p, max_length=tok.model_max_length, padding="max_length", truncation=True,
                        return_tensors="pt").input_ids.to(enc.device))[0] for p in prompts]
    del tok, enc
    torch.cuda.empty_cache()
    return torch.concat(embs, dim=0)


def main(args):
    if args.prompt_embeds:
        prompt_embeds = torch.load(args.prompt_embeds).to(args.device, dtype=args.weight_dtype)
    elif args.sd_path.startswith("random://"):
        g = torch.Generator().manual_seed(1235)
        dim = 1024 if "sd21" in args.sd_path else 128
        prompt_embeds = torch.randn(1, 77, dim, generator=g).to(args.device, dtype=args.weight_dtype)
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
