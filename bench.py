#!/usr/bin/env python
"""bench.py — OMGSR one-step SR throughput on B200 (BASELINE.json metric: SR images/s).

Workload (N=1): BASELINE.json configs[1] — OMGSR-S-512, bf16, batch 16 per GPU: VAE encode -> UNet forward at
the mid-timestep -> x0 step -> VAE decode, SD-2.1 shapes, random-init weights with merged LoRA, synthetic LQ
images. N>1: image-parallel replicas (one process per GPU, no data-path collective) => weak scaling.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference ...                   # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0). `value` = device-timed images/s with inputs resident in HBM; `e2e` = the same
through the public module call with pinned-host inputs/outputs inside the timed region.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

S512_FLOP_PER_IMAGE = 4.435e12  # enc 1.1167 + UNet 0.8043 + dec 2.5145 TFLOP (2*MAC)
F1024_FLOP_PER_IMAGE = 89.74e12  # enc 4.8826 + DiT 74.3847 + dec 10.4722 TFLOP
WORKLOAD = "OMGSR-S-512 bf16 batch 16 (128->512 SR: VAE enc + UNet@t=273 + VAE dec, SD-2.1 shapes)"
WORKLOAD_F = "OMGSR-F-1024 bf16 batch 4 per GPU (256->1024 SR: FLUX VAE enc + FLUX.1-dev-shape DiT@sigma=0.505 + VAE dec)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _dist():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    return ws, rank, lr


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port; diffusers/peft are not installable offline) on CPU
# ----------------------------------------------------------------------------------------------------
def cpu_reference_images_per_s(n_images, threads, budget_s=1e9):
    from infer.omgsr_s_infer_model import build_s_state
    from omgsr_b200 import weights as W
    from oracle.omgsr_s import OracleS
    torch.set_num_threads(threads)
    vae_sd, unet_sd, vae_cfg, unet_cfg = build_s_state("random://sd21?seed=0", "random://lora?seed=1")
    oracle = OracleS(vae_sd, unet_sd, vae_cfg, unet_cfg, 273, W.ddpm_alphas_cumprod())
    g = torch.Generator().manual_seed(1234)
    lq = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    ctx = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(1235))
    eps = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(1236))
    times = []
    t_start = time.time()
    with torch.no_grad():
        for _ in range(n_images):
            t0 = time.time()
            oracle.forward(lq, ctx, 64, 32, eps)
            times.append(time.time() - t0)
            if time.time() - t_start > budget_s:
                break
    return times


def run_reference(args):
    ws, rank, _ = _dist()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # one step = one 512x512 image at batch 1 (a bounded sample of the batch-16 workload); the run is
    # capped at ~4 minutes of CPU time and reports the steps it completed
    times = cpu_reference_images_per_s(args.warmup + args.steps, threads, budget_s=240.0)
    w = min(args.warmup, max(len(times) - 1, 0))
    timed = times[w:]
    per = sum(timed) / len(timed)
    val = 1.0 / per
    line = {
        "impl": "reference", "metric": "SR images/sec (OMGSR-S 128->512)", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": w, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "1 image (batch 1) per step on the host CPU"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{len(timed)} x one 512x512 image, fp32, oracle port of the reference path"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------
def run_ours_f1024(args):
    """BASELINE config 4 (not the default N=1 line): OMGSR-F-1024, batch 4 per GPU, image-parallel replicas."""
    ws, rank, lr = _dist()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the omgsr_b200 hot path has no CPU fallback")
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    import torch.distributed as dist
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    from infer.infer_omgsr_f import _prepare_latent_image_ids
    from infer.omgsr_f_infer_model import OMGSR_F_Infer
    from omgsr_b200 import capi
    B = args.batch if args.batch != 16 else 4
    with _quiet():
        model = OMGSR_F_Infer("random://flux_dev?seed=0", "random://lora?seed=1&rank=64", dev)
    g = torch.Generator().manual_seed(1234 + rank)
    host_in = [torch.rand(B, 3, 1024, 1024, generator=g).mul_(2).sub_(1).bfloat16().pin_memory() for _ in range(2)]
    dev_in = [t.to(dev) for t in host_in]
    g2 = torch.Generator().manual_seed(1235)
    pe = torch.randn(1, 512, 4096, generator=g2).to(dev).bfloat16()
    pp = torch.randn(1, 768, generator=g2).to(dev).bfloat16()
    tid = torch.zeros(512, 3, device=dev, dtype=torch.bfloat16)
    ids = _prepare_latent_image_ids(64, 64, dev, torch.bfloat16)
    host_out = torch.empty(B, 3, 1024, 1024, dtype=torch.bfloat16).pin_memory()

    def step(i, x=None):
        with torch.no_grad(), _quiet():
            return model(dev_in[i % 2] if x is None else x, pe, pp, tid, ids, 128, 64)[0]

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(lr) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    launches = capi.launch_count() - l0
    ms = e0.elapsed_time(e1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = step(i, host_in[i % 2].to(dev, non_blocking=True))
        host_out.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if ws > 1:
        t = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0].item()), float(t[1].item())
    ms_per_step = ms / args.steps
    burst, sustained, hbm, how = _peaks()
    step_tflops = F1024_FLOP_PER_IMAGE * B / (ms_per_step * 1e-3) / 1e12
    if rank == 0:
        print(json.dumps({
            "metric": "SR images/sec (OMGSR-F 256->1024)", "value": ws * B / (ms_per_step * 1e-3), "unit": "images/s",
            "n_gpus": ws, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD_F, "batch_per_gpu": B, "parallelism": f"image-parallel x{ws}",
                       "l2": "2 rotating input buffers; 24 GB of weights + GB-sized activations per step >> L2",
                       "step_tflops_algorithmic": step_tflops, "step_frac_of_sustained_peak": step_tflops / sustained},
            "clocks": clocks.summary(),
            "e2e": {"value": ws * B * args.steps / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": B * 3 * 1024 * 1024 * 2, "d2h_bytes_per_step": B * 3 * 1024 * 1024 * 2},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "whole step (DiT GEMMs + flash attention + VAE convs)",
                         "achieved": step_tflops, "peak": sustained, "unit": "TFLOP/s", "frac": step_tflops / sustained,
                         "traffic": None, "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({how})"},
            "cpu_baseline": None}))
    if ws > 1:
        dist.destroy_process_group()
    return 0


def run_ours(args):
    ws, rank, lr = _dist()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the omgsr_b200 hot path has no CPU fallback")
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    import torch.distributed as dist
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    from infer.omgsr_s_infer_model import OMGSR_S_Infer
    from omgsr_b200 import capi

    B, H, W_ = args.batch, 512, 512
    with _quiet():
        model = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, dev, torch.bfloat16)
    g = torch.Generator().manual_seed(1234 + rank)
    n_buf = 3  # rotate input buffers; the per-step activation working set (GBs) is >> the 126 MB L2 anyway
    host_in = [torch.rand(B, 3, H, W_, generator=g).mul_(2).sub_(1).bfloat16().pin_memory() for _ in range(n_buf)]
    dev_in = [t.to(dev) for t in host_in]
    ctx = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(1235)).to(dev).bfloat16()
    host_out = torch.empty(B, 3, H, W_, dtype=torch.bfloat16).pin_memory()

    def step(i):
        with torch.no_grad():
            z = model._encode(dev_in[i % n_buf])
            return model._forward_no_tile(z, ctx)

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(lr) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    launches = capi.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if ws > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = ws * B / (ms_per_step * 1e-3)

    # ---- e2e through the public module call, pinned host buffers inside the timed region ----
    def e2e_step(i):
        with torch.no_grad(), _quiet():
            x = host_in[i % n_buf].to(dev, non_blocking=True)
            out, _ = model(x, ctx, 64, 32)
            host_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if ws > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_val = ws * B * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel class, timed live on its launch stream ----
    burst, sustained, hbm, how = _peaks()
    roof = dominant_kernel_roofline(capi, dev, B, burst, how)
    step_tflops = S512_FLOP_PER_IMAGE * B / (ms_per_step * 1e-3) / 1e12

    if rank == 0:
        cpu = None
        if ws == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            times = cpu_reference_images_per_s(1, threads)
            cpu = {"value": 1.0 / times[0], "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": "one 512x512 image (batch 1), fp32, oracle port of the reference path, "
                             f"{times[0]:.1f} s, no warm-up"}
        line = {
            "metric": "SR images/sec (OMGSR-S 128->512)", "value": value, "unit": "images/s", "n_gpus": ws,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "parallelism": f"image-parallel x{ws}",
                       "l2": "3 rotating input buffers; per-step activations (>1 GB tensors) exceed the 126 MB L2",
                       "step_tflops_algorithmic": step_tflops,
                       "step_frac_of_sustained_peak": step_tflops / sustained},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W_ * 2,
                    "d2h_bytes_per_step": B * 3 * H * W_ * 2},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()
    return 0


HALO_GN_TRAFFIC = 2.1099e9  # bytes per launch, updated from profiles/r01/ncu_halo_pair128_gn.txt


def dominant_kernel_roofline(capi, dev, B, peak_tflops, how):
    """the launch class that dominates the step: the VAE's GroupNorm+SiLU -> 3x3 conv 128->128 at 512x512 (11 launches
    per step, 1.24 TFLOP each at batch 16), served by conv_halo_tc2_kernel<2,true> (CTA-pair input-halo kernel with the
    normalisation applied to the operand tile in shared memory). Algorithmic FLOPs = 2*Cin*Cout*9*B*H*W."""
    from omgsr_b200.pack import pack_conv_weight
    C, H = 128, 512
    x = torch.randn(B, H, H, C, device=dev, dtype=torch.bfloat16)
    w = pack_conv_weight(torch.randn(C, C, 3, 3) / 34.0).to(dev)
    b = torch.zeros(C, device=dev)
    coef = capi.groupnorm_coef(capi.groupnorm_stats(x, 32), torch.ones(C, device=dev), torch.zeros(C, device=dev), B,
                               H * H, 32, 1e-6)
    out = torch.empty_like(x)
    # the burst peak is a cold-start figure: let the board leave the power-capped state of the timed steps first
    torch.cuda.synchronize()
    time.sleep(3.0)
    for _ in range(3):
        capi.conv2d_gn(x, coef, w, b, out=out)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        capi.conv2d_gn(x, coef, w, b, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * C * C * 9 * B * H * H
    ach = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "conv_halo_tc2_kernel<2,true> (GroupNorm+SiLU -> 3x3, 128->128 ch, 512x512, "
                                         "batch %d)" % B,
            "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
            # dram__bytes_read.sum + dram__bytes_write.sum of this very launch from one `ncu --set full` capture
            # (profiles/r01/ncu_halo_pair128_gn.txt); algorithmic input + output = 2 x 1.0737 GB
            "traffic": HALO_GN_TRAFFIC if (B, C, H) == (16, 128, 512) else None, "traffic_unit": "bytes/launch",
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops burst ({how})", "ms_per_launch": ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="s512", choices=["s512", "f1024"],
                    help="s512 = BASELINE configs[1] (default, the N=1 line); f1024 = configs[3] (OMGSR-F-1024, batch 4/GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "f1024":
        return run_ours_f1024(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
