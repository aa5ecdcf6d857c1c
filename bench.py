#!/usr/bin/env python
"""bench.py — OMGSR one-step SR throughput on B200 (BASELINE.json metric: SR images/s at 1024^2 (OMGSR-F) & 512^2
(OMGSR-S), 1/2/4/8 B200).

Main line (N=1): BASELINE.json configs[1] — OMGSR-S-512, bf16, batch 16 per GPU: VAE encode -> UNet forward at the
mid-timestep -> x0 step -> VAE decode, SD-2.1 shapes, random-init weights with merged LoRA, synthetic LQ images.
The same JSON line carries an `f1024` sub-object (BASELINE configs[3]: OMGSR-F-1024, FLUX.1-dev shapes, batch 4 per
GPU) with its own value / e2e / roofline / cpu_baseline, and for both workloads a `gpu_eager_baseline`: the plain-PyTorch
restatement of the reference path (oracle/) run in bf16 on the SAME GPU through cuDNN / cuBLAS / SDPA — the reference's
own GPU dataflow (infer/omgsr_s_infer_model.py:75-85,173; infer/omgsr_f_infer_model.py:174-212) and the bar
SURVEY.md §2.1 names. N>1: image-parallel replicas (one process per GPU, no data-path collective) => weak scaling.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (S-512 line + f1024 sub-object)
    python bench.py --workload s512|f1024|c5               # one workload only (c5 = F 512->2048, one image, N ranks)
    python bench.py --impl reference ...                   # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0). `value` = device-timed images/s with inputs resident in HBM; `e2e` = the same
through the public module call with pinned-host inputs/outputs inside the timed region.

oracle/ is imported ONLY by the baseline legs (cpu_baseline, gpu_eager_baseline, --impl reference); the reference arm
imports nothing from omgsr_b200/ or infer/ (no CUDA library is loaded in that process).
"""
import argparse
import contextlib
import gc
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
F2048_FLOP_PER_IMAGE = 744.1e12  # enc 26.127 + 9 x 74.3847 + dec 48.486 TFLOP
WORKLOAD = "OMGSR-S-512 bf16 batch 16 (128->512 SR: VAE enc + UNet@t=273 + VAE dec, SD-2.1 shapes)"
WORKLOAD_F = "OMGSR-F-1024 bf16 batch 4 per GPU (256->1024 SR: FLUX VAE enc + FLUX.1-dev-shape DiT@sigma=0.505 + VAE dec)"
WORKLOAD_C5 = "OMGSR-F 512->2048 single image, 9 latent tiles + tiled VAE, spatially sharded over the ranks"


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


def _free():
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------
# baselines: the reference algorithm (oracle port; diffusers/peft are not installable offline)
#   * on the host CPU cores in fp32 (cpu_baseline / --impl reference)
#   * on the same GPU in bf16 through PyTorch eager: cuDNN / cuBLAS / SDPA (gpu_eager_baseline)
# Nothing below imports omgsr_b200/ or infer/.
# ----------------------------------------------------------------------------------------------------
def _oracle_s(device="cpu", dtype=torch.float32):
    import omgsr_synth as S
    from oracle.omgsr_s import OracleS
    from oracle.synth import s_state
    vae_sd, unet_sd, vae_cfg, unet_cfg = s_state("sd21", seed=0, lora_seed=1)
    o = OracleS(vae_sd, unet_sd, vae_cfg, unet_cfg, 273, S.ddpm_alphas_cumprod())
    return o.to(device, dtype) if (str(device) != "cpu" or dtype != torch.float32) else o


def _s_inputs(B, device="cpu", dtype=torch.float32):
    lq = torch.rand(B, 3, 512, 512, generator=torch.Generator().manual_seed(1234)) * 2 - 1
    ctx = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(1235))
    eps = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(1236))
    return lq.to(device, dtype), ctx.to(device, dtype), eps.to(device, dtype)


def cpu_reference_s(n_images, threads, budget_s=1e9):
    """per-image seconds of the fp32 oracle port on the host: one 512x512 image (batch 1) per step."""
    torch.set_num_threads(threads)
    oracle = _oracle_s()
    lq, ctx, eps = _s_inputs(1)
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


def cpu_reference_f_sample(threads, reps=1):
    """Bounded CPU sample of OMGSR-F-1024 (fp32 oracle port): the full FLUX VAE encoder + decoder on one 1024^2 image,
    plus ONE double and ONE single DiT block at FLUX.1-dev width on the 4608-token joint sequence (timed through
    oracle.nets.flux_forward with a 1 + 1 block config, embedders included), scaled to 19 + 38 blocks. The whole DiT
    in fp32 would need 47.6 GB of host weights and ~90 TFLOP per image."""
    import omgsr_synth as S
    from oracle import nets
    torch.set_num_threads(threads)
    vae_cfg = S.FLUX_VAE
    vae_sd = S.random_state_dict(S.vae_shapes(vae_cfg), 0)
    g = torch.Generator().manual_seed(1234)
    lq = torch.rand(1, 3, 1024, 1024, generator=g) * 2 - 1
    eps = torch.randn(1, 16, 128, 128, generator=g)

    def flux_time(nd, ns):
        cfg = S.FluxConfig(num_layers=nd, num_single_layers=ns)
        sd = S.random_state_dict(S.flux_shapes(cfg), 1)
        x = torch.randn(1, 4096, 64, generator=g)
        pe, pp = torch.randn(1, 512, 4096, generator=g), torch.randn(1, 768, generator=g)
        ids = torch.zeros(4096, 3)
        best = 1e9
        for _ in range(reps):
            t0 = time.time()
            nets.flux_forward(sd, cfg, x, torch.tensor([0.5051]), torch.ones(1), pp, pe, torch.zeros(512, 3), ids)
            best = min(best, time.time() - t0)
        return best

    with torch.no_grad():
        t0 = time.time()
        z = nets.posterior_sample(nets.vae_encode_moments(vae_sd, vae_cfg, lq), eps)
        t_enc = time.time() - t0
        t0 = time.time()
        nets.vae_decode(vae_sd, vae_cfg, z)
        t_dec = time.time() - t0
        t_00 = flux_time(0, 0)   # embedders + output layer only
        t_10 = flux_time(1, 0)
        t_01 = flux_time(0, 1)
    t_double, t_single = max(t_10 - t_00, 0.0), max(t_01 - t_00, 0.0)
    total = t_enc + t_dec + t_00 + 19 * t_double + 38 * t_single
    return total, {"vae_enc_s": t_enc, "vae_dec_s": t_dec, "double_block_s": t_double, "single_block_s": t_single,
                   "embed_out_s": t_00}


def gpu_eager_s(dev, B, steps=5, warmup=3):
    """PyTorch-eager bf16 on this GPU: oracle.OracleS (== the reference's dataflow, infer/omgsr_s_infer_model.py:173,
    75-85, over a plain-torch AutoencoderKL / UNet) with cudnn.benchmark on. Returns images/s at batch B and batch 1."""
    torch.backends.cudnn.benchmark = True
    oracle = _oracle_s(dev, torch.bfloat16)
    out = {}
    with torch.no_grad():
        for b in (B, 1):
            lq, ctx, eps = _s_inputs(b, dev, torch.bfloat16)
            for _ in range(warmup):
                oracle.forward(lq, ctx, 64, 32, eps)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                oracle.forward(lq, ctx, 64, 32, eps)
            e1.record()
            torch.cuda.synchronize()
            out[b] = e0.elapsed_time(e1) / steps
    del oracle
    _free()
    return {"value": B / (out[B] * 1e-3), "unit": "images/s", "ms_per_step": out[B], "batch": B,
            "value_batch1": 1.0 / (out[1] * 1e-3), "ms_batch1": out[1], "dtype": "bf16", "steps": steps,
            "impl": "torch-eager cuDNN/cuBLAS/SDPA (oracle.OracleS on cuda, cudnn.benchmark=True)"}


def gpu_eager_f(dev, B, steps=3, warmup=2):
    """PyTorch-eager bf16 OMGSR-F-1024 on this GPU: oracle.OracleF (infer/omgsr_f_infer_model.py:15-18,174-212)."""
    import omgsr_synth as S
    from oracle.omgsr_f import OracleF
    from oracle.synth import f_state
    torch.backends.cudnn.benchmark = True
    vae_sd, flux_sd, vae_cfg, flux_cfg = f_state("flux_dev", 0, 1, 64, torch.bfloat16, dev)
    oracle = OracleF(vae_sd, flux_sd, vae_cfg, flux_cfg).to(dev, torch.bfloat16)
    g = torch.Generator().manual_seed(1234)
    lq = (torch.rand(B, 3, 1024, 1024, generator=g) * 2 - 1).to(dev, torch.bfloat16)
    eps = torch.randn(B, 16, 128, 128, generator=g).to(dev, torch.bfloat16)
    pe = torch.randn(1, 512, 4096, generator=g).to(dev, torch.bfloat16)
    pp = torch.randn(1, 768, generator=g).to(dev, torch.bfloat16)
    tid = torch.zeros(512, 3, device=dev)
    ids = torch.zeros(64, 64, 3, device=dev)
    ids[..., 1] += torch.arange(64, device=dev)[:, None]
    ids[..., 2] += torch.arange(64, device=dev)[None, :]
    ids = ids.reshape(4096, 3)
    with torch.no_grad():
        for _ in range(warmup):
            oracle.forward(lq, pe, pp, tid, ids, 128, 64, eps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            oracle.forward(lq, pe, pp, tid, ids, 128, 64, eps)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del oracle, vae_sd, flux_sd
    _free()
    return {"value": B / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "batch": B, "dtype": "bf16", "steps": steps,
            "impl": "torch-eager cuDNN/cuBLAS/SDPA (oracle.OracleF on cuda, cudnn.benchmark=True)"}


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port, fp32) on all host cores. Rank 0 only."""
    ws, rank, _ = _dist()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    if args.workload == "f1024":
        total, parts = cpu_reference_f_sample(threads)
        val = 1.0 / total
        sample = ("one 1024x1024 image: full FLUX VAE enc + dec, 1 double + 1 single DiT block at FLUX.1-dev width on "
                  "4608 tokens scaled to 19 + 38 blocks, fp32, oracle port")
        line = {
            "impl": "reference", "metric": "SR images/sec (OMGSR-F 256->1024)", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": total * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_F, "sample": sample, "parts_s": parts},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0
    # one step = one 512x512 image at batch 1 (a bounded sample of the batch-16 workload); the run is
    # capped at ~4 minutes of CPU time and reports the steps it completed
    times = cpu_reference_s(args.warmup + args.steps, threads, budget_s=240.0)
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
class _Ctx:
    """per-process distributed context (one process per GPU)."""

    def __init__(self):
        self.ws, self.rank, self.lr = _dist()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the omgsr_b200 hot path has no CPU fallback")
        torch.cuda.set_device(self.lr)
        self.dev = torch.device("cuda", self.lr)
        import torch.distributed as dist
        self.dist = dist
        if self.ws > 1 and not dist.is_initialized():
            # NCCL prints its version banner on STDOUT when the first communicator comes up: keep stdout to the one JSON
            # line the driver parses (the banner goes to stderr instead)
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                t = torch.zeros(1, device=self.dev)
                dist.all_reduce(t)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        if self.ws > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_(self, *vals):
        if self.ws == 1:
            return vals if len(vals) > 1 else vals[0]
        t = torch.tensor(list(vals), device=self.dev, dtype=torch.float32)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = tuple(float(v) for v in t.tolist())
        return out if len(out) > 1 else out[0]

    def close(self):
        if self.ws > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


def _timed(ctx, step, steps, clocks_index=None):
    """K device-timed steps bracketed by barrier + synchronize; max over ranks. Returns (ms_total, clocks summary)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(ctx.lr) as clocks:
        ctx.barrier()
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        ctx.barrier()
    return ctx.max_(e0.elapsed_time(e1)), clocks.summary()


def bench_f1024(ctx, args, with_baselines=True):
    """BASELINE config 4: OMGSR-F-1024, batch 4 per GPU, image-parallel replicas."""
    from infer.infer_omgsr_f import _prepare_latent_image_ids
    from infer.omgsr_f_infer_model import OMGSR_F_Infer
    from omgsr_b200 import capi
    dev, ws, rank = ctx.dev, ctx.ws, ctx.rank
    B = args.batch_f
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

    steps = max(3, min(args.steps, args.steps_f))
    for i in range(3):
        step(i)
    ctx.barrier()
    l0 = capi.launch_count()
    ms, clocks = _timed(ctx, step, steps)
    launches = capi.launch_count() - l0
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        out = step(i, host_in[i % 2].to(dev, non_blocking=True))
        host_out.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
    ctx.barrier()
    e2e_ms = ctx.max_((time.perf_counter() - t0) * 1e3)
    ms_per_step = ms / steps
    burst, sustained, hbm, how = _peaks()
    step_tflops = F1024_FLOP_PER_IMAGE * B / (ms_per_step * 1e-3) / 1e12
    roof = f_kernel_roofline(capi, dev, B, burst, how) if rank == 0 else None
    del model, dev_in
    _free()
    res = {
        "metric": "SR images/sec (OMGSR-F 256->1024)", "value": ws * B / (ms_per_step * 1e-3), "unit": "images/s",
        "n_gpus": ws, "steps": steps, "warmup": 3, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD_F, "batch_per_gpu": B, "parallelism": f"image-parallel x{ws}",
                   "l2": "2 rotating input buffers; 24 GB of weights + GB-sized activations per step >> L2",
                   "step_tflops_algorithmic": step_tflops, "step_frac_of_sustained_peak": step_tflops / sustained},
        "clocks": clocks,
        "e2e": {"value": ws * B * steps / (e2e_ms * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": B * 3 * 1024 * 1024 * 2, "d2h_bytes_per_step": B * 3 * 1024 * 1024 * 2},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": None, "gpu_eager_baseline": None,
    }
    if rank == 0 and with_baselines:
        if not args.no_eager:
            try:
                res["gpu_eager_baseline"] = gpu_eager_f(dev, B)
            except Exception as ex:  # a baseline that fails must not take the measured line with it
                res["gpu_eager_baseline"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
        if ws == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            total, parts = cpu_reference_f_sample(threads)
            res["cpu_baseline"] = {
                "value": 1.0 / total, "unit": "images/s", "cores": threads, "kind": "port",
                "sample": "one 1024x1024 image: full FLUX VAE enc + dec (fp32) + 1 double + 1 single DiT block at "
                          "FLUX.1-dev width on 4608 tokens, scaled to 19 + 38 blocks; oracle port, no warm-up",
                "parts_s": parts}
    return res


# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/r02/)
HALO_GN_TRAFFIC = 2.1171e9   # conv_halo_rr_kernel<1> (GroupNorm+SiLU -> 3x3 128->128 @512^2 x16): 1.0882 GB read + 1.0289 GB
                             # written, profiles/r02/ncu_r02_halo_rr_gn_final.txt (algorithmic: input + output = 2 x 1.0737 GB)
F_GEMM_TRAFFIC = 2.4589e9    # conv_gemm_tc2_kernel<256> (18432 x 12288 x 3072 + GELU): 2.0158 GB read + 0.4431 GB written,
                             # profiles/r02/ncu_r02b_f_gemm.txt (algorithmic A + W + D = 0.642 GB; the 75 MB weight
                             # matrix is re-streamed per 256-row tile pair and only partly stays in L2)


def f_kernel_roofline(capi, dev, B, peak_tflops, how):
    """dominant kernel of the F-1024 step: the DiT's 3072 -> 12288 GELU-tanh linear (ff.net.0 / proj_mlp: 57 + 19
    launches per step) on the joint sequence, conv_gemm_tc2_kernel<256>. Algorithmic FLOPs = 2*M*N*K."""
    M, N, K = B * 4608, 12288, 3072
    a = torch.randn(B, 4608, K, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / 55.0).bfloat16()
    bias = torch.zeros(N, device=dev)
    out = torch.empty(B, 4608, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        capi.gemm_batched(a, w, bias, act=capi.ACT_GELU_TANH, out=out)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        capi.gemm_batched(a, w, bias, act=capi.ACT_GELU_TANH, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ach = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    # second kernel class: the joint attention (24 heads, d = 128, 4608 tokens), 57 launches per step
    q = torch.randn(B, 4608, 3 * 3072, device=dev, dtype=torch.bfloat16)
    qq, kk, vv = (q[..., j * 3072:(j + 1) * 3072].unflatten(-1, (24, 128)) for j in range(3))
    o = torch.empty(B, 4608, 24, 128, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        capi.flash_attn(qq, kk, vv, 128 ** -0.5, out=o)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        capi.flash_attn(qq, kk, vv, 128 ** -0.5, out=o)
    e1.record()
    torch.cuda.synchronize()
    ms_a = e0.elapsed_time(e1) / n
    ach_a = 4.0 * B * 24 * 4608 * 4608 * 128 / (ms_a * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "conv_gemm_tc2_kernel<256> (DiT linear 3072 -> 12288 + GELU-tanh, M = %d)" % M,
            "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
            "traffic": F_GEMM_TRAFFIC if B == 4 else None, "traffic_unit": "bytes/launch",
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops burst ({how})", "ms_per_launch": ms,
            "attention": {"kernel": "flash_attn5_kernel<128> (joint attention, 24 heads x 4608 tokens, d = 128, batch %d)" % B,
                          "achieved": ach_a, "frac": ach_a / peak_tflops, "ms_per_launch": ms_a}}


def bench_s512(ctx, args):
    from infer.omgsr_s_infer_model import OMGSR_S_Infer
    from omgsr_b200 import capi
    dev, ws, rank = ctx.dev, ctx.ws, ctx.rank
    B, H, W_ = args.batch, 512, 512
    with _quiet():
        model = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, dev, torch.bfloat16)
    g = torch.Generator().manual_seed(1234 + rank)
    n_buf = 3  # rotate input buffers; the per-step activation working set (GBs) is >> the 126 MB L2 anyway
    host_in = [torch.rand(B, 3, H, W_, generator=g).mul_(2).sub_(1).bfloat16().pin_memory() for _ in range(n_buf)]
    dev_in = [t.to(dev) for t in host_in]
    ctxt = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(1235)).to(dev).bfloat16()
    host_out = torch.empty(B, 3, H, W_, dtype=torch.bfloat16).pin_memory()

    def step(i):
        with torch.no_grad():
            z = model._encode(dev_in[i % n_buf])
            return model._forward_no_tile(z, ctxt)

    for i in range(max(args.warmup, 3)):
        step(i)
    ctx.barrier()
    l0 = capi.launch_count()
    ms, clocks = _timed(ctx, step, args.steps)
    launches = capi.launch_count() - l0
    ms_per_step = ms / args.steps
    value = ws * B / (ms_per_step * 1e-3)

    # ---- e2e through the public module call, pinned host buffers inside the timed region ----
    def e2e_step(i):
        with torch.no_grad(), _quiet():
            x = host_in[i % n_buf].to(dev, non_blocking=True)
            out, _ = model(x, ctxt, 64, 32)
            host_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
    for i in range(2):
        e2e_step(i)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    ctx.barrier()
    e2e_ms = ctx.max_((time.perf_counter() - t0) * 1e3)
    e2e_val = ws * B * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel class, timed live on its launch stream ----
    burst, sustained, hbm, how = _peaks()
    roof = dominant_kernel_roofline(capi, dev, B, burst, how)
    step_tflops = S512_FLOP_PER_IMAGE * B / (ms_per_step * 1e-3) / 1e12
    del model, dev_in
    _free()
    line = {
        "metric": "SR images/sec (OMGSR-S 128->512)", "value": value, "unit": "images/s", "n_gpus": ws,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "parallelism": f"image-parallel x{ws}",
                   "l2": "3 rotating input buffers; per-step activations (>1 GB tensors) exceed the 126 MB L2",
                   "step_tflops_algorithmic": step_tflops,
                   "step_frac_of_sustained_peak": step_tflops / sustained},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W_ * 2,
                "d2h_bytes_per_step": B * 3 * H * W_ * 2},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": None, "gpu_eager_baseline": None,
    }
    if rank == 0:
        if not args.no_eager:
            try:
                line["gpu_eager_baseline"] = gpu_eager_s(dev, B)
            except Exception as ex:
                line["gpu_eager_baseline"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
        if ws == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            times = cpu_reference_s(1, threads)
            line["cpu_baseline"] = {"value": 1.0 / times[0], "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": "one 512x512 image (batch 1), fp32, oracle port of the reference path, "
                                              f"{times[0]:.1f} s, no warm-up"}
    return line


PAIR256_TRAFFIC = 0.5127e9   # conv_gemm_tc2_kernel<256> (3x3 512->512 @128^2 x16): 0.2732 GB read + 0.2394 GB written,
                             # profiles/r02/ncu_r02c_conv512.txt (algorithmic input + output = 2 x 0.2684 GB)


def _time_launch(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def dominant_kernel_roofline(capi, dev, B, peak_tflops, how):
    """Roofline of the step's dominant kernel, timed live on its launch stream.

    By share of the step (ncu launch list, profiles/r02/launches_s512_b16_final_summary.txt) the dominant KERNEL at the end
    of round 2 is conv_halo_rr_kernel<1> - the rolling-row 3x3 convolution with the GroupNorm + SiLU of its input fused
    into the operand path (32 % of the step in 21 launches since it also serves the 256-channel level) - probed on its
    largest recurring shape, 128->128 channels at 512x512 (1.24 TFLOP per launch at batch 16; this is also the kernel
    round 1 reported and its verdict asked to lift). Second: conv_gemm_tc2_kernel<256> (CTA-pair implicit-GEMM conv /
    GEMM, N = 256 tiles: 28 % in 72 launches), probed on the VAE's 3x3 512->512 convolution at 128x128
    (1.24 TFLOP per launch; its 256-channel layers moved to the rolling-row kernel), reported as `second`.
    Algorithmic FLOPs = 2*Cin*Cout*9*B*H*W in both cases."""
    from omgsr_b200.pack import pack_conv_weight
    # the burst peak is a cold-start figure: let the board leave the power-capped state of the timed steps first
    torch.cuda.synchronize()
    time.sleep(3.0)
    out = {}
    for key, C, H in (("pair256", 512, 128), ("halo", 128, 512)):
        x = torch.randn(B, H, H, C, device=dev, dtype=torch.bfloat16)
        w = pack_conv_weight(torch.randn(C, C, 3, 3) / (3.0 * C ** 0.5)).to(dev)
        b = torch.zeros(C, device=dev)
        y = torch.empty_like(x)
        if key == "halo":
            coef = capi.groupnorm_coef(capi.groupnorm_stats(x, 32), torch.ones(C, device=dev),
                                       torch.zeros(C, device=dev), B, H * H, 32, 1e-6)
            ms = _time_launch(lambda: capi.conv2d_gn(x, coef, w, b, out=y))
            label = "GroupNorm+SiLU -> 3x3, 128->128 ch, 512x512"
            traffic = HALO_GN_TRAFFIC if (B, C, H) == (16, 128, 512) else None
        else:
            ms = _time_launch(lambda: capi.conv2d(x, w, b, out=y))
            label = "3x3, 512->512 ch, 128x128"
            traffic = PAIR256_TRAFFIC if (B, C, H) == (16, 512, 128) else None
        flops = 2.0 * C * C * 9 * B * H * H
        ach = flops / (ms * 1e-3) / 1e12
        out[key] = {"bound": "tensor", "kernel": f"{capi.last_conv_variant()} ({label}, batch {B})", "achieved": ach,
                    "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this very launch from one `ncu --set full` capture
                    "traffic": traffic, "traffic_unit": "bytes/launch",
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops burst ({how})", "ms_per_launch": ms}
        del x, y
    roof = out["halo"]
    roof["share_of_step"] = ("32 % (conv_halo_rr_kernel<1, *>: fused input GroupNorm, 21 launches; "
                             "profiles/r02/launches_s512_b16_final_summary.txt)")
    roof["second"] = out["pair256"]
    roof["second"]["share_of_step"] = "28 % (conv_gemm_tc2_kernel<256>, 72 launches)"
    return roof


def bench_c5(ctx, args):
    """BASELINE config 5: OMGSR-F 512->2048, ONE image sharded spatially over all ranks (strong scaling)."""
    from infer.infer_omgsr_f import _prepare_latent_image_ids
    from infer.omgsr_f_infer_model import OMGSR_F_Infer
    from omgsr_b200 import capi
    from omgsr_b200 import dist as D
    dev, ws, rank = ctx.dev, ctx.ws, ctx.rank
    with _quiet():
        model = OMGSR_F_Infer("random://flux_dev?seed=0", "random://lora?seed=1&rank=64", dev)
        model._init_tiled_vae(encoder_tile_size=1024, decoder_tile_size=224)
    if ws > 1:
        D.enable_tile_parallel(True)
    g = torch.Generator().manual_seed(1234)
    x = (torch.rand(1, 3, 2048, 2048, generator=g) * 2 - 1).bfloat16()
    host_in = x.pin_memory()
    xd = x.to(dev)
    g2 = torch.Generator().manual_seed(1235)
    pe = torch.randn(1, 512, 4096, generator=g2).to(dev).bfloat16()
    pp = torch.randn(1, 768, generator=g2).to(dev).bfloat16()
    tid = torch.zeros(512, 3, device=dev, dtype=torch.bfloat16)
    ids = _prepare_latent_image_ids(64, 64, dev, torch.bfloat16)
    host_out = torch.empty(1, 3, 2048, 2048, dtype=torch.float32).pin_memory()

    def step(i, inp=None):
        with torch.no_grad(), _quiet():
            return model(xd if inp is None else inp, pe, pp, tid, ids, 128, 64)[0]

    steps = max(3, min(args.steps, 5))
    for i in range(3):
        step(i)
    ctx.barrier()
    l0 = capi.launch_count()
    ms, clocks = _timed(ctx, step, steps)
    launches = capi.launch_count() - l0
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        out = step(i, host_in.to(dev, non_blocking=True))
        if rank == 0:
            host_out.copy_(out.float(), non_blocking=True)
        torch.cuda.synchronize()
    ctx.barrier()
    e2e_ms = ctx.max_((time.perf_counter() - t0) * 1e3)
    ms_per_step = ms / steps
    burst, sustained, hbm, how = _peaks()
    tfl = F2048_FLOP_PER_IMAGE / (ms_per_step * 1e-3) / 1e12
    return {
        "metric": "SR images/sec (OMGSR-F 512->2048, single image)", "value": 1.0 / (ms_per_step * 1e-3),
        "unit": "images/s", "n_gpus": ws, "steps": steps, "warmup": 3, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD_C5, "parallelism": f"tile-parallel x{ws} (latent tiles + VAE tiles)",
                   "l2": "24 GB of weights + GB-sized activations per step >> L2",
                   "step_tflops_algorithmic": tfl, "step_frac_of_sustained_peak": tfl / (sustained * ws)},
        "clocks": clocks,
        "e2e": {"value": steps / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": 3 * 2048 * 2048 * 2,
                "d2h_bytes_per_step": 3 * 2048 * 2048 * 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "whole step over all ranks", "achieved": tfl, "peak": sustained * ws,
                     "unit": "TFLOP/s", "frac": tfl / (sustained * ws), "traffic": None,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained x {ws} ({how})"},
        "cpu_baseline": None,
    }


def run_ours(args):
    ctx = _Ctx()
    line = None
    if args.workload in ("all", "s512"):
        line = bench_s512(ctx, args)
    if args.workload in ("all", "f1024"):
        f = bench_f1024(ctx, args)
        if line is None:
            line = f
        else:
            line["f1024"] = f
    if args.workload == "c5":
        line = bench_c5(ctx, args)
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="OMGSR-S-512 images per GPU per step")
    ap.add_argument("--batch-f", type=int, default=4, help="OMGSR-F-1024 images per GPU per step")
    ap.add_argument("--steps-f", type=int, default=5, help="cap on the timed OMGSR-F-1024 steps (375 ms each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager bf16 GPU baseline")
    ap.add_argument("--workload", default=None, choices=["all", "s512", "f1024", "c5"],
                    help="all (default) = the S-512 line (BASELINE configs[1]) with an f1024 sub-object (configs[3]); "
                         "c5 = configs[4], OMGSR-F 512->2048 single image over all ranks")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "s512" if args.impl == "reference" else "all"
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
