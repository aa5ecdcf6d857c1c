"""State-dict construction for the nets on the OMGSR inference path: synthetic specs (re-exported from `omgsr_synth`)
plus the PEFT-style LoRA merge the product path applies at load time.

`merge_lora` reproduces PEFT's `merge_and_unload` (W <- W + (alpha/r) * B A, alpha = r; reference:
infer/omgsr_s_infer_model.py:16-23, infer/omgsr_f_infer_model.py:115-118). The oracle carries its own restatement
(oracle/lora.py); tests/test_oracle_cpu.py checks that the two agree.
"""
import torch

from omgsr_synth import *  # noqa: F401,F403  (configs, shape tables, random init, ddpm_alphas_cumprod)
from omgsr_synth import _conv, _lin, _norm  # noqa: F401


def merge_lora(sd, adapter, scale=1.0):
    """W <- W + scale * B A  (PEFT merge_and_unload; lora_alpha = r => scale 1). Done in fp32 as the
    reference's OMGSR-S path does (merge before the bf16 cast, infer/omgsr_s_infer_model.py:22-25)."""
    out = dict(sd)
    for mod, (a, b) in adapter.items():
        w = sd[mod + ".weight"].float()
        if w.dim() == 4:
            delta = torch.einsum("or,rikl->oikl", b.float()[:, :, 0, 0], a.float())
        else:
            delta = b.float() @ a.float()
        out[mod + ".weight"] = (w + scale * delta).to(sd[mod + ".weight"].dtype)
    return out
