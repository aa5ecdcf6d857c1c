"""Oracle restatement of the LoRA load + merge step (TEST INFRASTRUCTURE ONLY).

The reference does `PeftModel.from_pretrained(...)` then `merge_and_unload()` (infer/omgsr_s_infer_model.py:16-23,
infer/omgsr_f_infer_model.py:115-118) with peft==0.17.0 (requirements.txt:7), which is neither under /root/reference
nor installable offline, and no second implementation of peft exists in this image (searched site-packages: only
vllm's runtime LoRA ops and transformers' peft *integration*, which imports peft). PARITY UNPINNED for this step.
What is restated here is PEFT's published merge rule, written independently of omgsr_b200.weights.merge_lora:

    Linear:  W <- W + (lora_alpha / r) * (B @ A)                      A [r, in],       B [out, r]
    Conv2d:  W <- W + (lora_alpha / r) * conv-composition(B, A)       A [r, Cin, k, k], B [Cout, r, 1, 1]
             i.e. dW[o, i, :, :] = sum_r B[o, r, 0, 0] * A[r, i, :, :]

with lora_alpha = r in both OMGSR configs (train/train_omgsr_s.py:74-79, configs/omgsr_s_512.yml:8-9,
configs/omgsr_f_1024.yml:8-9) so the factor is 1. The definition is anchored by a behavioural test
(tests/test_oracle_cpu.py::test_lora_merge_is_adapter_forward): base(x) + B(A(x)) == merged(x) for both layer kinds.
"""
import torch
import torch.nn.functional as F


def merge_lora(sd, adapter, lora_alpha=None):
    """adapter: {module path: (A, B)}; returns a new state dict with the deltas folded into `<module>.weight`."""
    out = dict(sd)
    for mod, (a, b) in adapter.items():
        r = a.shape[0]
        scale = 1.0 if lora_alpha is None else float(lora_alpha) / r
        w = sd[mod + ".weight"]
        a32, b32 = a.to(torch.float32), b.to(torch.float32)
        if w.dim() == 4:
            # composing a kxk conv (A) with a 1x1 conv (B) is a matrix product over the rank axis at every tap
            k = a32.shape[-1]
            delta = (b32.reshape(b32.shape[0], r) @ a32.reshape(r, -1)).reshape(w.shape[0], w.shape[1], k, k)
        else:
            delta = b32 @ a32
        out[mod + ".weight"] = (w.to(torch.float32) + scale * delta).to(w.dtype)
    return out


def adapter_forward(x, w, bias, a, b, stride=1, padding=0):
    """what a PEFT-wrapped layer computes BEFORE merging: base(x) + B(A(x)) (scale 1, dropout 0)."""
    if w.dim() == 4:
        return F.conv2d(x, w, bias, stride=stride, padding=padding) + \
            F.conv2d(F.conv2d(x, a, None, stride=stride, padding=padding), b, None)
    return F.linear(x, w, bias) + F.linear(F.linear(x, a), b)
