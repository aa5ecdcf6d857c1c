import torch


def bf16r(t):
    """round-trip through bf16 so oracle and kernel see identical inputs"""
    return t.to(torch.bfloat16).float()


def report(name, got, ref, tol=1e-2):
    """max-abs error relative to max|ref|; prints a localisation summary on failure."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    rel = err.max().item() / scale
    ok = bool(torch.isfinite(got).all()) and rel <= tol
    if not ok:
        bad = (err > tol * scale) | ~torch.isfinite(got)
        idx = bad.nonzero()
        print(f"\n[{name}] FAIL rel={rel:.3e} scale={scale:.3e} bad={int(bad.sum())}/{bad.numel()} "
              f"nan={int((~torch.isfinite(got)).sum())}")
        print("  first bad idx:", idx[:8].tolist())
        print("  last bad idx :", idx[-4:].tolist())
        for d in range(got.dim()):
            other = [i for i in range(got.dim()) if i != d]
            cnt = bad.sum(dim=other)
            nz = cnt.nonzero().flatten()
            print(f"  dim{d}: {len(nz)}/{got.shape[d]} indices bad; first {nz[:12].tolist()}")
        flat = idx[:4]
        for i in flat:
            t = tuple(i.tolist())
            print(f"   at {t}: got {got[t].item():.5f} ref {ref[t].item():.5f}")
    else:
        print(f"[{name}] ok rel={rel:.3e}")
    return ok, rel
