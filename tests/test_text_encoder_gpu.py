"""Text-conditioning step: the CLIP text tower through the sm_100a kernels against the fp32 oracle
(oracle/clip_text.py, itself pinned to transformers' CLIPTextModel in tests/test_oracle_cpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from omgsr_b200 import capi as c
    return c


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("name", ["TINY_CLIP", "TINY_CLIP_QUICK", "SD21_CLIP_H", "FLUX_CLIP_L"])
def test_clip_text_engine_matches_oracle(name):
    """last_hidden_state (what infer_omgsr_s.py:19-45 feeds the UNet as prompt_embeds) and pooler_output (what
    FluxPipeline.encode_prompt feeds the DiT) at the real SD-2.1 / CLIP-L sizes: max-relative error <= 2e-2 against
    the fp32 oracle on the same bf16-rounded weights; ids = the empty prompt and a random prompt with an early EOS."""
    import omgsr_synth as syn
    from omgsr_b200.engine_clip import B200CLIPText
    from oracle import clip_text
    cfg = getattr(syn, name)
    sd = syn.random_state_dict(syn.clip_text_shapes(cfg), seed=11)
    sd = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}   # both sides see bf16-representable weights
    g = torch.Generator().manual_seed(5)
    rnd = torch.randint(1, cfg.eos_token_id - 1, (2, cfg.max_positions), generator=g)
    rnd[:, 9] = cfg.eos_token_id
    rnd[:, 10:] = cfg.pad_token_id
    ids = torch.cat([syn.empty_prompt_ids(cfg), rnd])
    eng = B200CLIPText(sd, cfg, "cuda")
    h, pooled = eng(ids)
    torch.cuda.synchronize()
    with torch.no_grad():
        sd_dev = {k: v.cuda() for k, v in sd.items()}
        torch.backends.cuda.matmul.allow_tf32 = False
        h_ref, pooled_ref = clip_text.clip_text_forward(sd_dev, cfg, ids.cuda())
    assert h.shape == (3, cfg.max_positions, cfg.hidden_size) and pooled.shape == (3, cfg.hidden_size)
    assert torch.isfinite(h.float()).all()
    e_h, e_p = _rel(h, h_ref), _rel(pooled, pooled_ref)
    print(f"\n[clip text {name}] last_hidden_state max-rel {e_h:.3e}, pooled {e_p:.3e}")
    assert e_h <= 2e-2 and e_p <= 2e-2
    # causality: changing a LATER token must not change earlier rows
    ids2 = ids.clone()
    ids2[:, 40] = 7
    h2, _ = eng(ids2)
    assert torch.equal(h2[:, :40], h[:, :40])


def test_softmax_rows_causal_and_embed(capi):
    g = torch.Generator().manual_seed(3)
    s = torch.randn(2, 3, 77, 128, generator=g).cuda()
    p = capi.softmax_rows(s, 77, 0.125, causal=True).float().cpu()
    mask = torch.full((77, 77), float("-inf")).triu(1)
    ref = torch.softmax(s.cpu()[..., :77] * 0.125 + mask, -1)
    assert float((p[..., :77] - ref).abs().max()) < 4e-3
    assert float(p[..., 77:].abs().max()) == 0.0
    assert float(p[..., :77].triu(1).abs().max()) == 0.0
    tok = torch.randn(100, 64, generator=g).bfloat16().cuda()
    pos = torch.randn(10, 64, generator=g).bfloat16().cuda()
    ids = torch.randint(0, 100, (4, 10), generator=g).to(torch.int32).cuda()
    out = capi.embed_tokens(ids, tok, pos)
    ref = (tok.float()[ids.long()] + pos.float()[None]).bfloat16()
    assert torch.equal(out, ref)
