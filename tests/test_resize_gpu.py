"""GPU pre/post-processing kernels against PIL / the torch expressions of the reference driver
(infer/infer_omgsr_s.py:76-98). Integer work: bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

CASES = [(32, 48, 128, 192, 0), (67, 45, 268, 180, 0), (133, 90, 128, 88, 1), (64, 64, 61, 37, 1), (50, 70, 23, 31, 0),
         (40, 40, 40, 64, 0), (128, 128, 512, 512, 0), (515, 509, 512, 504, 1), (256, 256, 1024, 1024, 0)]
PIL_FILTER = {0: Image.BICUBIC, 1: Image.LANCZOS}


@pytest.fixture(scope="module")
def capi():
    from omgsr_b200 import capi as c
    return c


@pytest.mark.parametrize("H,W,h,w,flt", CASES)
def test_resize_u8_matches_pil(capi, H, W, h, w, flt):
    a = np.random.default_rng(H * 1000 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a, "RGB").resize((w, h), PIL_FILTER[flt]))
    got = capi.resize_u8(torch.from_numpy(a).cuda(), (h, w), flt).cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), f"max |diff| = {np.abs(got.astype(int) - ref.astype(int)).max()}"


def test_resize_u8_matches_oracle_full_chain(capi):
    """the whole driver chain on one seeded image: x4 BICUBIC, then LANCZOS to a multiple of 8"""
    from oracle import resize as oresize
    a = np.random.default_rng(3).integers(0, 256, (131, 127, 3), dtype=np.uint8)
    o = oresize.resize_u8(oresize.resize_u8(a, (524, 508), 0), (520, 504), 1)
    g = capi.resize_u8(capi.resize_u8(torch.from_numpy(a).cuda(), (524, 508), 0), (520, 504), 1)
    assert np.array_equal(g.cpu().numpy(), o)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_u8_to_input(capi, dtype):
    a = torch.from_numpy(np.random.default_rng(5).integers(0, 256, (72, 40, 3), dtype=np.uint8)).cuda()
    # torchvision to_tensor: HWC u8 -> CHW float / 255; then the reference's .to(dtype) * 2 - 1
    ref = (a.permute(2, 0, 1).float().div(255).unsqueeze(0).to(dtype) * 2 - 1)
    got = capi.u8_to_input(a, dtype)
    assert got.dtype == dtype and torch.equal(got, ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_output_to_u8(capi, dtype):
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(3, 64, 48, generator=g) * 2.4 - 1.2).to(dtype).cuda()  # includes values outside [-1, 1]
    ref = torch.clip(x * 0.5 + 0.5, 0, 1).float().mul(255).byte().permute(1, 2, 0)  # ToPILImage semantics
    got = capi.output_to_u8(x)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("size,process_size", [((96, 80), 512), ((131, 127), 256), ((40, 52), 512)])
def test_driver_gpu_chain_equals_pil_chain(size, process_size):
    """infer/_driver_common.py: preprocess_gpu == preprocess + pil_to_model_input, finish_gpu == finish"""
    from infer import _driver_common as dc
    a = np.random.default_rng(size[0]).integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)
    img = Image.fromarray(a, "RGB")
    lq_pil, flag_h, orig_h = dc.preprocess(img, process_size, 4)
    x_h = dc.pil_to_model_input(lq_pil, "cuda", torch.bfloat16)
    lq_u8, x_g, flag_g, orig_g = dc.preprocess_gpu(img, process_size, 4, "cuda", torch.bfloat16)
    assert (flag_h, orig_h) == (flag_g, orig_g)
    assert np.array_equal(lq_u8.cpu().numpy(), np.asarray(lq_pil))
    assert torch.equal(x_h, x_g)
    g = torch.Generator().manual_seed(1)
    out = (torch.rand(1, 3, x_g.shape[2], x_g.shape[3], generator=g) * 2.2 - 1.1).bfloat16().cuda()
    for method in ("adain", "wavelet", "nofix"):
        p_h = dc.finish(out, lq_pil, method, flag_h, orig_h, 4, None)
        p_g = dc.finish_gpu(out, lq_u8, method, flag_g, orig_g, 4, None)
        assert p_h.size == p_g.size
        assert np.array_equal(np.asarray(p_h), np.asarray(p_g)), method
