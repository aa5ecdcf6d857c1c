"""CLI driver for OMGSR-F (mirror of the reference's infer/infer_omgsr_f.py; completed in the next step)."""
import torch


def _prepare_latent_image_ids(height, width, device, dtype):
    """[height*width, 3] position ids (0, row, col) for the packed latent grid (reference :17-28)."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] += torch.arange(height)[:, None]
    ids[..., 2] += torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)
