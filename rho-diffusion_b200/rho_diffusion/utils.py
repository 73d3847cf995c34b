"""Host-side helpers on the hot path (reference ``rho_diffusion/utils.py``): process-group setup, checkpoint
save, sha512 label embeddings, parameter-space sampling and config number casting.  The plotting helpers of
the reference (utils.py:84-163) are visualisation and out of scope."""
from __future__ import annotations

import hashlib
import itertools
import json
import os

import numpy as np
import torch
import torch.distributed as dist


def ddp_setup(port: int = 29600, backend: str | None = None):
    """(world, rank, local_rank) over NCCL -- replaces the ccl/nccl set-up of utils.py:45-81.

    Rank discovery accepts both launchers: Intel-MPI style ``PMI_SIZE/PMI_RANK/MPI_LOCALRANKID`` (what the
    reference reads) and torchrun's ``WORLD_SIZE/RANK/LOCAL_RANK``.  A single process needs no group.
    """
    world = int(os.environ.get("PMI_SIZE", os.environ.get("WORLD_SIZE", 1)))
    rank = int(os.environ.get("PMI_RANK", os.environ.get("RANK", 0)))
    local_rank = int(os.environ.get("MPI_LOCALRANKID", os.environ.get("LOCAL_RANK", rank)))
    if world < 1:
        world, rank, local_rank = 1, 0, 0
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(port))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, init_method=f"tcp://{os.environ['MASTER_ADDR']}:{os.environ['MASTER_PORT']}",
                                world_size=world, rank=rank)
    if rank == 0:
        print("Running on %d workers" % world)
    return world, rank, local_rank


def _make_grid(images: torch.Tensor, nrow: int, padding: int = 2, pad_value: float = 1.0) -> torch.Tensor:
    """``torchvision.utils.make_grid(images, nrow, padding, pad_value, normalize=True)`` restated for hosts without torchvision:
    the batch is scaled to [0, 1] by its global min / max, single-channel images are repeated to three channels, and the tiles are
    laid out row-major with ``padding`` pixels of ``pad_value`` around each."""
    images = images.detach().float().cpu()
    if images.shape[1] == 1:
        images = images.repeat(1, 3, 1, 1)
    lo, hi = float(images.min()), float(images.max())
    images = (images.clamp(lo, hi) - lo) / max(hi - lo, 1e-5)
    n, c, h, w = images.shape
    xmaps = min(nrow, n)
    ymaps = -(-n // xmaps)
    H, W = h + padding, w + padding
    grid = torch.full((c, H * ymaps + padding, W * xmaps + padding), float(pad_value))
    for k in range(n):
        y, x = divmod(k, xmaps)
        grid[:, y * H + padding: y * H + padding + h, x * W + padding: x * W + padding + w] = images[k]
    return grid


def _write_png(grid: torch.Tensor, filename) -> None:
    """8-bit RGB PNG of a [3, H, W] tensor in [0, 1] with the rounding of ``torchvision.utils.save_image`` (x * 255 + 0.5, clamped);
    zlib / struct only."""
    import struct
    import zlib
    arr = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).contiguous().numpy()
    h, w, _ = arr.shape
    raw = b"".join(b"\x00" + arr[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(filename, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def plot_image_grid(img_grid: torch.Tensor, transpose: bool = False, filename=None):
    """utils.py:84-117: a [rows, cols, C, H, W] tensor of images, clamped to [-1, 1], as one normalised grid (2 px white padding)
    with ``cols`` tiles per line (``rows`` when ``transpose``); written to ``filename`` as a PNG or returned.  A [N, C, H, W]
    batch (what ``generate()`` returns here) is one row.  torchvision renders it when importable; otherwise the restatement above."""
    td = img_grid.detach().float().cpu().clamp(-1, 1)
    if td.ndim == 4:
        td = td.unsqueeze(0)
    if td.ndim != 5:
        raise ValueError("plot_image_grid expects [rows, cols, C, H, W] (or [N, C, H, W]) images, got %s" % (tuple(img_grid.shape),))
    nrow = td.shape[0] if transpose else td.shape[1]
    if transpose:
        td = td.swapaxes(0, 1)
    td = td.reshape(td.shape[0] * td.shape[1], *td.shape[2:])
    try:
        from torchvision.utils import make_grid, save_image
        rendered = make_grid(td, nrow=nrow, padding=2, pad_value=1, normalize=True).cpu()
        if filename is not None:
            save_image(rendered, fp=filename)
            return None
        return rendered
    except ImportError:
        rendered = _make_grid(td, nrow=nrow, padding=2, pad_value=1.0)
        if filename is not None:
            _write_png(rendered, filename)
            return None
        return rendered


def save_model_checkpoint(model, filename):
    """Backbone ``state_dict`` with the reference key names / fp32 PyTorch layouts (utils.py:166-167)."""
    torch.save({k: v.detach().float().cpu() for k, v in model.state_dict().items()}, filename)


def calculate_sha512_embedding(d: dict, l: int = 128) -> torch.Tensor:
    """utils.py:170-177: ASCII codes of the sha512 hex digest of the sorted-key JSON, / 128, each repeated l//128 times."""
    digest = hashlib.sha512(json.dumps(d, sort_keys=True).encode()).hexdigest()
    codes = np.frombuffer(digest.encode("ascii"), dtype=np.uint8)
    return torch.tensor(codes.repeat(l // 128) / 128, dtype=torch.float32)


def vector_to_embeddings(v, keyname):
    return torch.stack([calculate_sha512_embedding({keyname: int(x)}) for x in v])


def parameter_space_to_embeddings(param_dict: dict) -> torch.Tensor:
    """One sha512 embedding per point of the Cartesian product of the parameter space (utils.py:187-211)."""
    keys = list(param_dict.keys())
    return torch.stack([calculate_sha512_embedding(dict(zip(keys, combo))) for combo in itertools.product(*param_dict.values())])


def sample_from_discrete_parameter_space(param_dict: dict, batch_size: int, random=True, device=None) -> torch.Tensor:
    """Rows of the Cartesian product: the first ``batch_size`` rows, or uniformly random ones (utils.py:213-220)."""
    combos = torch.tensor(list(itertools.product(*param_dict.values())), device=device)
    if random:
        idx = torch.randint(low=0, high=combos.shape[0], size=(batch_size,), device=device)
    else:
        idx = torch.arange(batch_size, device=device)
    return combos[idx]


def number_cast_dict(input_dict: dict) -> dict:
    """Strings that parse as numbers become float, integral floats become int, element-wise for lists (utils.py:223-244)."""
    def cast(v):
        try:
            f = float(v)
        except (ValueError, TypeError):
            return v
        return int(f) if f.is_integer() else f

    return {k: [cast(x) for x in v] if isinstance(v, list) else cast(v) for k, v in input_dict.items()}


def right_pad_dims_to(x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    extra = x.ndim - t.ndim
    return t if extra <= 0 else t.view(*t.shape, *((1,) * extra))
