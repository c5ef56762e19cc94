"""Image transforms — drop-in for `imagescry/image/transforms.py`, computed by the sm_100a kernels.

Same names, signatures, argument meaning and error behaviour as the reference module
(`/root/reference/src/imagescry/image/transforms.py`): `normalize_per_channel` (:16-74),
`resize` (:78-126), `to_4d` (:130-164), `_calc_scale_factor` (:168-197).  Tensors must live on a
CUDA device; there is no CPU path.

Beyond the reference's surface, `preprocess_tiles` exposes the fused stage-1 pipeline directly
(optional NHWC uint8 input, resize folded into both passes, fp32 or bf16 output) and
`preprocess_patches` adds patch tiling of large images as index arithmetic inside the same passes.
"""

from __future__ import annotations

import math
from typing import Literal

import torch
from jaxtyping import Float, Num, Shaped, jaxtyped
from torch import Tensor

from imagescry_b200 import _lib
from imagescry_b200.typechecking import typechecker

_IN_DTYPES = {torch.uint8: _lib.DTYPE_U8, torch.float32: _lib.DTYPE_F32}
_OUT_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}


def _as_kernel_input(x: Tensor) -> Tensor:
    """uint8 and float32 go to the kernels as they are; every other numeric dtype is cast to
    float32 first — which is what the reference's `.float()` does (transforms.py:59,103)."""
    if x.dtype not in _IN_DTYPES:
        x = x.float()
    return x.contiguous()


def _stats(x: Tensor, layout: int, B: int, C: int, H: int, W: int, oh: int, ow: int) -> tuple[Tensor, Tensor]:
    lib = _lib.load()
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    std = torch.empty(C, dtype=torch.float32, device=x.device)
    ws_bytes = lib.isx_preprocess_stats_workspace_bytes(C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with _lib.on_device(x) as stream:
        rc = lib.isx_preprocess_stats(
            x.data_ptr(), _IN_DTYPES[x.dtype], layout, B, C, H, W, oh, ow, mean.data_ptr(), std.data_ptr(),
            ws.data_ptr(), ws_bytes, stream,
        )
    _lib.check(rc, "isx_preprocess_stats")
    return mean, std


def _apply(
    x: Tensor, layout: int, B: int, C: int, H: int, W: int, oh: int, ow: int, mean: Tensor, std: Tensor,
    stat_batch: int, eps: float, min_value: float | None, max_value: float | None, out_dtype: torch.dtype,
) -> Tensor:
    lib = _lib.load()
    out = torch.empty((B, C, oh, ow), dtype=out_dtype, device=x.device)
    with _lib.on_device(x, mean, std) as stream:
        rc = lib.isx_preprocess_apply(
            x.data_ptr(), _IN_DTYPES[x.dtype], layout, B, C, H, W, oh, ow, mean.data_ptr(), std.data_ptr(), stat_batch,
            float(eps), int(min_value is not None), float(min_value or 0.0), int(max_value is not None),
            float(max_value or 0.0), out.data_ptr(), _OUT_DTYPES[out_dtype], stream,
        )
    _lib.check(rc, "isx_preprocess_apply")
    return out


def _flatten_stats(t: Tensor, B: int, C: int, device: torch.device) -> tuple[Tensor, int]:
    """'#B C 1 1' statistics → contiguous fp32 (sb*C,) on `device`, sb in {1, B}."""
    t = t.to(device=device, dtype=torch.float32)
    sb = t.shape[0]
    return t.reshape(sb, C).contiguous(), sb


def preprocess_tiles(
    tiles: Tensor,
    *,
    layout: Literal["nchw", "nhwc"] = "nchw",
    output_hw: tuple[int, int] | None = None,
    channel_means: Tensor | None = None,
    channel_stds: Tensor | None = None,
    min_value: float | None = None,
    max_value: float | None = None,
    eps: float = 1e-6,
    out_dtype: torch.dtype = torch.float32,
) -> Tensor:
    """Fused stage 1: (HWC→CHW) → bilinear resize → per-channel normalise → clip, NCHW out.

    Equivalent to `normalize_per_channel(resize(tiles, output_hw), ...)` of the reference
    (transforms.py:16-126) with the resized image never materialised: the statistics pass and the
    apply pass both sample the uint8 input directly.
    """
    _lib.require_cuda(tiles, "tiles")
    if tiles.ndim != 4:
        raise ValueError(f"tiles must be 4-D, got shape {tuple(tiles.shape)}")
    x = _as_kernel_input(tiles)
    if layout == "nchw":
        B, C, H, W = x.shape
        lay = _lib.LAYOUT_NCHW
    elif layout == "nhwc":
        B, H, W, C = x.shape
        lay = _lib.LAYOUT_NHWC
    else:
        raise ValueError(f"Invalid layout: {layout}")
    oh, ow = (H, W) if output_hw is None else (int(output_hw[0]), int(output_hw[1]))
    if out_dtype not in _OUT_DTYPES:
        raise ValueError(f"out_dtype must be float32 or bfloat16, got {out_dtype}")
    if B == 0:
        return torch.empty((0, C, oh, ow), dtype=out_dtype, device=x.device)

    mean = std = None
    if channel_means is None or channel_stds is None:
        mean, std = _stats(x, lay, B, C, H, W, oh, ow)
    sb_m = sb_s = 1
    if channel_means is not None:
        mean, sb_m = _flatten_stats(channel_means, B, C, x.device)
    if channel_stds is not None:
        std, sb_s = _flatten_stats(channel_stds, B, C, x.device)
    sb = max(sb_m, sb_s)
    if sb > 1:
        mean = mean.reshape(-1, C).expand(sb, C).contiguous()
        std = std.reshape(-1, C).expand(sb, C).contiguous()
    return _apply(x, lay, B, C, H, W, oh, ow, mean, std, sb, eps, min_value, max_value, out_dtype)


def patch_grid(height: int, width: int, patch_size: int, stride: int | None = None) -> tuple[int, int]:
    """(ny, nx): full `patch_size` windows on a `stride` grid that fit an image of height x width."""
    stride = patch_size if stride is None else stride
    if patch_size < 1 or stride < 1:
        raise ValueError(f"patch_size and stride must be positive, got {patch_size} and {stride}")
    if patch_size > height or patch_size > width:
        raise ValueError(f"patch_size {patch_size} exceeds the image size {height}x{width}")
    return (height - patch_size) // stride + 1, (width - patch_size) // stride + 1


def preprocess_patches(
    images: Tensor,
    patch_size: int,
    *,
    stride: int | None = None,
    layout: Literal["nchw", "nhwc"] = "nhwc",
    output_hw: tuple[int, int] | None = None,
    channel_means: Tensor | None = None,
    channel_stds: Tensor | None = None,
    min_value: float | None = None,
    max_value: float | None = None,
    eps: float = 1e-6,
    out_dtype: torch.dtype = torch.float32,
) -> Tensor:
    """Patch tiling fused into stage 1: cut uint8 3-channel images (`n×H×W×3` or `n×3×H×W`) into
    `patch_size` windows on a `stride` grid (full windows only, ordered image, py, px) and preprocess
    the windows as a tile batch — `preprocess_tiles` on a patch tensor that is never built: the
    window offsets are index arithmetic inside the statistics and apply kernels' reads.

    Returns `(n·ny·nx)×3×outH×outW`.  The reference has no patch tiling; the
    semantics are the oracle's `preprocess_patches` restatement."""
    _lib.require_cuda(images, "images")
    if images.ndim != 4 or images.dtype != torch.uint8:
        raise ValueError(f"images must be a 4-D uint8 tensor, got {images.dtype} with shape {tuple(images.shape)}")
    x = images.contiguous()
    if layout == "nhwc":
        n, H, W, C = x.shape
        lay = _lib.LAYOUT_NHWC
    elif layout == "nchw":
        n, C, H, W = x.shape
        lay = _lib.LAYOUT_NCHW
    else:
        raise ValueError(f"Invalid layout: {layout}")
    if C != 3:
        raise ValueError(f"patch tiling supports 3-channel images, got {C} channels")
    stride = patch_size if stride is None else int(stride)
    ny, nx = patch_grid(H, W, patch_size, stride)
    oh, ow = (patch_size, patch_size) if output_hw is None else (int(output_hw[0]), int(output_hw[1]))
    if out_dtype not in _OUT_DTYPES:
        raise ValueError(f"out_dtype must be float32 or bfloat16, got {out_dtype}")
    B = n * ny * nx
    out = torch.empty((B, 3, oh, ow), dtype=out_dtype, device=x.device)
    if B == 0:
        return out
    lib = _lib.load()
    if channel_means is None or channel_stds is None:
        mean = torch.empty(3, dtype=torch.float32, device=x.device)
        std = torch.empty(3, dtype=torch.float32, device=x.device)
        ws_bytes = lib.isx_preprocess_stats_workspace_bytes(3)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        with _lib.on_device(x) as stream:
            rc = lib.isx_preprocess_patches_stats(
                x.data_ptr(), lay, n, 3, H, W, patch_size, stride, oh, ow, mean.data_ptr(), std.data_ptr(), ws.data_ptr(),
                ws_bytes, stream,
            )
        _lib.check(rc, "isx_preprocess_patches_stats")
    if channel_means is not None:
        mean = channel_means.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
    if channel_stds is not None:
        std = channel_stds.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
    if mean.numel() != 3 or std.numel() != 3:
        raise ValueError("patch tiling takes batch-wide statistics: channel_means / channel_stds must hold 3 values")
    with _lib.on_device(x, mean, std) as stream:
        rc = lib.isx_preprocess_patches_apply(
            x.data_ptr(), lay, n, 3, H, W, patch_size, stride, oh, ow, mean.data_ptr(), std.data_ptr(), float(eps),
            int(min_value is not None), float(min_value or 0.0), int(max_value is not None), float(max_value or 0.0),
            out.data_ptr(), _OUT_DTYPES[out_dtype], stream,
        )
    _lib.check(rc, "isx_preprocess_patches_apply")
    return out


@jaxtyped(typechecker=typechecker)
def normalize_per_channel(
    image_tensor: Num[Tensor, "B C H W"],
    *,
    channel_means: Float[Tensor, "#B C 1 1"] | None = None,
    channel_stds: Float[Tensor, "#B C 1 1"] | None = None,
    min_value: float | None = None,
    max_value: float | None = None,
    eps: float = 1e-6,
) -> Float[Tensor, "B C H W"]:
    """Normalize image pixels (per channel) to zero mean and unit variance — same contract as the
    reference (`transforms.py:16-74`): batch statistics over dims (0, 2, 3) with the unbiased std
    unless supplied, `(x - mean) / (std + eps)`, optional clip.  Returns a new fp32 tensor."""
    return preprocess_tiles(
        image_tensor, layout="nchw", channel_means=channel_means, channel_stds=channel_stds,
        min_value=min_value, max_value=max_value, eps=eps,
    )


@jaxtyped(typechecker=typechecker)
def resize(
    image_tensor: Num[Tensor, "... H1 W1"],
    output_size: int | tuple[int, int],
    *,
    side_ref: Literal["height", "width", "long", "short"] = "long",
) -> Float[Tensor, "... H2 W2"]:
    """Resize image tensor (bilinear, `align_corners=False`, no antialias) — same contract as the
    reference (`transforms.py:78-126`): 2-/3-/4-D input, float output, integer `output_size` scales
    the `side_ref` side with output dims `floor(in * scale)`."""
    _lib.require_cuda(image_tensor, "image_tensor")
    squeeze_dims = tuple(range(4 - image_tensor.ndim))
    x = _as_kernel_input(to_4d(image_tensor))
    B, C, H, W = x.shape
    oh, ow = resized_shape(H, W, output_size, side_ref)
    out = torch.empty((B, C, oh, ow), dtype=torch.float32, device=x.device)
    if out.numel() > 0:
        lib = _lib.load()
        with _lib.on_device(x) as stream:
            rc = lib.isx_resize_bilinear(
                x.data_ptr(), _IN_DTYPES[x.dtype], _lib.LAYOUT_NCHW, B, C, H, W, oh, ow, out.data_ptr(), stream
            )
        _lib.check(rc, "isx_resize_bilinear")
    return out.squeeze(squeeze_dims) if squeeze_dims else out


def resized_shape(
    height: int, width: int, output_size: int | tuple[int, int], side_ref: str = "long"
) -> tuple[int, int]:
    """Output (H2, W2) of `resize`: `floor(in * scale)` in python doubles for an integer size (what
    `interpolate(scale_factor=..., recompute_scale_factor=True)` computes), the tuple otherwise."""
    if isinstance(output_size, int):
        s = _calc_scale_factor(height, width, output_size, side_ref)
        return int(math.floor(height * s)), int(math.floor(width * s))
    return int(output_size[0]), int(output_size[1])


@jaxtyped(typechecker=typechecker)
def to_4d(image_tensor: Shaped[Tensor, "... H W"]) -> Shaped[Tensor, "B C H W"]:
    """Add phantom leading dimensions until the tensor is 4-D (`transforms.py:130-164`).

    Raises:
        ValueError: If `image_tensor` is not 2D, 3D, or 4D.
    """
    if image_tensor.ndim == 2:
        return image_tensor.unsqueeze(0).unsqueeze(0)
    if image_tensor.ndim == 3:
        return image_tensor.unsqueeze(0)
    if image_tensor.ndim == 4:
        return image_tensor
    raise ValueError(f"Invalid image tensor shape: {image_tensor.shape}")


def _calc_scale_factor(
    height: int, width: int, output_size: int, side_ref: Literal["height", "width", "long", "short"]
) -> float:
    """Scale factor for an integer `output_size` (`transforms.py:168-197`)."""
    if side_ref == "height":
        return output_size / height
    if side_ref == "width":
        return output_size / width
    if side_ref == "long":
        return output_size / max(height, width)
    if side_ref == "short":
        return output_size / min(height, width)
    raise ValueError(f"Invalid side_ref: {side_ref}")
