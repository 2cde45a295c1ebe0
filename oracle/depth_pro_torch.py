"""ORACLE (test infrastructure, never shipped, never on the product path).

Depth Pro's patch-encoder stage, restated: the reference runs `depth_pro.create_model_and_transforms` from the
un-vendored github.com/apple/ml-depth-pro package (models/depth_pro/onnx_export.py:2,15-29: `dinov2l16_384` patch /
image encoders, 1536 x 1536 input).  Its encoder (`depth_pro/network/encoder.py`: `split`, `merge`, the two hooked
blocks) is public; transformers carries an independent implementation (`DepthProPatchEncoder`) that loads the released
checkpoint, and tests/test_oracle_depth_pro.py pins this file against it on copied weights.

  image [3, 1536, 1536] -> pyramid 1, 1/2, 1/4 (bilinear, align_corners=False) -> 25 + 9 + 1 crops of 384 x 384
  -> shared ViT/16 trunk -> per crop: normalised final tokens and the raw outputs of two hooked blocks, 24 x 24 x D
  -> merged maps: 96 x 96 (final, full-resolution crops, 3 tokens trimmed at every inner edge), 48 x 48 (half
     resolution, 6 trimmed), 24 x 24 (quarter resolution), and 96 x 96 for each hook (full-resolution crops only).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from oracle import dav2_torch as O

GRID = 24           # tokens per crop side (384 / 16)
MERGE_PADDING = 3   # tokens trimmed per inner edge at full resolution; x2, x4 at the lower levels (capped at GRID // 4)


def merge_crops(tokens: torch.Tensor, per_side: int, padding: int) -> torch.Tensor:
    """tokens [per_side**2, GRID*GRID, D] (row-major crops) -> [S, S, D] with S = per_side*GRID - 2*padding*(per_side-1).
    Inner edges of every crop lose `padding` tokens (upstream `merge`, transformers `merge_patches`)."""
    n, t, d = tokens.shape
    assert n == per_side * per_side and t == GRID * GRID
    padding = min(GRID // 4, padding) if per_side > 1 else 0
    grid = tokens.reshape(per_side, per_side, GRID, GRID, d)
    rows = []
    for h in range(per_side):
        top = padding if h != 0 else 0
        bottom = GRID - (padding if h != per_side - 1 else 0)
        cols = []
        for w in range(per_side):
            left = padding if w != 0 else 0
            right = GRID - (padding if w != per_side - 1 else 0)
            cols.append(grid[h, w, top:bottom, left:right])
        rows.append(torch.cat(cols, dim=1))
    return torch.cat(rows, dim=0)


def merged_features(taps: Sequence[torch.Tensor], hook_taps: Sequence[int] = (1, 0), final_tap: int = 3) -> List[torch.Tensor]:
    """taps: the trunk's four outputs [35, 576, D] each (crop order: 25 full-resolution, 9 half, 1 quarter).
    -> [f24, f48, f96, hook_a 96, hook_b 96] as [S, S, D], the order transformers' DepthProPatchEncoder returns them
    (low resolution first, then the hooks in the configured order -- (11, 5) for the released model = taps (1, 0))."""
    fin = taps[final_tap]
    out = [merge_crops(fin[34:35], 1, 0), merge_crops(fin[25:34], 3, 2 * MERGE_PADDING), merge_crops(fin[0:25], 5, MERGE_PADDING)]
    for t in hook_taps:
        out.append(merge_crops(taps[t][0:25], 5, MERGE_PADDING))
    return out


@torch.no_grad()
def patch_encoder_features(sd, image: torch.Tensor, encoder: str, hook_taps: Sequence[int] = (1, 0)) -> List[torch.Tensor]:
    """image [3, S, S] float32 (normalised) -> the five merged maps, through the oracle trunk."""
    from monocular_depth_estimation_trt_b200 import sharding as S      # crop geometry only (pure host code)
    crops = S.make_crops(image)
    taps = O.encoder_taps(sd, crops, O.MODEL_CONFIGS[encoder], norm_mask=0x8)
    return merged_features(taps, hook_taps)
