"""torch-CPU restatement of the SuperPoint-style detector/descriptor stack (TEST ORACLE).

Replaces ``cv2.SIFT.detectAndCompute`` (ros/gisnav/gisnav/core/pose_node.py:230;
twist_node.py:227) as BASELINE.json's north_star prescribes; architecture per the published
SuperPoint model: encoder 1-64-64 | 64-64 | 128-128 | 128-128 with 2x2 max-pool
after the first three blocks, detector head 128-256-65, descriptor head 128-256-256.

Numerics contract shared with the CUDA path ("bf16 operands, fp32 accumulate"): every conv reads
its input activations and weights rounded to bfloat16 (round-to-nearest-even), accumulates in
fp32, adds an fp32 bias and applies ReLU in fp32.  Head outputs (65 logits, 256-d raw descriptors)
stay fp32.  ``quantize=False`` gives the plain fp32 network for reporting the bf16 deviation.

``quantize="x3"`` is the contract of the library's fp32-faithful mode (``conv_precision = 1``): every
conv operand is split into two bf16 terms, ``v = hi + lo`` with ``hi = bf16(v)``, ``lo = bf16(v - hi)``
(16 significant bits), and the product keeps the three leading terms
``a_hi w_hi + a_hi w_lo + a_lo w_hi`` accumulated in fp32 — what three tcgen05 MMAs into one TMEM
accumulator compute.  Relative error per product ~2^-16 against 2^-8 for plain bf16 operands.  conv1a (Cin = 1) is
plain fp32 in that mode (CUDA cores); the two 1x1 heads use the split like the 3x3 layers.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def _q(x: torch.Tensor, quantize) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if quantize else x


def split_hi_lo(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """v -> (hi, lo), both bf16-representable fp32 tensors with v ~= hi + lo to 2^-17 relative."""
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def _conv(x, p, name, relu=True, quantize=True):
    w = torch.from_numpy(p[name + ".weight"])
    b = torch.from_numpy(p[name + ".bias"])
    pad = w.shape[-1] // 2
    if quantize == "x3" and name == "conv1a":
        # Cin = 1, nine FMAs per output: plain fp32 on the CUDA cores
        y = F.conv2d(x, w, b, padding=pad)
    elif quantize == "x3":
        xh, xl = split_hi_lo(x)
        wh, wl = split_hi_lo(w)
        # the two small terms first, then the leading one and the bias (fp32 accumulation throughout)
        y = (F.conv2d(xh, wl, None, padding=pad) + F.conv2d(xl, wh, None, padding=pad)) + F.conv2d(xh, wh, b, padding=pad)
    else:
        y = F.conv2d(_q(x, quantize), _q(w, quantize), b, padding=pad)
    return F.relu(y) if relu else y


@torch.no_grad()
def forward_dense(image_u8: np.ndarray, params: Dict[str, np.ndarray], quantize: bool = True
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """u8 [H,W] (H,W multiples of 8) -> (score map f32 [H,W], dense descriptors f32 [H/8,W/8,256]).

    Score map = softmax over the 65 detector channels, dustbin dropped, 8x8 depth-to-space.
    Dense descriptors are L2-normalised over channels and returned channels-last.
    """
    h, w = image_u8.shape
    assert h % 8 == 0 and w % 8 == 0, "image sides must be multiples of 8"
    x = torch.from_numpy(image_u8.astype(np.float32) / np.float32(255.0))[None, None]
    c = lambda n, t, relu=True: _conv(t, params, n, relu, quantize)  # noqa: E731
    x = c("conv1a", x)
    x = F.max_pool2d(c("conv1b", x), 2)
    x = c("conv2a", x)
    x = F.max_pool2d(c("conv2b", x), 2)
    x = c("conv3a", x)
    x = F.max_pool2d(c("conv3b", x), 2)
    x = c("conv4a", x)
    x = c("conv4b", x)
    semi = c("convPb", c("convPa", x), relu=False)  # [1,65,h/8,w/8]
    prob = F.softmax(semi, dim=1)[:, :-1]
    hc, wc = h // 8, w // 8
    score = prob.permute(0, 2, 3, 1).reshape(1, hc, wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(h, w)
    desc = c("convDb", c("convDa", x), relu=False)  # [1,256,hc,wc]
    desc = F.normalize(desc, p=2, dim=1)
    return score.numpy().copy(), desc[0].permute(1, 2, 0).contiguous().numpy().copy()


@torch.no_grad()
def forward_layers(image_u8: np.ndarray, params: Dict[str, np.ndarray], quantize: bool = True
                   ) -> Dict[str, np.ndarray]:
    """Per-layer activations (channels-last, after ReLU and pooling) for stage-isolated parity."""
    out: Dict[str, np.ndarray] = {}
    x = torch.from_numpy(image_u8.astype(np.float32) / np.float32(255.0))[None, None]
    c = lambda n, t, relu=True: _conv(t, params, n, relu, quantize)  # noqa: E731
    cl = lambda t: t[0].permute(1, 2, 0).contiguous().numpy().copy()  # noqa: E731
    x = c("conv1a", x); out["conv1a"] = cl(x)
    x = F.max_pool2d(c("conv1b", x), 2); out["pool1"] = cl(x)
    x = c("conv2a", x); out["conv2a"] = cl(x)
    x = F.max_pool2d(c("conv2b", x), 2); out["pool2"] = cl(x)
    x = c("conv3a", x); out["conv3a"] = cl(x)
    x = F.max_pool2d(c("conv3b", x), 2); out["pool3"] = cl(x)
    x = c("conv4a", x); out["conv4a"] = cl(x)
    x = c("conv4b", x); out["conv4b"] = cl(x)
    pa = c("convPa", x); out["convPa"] = cl(pa)
    out["semi"] = cl(c("convPb", pa, relu=False))
    da = c("convDa", x); out["convDa"] = cl(da)
    out["desc_raw"] = cl(c("convDb", da, relu=False))
    return out
