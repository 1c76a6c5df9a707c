"""LightGlue transformer layers in front of the assignment head (TEST ORACLE).

The reference matcher is ``kornia.feature.LightGlueMatcher("sift", {"n_layers": 9, "depth_confidence": -1,
"width_confidence": -1, "filter_threshold": 0.5, ...})`` (ros/gisnav/gisnav/core/pose_node.py:109-121,
called at :285-287): nine layers of self + cross attention over the two keypoint sets, no early exit,
no pruning, then the assignment head restated in ``matcher_ref``.  kornia 0.7.2 and its weights are
absent here; the layer arithmetic is restated from the published LightGlue model
and PINNED against the independent ``transformers`` 5.5 implementation with copied weights
(transformers/models/lightglue/modeling_lightglue.py:86-343; tests/test_oracle_cpu.py).

Per image: keypoints are centred and scaled by max(w, h)/2; a bias-free Linear(2 -> 32) gives the rotary
angles, repeated pairwise to the 64-wide head.  Per layer and image:

    self:  q,k,v = Linear(x); q,k rotated by the keypoint angles; a = softmax(q k^T / 8) v  (4 heads x 64)
           x += fc2(gelu(LayerNorm(fc1([x, Wo a + bo]))))
    cross: q = Linear_q(x), k,v = Linear_k,v(x_other) (no rotation), same MLP pattern with its own weights.

Numerics contract with the CUDA path (``emulate_bf16=True``): tensor-core operands are bf16 — x as a
GEMM input, q/k/v, the unnormalised softmax weights exp(s - max), the attention output, Wo's output and the MLP hidden
activation — while the residual stream, all accumulations, the softmax statistics, LayerNorm and GELU
are fp32.  With ``emulate_bf16=False`` everything is fp32 (the form pinned against transformers).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

HEADS = 4
HEAD_DIM = 64
DIM = 256


def normalize_keypoints(kp: np.ndarray, height: int, width: int) -> np.ndarray:
    """modeling_lightglue.py:461-480."""
    size = np.array([width, height], np.float32)
    shift = size / 2
    scale = size.max() / 2
    return ((np.asarray(kp, np.float32) - shift) / scale).astype(np.float32)


def rotary_tables(kp_norm: np.ndarray, pos_w: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
    """LightGluePositionalEncoder (modeling_lightglue.py:86-98): cos/sin [n,64], pairs share an angle."""
    proj = torch.from_numpy(np.ascontiguousarray(kp_norm, np.float32)) @ torch.from_numpy(pos_w).t()  # [n,32]
    emb = proj.repeat_interleave(2, dim=-1)
    return torch.cos(emb), torch.sin(emb)


def _rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x[..., ::2], x[..., 1::2]
    return torch.stack([-x2, x1], dim=-1).flatten(-2)


class _Q:
    def __init__(self, on: bool):
        self.on = on

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return x.to(torch.bfloat16).to(torch.float32) if self.on else x


def _lin(x, p, name, q):
    return x @ q(torch.from_numpy(p[name + ".weight"])).t() + torch.from_numpy(p[name + ".bias"])


def _attention(qh, kh, vh, q):
    """qh [n,256] (already scaled by 1/8), kh,vh [m,256] -> [n,256]; softmax over the m keys per head."""
    n, m = qh.shape[0], kh.shape[0]
    qh = qh.view(n, HEADS, HEAD_DIM).transpose(0, 1)
    kh = kh.view(m, HEADS, HEAD_DIM).transpose(0, 1)
    vh = vh.view(m, HEADS, HEAD_DIM).transpose(0, 1)
    s = qh @ kh.transpose(1, 2)
    # softmax with deferred normalisation (what the CUDA kernel does): the bf16 MMA operand is exp(s - max),
    # the fp32 row sum of the unrounded values divides the result; identical to softmax(s) @ v in exact arithmetic
    p = torch.exp(s - s.max(dim=-1, keepdim=True).values)
    z = p.sum(dim=-1, keepdim=True)
    return ((q(p) @ vh) / z).transpose(0, 1).reshape(n, DIM)


def _block(x, x_kv, p, prefix, q, rot=None, rot_kv=None):
    """One attention + MLP block; returns the new residual stream (fp32)."""
    xb, xkvb = q(x), q(x_kv)
    qq = _lin(xb, p, prefix + ".q", q)
    kk = _lin(xkvb, p, prefix + ".k", q)
    vv = _lin(xkvb, p, prefix + ".v", q)
    if rot is not None:
        cos, sin = rot
        cos_kv, sin_kv = rot_kv
        qq = (qq.view(-1, HEADS, HEAD_DIM) * cos[:, None] + _rotate_half(qq.view(-1, HEADS, HEAD_DIM)) * sin[:, None]).reshape(-1, DIM)
        kk = (kk.view(-1, HEADS, HEAD_DIM) * cos_kv[:, None] + _rotate_half(kk.view(-1, HEADS, HEAD_DIM)) * sin_kv[:, None]).reshape(-1, DIM)
    qq = q(qq * np.float32(HEAD_DIM ** -0.5))
    kk, vv = q(kk), q(vv)
    a = q(_attention(qq, kk, vv, q))
    o = q(_lin(a, p, prefix + ".o", q))
    h = _lin(torch.cat([xb, o], dim=-1), p, prefix + ".fc1", q)
    h = F.layer_norm(h, (2 * DIM,), torch.from_numpy(p[prefix + ".ln.weight"]), torch.from_numpy(p[prefix + ".ln.bias"]), 1e-5)
    h = q(F.gelu(h))
    return x + _lin(h, p, prefix + ".fc2", q)


@torch.no_grad()
def forward(desc0: np.ndarray, kp0: np.ndarray, hw0: Tuple[int, int], desc1: np.ndarray, kp1: np.ndarray,
            hw1: Tuple[int, int], lparams: Dict[str, np.ndarray], n_layers: int, emulate_bf16: bool = True,
            hidden: Optional[List] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Descriptors f32 [n,256] / [m,256] + pixel keypoints -> refined descriptors after ``n_layers`` layers."""
    q = _Q(emulate_bf16)
    x0 = torch.from_numpy(np.ascontiguousarray(desc0, np.float32))
    x1 = torch.from_numpy(np.ascontiguousarray(desc1, np.float32))
    if x0.shape[0] == 0 or x1.shape[0] == 0:
        return x0.numpy(), x1.numpy()
    r0 = rotary_tables(normalize_keypoints(kp0, *hw0), lparams["lg.pos.weight"])
    r1 = rotary_tables(normalize_keypoints(kp1, *hw1), lparams["lg.pos.weight"])
    for i in range(n_layers):
        s0 = _block(x0, x0, lparams, f"lg.{i}.self", q, r0, r0)
        s1 = _block(x1, x1, lparams, f"lg.{i}.self", q, r1, r1)
        x0 = _block(s0, s1, lparams, f"lg.{i}.cross", q)
        x1 = _block(s1, s0, lparams, f"lg.{i}.cross", q)
        if hidden is not None:
            hidden.append((s0.numpy().copy(), s1.numpy().copy(), x0.numpy().copy(), x1.numpy().copy()))
    return x0.numpy(), x1.numpy()


def infer_image_size(kp: np.ndarray) -> Tuple[int, int]:
    """(h, w) when the caller passes no image size: the reference calls the matcher without ``hw1``/``hw2``
    (pose_node.py:285-287), and kornia's LightGlueMatcher then takes the maximum keypoint coordinate per axis
    (upstream knowledge, not verifiable here: SURVEY.md Appendix A)."""
    m = np.asarray(kp, np.float32).max(axis=0)
    return float(m[1]), float(m[0])
