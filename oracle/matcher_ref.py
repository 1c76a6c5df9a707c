"""Matcher oracle (TEST ORACLE): the LightGlue *assignment head* the reference matcher ends with.

Reference call: ``dists, idx = self._matcher(desc_q, desc_r, lafs_q, lafs_r)`` with
``filter_threshold = CONFIDENCE_THRESHOLD = 0.5`` (ros/gisnav/gisnav/core/pose_node.py:60,
109-121,285-287); rows of ``idx`` index (query, reference) (pose_node.py:296-297).  kornia 0.7.2 is
not installed; the head's arithmetic is restated from the published LightGlue
model (cf. transformers/models/lightglue/modeling_lightglue.py:345-357,429-458):

    m = (W d + b) / D^(1/4);  S = m_a m_b^T;  z = w_m . d + b_m
    score = log_softmax_row(S) + log_softmax_col(S) + logsigmoid(z_a)_i + logsigmoid(z_b)_j
    m0 = argmax_j score, m1 = argmax_i score, keep i iff m1[m0[i]] == i and exp(score) > thr

Numerics contract with the CUDA path: descriptors, W and the projected m are rounded to bf16
(tensor-core operands), all sums are fp32.  Ties in argmax resolve to the lowest index.
The transformer layers in front of the head are out of scope this round.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def _q(x: torch.Tensor, quantize: bool = True) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if quantize else x


@torch.no_grad()
def project(desc: np.ndarray, params: Dict[str, np.ndarray], quantize: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """desc f32 [n,256] -> (m f32 [n,256] (bf16-valued when ``quantize``), matchability logit z f32 [n]).
    ``quantize=False`` is the plain fp32 head: the reference's tensors (pose_node.py:254-287) and the
    library's fp32-faithful mode (``precision = 1``)."""
    d = _q(torch.from_numpy(np.ascontiguousarray(desc, np.float32)), quantize)
    w = _q(torch.from_numpy(params["match.proj.weight"]), quantize)
    b = torch.from_numpy(params["match.proj.bias"])
    m = (d @ w.t() + b) * np.float32(1.0 / desc.shape[1] ** 0.25)
    wm = _q(torch.from_numpy(params["match.m.weight"]), quantize)
    z = d @ wm + torch.from_numpy(params["match.m.bias"])[0]
    return _q(m, quantize).numpy(), z.numpy()


def _split(x: torch.Tensor):
    hi = x.to(torch.bfloat16).to(torch.float32)
    return hi, (x - hi).to(torch.bfloat16).to(torch.float32)


@torch.no_grad()
def assignment_scores(desc_a, desc_b, params, quantize=True) -> np.ndarray:
    """``quantize``: True = bf16 operands (fast mode), False = plain fp32, "x3" = the fp32-faithful mode's head:
    fp32 projection, then S from split-bf16 operands (m = hi + lo; S = hi hi^T + hi lo^T + lo hi^T, fp32 accumulate —
    three tensor-core MMAs per product)."""
    if quantize == "x3":
        ma, za = project(desc_a, params, False)
        mb, zb = project(desc_b, params, False)
        ah, al = _split(torch.from_numpy(ma))
        bh, bl = _split(torch.from_numpy(mb))
        s = (ah @ bl.t() + al @ bh.t()) + ah @ bh.t()
        sc = (F.log_softmax(s, 1) + F.log_softmax(s, 0)
              + F.logsigmoid(torch.from_numpy(za))[:, None] + F.logsigmoid(torch.from_numpy(zb))[None, :])
        return sc.numpy()
    ma, za = project(desc_a, params, quantize)
    mb, zb = project(desc_b, params, quantize)
    s = torch.from_numpy(ma) @ torch.from_numpy(mb).t()
    sc = (F.log_softmax(s, 1) + F.log_softmax(s, 0)
          + F.logsigmoid(torch.from_numpy(za))[:, None] + F.logsigmoid(torch.from_numpy(zb))[None, :])
    return sc.numpy()


@torch.no_grad()
def match(desc_a: np.ndarray, desc_b: np.ndarray, params: Dict[str, np.ndarray],
          threshold: float = 0.5, quantize=True) -> Tuple[np.ndarray, np.ndarray]:
    """-> (scores f32 [k,1], idx int64 [k,2]) sorted by query index, like the reference call."""
    n, m = desc_a.shape[0], desc_b.shape[0]
    if n == 0 or m == 0:
        return np.zeros((0, 1), np.float32), np.zeros((0, 2), np.int64)
    sc = assignment_scores(desc_a, desc_b, params, quantize)
    m0 = sc.argmax(1)  # numpy argmax returns the first maximum
    m1 = sc.argmax(0)
    i = np.arange(n)
    mutual = m1[m0] == i
    ms = np.exp(sc[i, m0].astype(np.float32))
    valid = mutual & (ms > np.float32(threshold))
    idx = np.stack([i[valid], m0[valid]], axis=1).astype(np.int64)
    return ms[valid].reshape(-1, 1).astype(np.float32), idx
