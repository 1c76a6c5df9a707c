#!/usr/bin/env python
"""Train the detector/descriptor net + matcher head on procedural textures (CPU, offline).

No SuperPoint/LightGlue checkpoint exists in the reference tree or in this image and there is no
network, so the weights shipped in ``gisnav_b200/weights/`` are produced
by this script: self-supervised homography pairs of procedural ground textures, detector labels
from Shi-Tomasi corners of the *base* texture transferred geometrically into both views, and a
dual-softmax matching loss on descriptors sampled at the labelled points (the same criterion the
matcher kernel evaluates at inference).  Activations/weights are fake-quantised to bf16 in the
forward pass because the B200 path computes convs with bf16 operands and fp32 accumulation.

    python tools/train_weights.py --steps 4000 --out gisnav_b200/weights/gnb_superpoint_v1.bin
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gisnav_b200 import synth, weights as W  # noqa: E402


def fq(x: torch.Tensor) -> torch.Tensor:
    """bf16 fake-quant with straight-through gradient."""
    return x + (x.to(torch.bfloat16).to(torch.float32) - x).detach()


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        for name, cin, cout, k in W.CONV_LAYERS:
            setattr(self, name, nn.Conv2d(cin, cout, k, padding=k // 2))
        self.proj = nn.Linear(W.DESC_DIM, W.DESC_DIM)
        self.m = nn.Linear(W.DESC_DIM, 1)
        for mod in self.modules():
            if isinstance(mod, nn.Conv2d):
                nn.init.kaiming_normal_(mod.weight, nonlinearity="relu")
                nn.init.zeros_(mod.bias)
        with torch.no_grad():
            scale = (12.0 ** 0.5) * W.DESC_DIM ** 0.25
            self.proj.weight.copy_(torch.eye(W.DESC_DIM) * scale)
            self.proj.bias.zero_()
            self.m.weight.zero_()
            self.m.bias.fill_(3.0)

    def conv(self, name, x, relu=True):
        c = getattr(self, name)
        y = F.conv2d(fq(x), fq(c.weight), c.bias, padding=c.kernel_size[0] // 2)
        return F.relu(y) if relu else y

    def forward(self, img):
        x = self.conv("conv1a", img)
        x = F.max_pool2d(self.conv("conv1b", x), 2)
        x = self.conv("conv2a", x)
        x = F.max_pool2d(self.conv("conv2b", x), 2)
        x = self.conv("conv3a", x)
        x = F.max_pool2d(self.conv("conv3b", x), 2)
        x = self.conv("conv4a", x)
        x = self.conv("conv4b", x)
        semi = self.conv("convPb", self.conv("convPa", x), relu=False)
        desc = self.conv("convDb", self.conv("convDa", x), relu=False)
        desc = F.normalize(desc, p=2, dim=1)
        return semi, desc

    def export(self):
        p = {}
        for name, _, _, _ in W.CONV_LAYERS:
            c = getattr(self, name)
            p[name + ".weight"] = c.weight.detach().numpy()
            p[name + ".bias"] = c.bias.detach().numpy()
        p["match.proj.weight"] = self.proj.weight.detach().numpy()
        p["match.proj.bias"] = self.proj.bias.detach().numpy()
        p["match.m.weight"] = self.m.weight.detach().numpy().reshape(-1)
        p["match.m.bias"] = self.m.bias.detach().numpy().reshape(1)
        return W.pack(p)


def sample_desc(desc, kp, hw):
    """desc [1,256,h,w]; kp [n,2] (x,y) pixels -> [n,256], the inference-time sampling rule."""
    h, w = hw
    g = (kp - 3.5) / torch.tensor([w - 4.5, h - 4.5]) * 2 - 1
    d = F.grid_sample(desc, g.view(1, 1, -1, 2), mode="bilinear", align_corners=True)
    return F.normalize(d.reshape(desc.shape[1], -1).t(), p=2, dim=1)


class Data:
    def __init__(self, n_tex=10, tex_size=1024, seed=1234, view=160):
        self.view = view
        self.rng = np.random.default_rng(seed)
        self.tex, self.corners = [], []
        for i in range(n_tex):
            t = synth.ground_texture(tex_size, 100 + i, n_shapes=200 * 16 // 4)
            c = cv2.goodFeaturesToTrack(t, maxCorners=tex_size * tex_size // 200, qualityLevel=0.01,
                                        minDistance=8, blockSize=5)
            self.tex.append(t)
            self.corners.append(c.reshape(-1, 2).astype(np.float64))

    def _homog(self, center, scale, theta, persp, out_center):
        t1 = np.array([[1, 0, -center[0]], [0, 1, -center[1]], [0, 0, 1.0]])
        c, s = np.cos(theta) * scale, np.sin(theta) * scale
        a = np.array([[c, -s, 0], [s, c, 0], [persp[0], persp[1], 1.0]])
        t2 = np.array([[1, 0, out_center[0]], [0, 1, out_center[1]], [0, 0, 1.0]])
        return t2 @ a @ t1

    def pair(self):
        rng, v = self.rng, self.view
        i = int(rng.integers(len(self.tex)))
        tex, corners = self.tex[i], self.corners[i]
        n = tex.shape[0]
        cb = rng.uniform(200, n - 200, 2)
        views = []
        for kind in range(2):
            if kind == 0:  # tile-like: native scale, small rotation
                h = self._homog(cb + rng.uniform(-25, 25, 2), rng.uniform(0.9, 1.1),
                                np.radians(rng.uniform(-4, 4)), (0, 0), (v / 2, v / 2))
                sigma, blur = rng.uniform(0, 1.5), False
            else:  # frame-like: magnified, rotated, slight perspective, noisy
                h = self._homog(cb + rng.uniform(-20, 20, 2), rng.uniform(1.0, 1.9),
                                np.radians(rng.uniform(-27, 27)), rng.uniform(-4e-4, 4e-4, 2), (v / 2, v / 2))
                sigma, blur = rng.uniform(0.5, 4.0), rng.random() < 0.3
            img = cv2.warpPerspective(tex, h, (v, v), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
            img = img.astype(np.float32)
            if blur:
                img = cv2.GaussianBlur(img, (0, 0), rng.uniform(0.5, 1.2))
            gain, bias = rng.uniform(0.8, 1.2), rng.uniform(-15, 15)
            img = np.clip((img - 128) * gain + 128 + bias + rng.standard_normal(img.shape) * sigma, 0, 255)
            img = np.round(img).astype(np.uint8)
            p = cv2.perspectiveTransform(corners.reshape(-1, 1, 2), h).reshape(-1, 2)
            views.append((img, p))
        if rng.random() < 0.5:
            views = views[::-1]
        return views


def make_labels(p, v):
    """p [n,2] view coords of all base corners -> (cell label map [v/8,v/8], visible idx, xy)."""
    xi, yi = np.round(p[:, 0]).astype(int), np.round(p[:, 1]).astype(int)
    vis = (xi >= 0) & (xi < v) & (yi >= 0) & (yi < v)
    lab = np.full((v // 8, v // 8), 64, np.int64)
    idx = np.nonzero(vis)[0]
    # first corner wins per cell (corners are sorted by Shi-Tomasi strength)
    for j in idx[::-1]:
        lab[yi[j] // 8, xi[j] // 8] = (yi[j] % 8) * 8 + (xi[j] % 8)
    return lab, vis


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--batch", type=int, default=6)
    ap.add_argument("--view", type=int, default=160)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--out", default=W.DEFAULT_WEIGHTS_PATH)
    ap.add_argument("--save-every", type=int, default=200)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    cv2.setNumThreads(1)
    torch.manual_seed(args.seed)
    net = Net()
    if args.resume:
        p = W.unpack(open(args.resume, "rb").read())
        with torch.no_grad():
            for name, _, _, _ in W.CONV_LAYERS:
                getattr(net, name).weight.copy_(torch.from_numpy(p[name + ".weight"]))
                getattr(net, name).bias.copy_(torch.from_numpy(p[name + ".bias"]))
            net.proj.weight.copy_(torch.from_numpy(p["match.proj.weight"]))
            net.proj.bias.copy_(torch.from_numpy(p["match.proj.bias"]))
            net.m.weight.copy_(torch.from_numpy(p["match.m.weight"]).view(1, -1))
            net.m.bias.copy_(torch.from_numpy(p["match.m.bias"]))
    data = Data(view=args.view, seed=1234 + args.seed)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=args.lr, total_steps=args.steps, pct_start=0.05)
    v = args.view
    t0 = time.time()
    ema = None
    for step in range(1, args.steps + 1):
        batch = [data.pair() for _ in range(args.batch)]
        imgs = np.stack([vw[0] for pr in batch for vw in pr]).astype(np.float32) / 255.0
        semi, desc = net(torch.from_numpy(imgs)[:, None])
        loss_det = 0.0
        loss_match = 0.0
        loss_m = 0.0
        n_pairs_used = 0
        acc = []
        for b, pr in enumerate(batch):
            labs, viss = [], []
            for kview in range(2):
                lab, vis = make_labels(pr[kview][1], v)
                labs.append(lab)
                viss.append(vis)
                loss_det = loss_det + F.cross_entropy(semi[2 * b + kview][None], torch.from_numpy(lab)[None])
            ia = np.nonzero(viss[0])[0]
            ib = np.nonzero(viss[1])[0]
            if len(ia) < 8 or len(ib) < 8:
                continue
            # cap the number of keypoints per view for speed
            if len(ia) > 300:
                ia = np.sort(data.rng.choice(ia, 300, replace=False))
            if len(ib) > 300:
                ib = np.sort(data.rng.choice(ib, 300, replace=False))
            pa = pr[0][1][ia] + data.rng.uniform(-1.2, 1.2, (len(ia), 2))
            pb = pr[1][1][ib] + data.rng.uniform(-1.2, 1.2, (len(ib), 2))
            da = sample_desc(desc[2 * b][None], torch.from_numpy(np.round(pa)).float(), (v, v))
            db = sample_desc(desc[2 * b + 1][None], torch.from_numpy(np.round(pb)).float(), (v, v))
            da_q, db_q = fq(da), fq(db)
            ma = fq(net.proj(da_q) / W.DESC_DIM ** 0.25)
            mb = fq(net.proj(db_q) / W.DESC_DIM ** 0.25)
            s = ma @ mb.t()
            za, zb = net.m(da_q)[:, 0], net.m(db_q)[:, 0]
            sc = F.log_softmax(s, 1) + F.log_softmax(s, 0) + F.logsigmoid(za)[:, None] + F.logsigmoid(zb)[None, :]
            common, ja, jb = np.intersect1d(ia, ib, return_indices=True)
            if len(common) < 4:
                continue
            gt = sc[torch.from_numpy(ja), torch.from_numpy(jb)]
            loss_match = loss_match - gt.mean()
            has_a = torch.zeros(len(ia)); has_a[torch.from_numpy(ja)] = 1
            has_b = torch.zeros(len(ib)); has_b[torch.from_numpy(jb)] = 1
            # unmatched points should have low matchability
            loss_m = loss_m + F.binary_cross_entropy_with_logits(za, has_a) * 0.0 \
                + (-(F.logsigmoid(-za) * (1 - has_a)).sum() - (F.logsigmoid(-zb) * (1 - has_b)).sum()) \
                / max(1.0, float((1 - has_a).sum() + (1 - has_b).sum()))
            n_pairs_used += 1
            with torch.no_grad():
                m0 = sc.argmax(1)[torch.from_numpy(ja)]
                ok = (m0 == torch.from_numpy(jb)) & (gt.exp() > 0.5)
                acc.append(float(ok.float().mean()))
        loss = loss_det / (2 * args.batch) + (loss_match + 0.5 * loss_m) / max(1, n_pairs_used)
        opt.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()
        sched.step()
        lv = float(loss)
        ema = lv if ema is None else 0.98 * ema + 0.02 * lv
        if step % 20 == 0 or step == 1:
            print(f"step {step} loss {lv:.3f} ema {ema:.3f} det {float(loss_det) / (2 * args.batch):.3f} "
                  f"match {float(loss_match) / max(1, n_pairs_used):.3f} recall@.5 {np.mean(acc) if acc else 0:.3f} "
                  f"{(time.time() - t0) / step:.2f}s/step", flush=True)
        if step % args.save_every == 0 or step == args.steps:
            tmp = args.out + ".tmp"
            with open(tmp, "wb") as f:
                f.write(net.export())
            os.replace(tmp, args.out)
    print("done", time.time() - t0)


if __name__ == "__main__":
    main()
