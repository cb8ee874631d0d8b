"""Seeded synthetic speckle stacks for parity tests and bench.py.

Reference frames are fully developed speckle (|complex Gaussian field|^2, speckle
size ~2 px, unit mean, unit contrast -- the statistics of the reference's own
``utils.prep_simul`` at UMPA/utils.py:402-405); sample frames are the reference
warped by a smooth displacement field, attenuated by T and with reduced
visibility v.  torch is used so that the 25 x 2048^2 bench stacks can be made
on the GPU in a second; the noise itself always comes from a CPU generator so a
given seed yields the same stack on every device.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _gauss_blur(x, sigma):
    r = int(4 * sigma + .5)
    t = torch.arange(-r, r + 1, dtype=x.dtype, device=x.device)
    k = torch.exp(-.5 * (t / sigma) ** 2)
    k = k / k.sum()
    x = x[None, None]
    x = F.conv2d(F.pad(x, (r, r, 0, 0), mode="reflect"), k[None, None, None, :])
    x = F.conv2d(F.pad(x, (0, 0, r, r), mode="reflect"), k[None, None, :, None])
    return x[0, 0]


def shift_fields(H, W, max_shift, amplitude=None, device="cpu", dtype=torch.float64):
    """dx (column shift), dy (row shift), T, v maps of the synthetic sample."""
    A = min(1.7, (max_shift - 2) / 2.) if amplitude is None else amplitude
    y = torch.arange(H, dtype=dtype, device=device)[:, None].expand(H, W)
    x = torch.arange(W, dtype=dtype, device=device)[None, :].expand(H, W)
    dx = A * torch.sin(2 * math.pi * y / H)
    dy = .7 * A * torch.cos(2 * math.pi * x / W)
    T = .8 + .1 * torch.cos(2 * math.pi * x / W)
    v = .9 + .1 * torch.sin(2 * math.pi * y / H)
    return dx, dy, T, v


def speckle_stack(Na, H, W, seed=0, max_shift=4, dark_field=True, amplitude=None, noise=0.,
                  speckle_sigma=2., contrast=1., device="cpu", as_numpy=True):
    """Returns dict(sam, ref [Na,H,W] float64, dx, dy, T, v truth maps).

    sam_k(y,x) = T * [ v * ref_k(y+dy, x+dx) + (1-v) * local_mean(ref_k) ]   (+ noise)
    contrast < 1 adds a constant pedestal to ref (low-visibility speckle).
    """
    dev = torch.device(device)
    dt = torch.float64
    dx, dy, T, v = shift_fields(H, W, max_shift, amplitude, dev, dt)
    if not dark_field:
        v = torch.ones_like(v)
    yy = torch.arange(H, dtype=dt, device=dev)[:, None] + dy
    xx = torch.arange(W, dtype=dt, device=dev)[None, :] + dx
    grid = torch.stack([2 * xx / (W - 1) - 1, 2 * yy / (H - 1) - 1], dim=-1)[None]
    sam = torch.empty((Na, H, W), dtype=dt, device=dev)
    ref = torch.empty((Na, H, W), dtype=dt, device=dev)
    for k in range(Na):
        g = torch.Generator(device="cpu").manual_seed(1000 * seed + k)
        n = torch.randn((2, H, W), generator=g, dtype=dt).to(dev)
        re, im = _gauss_blur(n[0], speckle_sigma), _gauss_blur(n[1], speckle_sigma)
        r = re * re + im * im
        r = r / r.mean()
        if contrast != 1.:
            r = 1. + contrast * (r - 1.)
        warped = F.grid_sample(r[None, None], grid, mode="bicubic", padding_mode="reflection",
                               align_corners=True)[0, 0]
        loc = _gauss_blur(r, 3. * speckle_sigma)
        s = T * (v * warped + (1. - v) * loc)
        if noise > 0.:
            s = s + noise * torch.randn((H, W), generator=g, dtype=dt).to(dev)
            r = r + noise * torch.randn((H, W), generator=g, dtype=dt).to(dev)
        sam[k], ref[k] = s, r
    out = dict(sam=sam, ref=ref, dx=dx, dy=dy, T=T, v=v)
    if as_numpy:
        out = {k_: np.ascontiguousarray(a.cpu().numpy()) for k_, a in out.items()}
    return out


def blur_abc(N0, N1, as_numpy=True):
    """Spatially varying blur-kernel parameters for UMPAModelDFKernel."""
    y = torch.arange(N0, dtype=torch.float64)[:, None].expand(N0, N1)
    x = torch.arange(N1, dtype=torch.float64)[None, :].expand(N0, N1)
    a = .5 + .25 * torch.sin(2 * math.pi * x / N1)
    b = .1 * torch.cos(2 * math.pi * y / N0)
    abc = torch.stack([a, b, a], dim=-1).contiguous()
    return abc.numpy() if as_numpy else abc
