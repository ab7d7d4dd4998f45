"""Synthetic stereo data of the shapes BASELINE.json names (configs 3-5).  Test/bench data only.

Images: a multi-scale random texture (blocky cells at several scales, 3x3 box-filtered so that the FAST
statistics resemble the reference's KITTI images at threshold 15: ~18 % compass candidates, ~7 % corners
before NMS, ~7 k keypoints after NMS vs 21 % / 5.7 % / 6.3 k on test_data/kitti) in a "world" strip wider
than the image; the left image is a crop, the right image is the same
crop displaced per 8-row band by a disparity in [2, max_disp) (rectified stereo: matches lie on the
same row), plus independent +-noise on both so that descriptors differ.  Deterministic per
(seed, device type); generated with torch so that 10k-pair batches are produced on the GPU in
seconds.  torch is plumbing here (allocation + RNG), never the measured path.
"""
import numpy as np
import torch


def stereo_pairs(n_pairs, rows=376, cols=1241, seed=0, device="cpu", max_disp=96, noise=2,
                 scales=((32, 70.0), (16, 50.0), (8, 45.0)), band=8, box=3):
    """returns u8 tensor [n_pairs, 2, rows, cols] (left, right)"""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    W = cols + max_disp + 8
    tex = torch.zeros((n_pairs, 1, rows, W), device=dev, dtype=torch.float32)
    for s, amp in scales:
        grid = torch.rand((n_pairs, 1, rows // s + 2, W // s + 2), generator=g, device=dev)
        up = torch.repeat_interleave(torch.repeat_interleave(grid, s, dim=2), s, dim=3)
        tex += amp * up[:, :, :rows, :W]
    if box > 1:
        pad = (box // 2, box - 1 - box // 2, box // 2, box - 1 - box // 2)
        tex = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(tex, pad, mode="replicate"), box, stride=1)
    tex = tex[:, 0]
    tex = tex - tex.amin(dim=(1, 2), keepdim=True)
    tex = tex * (235.0 / tex.amax(dim=(1, 2), keepdim=True)) + 10.0
    n_bands = (rows + band - 1) // band
    disp = torch.randint(2, max_disp, (n_pairs, n_bands), generator=g, device=dev)
    disp_rows = torch.repeat_interleave(disp, band, dim=1)[:, :rows]              # [n, rows]
    base = torch.arange(cols, device=dev)[None, None, :]
    left = tex[:, :, :cols]
    idx = (base + disp_rows[:, :, None]).expand(n_pairs, rows, cols)
    right = torch.gather(tex, 2, idx)
    out = torch.empty((n_pairs, 2, rows, cols), device=dev, dtype=torch.uint8)
    for k, im in enumerate((left, right)):
        nz = torch.randint(-noise, noise + 1, im.shape, generator=g, device=dev).to(torch.float32)
        out[:, k] = torch.clamp(torch.round(im + nz), 0, 255).to(torch.uint8)
    return out


def hamming_sets(nq, nt, seed=0, planted=0.10, max_flips=20):
    """BASELINE config 5: iid Bernoulli(1/2) 256-bit descriptors (uint8 [n,32]); a `planted` fraction of
    the query rows are near-duplicates (<= max_flips flipped bits) of random train rows."""
    rq, rt = np.random.default_rng(seed), np.random.default_rng(seed + 1)
    q = rq.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rt.integers(0, 256, (nt, 32), dtype=np.uint8)
    n_pl = int(planted * nq)
    if n_pl and nt:
        rows = rq.choice(nq, n_pl, replace=False)
        src = rq.integers(0, nt, n_pl)
        dup = t[src].copy()
        for i in range(n_pl):
            bits = rq.choice(256, int(rq.integers(0, max_flips + 1)), replace=False)
            for b in bits:
                dup[i, b >> 3] ^= np.uint8(1 << (b & 7))
        q[rows] = dup
    return q, t
