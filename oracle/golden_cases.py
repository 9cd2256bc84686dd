"""TEST INFRASTRUCTURE ONLY -- the table of golden cases and the digest format.

``oracle/make_golden.py`` runs the REAL reference (loaded by path, build container
only) on these cases and writes ``tests/golden/<name>.pt``; the tests rebuild the
inputs from the same kwargs (seeded CPU generators) and compare.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from erd_b200.synth import Batch, make_batch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# name -> (make_batch kwargs, store_full)
CASES: Dict[str, tuple] = {
    # small enough to commit inputs and every output tensor
    'tiny_40_40': (dict(num_imgs=2, img_hw=(96, 128), ori=40, seed=11, num_gt=3, gt_size_pow=1.5), True),
    # mixed padded shapes (invalid anchors), one image without GT, trained-like teacher
    'mixed_pad': (dict(num_imgs=3, img_hw=(320, 480), ori=40, seed=12, num_gt=[3, 0, 4], mode='trained',
                       pad_shapes=[(320, 480), (256, 480), (320, 352)],
                       img_shapes=[(320, 470), (250, 480), (311, 350)], gt_size_pow=2.0), False),
    # BASELINE.json configs[0]: 2 synthetic 800x1333 images, 40+40
    'cfg1_40_40': (dict(num_imgs=2, img_hw=(800, 1333), ori=40, seed=1234, gt_size_pow=1.0), False),
    # 40+40 with small boxes so all five levels hold positives, trained-like teacher (NMS busy)
    'cfg1_trained': (dict(num_imgs=2, img_hw=(800, 1333), ori=40, seed=77, num_gt=(5, 9), mode='trained',
                          gt_size_pow=2.5), False),
    # BASELINE.json configs[3]: 70+10 split
    'cfg4_70_10': (dict(num_imgs=2, img_hw=(800, 1333), ori=70, seed=4321, mode='trained', gt_size_pow=2.0), False),
    # BASELINE.json configs[4] (one image of it): dense scene, 100 GT, 1600x1600
    'dense_1600': (dict(num_imgs=1, img_hw=(1600, 1600), ori=40, seed=55, num_gt=100, mode='trained',
                        gt_size_pow=3.0), False),
}

N_SAMPLES = 4096


def case_batch(name: str) -> Batch:
    return make_batch(**CASES[name][0])


def grad_digest(t: torch.Tensor, seed: int) -> dict:
    """Summary of a gradient tensor small enough to commit: moments in float64,
    N_SAMPLES seeded positions, and every non-zero position when the tensor is sparse."""
    flat = t.detach().reshape(-1).cpu()
    d = flat.double()
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, flat.numel(), (min(N_SAMPLES, flat.numel()),), generator=g)
    out = dict(shape=tuple(t.shape), sum=float(d.sum()), abssum=float(d.abs().sum()),
               sqsum=float((d * d).sum()), idx=idx, val=flat[idx].clone(),
               nnz=int((flat != 0).sum()))
    if out['nnz'] * 20 < flat.numel():
        nz = (flat != 0).nonzero().squeeze(1)
        out['nz_idx'] = nz
        out['nz_val'] = flat[nz].clone()
    return out


def load_golden(name: str) -> dict:
    return torch.load(os.path.join(GOLDEN_DIR, name + '.pt'), weights_only=False)
