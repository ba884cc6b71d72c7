"""Synthetic-bag replacement for the reference's ``data/dataset.py`` (absent from its tree: ``main.py:10`` does
``from data.dataset import *`` and builds ``TCGA_Dataset / IvYGAP_Dataset / CPTAC_Dataset(excel_wsi=..., args=args)``,
``main.py:103-127,307-346``).

Items follow the 6-tuple the trainers unpack (``train_test.py:817``):
``(x_path [N, 1024], x_path_20x, x_omic [431], x_omic_tumor [59], x_omic_immune [361], label [12])`` with the label columns of
``train_test.py:820``: 0 IDH, 1 1p19q, 2 CDKN, 3 His, 4 Grade, 5 Diag (WHO 2021), 6 His_2class, 7 Subtype, 8 survival bin,
9 censor, 10 unused, 11 survival time.  Bags are ``dml_b200.synth.synthetic_bag`` (per-feature standardised relu(randn), SURVEY
8d) drawn per index, so every rank and every epoch sees the same bag for the same index without any file.
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import synth

LABEL_COLUMNS = 12


class SyntheticBagDataset(Dataset):
    def __init__(self, num_bags: int, n_patches: Union[int, Tuple[int, int]] = 2500, seed: int = 42,
                 bag_dtype: torch.dtype = torch.float32, two_scales: bool = False, cache: bool = True):
        """n_patches: a fixed bag length (the reference's 2 500) or a (lo, hi) range drawn per bag, rounded to even.
        cache: keep generated items in host memory (drawing 16 384 x 1 024 normals costs ~0.1 s per bag on the CPU)."""
        self.num_bags, self.seed, self.bag_dtype, self.two_scales = int(num_bags), int(seed), bag_dtype, two_scales
        self._cache = {} if cache else None
        if isinstance(n_patches, int):
            self.lengths = [n_patches] * self.num_bags
        else:
            lo, hi = n_patches
            rng = np.random.Generator(np.random.PCG64(seed))
            self.lengths = [int(v) // 2 * 2 for v in rng.integers(lo, hi + 1, size=self.num_bags)]

    def __len__(self):
        return self.num_bags

    def __getitem__(self, i):
        if self._cache is not None and i in self._cache:
            return self._cache[i]
        item = self._make(i)
        if self._cache is not None:
            self._cache[i] = item
        return item

    def _make(self, i):
        n = self.lengths[i]
        s = self.seed * 100003 + i
        b = synth.synthetic_bag(n, seed=s, B=1)
        rng = np.random.Generator(np.random.PCG64(s))
        label = torch.zeros(LABEL_COLUMNS, dtype=torch.float32)
        label[0:4] = torch.from_numpy(rng.integers(0, 2, size=4)).float()
        label[4] = float(b["label_grade"][0])
        label[5] = float(b["label_diag"][0])
        label[6] = float(rng.integers(0, 2))
        label[7] = float(rng.integers(0, 3))
        label[8] = float(b["label_surv"][0])
        label[9] = float(b["censor"][0])
        label[11] = float(rng.random() * 100.0)
        x_path = b["x_path"][0].to(self.bag_dtype)
        x20 = x_path if self.two_scales else torch.zeros(1)
        x_omic = synth.normal((431,), s, "x_omic")
        return x_path, x20, x_omic, b["x_omic_tumor"][0], b["x_omic_immune"][0], label


def _from_reference_args(excel_wsi, args):
    n = len(excel_wsi) if isinstance(excel_wsi, Sequence) and not isinstance(excel_wsi, str) else int(getattr(args, "synthetic_bags", 64))
    return dict(num_bags=n, n_patches=getattr(args, "synthetic_patches", 2500), seed=int(getattr(args, "seed", 42)))


class TCGA_Dataset(SyntheticBagDataset):
    """Same constructor as the reference's dataset classes: one synthetic bag per row of the slide list."""

    def __init__(self, excel_wsi=None, args=None):
        super().__init__(**_from_reference_args(excel_wsi, args))


class IvYGAP_Dataset(TCGA_Dataset):
    pass


class CPTAC_Dataset(TCGA_Dataset):
    pass


__all__ = ["SyntheticBagDataset", "TCGA_Dataset", "IvYGAP_Dataset", "CPTAC_Dataset", "LABEL_COLUMNS"]
