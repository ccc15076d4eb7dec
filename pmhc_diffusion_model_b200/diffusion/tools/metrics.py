"""`MetricsRecord` — drop-in for diffusion/tools/metrics.py:8-40.

Same API and CSV format; the per-key sums stay on the device (the reference does a `.item()` host sync per
key per batch, metrics.py:17) and are read back once, in `mean()`.
"""
import csv
import os
from typing import Dict

import torch


class MetricsRecord:
    def __init__(self):
        self._sums = {}
        self._size = 0

    def add_batch(self, results: Dict[str, torch.Tensor]):
        batch_size = 0
        for key, data in results.items():
            s = data.detach().sum()
            self._sums[key] = self._sums[key] + s if key in self._sums else s
            batch_size = data.shape[0]
        self._size += batch_size

    def mean(self) -> Dict[str, float]:
        return {key: float(sum_) / self._size for key, sum_ in self._sums.items()}

    def save(self, path: str, epoch_number: int):
        keys = list(self._sums.keys())
        add_header = not os.path.isfile(path)
        with open(path, "at") as f:
            w = csv.writer(f, delimiter=",")
            if add_header:
                w.writerow(["epoch"] + keys)
            m = self.mean()
            w.writerow([epoch_number] + [round(m[key], 3) for key in keys])
