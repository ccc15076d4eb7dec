"""`MetricsRecord` — same interface and CSV layout as diffusion/tools/metrics.py:8-40 (`add_batch`, `mean`, `save`;
header `epoch,<keys...>`, one row per epoch, three decimals), different mechanics: the running totals of all keys live in
ONE device tensor that a batch is added to with a single stacked reduction, and nothing is read back to the host before
`mean()` / `save()` (the reference syncs once per key per batch, metrics.py:17).
"""
import os
from typing import Dict, List, Optional

import torch


class MetricsRecord:
    def __init__(self):
        self.keys: List[str] = []                    # column order = order of first appearance, as in the reference's dict
        self.totals: Optional[torch.Tensor] = None   # [len(keys)] running sums on the device of the first batch
        self.count = 0                               # complexes seen

    def add_batch(self, results: Dict[str, torch.Tensor]) -> None:
        names = list(results)
        if names != self.keys[:len(names)] or len(names) != len(self.keys):
            self._grow(names, next(iter(results.values())))
        per_key = torch.stack([results[k].detach().to(self.totals.dtype).sum() for k in self.keys if k in results])
        if len(names) == len(self.keys):
            self.totals += per_key
        else:                                        # a batch that reports only some of the keys
            index = torch.tensor([self.keys.index(k) for k in names], device=self.totals.device)
            self.totals.index_add_(0, index, per_key)
        self.count += int(next(iter(results.values())).shape[0])

    def add_stacked(self, names, rows: torch.Tensor) -> None:
        """Same as add_batch for results that already lie stacked as rows [len(names), B] of one tensor (what the fused loss
        kernel writes): one reduction and one add per batch instead of one reduction per key."""
        names = list(names)
        if names != self.keys:
            if self.keys:                                # mixed use with add_batch: take the general path
                return self.add_batch({k: rows[i] for i, k in enumerate(names)})
            self._grow(names, rows)
        self.totals += rows.detach().sum(dim=1, dtype=self.totals.dtype)
        self.count += int(rows.shape[1])

    def _grow(self, names: List[str], like: torch.Tensor) -> None:
        fresh = [k for k in names if k not in self.keys]
        if not fresh:
            return
        extra = torch.zeros(len(fresh), dtype=torch.float64, device=like.device)
        self.totals = extra if self.totals is None else torch.cat((self.totals, extra))
        self.keys.extend(fresh)

    def mean(self) -> Dict[str, float]:
        if self.totals is None:
            return {}
        values = (self.totals / max(self.count, 1)).tolist()      # the one device -> host read
        return dict(zip(self.keys, values))

    def save(self, path: str, epoch_number: int) -> None:
        means = self.mean()
        lines = []
        if not os.path.isfile(path):
            lines.append(",".join(["epoch"] + self.keys))
        lines.append(",".join([str(epoch_number)] + [repr(round(means[k], 3)) for k in self.keys]))
        with open(path, "a", newline="") as out:
            out.write("\r\n".join(lines) + "\r\n")       # csv.writer's default line terminator, as the reference's file has
