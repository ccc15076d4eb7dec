"""Shared fixture loading for the parity tests (test infrastructure)."""
import os

import torch

from oracle import egnn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name: str) -> dict:
    case = torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
    tag = case.get("params")
    if tag == "shipped":
        case["params"] = torch.load(os.path.join(GOLDEN, "shipped_params.pt"), map_location="cpu")
    elif isinstance(tag, tuple) and tag[0] == "random":
        case["params"] = orc.random_params(seed=tag[1])
    return case


def noise_dict(n: dict) -> dict:
    """Fixture noise (plain tensors) -> oracle noise dict."""
    return {"frames": {"quats": n["q"], "trans": n["x"]}, "torsions": n["tors"]}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(1, max|b|): the relative error used for the 1e-4 / 1e-2 gates."""
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def real_rows(batch: dict) -> torch.Tensor:
    return batch["mask"].bool()
