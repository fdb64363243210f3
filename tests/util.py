"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import os

import torch

from oracle import avmoe_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)
    fx["spec"] = O.AdapterSpec(**fx["spec"])
    return fx


def rel_err(a: torch.Tensor, b: torch.Tensor, floor: float = 0.0) -> float:
    """max|a-b| / max(max|b|, floor): normalised max error (b is the reference side)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(b.abs().max().item(), floor, 1e-30)
    return (a - b).abs().max().item() / denom


def fro_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run_oracle_on_fixture(fx, dtype=torch.float32):
    """Oracle fwd+bwd on a golden fixture's inputs.  Returns dict with out/idx/probs/lb/dx/do/dparams/buffers."""
    spec = fx["spec"]
    sd = {k: (v.to(dtype).clone().requires_grad_() if v.is_floating_point() else v.clone())
          for k, v in fx["state_dict"].items()}
    x = fx["x"].to(dtype).clone().requires_grad_()
    o = fx["o"].to(dtype).clone().requires_grad_()
    nb = {}
    res = O.moe_adapter_forward(x, o, sd, spec, training=fx["training"], new_buffers=nb)
    loss = (res["out"] * fx["gout"].to(dtype)).sum()
    if fx["lb_loss"] is not None:
        loss = loss + res["lb_loss"]
    if fx["gprobs"] is not None:
        loss = loss + (res["probs"] * fx["gprobs"].to(dtype)).sum()
    loss.backward()
    return {"out": res["out"].detach(), "idx": res["idx"], "probs": res["probs"].detach(),
            "lb_loss": None if res["lb_loss"] is None else res["lb_loss"].detach(),
            "dx": x.grad, "do": o.grad,
            "dparams": {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None},
            "buffers": nb}
