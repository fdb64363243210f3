"""TEST INFRASTRUCTURE — generate tests/golden/*.pt FROM THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

For every variant x {audio, visual} adapter at a miniature shape, instantiates the
reference ``MoEAdapter`` (oracle/ref_harness.py), loads seeded weights, runs forward and
backward in fp32 on CPU and stores inputs, weights, outputs, input/parameter gradients and
the post-forward BatchNorm buffers.  The fixtures are what pins ``oracle/avmoe_oracle.py``
and the CUDA path to the reference on the GPU box, where /root/reference does not exist.

Notes
  * gates are drawn ~N(0.5,0.1) (reference zero-init makes the output identically 0).
  * AVS router noise is off (``is_training=False``) and MHA dropout is set to 0: both are
    RNG-dependent (SURVEY.md §8c caveat 3).
  * With is_bn=1 the reference's CPU backward trips a torch batch_norm_backward layout bug;
    ``make_bn_grads_contiguous`` avoids it without touching reference math (see its docstring).
"""
from __future__ import annotations

import dataclasses
import os
import sys

import torch

from . import avmoe_oracle as O
from . import ref_harness as H

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# miniature shapes: (C, N, Cx, Nx) for the "audio" adapter; visual swaps the pairs
MINI = dict(C=16, N=24, Cx=32, Nx=20)
MINI_HP = {
    "ave": dict(r=4, G=2, T=5),
    "avqa": dict(r=4, G=4, T=2),
    "avvp": dict(r=4, G=2, T=5),
    "s4": dict(r=4, G=2, T=7),
    "ms3": dict(r=2, G=2, T=7),
}
BT = 5


def golden_cases():
    for variant in ("ave", "avqa", "avvp", "s4", "ms3"):
        for modality in ("audio", "visual"):
            kw = dict(O.VARIANT_DEFAULTS[variant])
            kw.update(MINI_HP[variant])
            if modality == "audio":
                shp = dict(C=MINI["C"], N=MINI["N"], Cx=MINI["Cx"], Nx=MINI["Nx"])
            else:
                shp = dict(C=MINI["Cx"], N=MINI["Nx"], Cx=MINI["C"], Nx=MINI["N"])
            yield f"{variant}_{modality}", O.AdapterSpec(variant=variant, **shp, **kw)


def generate_one(name: str, spec: O.AdapterSpec, training: bool = True):
    seed = abs(hash(name)) % 1000 if False else sum(ord(c) for c in name)
    sd = O.make_state_dict(spec, seed=seed)
    x = O.make_tokens(BT, spec.N, spec.C, seed + 1).requires_grad_()
    o = O.make_tokens(BT, spec.Nx, spec.Cx, seed + 2).requires_grad_()
    ref = H.build_reference(spec, sd)
    for m in ref.modules():
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    H.make_bn_grads_contiguous(ref)
    res = H.run_reference(ref, spec, x, o, training=training)
    g = torch.Generator().manual_seed(seed + 3)
    gout = torch.randn(res["out"].shape, generator=g)
    loss = (res["out"] * gout).sum()
    if res["lb_loss"] is not None:
        loss = loss + res["lb_loss"]
    gprobs = None
    if res["probs"] is not None:
        gprobs = torch.randn(res["probs"].shape, generator=g)
        loss = loss + (res["probs"] * gprobs).sum()
    loss.backward()
    fx = {
        "name": name, "spec": dataclasses.asdict(spec), "training": training,
        "state_dict": {k: v.clone() for k, v in sd.items()},
        "x": x.detach().clone(), "o": o.detach().clone(), "gout": gout, "gprobs": gprobs,
        "out": res["out"].detach().clone(),
        "idx": None if res["idx"] is None else res["idx"].clone(),
        "probs": None if res["probs"] is None else res["probs"].detach().clone(),
        "lb_loss": None if res["lb_loss"] is None else res["lb_loss"].detach().clone(),
        "dx": x.grad.clone(), "do": o.grad.clone(),
        "dparams": {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None},
        "buffers_after": {k: v.clone() for k, v in ref.state_dict().items() if "running_" in k or "num_batches" in k},
    }
    return fx


def main():
    if not H.ref_loader.available():
        print("reference tree not available; nothing generated", file=sys.stderr)
        return 1
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, spec in golden_cases():
        fx = generate_one(name, spec)
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: out|max|={fx['out'].abs().max():.4f} idx={None if fx['idx'] is None else fx['idx'].tolist()} "
              f"-> {os.path.getsize(path)/1024:.0f} KiB")
    # eval-mode fixture for the BatchNorm variant (running statistics path)
    name, spec = "ave_audio", dict(golden_cases())["ave_audio"]
    fx = generate_one(name + "_eval", spec, training=False)
    torch.save(fx, os.path.join(GOLDEN_DIR, name + "_eval.pt"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
