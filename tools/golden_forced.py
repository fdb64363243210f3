"""Diagnostic: golden fixtures through the forced GEMM paths on fp32, tensor-core vs exact SIMT GEMMs (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avmoe_b200 import ops
from tests import util
from tests.test_functional_emulated import build_module, run_module_on_fixture

names = sys.argv[1:] or util.golden_names()
for impl in (0, 1):
    ops._GEMM_IMPL_OVERRIDE = impl
    ops.USE_GEMM_ATTENTION = ops.USE_GEMM_EXPERTS = "always"
    for n in names:
        fx = util.load_golden(n)
        mod = build_module(fx["spec"], fx["state_dict"], kernels=None)
        got = run_module_on_fixture(mod, fx, device="cuda")
        row = {k: round(util.rel_err(got[k], fx[k]), 5) for k in ("out", "dx", "do") if k in got and k in fx}
        worst = max(((util.rel_err(got["dparams"][k], g), k) for k, g in fx["dparams"].items()), default=(0, ""))
        print(f"impl={impl} {n:14s} {row} worst param grad {worst[0]:.4f} {worst[1]}", flush=True)
