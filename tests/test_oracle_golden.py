"""CPU: the oracle restatement reproduces every golden vector generated from the reference."""
import pytest
import torch

from tests import util

TOL = 2e-5  # fp32 vs fp32, different op order


@pytest.mark.parametrize("name", util.golden_names())
def test_oracle_matches_reference_golden(name):
    fx = util.load_golden(name)
    got = util.run_oracle_on_fixture(fx)
    assert util.rel_err(got["out"], fx["out"]) < TOL
    if fx["idx"] is not None:
        assert torch.equal(got["idx"], fx["idx"]), "expert selection must be bit-exact"
    if fx["probs"] is not None:
        assert util.rel_err(got["probs"], fx["probs"]) < TOL
    if fx["lb_loss"] is not None:
        assert abs(got["lb_loss"].item() - fx["lb_loss"].item()) < 1e-5 * max(1.0, abs(fx["lb_loss"].item()))
    assert util.rel_err(got["dx"], fx["dx"]) < 10 * TOL
    assert util.rel_err(got["do"], fx["do"]) < 10 * TOL
    gmax = max(v.abs().max().item() for v in fx["dparams"].values())
    for k, g in fx["dparams"].items():
        assert k in got["dparams"], k
        assert util.rel_err(got["dparams"][k], g, floor=1e-3 * gmax) < 10 * TOL, k
    for k, v in fx["buffers_after"].items():
        if fx["training"] and k in got["buffers"]:
            assert util.rel_err(got["buffers"][k].float(), v.float(), floor=1e-6) < TOL, k


def test_golden_set_is_complete():
    names = set(util.golden_names())
    for variant in ("ave", "avqa", "avvp", "s4", "ms3"):
        for modality in ("audio", "visual"):
            assert f"{variant}_{modality}" in names
    assert "ave_audio_eval" in names
