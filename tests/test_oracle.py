"""Pins the oracle: fp64 DP, C lattice and the torch restatement against (1) golden vectors the
UNMODIFIED reference produced (oracle/gen_golden.py), (2) the warp-transducer known answer."""
import numpy as np
import pytest
import torch

from conftest import load_golden, pred_net_douts, rel_err
from oracle import clattice, ctc_dp, rnnt_dp, torch_path

RNNT_CASES = ["ref_rnnt_small_full", "ref_rnnt_small_ragged", "ref_rnnt_small_auxctc", "ref_rnnt_medium_ragged"]
CTC_CASES = ["ref_ctc_small_full", "ref_ctc_small_ragged", "ref_ctc_medium_ragged"]


def test_known_answer_dp():
    g = load_golden("known_answer_warp_transducer")
    lp = rnnt_dp.log_softmax(g["acts"].astype(np.float64))
    costs, glp = rnnt_dp.rnnt_loss_dense(lp, g["labels"], g["T"], g["U"], blank=0)
    assert abs(costs[0] - float(g["cost_published"])) < 2e-6
    dz = glp - np.exp(lp) * glp.sum(-1, keepdims=True)
    assert np.abs(dz[0, 0, 0] - g["grad_row_000_published"]).max() < 1e-6
    assert np.abs(dz[0, 1, 2] - g["grad_row_012_published"]).max() < 1e-6
    assert abs(float(g["cost_here"][0]) - float(g["cost_published"])) < 2e-6
    assert np.abs(g["grad_here"] - dz).max() < 1e-6


def test_aligner_smoke_fixture_dp_and_c():
    g = load_golden("ref_rnnt_aligner_smoke")
    costs, glp = rnnt_dp.rnnt_loss_dense(g["log_probs"], g["labels"], g["T"], g["U"], blank=0)
    assert np.abs(costs - g["costs"]).max() < 1e-5 * np.abs(g["costs"]).max()
    assert rel_err(glp, g["grad"]) < 1e-5
    # C restatement on the gathered pairs
    for b in range(2):
        T, U = int(g["T"][b]), int(g["U"][b])
        lp = g["log_probs"][b].astype(np.float64)
        lp2 = np.zeros(lp.shape[:2] + (2,))
        lp2[..., 0] = lp[..., 0]
        lp2[:, :U, 1] = lp[:, np.arange(U), g["labels"][b, :U]]
        cost, gam = clattice.rnnt_lattice(lp2, T, U)
        assert abs(cost - costs[b]) < 1e-9
        assert np.abs(-gam[..., 0] - glp[b][..., 0]).max() < 1e-9


@pytest.mark.parametrize("name", RNNT_CASES)
def test_rnnt_dp_vs_reference_golden(name):
    g = load_golden(name)
    douts = pred_net_douts(g).numpy()
    r = rnnt_dp.joint_loss_and_grads(
        g["eouts"], douts, g["param.w_enc.weight"], g["param.w_enc.bias"],
        g["param.w_dec.weight"], g["param.w_dec.bias"], g["param.output.weight"], g["param.output.bias"],
        g["ys"], g["elens"], g["ylens"], blank=int(g["hp.blank_id"]))
    assert abs(r["loss"] - float(g["lossdict.loss_rnnt"])) <= 1e-5 * abs(r["loss"])
    # logits themselves
    _, _, _, z = rnnt_dp.joint_logits(g["eouts"], douts, g["param.w_enc.weight"], g["param.w_enc.bias"],
                                      g["param.w_dec.weight"], g["param.w_dec.bias"],
                                      g["param.output.weight"], g["param.output.bias"])
    assert rel_err(z, g["logits"]) < 1e-5
    assert rel_err(r["d_w_out"], g["grad.output.weight"]) < 1e-4
    assert rel_err(r["d_b_out"], g["grad.output.bias"]) < 1e-4
    assert rel_err(r["d_w_dec"], g["grad.w_dec.weight"]) < 1e-4
    assert rel_err(r["d_b_dec"], g["grad.w_dec.bias"]) < 1e-4
    if float(g["meta_mtl_ctc_weight"]) == 0:      # eouts / w_enc also receive aux-CTC grads otherwise
        assert rel_err(r["d_w_enc"], g["grad.w_enc.weight"]) < 1e-4
        assert rel_err(r["d_eouts"], g["grad_eouts"]) < 1e-4
    else:
        c = ctc_dp.ctc_head_loss_and_grads(g["eouts"], g["param.ctc.output.weight"], g["param.ctc.output.bias"],
                                           g["ys"], g["elens"], g["ylens"], blank=int(g["hp.blank_id"]))
        w = float(g["meta_mtl_ctc_weight"])
        assert abs(c["loss"] - float(g["lossdict.loss_ctc"])) <= 1e-5 * abs(c["loss"])
        assert abs(r["loss"] + w * c["loss"] - float(g["loss_total"])) <= 1e-5 * float(g["loss_total"])
        assert rel_err(r["d_eouts"] + w * c["d_eouts"], g["grad_eouts"]) < 1e-4
        assert rel_err(w * c["d_w"], g["grad.ctc.output.weight"]) < 1e-4


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_dp_vs_reference_golden(name):
    g = load_golden(name)
    c = ctc_dp.ctc_head_loss_and_grads(g["eouts"], g["param.output.weight"], g["param.output.bias"],
                                       g["ys"], g["elens"], g["ylens"], blank=int(g["hp.blank_id"]))
    assert abs(c["loss"] - float(g["loss_total"])) <= 1e-5 * abs(c["loss"])
    assert rel_err(c["logits"], g["logits"]) < 1e-5
    assert rel_err(c["d_eouts"], g["grad_eouts"]) < 1e-4
    assert rel_err(c["d_w"], g["grad.output.weight"]) < 1e-4
    assert rel_err(c["d_b"], g["grad.output.bias"]) < 1e-4


def test_ctc_aligner_smoke_fixture():
    g = load_golden("ref_ctc_aligner_smoke")
    loss, nll, grad = ctc_dp.ctc_loss_and_grad(g["logits"], g["ys"], g["elens"], g["ylens"], blank=0)
    assert abs(loss - float(g["loss"])) < 1e-5 * abs(loss)
    assert rel_err(grad, g["grad"]) < 1e-5
    # C restatement
    for b in range(2):
        T, U = int(g["elens"][b]), int(g["ylens"][b])
        lp = ctc_dp.log_softmax(g["logits"][b, :T].astype(np.float64))
        nll_c, occ = clattice.ctc_lattice(lp, g["ys"][b, :U], blank=0)
        assert abs(nll_c - nll[b]) < 1e-9
        assert np.abs((np.exp(lp) - occ) / 2 - grad[b, :T]).max() < 1e-9


def test_ctc_zero_infinity_and_padding():
    g = load_golden("ref_ctc_small_ragged")
    logits = g["logits"]
    loss, nll, grad = ctc_dp.ctc_loss_and_grad(logits, g["ys"], g["elens"], g["ylens"], blank=0)
    assert nll[2] == 0.0 and np.all(grad[2] == 0)            # T=4 < U=6: infeasible -> zeroed
    assert np.all(grad[1, int(g["elens"][1]):] == 0)          # padded frames
    rows = grad[0, : int(g["elens"][0])].sum(-1)
    assert np.abs(rows).max() < 1e-12                         # rows sum to zero


def test_torch_path_matches_dp_random():
    rng = np.random.default_rng(7)
    B, T, U, V, He, Hd, J = 3, 11, 4, 13, 6, 5, 8
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    eouts, douts = f(B, T, He), np.tanh(f(B, U + 1, Hd))
    w_enc, b_enc, w_dec, b_dec, w_out, b_out = f(J, He) * .3, f(J) * .1, f(J, Hd) * .3, f(J) * .1, f(V, J) * .3, f(V) * .1
    ys = rng.integers(1, V, (B, U))
    tl, ul = np.array([11, 8, 5]), np.array([4, 2, 0])
    r = rnnt_dp.joint_loss_and_grads(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out, ys, tl, ul)
    te = [torch.from_numpy(a).requires_grad_() for a in (eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out)]
    loss = torch_path.rnnt_joint_loss(*te, torch.from_numpy(ys), torch.from_numpy(tl), torch.from_numpy(ul))
    loss.backward()
    assert abs(float(loss) - r["loss"]) < 1e-5 * r["loss"]
    for t, k in zip(te, ["d_eouts", "d_douts", "d_w_enc", "d_b_enc", "d_w_dec", "d_b_dec", "d_w_out", "d_b_out"]):
        assert rel_err(t.grad.numpy(), r[k]) < 1e-4, k
