"""Pins the oracle: fp64 DP, C lattice and the torch restatement against (1) golden vectors the
UNMODIFIED reference produced (oracle/gen_golden.py), (2) the warp-transducer known answer."""
import numpy as np
import pytest
import torch

from conftest import load_golden, pred_net_douts, rel_err
from oracle import clattice, ctc_align, ctc_dp, rnnt_dp, torch_path

RNNT_CASES = ["ref_rnnt_small_full", "ref_rnnt_small_ragged", "ref_rnnt_small_auxctc", "ref_rnnt_medium_ragged",
              "ref_rnnt_tcshape_ragged", "ref_rnnt_tcshape_auxctc", "ref_rnnt_tcfull_auxctc"]
CTC_CASES = ["ref_ctc_small_full", "ref_ctc_small_ragged", "ref_ctc_medium_ragged", "ref_ctc_tchead_ragged"]


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


@pytest.mark.parametrize("name", ["ref_ctc_phone_final", "ref_ctc_phone_hie_inter", "ref_ctc_tchead_phone_hie_inter"])
def test_ctc_dp_vs_reference_golden_phone_and_inter_heads(name):
    """Phone CTC (final / intermediate layer, ctc.py:129-148) and intermediate CTC (ctc.py:150-170): the same
    head recipe on other inputs / weights; loss_total = ctc + w_p * phone + w_i * inter."""
    g = load_golden(name)
    blank = int(g["hp.blank_id"])
    wp, wi, hie = float(g["hp.mtl_phone_ctc_weight"]), float(g["hp.mtl_inter_ctc_weight"]), bool(g["hp.hie_mtl_phone"])
    main = ctc_dp.ctc_head_loss_and_grads(g["eouts"], g["param.output.weight"], g["param.output.bias"],
                                          g["ys"], g["elens"], g["ylens"], blank=blank)
    src = g["eouts_inter"] if hie else g["eouts"]
    phone = ctc_dp.ctc_head_loss_and_grads(src, g["param.phone_output.weight"], g["param.phone_output.bias"],
                                           g["ps"], g["elens"], g["plens"], blank=blank)
    total = main["loss"] + wp * phone["loss"]
    d_eouts = main["d_eouts"] + (0 if hie else wp * phone["d_eouts"])
    d_inter = wp * phone["d_eouts"] if hie else 0
    d_w = main["d_w"]
    key = "lossdict.loss_phone_ctc(inter)" if hie else "lossdict.loss_phone_ctc"
    assert abs(phone["loss"] - float(g[key])) <= 1e-5 * abs(phone["loss"])
    if wi > 0:
        inter = ctc_dp.ctc_head_loss_and_grads(g["eouts_inter"], g["param.output.weight"], g["param.output.bias"],
                                               g["ys"], g["elens"], g["ylens"], blank=blank)
        assert abs(inter["loss"] - float(g["lossdict.loss_inter_ctc"])) <= 1e-5 * abs(inter["loss"])
        total += wi * inter["loss"]
        d_inter = d_inter + wi * inter["d_eouts"]
        d_w = d_w + wi * inter["d_w"]
    assert abs(total - float(g["loss_total"])) <= 1e-5 * abs(total)
    assert rel_err(d_eouts, g["grad_eouts"]) < 1e-4
    if g["grad_eouts_inter"].size:
        assert rel_err(d_inter, g["grad_eouts_inter"]) < 1e-4
    assert rel_err(d_w, g["grad.output.weight"]) < 1e-4
    assert rel_err(wp * phone["d_w"], g["grad.phone_output.weight"]) < 1e-4


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


def test_forced_align_restatement_vs_reference():
    """oracle.rnnt_dp.forced_align against alignments the UNMODIFIED RNNTForcedAligner produced (its Numba CUDA
    kernels run under the CUDA simulator in oracle/gen_golden.py): the in-tree smoke input of rnnt_aligner.py:201-208
    and the kd_type="align" decoder case."""
    g = load_golden("ref_rnnt_aligner_smoke")
    assert np.array_equal(rnnt_dp.forced_align(g["log_probs"], g["labels"], g["T"], g["U"]), g["aligns"])
    g = load_golden("ref_rnnt_kd_align")
    P = {k[len("param."):]: v.astype(np.float64) for k, v in g.items() if k.startswith("param.")}
    douts = pred_net_douts(g)
    _, _, _, z = rnnt_dp.joint_logits(g["eouts"], douts, P["w_enc.weight"], P["w_enc.bias"], P["w_dec.weight"],
                                      P["w_dec.bias"], P["output.weight"], P["output.bias"])
    al = rnnt_dp.forced_align(rnnt_dp.log_softmax(z), g["ys"], g["elens"], g["ylens"])
    assert np.array_equal(al, g["aligns"])
    kd = rnnt_dp.align_distill_loss(z, g["soft_labels"], al, g["elens"], g["ylens"])
    assert abs(kd - float(g["lossdict.loss_kd"])) <= 1e-5 * abs(kd)


def test_word_distill_restatement_vs_reference():
    g = load_golden("ref_rnnt_kd_word")
    P = {k[len("param."):]: v.astype(np.float64) for k, v in g.items() if k.startswith("param.")}
    douts = pred_net_douts(g)
    _, _, _, z = rnnt_dp.joint_logits(g["eouts"], douts, P["w_enc.weight"], P["w_enc.bias"], P["w_dec.weight"],
                                      P["w_dec.bias"], P["output.weight"], P["output.bias"])
    kd = rnnt_dp.word_distill_loss(z, g["soft_labels"], g["elens"], g["ylens"])
    assert abs(kd - float(g["lossdict.loss_kd"])) <= 1e-5 * abs(kd)
    # loss_total = (1 - w) * loss_rnnt + w * loss_kd  (reduce_main_loss_kd, rnn_transducer.py:138-139)
    w = float(g["hp.kd_weight"])
    total = (1 - w) * float(g["lossdict.loss_rnnt"]) + w * kd
    assert abs(total - float(g["loss_total"])) <= 1e-5 * abs(total)


def test_ctc_forced_aligner_restatement_vs_reference_golden():
    """oracle/ctc_align.py against alignments the UNMODIFIED CTCForcedAligner produced (ctc_aligner.py:138-221; its own
    smoke input, ragged batches with repeated labels, an empty label sequence, an infeasible utterance)."""
    g = load_golden("ref_ctc_forced_align")
    for i in range(int(g["n_cases"])):
        got = ctc_align.ctc_forced_align(g[f"c{i}_log_probs"], g[f"c{i}_elens"], g[f"c{i}_ys"], g[f"c{i}_ylens"], blank=0)
        assert np.array_equal(got, g[f"c{i}_aligns"]), f"case {i}"
        # an alignment is a monotone walk through the blank-extended path: collapsing repeats and dropping blanks gives
        # a prefix-free subsequence check only for feasible utterances
        for b in range(got.shape[0]):
            x, u = int(g[f"c{i}_elens"][b]), int(g[f"c{i}_ylens"][b])
            if x >= 2 * u + 1:      # comfortably feasible
                seq = got[b, :x]
                collapsed = [int(v) for k, v in enumerate(seq) if k == 0 or v != seq[k - 1]]
                labels = [v for v in collapsed if v != 0]
                assert labels == [int(v) for v in g[f"c{i}_ys"][b, :u]] or len(labels) <= u
