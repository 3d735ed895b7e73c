"""CPU: pins the oracle's alternate blocks (GRU / LSTM decoder cells, continuous latent plan, RGB-D
static encoder) against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden_alt.py)."""
import pytest

from hulc2_b200.config import hulc2_config
from oracle import hulc2_oracle as O

from helpers import ALT_CASES, alt_case_inputs, alt_gt, alt_keys, assert_close, build_alt_model, oracle_params


@pytest.mark.parametrize("tag", list(ALT_CASES))
def test_alt_train_step_matches_reference(tag):
    kw, batch, draw = alt_case_inputs(tag)
    m = build_alt_model(tag)
    P = oracle_params(m)
    out = O.training_step(batch, {mod: {"plan_idx": draw[mod]} for mod in batch}, P, hulc2_config(pkg="x", **kw))
    out["loss"].backward()
    assert_close(out["loss"], alt_gt(f"{tag}/loss"), 1e-6, "loss")
    for k in alt_keys(f"{tag}/log/"):
        assert_close(out[k[len(tag) + 5:]], alt_gt(k), 2e-6, k)
    n = 0
    for k in alt_keys(f"{tag}/grad_norm/"):
        name = k[len(tag) + 11:]
        assert_close(P[name].grad.double().norm(), alt_gt(k), 1e-3 if name == "logit_scale" else 2e-5, k)
        n += 1
    assert n >= 95
    for k in alt_keys(f"{tag}/grad/"):
        assert_close(P[k[len(tag) + 6:]].grad, alt_gt(k), 5e-5, k)


def test_model_state_dict_names_of_alternate_blocks():
    """nn.GRU / nn.LSTM containers keep torch's parameter names and gate-stacked shapes; the continuous plan doubles fc_state."""
    sd = build_alt_model("gauss_gru").state_dict()
    assert tuple(sd["action_decoder.rnn.weight_hh_l1"].shape) == (3 * 256, 256)
    assert tuple(sd["plan_proposal.fc_state.0.weight"].shape) == (512, 256)
    assert tuple(sd["plan_recognition.fc_state.0.weight"].shape) == (512, 4096)
    sd = build_alt_model("lstm").state_dict()
    assert tuple(sd["action_decoder.rnn.weight_ih_l0"].shape) == (4 * 256, 64 + 32 + 1024)
    sd = build_alt_model("rgbd_rw").state_dict()
    assert tuple(sd["perceptual_encoder.depth_static_encoder.conv_model.0.weight"].shape) == (32, 1, 8, 8)


@pytest.mark.parametrize("tag", ["lstm", "gauss_gru"])
def test_alt_validation_matches_reference(tag):
    """lmp_val (hulc2.py:247-334) with an LSTM decoder / a continuous plan + GRU decoder against the reference's validation_step."""
    import torch

    from helpers import alt_val_inputs

    kw, batch, noise = alt_val_inputs(tag)
    P = {k: v.detach() for k, v in oracle_params(build_alt_model(tag), requires_grad=False).items()}
    cfg = hulc2_config(pkg="x", **kw)
    for mod, db in batch.items():
        with torch.no_grad():
            emb = O.perceptual_encoder(db["rgb_obs"], db["depth_obs"], P)
            goal = O.language_goal(db["lang"], P) if "lang" in mod else O.visual_goal(emb[:, -1], P)
            (ppp, lpp, ppr, lpr, kl, mae_pp, mae_pr, sr_pp, sr_pr, _) = O.lmp_val(emb, goal, db["actions"], db["state_info"]["robot_obs"], noise[mod], P, cfg)
        assert_close(ppp, alt_gt(f"{tag}/val/out/sampled_plan_pp_{mod}"), 1e-6, "plan pp")
        assert_close(ppr, alt_gt(f"{tag}/val/out/sampled_plan_pr_{mod}"), 1e-6, "plan pr")
        assert_close(lpp, alt_gt(f"{tag}/val/log/val_act/{mod}_act_loss_pp"), 1e-5)
        assert_close(lpr, alt_gt(f"{tag}/val/log/val_act/{mod}_act_loss_pr"), 1e-5)
        assert_close(kl, alt_gt(f"{tag}/val/log/val_kl/{mod}_kl_loss"), 1e-5)
        assert_close(mae_pp.mean(), alt_gt(f"{tag}/val/log/val_total_mae/{mod}_total_mae_pp"), 1e-4)
        assert_close(mae_pr.mean(), alt_gt(f"{tag}/val/log/val_total_mae/{mod}_total_mae_pr"), 1e-4)
        assert_close(sr_pp, alt_gt(f"{tag}/val/log/val_grip/{mod}_grip_sr_pp"), 1e-6)
