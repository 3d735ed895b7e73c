"""Checkpoint surface: every evaluation caller of the reference does
``Hulc2.load_from_checkpoint(ckpt[, perceptual_encoder=...]).freeze()`` and then
``model.action_decoder._setup_action_bounds(dir, None, None, True)`` (hulc2/evaluation/manager_lmp.py:91-109), and training
may warm-start through ``initialize_pretrained_weights`` (hulc2/utils/utils.py:36-45).  A Lightning-format checkpoint
(``state_dict`` + ``hyper_parameters`` with the REFERENCE's ``_target_`` paths) written from the unmodified reference model
must load into the mirror classes; host-side logic only, no CUDA needed."""
import copy
import os
import warnings

import pytest
import torch

from hulc2_b200._compat import instantiate
from hulc2_b200.config import hulc2_config
from hulc2_b200.models.hulc2 import Hulc2
from hulc2_b200.utils.checkpoint import initialize_pretrained_weights, save_checkpoint, to_plain
from oracle.ref_import import reference_available

SMALL = dict(hidden_size=32)


def _small_cfg(pkg, max_window_size=32):
    cfg = hulc2_config(pkg=pkg, dropout_p=0.1, max_window_size=max_window_size, **SMALL)
    cfg["plan_recognition"].update(encoder_hidden_size=64, fc_hidden_size=96)
    cfg["proj_vis_lang"]["im_dim"] = 96
    cfg["distribution"].update(category_size=8, class_size=8)
    return cfg


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_reference_written_checkpoint_loads_into_the_mirror(tmp_path):
    from oracle.ref_import import make_reference_model

    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = make_reference_model(_small_cfg("hulc2"))
    # what pytorch_lightning's ModelCheckpoint writes for the reference module (hyper_parameters = constructor arguments,
    # filled in by setup_input_sizes, with the reference's own _target_ paths)
    hp = to_plain(_small_cfg("hulc2"))
    for k in ("_target_", "_recursive_"):
        hp.pop(k)
    latent = ref.perceptual_encoder.latent_size
    hp["plan_proposal"].update(perceptual_features=latent, plan_features=64)
    hp["plan_recognition"].update(in_features=latent, plan_features=64)
    hp["visual_goal"]["in_features"] = latent
    hp["action_decoder"].update(perceptual_features=latent, plan_features=64)
    ckpt = {"epoch": 7, "global_step": 1234, "pytorch-lightning_version": "1.8.6", "state_dict": ref.state_dict(), "hyper_parameters": hp}
    path = tmp_path / "saved_models" / "epoch=7.ckpt"
    path.parent.mkdir()
    torch.save(ckpt, str(path))

    m = Hulc2.load_from_checkpoint(str(path))
    assert type(m).__module__ == "hulc2_b200.models.hulc2"
    assert type(m.action_decoder).__module__.startswith("hulc2_b200.") and type(m.plan_recognition).__module__.startswith("hulc2_b200.")
    sd_ref, sd = ref.state_dict(), m.state_dict()
    assert list(sd) == list(sd_ref)
    for k in sd_ref:
        assert sd[k].shape == sd_ref[k].shape and torch.equal(sd[k].cpu(), sd_ref[k]), k
    assert m.current_epoch == 7 and m.global_step == 1234
    assert m.kl_beta == 0.01 and m.plan_recognition.fc.out_features == 96           # hyper-parameters came from the checkpoint
    m.freeze()
    assert not m.training and not any(p.requires_grad for p in m.parameters())

    # keyword override, as manager_lmp.py:93-103 does for checkpoints older than the spatial_softmax_temp kwarg
    pe = copy.deepcopy(hp["perceptual_encoder"])
    pe["rgb_static"].pop("spatial_softmax_temp")
    pe["rgb_static"]["spatial_softmax_temp"] = 1.0
    m2 = Hulc2.load_from_checkpoint(str(path), perceptual_encoder=pe)
    assert torch.equal(m2.state_dict()["logit_scale"], sd_ref["logit_scale"])

    # action bounds from <package>/<dataset_dir>/training/statistics.yaml (logistic_decoder_rnn.py:154-179)
    m.action_decoder._setup_action_bounds("does/not/exist", [1.0] * 7, [-1.0] * 7, True)       # missing file: keeps the config's bounds
    assert float(m.action_decoder.action_max_bound.max()) == 1.0
    m.action_decoder._setup_action_bounds("dataset", [0.5] * 6 + [1.0], [-0.25] * 6 + [-1.0], False)
    assert float(m.action_decoder.action_max_bound.max()) == 0.5 and float(m.action_decoder.action_min_bound.min()) == -0.25
    assert m.action_decoder.action_max_bound.shape == (1, 1, 6, 10) and m.action_decoder.gripper_bounds.tolist() == [-1.0, 1.0]


def test_round_trip_and_pretrained_weight_slicing(tmp_path):
    torch.manual_seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        big = instantiate(_small_cfg("hulc2_b200", max_window_size=48))
        small = instantiate(_small_cfg("hulc2_b200", max_window_size=32))
    path = tmp_path / "pre.ckpt"
    save_checkpoint(big, path, epoch=2, global_step=99)
    again = Hulc2.load_from_checkpoint(str(path))
    assert all(torch.equal(a, b) for a, b in zip(big.state_dict().values(), again.state_dict().values()))
    assert again.plan_recognition.position_embeddings.weight.shape[0] == 48
    # a 48-step pretraining checkpoint warm-starts a 32-step model: position embeddings are cut (utils.py:38-40)
    with pytest.raises(RuntimeError):
        small.load_state_dict(torch.load(str(path), weights_only=False)["state_dict"])
    initialize_pretrained_weights(small, {"pretrain_chk": str(path)})
    assert torch.equal(small.plan_recognition.position_embeddings.weight, big.plan_recognition.position_embeddings.weight[:32])
    assert torch.equal(small.action_decoder.rnn.weight_hh_l1, big.action_decoder.rnn.weight_hh_l1)
    # pretrain_exclude_pr leaves the plan-recognition network at its own initialisation
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        other = instantiate(_small_cfg("hulc2_b200", max_window_size=32))
    before = other.plan_recognition.fc.weight.clone()
    initialize_pretrained_weights(other, {"pretrain_chk": str(path), "pretrain_exclude_pr": True})
    assert torch.equal(other.plan_recognition.fc.weight, before)
    assert torch.equal(other.plan_proposal.fc_model[0].weight, big.plan_proposal.fc_model[0].weight)


def test_optimizer_state_resumes_the_device_step_counter():
    """FusedAdam.load_state_dict must refresh the capturable device counter (bias correction after a resume)."""
    from hulc2_b200.optim import FusedAdam

    lin = torch.nn.Linear(4, 3)
    opt = FusedAdam(lin.parameters(), lr=1e-3)
    opt._arenas[0]["step"] = 5
    sd = opt.state_dict()
    opt2 = FusedAdam(torch.nn.Linear(4, 3).parameters(), lr=1e-3)
    opt2._step_dev = torch.zeros(1, dtype=torch.int64)
    opt2.load_state_dict(sd)
    assert opt2._arenas[0]["step"] == 5 and int(opt2._step_dev) == 5
    # parameters rebound away from the arena (model.to()/.half() after construction) are detected instead of silently ignored
    lin.weight.data = lin.weight.data.clone()
    with pytest.raises(RuntimeError):
        FusedAdam._check_bound(opt._arenas[0])
