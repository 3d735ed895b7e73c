"""Generates tests/golden/hulc2_alt_golden.npz from the UNMODIFIED reference: the alternate blocks reachable from the
same Hydra groups -- GRU / LSTM decoder cells (hulc2/models/decoders/utils/rnn.py:17-36), the
continuous latent plan (hulc2/utils/distributions.py:28-29,55-59) and the RGB-D static encoder
(hulc2/models/perceptual_encoders/concat_encoders.py:74-80).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_alt.py
Inputs / weights / noise are regenerated from the seeds below (hulc2_b200.synthetic); only outputs are stored.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from hulc2_b200.config import hulc2_config  # noqa: E402
from hulc2_b200.synthetic import synthetic_state_dict  # noqa: E402
from oracle.ref_import import make_reference_model  # noqa: E402
from oracle.ref_noise import supplied_categories, supplied_normals  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import ALT_CASES as CASES, ALT_VAL_TAGS, ALT_WEIGHT_SEED as WEIGHT_SEED, NON_LEARNED, alt_case_inputs, alt_val_inputs  # noqa: E402
from oracle.ref_noise import supplied_uniforms  # noqa: E402


def main():
    G = {}
    for tag in CASES:
        kw, batch, draw = alt_case_inputs(tag)
        m = make_reference_model(hulc2_config(pkg="hulc2", **kw))
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
        m.load_state_dict(synthetic_state_dict(shapes, seed=WEIGHT_SEED, skip=NON_LEARNED), strict=False)
        ctx = supplied_categories if kw["distribution"] == "discrete" else supplied_normals
        with ctx([draw[mod] for mod in batch]):
            loss = m.training_step(batch, 0)
        loss.backward()
        G[f"{tag}/loss"] = loss.detach().numpy()
        for k, v in m.logged.items():
            G[f"{tag}/log/{k}"] = v.detach().numpy()
        for n, p in m.named_parameters():
            if p.grad is not None:
                G[f"{tag}/grad_norm/{n}"] = p.grad.double().norm().numpy()
        for n in ("action_decoder.rnn.bias_hh_l0", "action_decoder.rnn.bias_ih_l1", "plan_recognition.fc_state.0.bias",
                  "plan_proposal.fc_state.0.bias", "perceptual_encoder.depth_static_encoder.conv_model.0.bias"):
            p = dict(m.named_parameters()).get(n)
            if p is not None and p.grad is not None:
                G[f"{tag}/grad/{n}"] = p.grad.numpy()
        print(tag, float(loss), sum(1 for k in G if k.startswith(f"{tag}/grad_norm/")))
    for tag in ALT_VAL_TAGS:                     # validation_step (hulc2.py:510-598) with the alternate blocks
        kw, batch, noise = alt_val_inputs(tag)
        m = make_reference_model(hulc2_config(pkg="hulc2", **kw))
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
        m.load_state_dict(synthetic_state_dict(shapes, seed=WEIGHT_SEED, skip=NON_LEARNED), strict=False)
        m.eval()

        class _DM:
            modalities = ["vis", "lang"]

        class _Tr:
            datamodule = _DM()

        m.trainer = _Tr()
        draws, unis = [], []
        for mod in batch:
            draws += [noise[mod]["plan_idx_pp"], noise[mod]["plan_idx_pr"]]
            unis += [noise[mod][k] for k in ("u1_pp", "u2_pp", "u1_pr", "u2_pr")]
        ctx = supplied_categories if kw["distribution"] == "discrete" else supplied_normals
        with torch.no_grad(), ctx(draws), supplied_uniforms(unis):
            out_d = m.validation_step(batch, 0)
        for k, v in m.logged.items():
            if k.startswith("val"):
                G[f"{tag}/val/log/{k}"] = v.detach().numpy()
        for k, v in out_d.items():
            G[f"{tag}/val/out/{k}"] = v.detach().numpy()
        print(tag, "val", sum(1 for k in G if k.startswith(f"{tag}/val/")))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hulc2_alt_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
