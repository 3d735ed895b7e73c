"""CPU: the C-ABI library loads and exports every symbol include/hulc2_b200.h declares (no compute calls
without a GPU); host-side logic (config surface, state_dict contract, synthetic data, noise plumbing)."""
import ctypes
import os
import re

import pytest
import torch

import hulc2_b200
from hulc2_b200 import _lib
from hulc2_b200.config import hulc2_config
from hulc2_b200.synthetic import synthetic_batch, synthetic_obs

from helpers import build_model, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "hulc2_b200.h")).read()
    return sorted(set(re.findall(r"\b(hulc2_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in include/hulc2_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), set(declared) ^ set(_lib.EXPORTED_SYMBOLS)
    bound = _lib.load_library(require_cuda=False)
    assert bound.hulc2_version() == 100
    assert bound.hulc2_last_error() is not None


def test_struct_layouts_match_header():
    # field order of the ctypes mirrors == declaration order in the header
    hdr = open(os.path.join(ROOT, "include", "hulc2_b200.h")).read()
    for cname, cls in (("hulc2_gemm_args", _lib.GemmArgs), ("hulc2_conv_args", _lib.ConvArgs)):
        body = hdr[: hdr.index("} " + cname)]
        body = body[body.rindex("typedef struct {") :]
        names = re.findall(r"[\s\*]([A-Za-z_][A-Za-z0-9_]*)\s*[;,]", body)
        assert names == [f[0] for f in cls._fields_], (cname, names)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from hulc2_b200 import ops

    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(2, 3), torch.zeros(4, 3), torch.zeros(4))
    m = build_model("calvin", hidden_size=64)
    batch = synthetic_batch(1, S=2)
    with pytest.raises(RuntimeError):
        m.training_step(batch, 0)


def test_state_dict_contract():
    m = build_model("calvin")
    sd = m.state_dict()
    assert len(sd) == 116
    assert sum(p.numel() for p in m.parameters()) == 47053815
    assert sd["plan_recognition.transformer_encoder.layers.0.self_attn.in_proj_weight"].shape == (384, 128)
    assert sd["action_decoder.rnn.weight_ih_l0"].shape == (2048, 1120)
    assert sd["perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"].shape == (128, 3136)
    assert sd["action_decoder.action_max_bound"].shape == (1, 1, 6, 10)
    names = {k[len("calvin_B2/gnorm/") :] for k in golden().files if k.startswith("calvin_B2/gnorm/")}
    mine = {n for n, _ in m.named_parameters()}
    assert names <= mine and mine - names == {"plan_recognition.layernorm.weight", "plan_recognition.layernorm.bias"}
    rw = build_model("real_world", (150, 200))
    assert sum(p.numel() for p in rw.parameters()) == 46647990
    assert rw.action_decoder.rnn.weight_ih_l0.shape == (2048, 1184)
    assert not hasattr(rw, "proj_vis_lang") or rw.use_clip_auxiliary_loss is False


def test_same_seed_same_init_as_torch_modules():
    # parameter containers are the torch modules of the reference, created in the reference's order
    torch.manual_seed(0)
    from hulc2_b200._compat import instantiate

    a = instantiate(hulc2_config(hidden_size=64))
    torch.manual_seed(0)
    b = instantiate(hulc2_config(hidden_size=64))
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)


def test_install_as_hulc2_aliases_reference_targets():
    import importlib
    import sys

    had = {k: v for k, v in sys.modules.items() if k == "hulc2" or k.startswith("hulc2.")}
    for k in had:
        del sys.modules[k]
    try:
        hulc2_b200.install_as_hulc2()
        mod = importlib.import_module("hulc2.models.hulc2")
        assert mod.Hulc2 is importlib.import_module("hulc2_b200.models.hulc2").Hulc2
        from hulc2_b200._compat import instantiate

        m = instantiate(hulc2_config(pkg="hulc2", hidden_size=64))   # reference _target_ strings, mirror classes
        assert type(m).__module__ == "hulc2_b200.models.hulc2"
    finally:
        for k in [k for k in sys.modules if k == "hulc2" or k.startswith("hulc2.")]:
            del sys.modules[k]
        sys.modules.update(had)


def test_synthetic_batch_contract_and_determinism():
    b1, b2 = synthetic_batch(2, seed=1), synthetic_batch(2, seed=1)
    assert set(b1) == {"vis", "lang"}
    v = b1["vis"]
    assert v["rgb_obs"]["rgb_static"].shape == (2, 32, 3, 200, 200) and v["rgb_obs"]["rgb_gripper"].shape == (2, 32, 3, 84, 84)
    assert v["actions"].shape == (2, 32, 7) and v["state_info"]["robot_obs"].shape == (2, 32, 15)
    assert set(v["actions"][..., 6].unique().tolist()) <= {-1.0, 1.0}
    assert b1["lang"]["lang"].shape == (2, 384) and b1["lang"]["use_for_aux_lang_loss"].dtype == torch.bool
    assert torch.equal(b1["vis"]["rgb_obs"]["rgb_static"], b2["vis"]["rgb_obs"]["rgb_static"])
    obs, goal = synthetic_obs(3)
    assert obs["rgb_obs"]["rgb_static"].shape == (3, 1, 3, 200, 200) and goal["lang"].shape == (3, 384)


def test_unsupported_config_values_raise_not_fallback():
    from hulc2_b200._compat import instantiate

    cfg = hulc2_config(hidden_size=64)
    cfg["action_decoder"]["discrete_gripper"] = False
    with pytest.raises(NotImplementedError):
        instantiate(cfg)
    cfg = hulc2_config(hidden_size=64)
    cfg["visual_goal"]["activation_function"] = "ELU"
    with pytest.raises(NotImplementedError):
        instantiate(cfg)
