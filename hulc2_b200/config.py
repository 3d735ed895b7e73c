"""Plain-dict restatement of the reference's Hydra model composition.

Mirrors the YAML tree under ``conf/model/**`` of the reference (cited per key
below) so the model can be built without Hydra/OmegaConf installed.  With
``pkg="hulc2"`` the ``_target_`` paths are the reference's own (what
``conf/model/*.yaml`` holds); with ``pkg="hulc2_b200"`` they point at this
package's mirror classes -- the only override a user needs to switch
(``model._target_`` and the nested ``_target_``s), all other keys are unchanged.
"""
from __future__ import annotations

from typing import Optional


def vision_static_cfg(pkg: str, input_height: int = 200, input_width: int = 200, num_c: int = 3) -> dict:
    # conf/model/perceptual_encoder/rgb_static/default.yaml:1-10 (input_height 150 -> 200 for CALVIN)
    return {
        "_target_": f"{pkg}.models.perceptual_encoders.vision_network.VisionNetwork",
        "input_width": input_width,
        "input_height": input_height,
        "activation_function": "ReLU",
        "dropout_vis_fc": 0.0,
        "l2_normalize_output": False,
        "visual_features": 64,
        "num_c": num_c,
        "use_sinusoid": False,
        "spatial_softmax_temp": 1.0,
    }


def vision_gripper_cfg(pkg: str, num_c: int = 3) -> dict:
    # conf/model/perceptual_encoder/rgb_gripper/default.yaml:1-9
    return {
        "_target_": f"{pkg}.models.perceptual_encoders.vision_network_gripper.VisionNetwork",
        "input_width": 84,
        "input_height": 84,
        "activation_function": "ReLU",
        "dropout_vis_fc": 0.0,
        "l2_normalize_output": False,
        "visual_features": 64,
        "conv_encoder": "nature_cnn",
        "num_c": num_c,
    }


def hulc2_config(
    pkg: str = "hulc2_b200",
    variant: str = "calvin",
    static_hw=(200, 200),
    dropout_p: float = 0.1,
    depth_static: bool = False,
    max_window_size: int = 32,
    rnn_model: str = "rnn_decoder",
    distribution: str = "discrete",
    hidden_size: int = 2048,
) -> dict:
    """Model config for ``variant`` in {"calvin", "real_world"}.

    calvin     = conf/model/calvin_hulc++.yaml with perceptual_encoder/rgb_static=default
                 (conv encoder), language_encoder=none (precomputed [B,384]).
    real_world = conf/model/real_world_hulc++.yaml (no clip loss, decoder slice [0,128],
                 gripper_control false); static frame 150x200 by default via ``static_hw``.
    """
    assert variant in ("calvin", "real_world")
    calvin = variant == "calvin"
    h, w = static_hw
    if distribution == "discrete":
        dist = {  # conf/model/distribution/discrete.yaml
            "_target_": f"{pkg}.utils.distributions.Distribution",
            "dist": "discrete",
            "category_size": 32,
            "class_size": 32,
        }
    else:
        dist = {  # conf/model/distribution/continuous.yaml
            "_target_": f"{pkg}.utils.distributions.Distribution",
            "dist": "continuous",
            "plan_features": 256,
        }
    cfg = {
        "_target_": f"{pkg}.models.hulc2.Hulc2",
        "_recursive_": False,
        "perceptual_encoder": {  # conf/model/perceptual_encoder/gripper_cam.yaml (+static_RGBD.yaml)
            "_target_": f"{pkg}.models.perceptual_encoders.concat_encoders.ConcatEncoders",
            "_recursive_": False,
            "rgb_static": vision_static_cfg(pkg, h, w),
            "rgb_gripper": vision_gripper_cfg(pkg),
            "depth_static": vision_static_cfg(pkg, h, w, num_c=1) if depth_static else None,
            "depth_gripper": None,
            "proprio": None,
            "tactile": None,
        },
        "plan_proposal": {  # conf/model/plan_proposal/default.yaml
            "_target_": f"{pkg}.models.plan_encoders.plan_proposal_net.PlanProposalNetwork",
            "perceptual_features": None,
            "latent_goal_features": 32,
            "plan_features": None,
            "activation_function": "ReLU",
            "hidden_size": hidden_size,
        },
        "plan_recognition": {  # conf/model/plan_recognition/transformers.yaml
            "_target_": f"{pkg}.models.plan_encoders.plan_recognition_net.PlanRecognitionTransformersNetwork",
            "num_heads": 8,
            "num_layers": 2,
            "encoder_hidden_size": 2048,
            "fc_hidden_size": 4096,
            "in_features": None,
            "plan_features": None,
            "action_space": 7,
            "dropout_p": dropout_p,
            "encoder_normalize": False,
            "positional_normalize": False,
            "position_embedding": True,
            "max_position_embeddings": max_window_size,
        },
        "distribution": dist,
        "visual_goal": {  # conf/model/visual_goal/default.yaml
            "_target_": f"{pkg}.models.encoders.goal_encoders.VisualGoalEncoder",
            "in_features": None,
            "hidden_size": hidden_size,
            "latent_goal_features": 32,
            "l2_normalize_goal_embeddings": False,
            "activation_function": "ReLU",
        },
        "language_encoder": None,  # conf/model/language_encoder/none.yaml
        "language_goal": {  # conf/model/language_goal/default.yaml
            "_target_": f"{pkg}.models.encoders.goal_encoders.LanguageGoalEncoder",
            "in_features": 384,
            "hidden_size": hidden_size,
            "latent_goal_features": 32,
            "l2_normalize_goal_embeddings": False,
            "activation_function": "ReLU",
            "word_dropout_p": 0.0,
        },
        "action_decoder": {  # conf/model/action_decoder/logistic_decoder_rnn_{calvin,real_world}.yaml
            "_target_": f"{pkg}.models.decoders.logistic_decoder_rnn.LogisticDecoderRNN",
            "n_mixtures": 10,
            "hidden_size": hidden_size,
            "out_features": 7,
            "log_scale_min": -7.0,
            "act_max_bound": [1.0] * 7,
            "act_min_bound": [-1.0] * 7,
            "dataset_dir": "dataset",
            "load_action_bounds": False,
            "num_classes": 10,
            "latent_goal_features": 32,
            "plan_features": None,
            "perceptual_features": None,
            "gripper_alpha": 1.0,
            "perceptual_emb_slice": [64, 128] if calvin else [0, 128],
            "policy_rnn_dropout_p": 0.0,
            "num_layers": 2,
            "rnn_model": rnn_model,
            "gripper_control": calvin,
            "discrete_gripper": True,
        },
        "kl_beta": 0.01,  # conf/loss/default.yaml:1
        "kl_balancing_mix": 0.8,  # conf/loss/default.yaml:3
        "replan_freq": 30,
        "use_clip_auxiliary_loss": calvin,
        "clip_auxiliary_loss_beta": 3.0,  # conf/loss/default.yaml:6
        "optimizer": {"_target_": "torch.optim.Adam", "lr": 2e-4},  # conf/model/optimizer/adam.yaml
        "lr_scheduler": {"_target_": "transformers.get_constant_schedule"},  # conf/model/lr_scheduler/constant.yaml
        "proj_vis_lang": {  # conf/model/proj_vis_lang/default.yaml
            "_target_": f"{pkg}.models.auxiliary_loss_networks.proj_vis_lang.ProjVisLang",
            "im_dim": 4096,
            "lang_dim": 32,
            "output_dim": 32,
            "proj_lang": True,
        }
        if calvin
        else None,
    }
    return cfg


def strip_none(cfg: Optional[dict]) -> Optional[dict]:
    return cfg
