"""Hulc2 policy module (mirror of hulc2/models/hulc2.py:27-719).

Same constructor (DictConfigs instantiated inside ``__init__`` with ``_recursive_: false``), same
state_dict names, same public methods used by the reference's callers: ``training_step``,
``validation_step``, ``configure_optimizers``, ``set_kl_beta``, ``reset``/``step``,
``predict_with_plan``, ``get_pp_plan_lang``/``get_pp_plan_vision``, ``lmp_train``/``lmp_val``,
``compute_kl_loss``, ``clip_auxiliary_loss``.  All math runs in the CUDA library (``hulc2_b200.ops``);
the host syncs of the reference (``torch.any`` on the aux mask, boolean row selection, NaN asserts)
are replaced by static-shape masked kernels.
"""
import logging
from typing import Any, Dict, NamedTuple, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .. import noise, ops
from .._compat import DictConfig, LightningModule, as_config, instantiate, rank_zero_info, rank_zero_only
from ..utils.distributions import State
from .decoders.action_decoder import ActionDecoder

logger = logging.getLogger(__name__)


@rank_zero_only
def log_rank_0(*args, **kwargs):
    logger.info(*args, **kwargs)


class Hulc2(LightningModule):
    def __init__(
        self,
        perceptual_encoder: DictConfig,
        plan_proposal: DictConfig,
        plan_recognition: DictConfig,
        language_encoder: DictConfig,
        language_goal: DictConfig,
        visual_goal: DictConfig,
        action_decoder: DictConfig,
        kl_beta: float,
        kl_balancing_mix: float,
        optimizer: DictConfig,
        lr_scheduler: DictConfig,
        distribution: DictConfig,
        use_clip_auxiliary_loss: bool,
        clip_auxiliary_loss_beta: float,
        replan_freq: int = 30,
        proj_vis_lang: Optional[DictConfig] = None,
    ):
        super().__init__()
        perceptual_encoder, plan_proposal, plan_recognition = map(as_config, (perceptual_encoder, plan_proposal, plan_recognition))
        language_encoder, language_goal, visual_goal = map(as_config, (language_encoder, language_goal, visual_goal))
        action_decoder, optimizer, lr_scheduler = map(as_config, (action_decoder, optimizer, lr_scheduler))
        distribution, proj_vis_lang = as_config(distribution), as_config(proj_vis_lang)

        self.perceptual_encoder = instantiate(perceptual_encoder, device=self.device)
        self.setup_input_sizes(self.perceptual_encoder, plan_proposal, plan_recognition, visual_goal, action_decoder, distribution)
        # plan networks
        self.dist = instantiate(distribution)
        self.plan_proposal = instantiate(plan_proposal, dist=self.dist)
        self.plan_recognition = instantiate(plan_recognition, dist=self.dist)
        # goal encoders
        self.visual_goal = instantiate(visual_goal)
        self.lang_encoder = instantiate(language_encoder) if language_encoder else None
        self.language_goal = instantiate(language_goal, lang_net=self.lang_encoder) if language_goal else None
        # policy network
        self.action_decoder: ActionDecoder = instantiate(action_decoder)
        # auxiliary losses
        self.use_clip_auxiliary_loss = use_clip_auxiliary_loss
        self.clip_auxiliary_loss_beta = clip_auxiliary_loss_beta
        if use_clip_auxiliary_loss:
            self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
            self.proj_vis_lang = instantiate(proj_vis_lang)

        self.kl_beta = kl_beta
        self.kl_balancing_mix = kl_balancing_mix
        self.modality_scope = "vis"
        self.optimizer_config = optimizer
        self.lr_scheduler = lr_scheduler
        self.save_hyperparameters()

        # for inference
        self.rollout_step_counter = 0
        self.replan_freq = replan_freq
        self.latent_goal = None
        self.plan = None

    # ------------------------------------------------------------------ checkpoints
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, hparams_file=None, strict: bool = True, **kwargs):
        """``LightningModule.load_from_checkpoint`` for Lightning-format checkpoints written by the REFERENCE model or by this
        mirror (``{"state_dict", "hyper_parameters", ...}``): the saved constructor arguments are re-used (keyword arguments
        override them, e.g. ``perceptual_encoder=...`` in evaluation/manager_lmp.py:91-103), ``_target_`` strings of the
        reference's model modules are pointed at this package's mirror classes, and the ``state_dict`` -- whose names and
        shapes are the reference's -- is loaded into them.  Weights of a frozen language encoder stored in the checkpoint
        (``lang_encoder.*``, absent here with ``language_encoder=none``) are skipped."""
        from ..utils.checkpoint import read_checkpoint, retarget

        ckpt = read_checkpoint(checkpoint_path, map_location)
        hp = dict(ckpt.get("hyper_parameters") or {})
        hp.update(kwargs)
        hp = {k: retarget(v) for k, v in hp.items() if k not in ("_target_", "_recursive_")}
        model = cls(**hp)
        sd = {k: v for k, v in ckpt["state_dict"].items() if model.lang_encoder is not None or not k.startswith("lang_encoder.")}
        model.load_state_dict(sd, strict=strict)
        for k in ("epoch", "global_step"):
            if k in ckpt:
                try:
                    setattr(model, "current_epoch" if k == "epoch" else k, int(ckpt[k]))
                except AttributeError:           # read-only properties under real Lightning
                    pass
        return model

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        ops.invalidate_weight_mirrors()          # bf16 operand mirrors of the old weights are stale
        return out

    # ------------------------------------------------------------------ setup
    @staticmethod
    def setup_input_sizes(perceptual_encoder, plan_proposal, plan_recognition, visual_goal, action_decoder, distribution):
        """hulc2.py:126-158: fill the ``???`` feature sizes from the encoder / distribution."""
        plan_proposal.perceptual_features = perceptual_encoder.latent_size
        plan_recognition.in_features = perceptual_encoder.latent_size
        visual_goal.in_features = perceptual_encoder.latent_size
        action_decoder.perceptual_features = perceptual_encoder.latent_size
        if distribution.dist == "discrete":
            n = distribution.class_size * distribution.category_size
        else:
            n = distribution.plan_features
        plan_proposal.plan_features = n
        plan_recognition.plan_features = n
        action_decoder.plan_features = n

    @property
    def num_training_steps(self) -> int:
        return int(self.trainer.estimated_stepping_batches)  # type: ignore

    def compute_warmup(self, num_training_steps: int, num_warmup_steps: Union[int, float]) -> Tuple[int, int]:
        if num_training_steps < 0:
            num_training_steps = self.num_training_steps
        if isinstance(num_warmup_steps, float):
            num_warmup_steps *= num_training_steps
        return num_training_steps, int(num_warmup_steps)

    def configure_optimizers(self):
        """hulc2.py:185-198.  ``torch.optim.Adam`` in the config is served by the library's fused Adam
        (same update rule, one kernel over a flat parameter arena); other optimizers are instantiated as given."""
        from ..optim import FusedAdam

        cfg = dict(self.optimizer_config)
        if cfg.get("_target_") == "torch.optim.Adam":
            cfg.pop("_target_")
            optimizer = FusedAdam(self.parameters(), **cfg)
        else:
            optimizer = instantiate(self.optimizer_config, params=self.parameters())
        if self.lr_scheduler and "num_warmup_steps" in self.lr_scheduler:
            self.lr_scheduler.num_training_steps, self.lr_scheduler.num_warmup_steps = self.compute_warmup(
                num_training_steps=self.lr_scheduler.num_training_steps, num_warmup_steps=self.lr_scheduler.num_warmup_steps
            )
            rank_zero_info(f"Inferring number of training steps, set to {self.lr_scheduler.num_training_steps}")
        if not self.lr_scheduler:
            return {"optimizer": optimizer}
        scheduler = instantiate(self.lr_scheduler, optimizer)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "step", "frequency": 1}}

    # ------------------------------------------------------------------ losses
    def compute_kl_loss(self, pp_state: State, pr_state: State) -> torch.Tensor:
        """hulc2.py:444-466 (KL balancing, alpha -> prior, 1-alpha -> posterior), scaled by kl_beta."""
        if self.dist.dist == "continuous":
            return ops.GaussKLFunction.apply(pp_state.mean, pp_state.std, pr_state.mean, pr_state.std,
                                             float(self.kl_balancing_mix), float(self.kl_beta))
        return ops.KLFunction.apply(pp_state.logit, pr_state.logit, self.dist.category_size, self.dist.class_size,
                                    float(self.kl_balancing_mix), float(self.kl_beta))

    def set_kl_beta(self, kl_beta):
        self.kl_beta = kl_beta

    def clip_auxiliary_loss(self, seq_vis_feat, encoded_lang, use_for_aux_loss):
        """hulc2.py:472-508; rows are selected by mask inside the kernel (returns 0 when none is selected)."""
        image_features, lang_features = self.proj_vis_lang(seq_vis_feat, encoded_lang)
        use = None
        if use_for_aux_loss is not None:
            use = use_for_aux_loss.contiguous().view(torch.uint8) if use_for_aux_loss.dtype == torch.bool else use_for_aux_loss.to(torch.uint8)
        return ops.InfoNCEFunction.apply(image_features, lang_features, self.logit_scale, use)

    def lmp_train(self, perceptual_emb, latent_goal, train_acts, robot_obs):
        """hulc2.py:200-245."""
        pp_state = self.plan_proposal(perceptual_emb[:, 0], latent_goal)
        pp_dist = self.dist.get_dist(pp_state)
        pr_state, seq_feat = self.plan_recognition(perceptual_emb)
        pr_dist = self.dist.get_dist(pr_state)
        sampled_plan = pr_dist.rsample()
        if self.dist.dist == "discrete":
            sampled_plan = torch.flatten(sampled_plan, start_dim=-2, end_dim=-1)
        action_loss = self.action_decoder.loss(sampled_plan, perceptual_emb, latent_goal, train_acts, robot_obs)
        kl_loss = self.compute_kl_loss(pp_state, pr_state)
        total_loss = ops.weighted_sum((1.0, 1.0), (action_loss, kl_loss))
        return kl_loss, action_loss, total_loss, pp_dist, pr_dist, seq_feat

    def lmp_val(self, perceptual_emb, latent_goal, actions, robot_obs):
        """hulc2.py:247-334."""
        pp_state = self.plan_proposal(perceptual_emb[:, 0], latent_goal)
        pp_dist = self.dist.get_dist(pp_state)
        sampled_plan_pp = self.dist.sample_latent_plan(pp_dist)
        action_loss_pp, sample_act_pp = self.action_decoder.loss_and_act(sampled_plan_pp, perceptual_emb, latent_goal, actions, robot_obs)
        mae_pp, gripper_sr_pp = _val_metrics(sample_act_pp, actions)
        pr_state, seq_feat = self.plan_recognition(perceptual_emb)
        pr_dist = self.dist.get_dist(pr_state)
        sampled_plan_pr = self.dist.sample_latent_plan(pr_dist)
        action_loss_pr, sample_act_pr = self.action_decoder.loss_and_act(sampled_plan_pr, perceptual_emb, latent_goal, actions, robot_obs)
        mae_pr, gripper_sr_pr = _val_metrics(sample_act_pr, actions)
        kl_loss = self.compute_kl_loss(pp_state, pr_state)
        return (sampled_plan_pp, action_loss_pp, sampled_plan_pr, action_loss_pr, kl_loss, mae_pp, mae_pr,
                gripper_sr_pp, gripper_sr_pr, seq_feat)

    # ------------------------------------------------------------------ train / val steps
    def training_step(self, batch: Dict[str, Dict], batch_idx: int) -> torch.Tensor:  # type: ignore
        """hulc2.py:336-442.  batch = {"vis": {...}, "lang": {...}} as documented there.

        When the modalities carry the same cameras (every shipped config), their windows run through the network as ONE
        batch (``_training_step_batched``): same per-modality losses, logged scalars and gradients as the reference's
        modality loop below, half the kernel launches and M=128 instead of M=64 rows for every contraction."""
        ops.invalidate_weight_mirrors()
        if self.batch_modalities and self._can_batch_modalities(batch):
            return self._training_step_batched(batch)
        n_mod = len(batch)
        terms, weights = [], []
        kls, acts = [], []
        clip = None
        batch_size: Dict[str, Any] = {}
        total_bs = 0
        for self.modality_scope, dataset_batch in batch.items():
            perceptual_emb = self.perceptual_encoder(dataset_batch["rgb_obs"], dataset_batch["depth_obs"], dataset_batch["robot_obs"])
            if "lang" in self.modality_scope:
                latent_goal = self.language_goal(dataset_batch["lang"])
            else:
                latent_goal = self.visual_goal(perceptual_emb[:, -1])
            kl, act_loss, mod_loss, pp_dist, pr_dist, seq_feat = self.lmp_train(
                perceptual_emb, latent_goal, dataset_batch["actions"], dataset_batch["state_info"]["robot_obs"]
            )
            if "lang" in self.modality_scope:
                use = dataset_batch.get("use_for_aux_lang_loss")
                batch_size["aux_lang"] = use.shape[0] if use is not None else 1
                if self.use_clip_auxiliary_loss:
                    c = self.clip_auxiliary_loss(seq_feat, latent_goal, use)
                    clip = c if clip is None else ops.weighted_sum((1.0, 1.0), (clip, c))
            terms += [act_loss, kl]
            weights += [1.0 / n_mod, 1.0 / n_mod]
            kls.append(kl)
            acts.append(act_loss)
            bs = dataset_batch["actions"].shape[0]
            batch_size[self.modality_scope] = bs
            total_bs += bs
            self.log(f"train/kl_loss_scaled_{self.modality_scope}", kl.detach(), on_step=False, on_epoch=True, batch_size=bs)
            self.log(f"train/action_loss_{self.modality_scope}", act_loss.detach(), on_step=False, on_epoch=True, batch_size=bs)
            self.log(f"train/total_loss_{self.modality_scope}", mod_loss.detach(), on_step=False, on_epoch=True, batch_size=bs)
        if self.use_clip_auxiliary_loss and clip is not None:
            terms.append(clip)
            weights.append(float(self.clip_auxiliary_loss_beta))
        total_loss = ops.weighted_sum(weights, terms)
        with torch.no_grad():
            if self.use_clip_auxiliary_loss and clip is not None:
                self.log("train/lang_clip_loss", ops.weighted_sum((float(self.clip_auxiliary_loss_beta),), (clip.detach(),)),
                         on_step=False, on_epoch=True, batch_size=batch_size.get("aux_lang", 1), sync_dist=True)
            w = (1.0 / n_mod,) * n_mod
            self.log("train/kl_loss", ops.weighted_sum(w, [k.detach() for k in kls]), on_step=False, on_epoch=True, batch_size=total_bs)
            self.log("train/action_loss", ops.weighted_sum(w, [a.detach() for a in acts]), on_step=False, on_epoch=True, batch_size=total_bs)
        self.log("train/total_loss", total_loss.detach(), on_step=False, on_epoch=True, batch_size=total_bs)
        return total_loss

    # ------------------------------------------------------------------ all modalities in one pass
    batch_modalities = True

    def _can_batch_modalities(self, batch) -> bool:
        if len(batch) < 2 or not hasattr(self.action_decoder, "loss_modalities"):
            return False
        mods = list(batch.values())

        def cams(d):
            return sorted((k, tuple(v.shape[1:])) for k, v in (d or {}).items() if v is not None)

        return all(cams(m["rgb_obs"]) == cams(mods[0]["rgb_obs"]) and cams(m["depth_obs"]) == cams(mods[0]["depth_obs"])
                   for m in mods[1:])

    def _training_step_batched(self, batch: Dict[str, Dict]) -> torch.Tensor:
        """Same arithmetic as the loop in ``training_step`` (hulc2.py:336-442 with lmp_train :200-245 inlined): windows
        are independent, so encoders, plan networks and the decoder see the concatenated batch; every loss is still a
        mean over its own modality's windows (segment reductions), and the InfoNCE term only sees the language rows."""
        names = list(batch.keys())
        mods = [batch[n] for n in names]
        n_mod = len(mods)
        sizes = [m["actions"].shape[0] for m in mods]
        offs = [sum(sizes[:i]) for i in range(n_mod)]
        noise.fuse_supplied(n_mod)
        emb = self.perceptual_encoder.forward_modalities([m["rgb_obs"] for m in mods], [m["depth_obs"] for m in mods], None)
        goals = []
        for name, m, o, n in zip(names, mods, offs, sizes):
            self.modality_scope = name
            goals.append(self.language_goal(m["lang"]) if "lang" in name else self.visual_goal(emb[o : o + n, -1]))
        latent_goal = ops.concat_rows(goals)
        pp_state = self.plan_proposal(emb[:, 0], latent_goal)
        pr_state, seq_feat = self.plan_recognition(emb)
        sampled_plan = self.dist.get_dist(pr_state).rsample()
        if self.dist.dist == "discrete":
            sampled_plan = torch.flatten(sampled_plan, start_dim=-2, end_dim=-1)
        act_losses = self.action_decoder.loss_modalities(
            sampled_plan, emb, latent_goal, [m["actions"] for m in mods], [m["state_info"]["robot_obs"] for m in mods])
        if self.dist.dist == "discrete":
            kl_losses = ops.KLFunction.apply(pp_state.logit, pr_state.logit, self.dist.category_size, self.dist.class_size,
                                             float(self.kl_balancing_mix), float(self.kl_beta), tuple(sizes))
        else:
            kl_losses = ops.GaussKLFunction.apply(pp_state.mean, pp_state.std, pr_state.mean, pr_state.std,
                                                  float(self.kl_balancing_mix), float(self.kl_beta), tuple(sizes))
        clip = None
        batch_size: Dict[str, Any] = {}
        terms, weights, kls, acts = [], [], [], []
        for i, (name, m, o, n) in enumerate(zip(names, mods, offs, sizes)):
            self.modality_scope = name
            kl, act_loss = kl_losses[i], act_losses[i]
            if "lang" in name:
                use = m.get("use_for_aux_lang_loss")
                batch_size["aux_lang"] = use.shape[0] if use is not None else 1
                if self.use_clip_auxiliary_loss:
                    c = self.clip_auxiliary_loss(seq_feat[o : o + n], goals[i], use)
                    clip = c if clip is None else ops.weighted_sum((1.0, 1.0), (clip, c))
            terms += [act_loss, kl]
            weights += [1.0 / n_mod, 1.0 / n_mod]
            kls.append(kl)
            acts.append(act_loss)
            batch_size[name] = n
            self.log(f"train/kl_loss_scaled_{name}", kl.detach(), on_step=False, on_epoch=True, batch_size=n)
            self.log(f"train/action_loss_{name}", act_loss.detach(), on_step=False, on_epoch=True, batch_size=n)
            self.log(f"train/total_loss_{name}", ops.weighted_sum((1.0, 1.0), (act_loss.detach(), kl.detach())),
                     on_step=False, on_epoch=True, batch_size=n)
        total_bs = sum(sizes)
        if self.use_clip_auxiliary_loss and clip is not None:
            terms.append(clip)
            weights.append(float(self.clip_auxiliary_loss_beta))
        total_loss = ops.weighted_sum(weights, terms)
        with torch.no_grad():
            if self.use_clip_auxiliary_loss and clip is not None:
                self.log("train/lang_clip_loss", ops.weighted_sum((float(self.clip_auxiliary_loss_beta),), (clip.detach(),)),
                         on_step=False, on_epoch=True, batch_size=batch_size.get("aux_lang", 1), sync_dist=True)
            w = (1.0 / n_mod,) * n_mod
            self.log("train/kl_loss", ops.weighted_sum(w, [k.detach() for k in kls]), on_step=False, on_epoch=True, batch_size=total_bs)
            self.log("train/action_loss", ops.weighted_sum(w, [a.detach() for a in acts]), on_step=False, on_epoch=True, batch_size=total_bs)
        self.log("train/total_loss", total_loss.detach(), on_step=False, on_epoch=True, batch_size=total_bs)
        return total_loss

    def _validation_step_batched(self, batch: Dict[str, Dict]) -> Dict[str, torch.Tensor]:
        """hulc2.py:510-598 with lmp_val (:247-334) inlined, all modalities as ONE batch (as ``_training_step_batched``): one
        encoder pass, one plan-proposal / plan-recognition pass, two decoder passes (proposal plan, recognition plan), segment
        losses, and the MAE / gripper-success reductions of every modality in one kernel each.  Same logged names and values
        as the modality loop in ``validation_step``; no host synchronisation, so the whole step can be captured in a graph."""
        names = list(batch.keys())
        mods = [batch[n] for n in names]
        n_mod = len(mods)
        n_log = len(getattr(getattr(self.trainer, "datamodule", None), "modalities", None) or batch)
        sizes = [m["actions"].shape[0] for m in mods]
        offs = [sum(sizes[:i]) for i in range(n_mod)]
        noise.fuse_supplied(n_mod)
        emb = self.perceptual_encoder.forward_modalities([m["rgb_obs"] for m in mods], [m["depth_obs"] for m in mods], None)
        goals = []
        for name, m, o, n in zip(names, mods, offs, sizes):
            self.modality_scope = name
            goals.append(self.language_goal(m["lang"]) if "lang" in name else self.visual_goal(emb[o : o + n, -1]))
        latent_goal = ops.concat_rows(goals)
        actions = [m["actions"] for m in mods]
        robot = [m["state_info"]["robot_obs"] for m in mods]
        # draw order of lmp_val: proposal plan, its action sample, recognition plan, its action sample
        pp_state = self.plan_proposal(emb[:, 0], latent_goal)
        plan_pp = self.dist.sample_latent_plan(self.dist.get_dist(pp_state))
        loss_pp, acts_pp = self.action_decoder.loss_and_act_modalities(plan_pp, emb, latent_goal, actions, robot)
        pr_state, seq_feat = self.plan_recognition(emb)
        plan_pr = self.dist.sample_latent_plan(self.dist.get_dist(pr_state))
        loss_pr, acts_pr = self.action_decoder.loss_and_act_modalities(plan_pr, emb, latent_goal, actions, robot)
        if self.dist.dist == "discrete":
            kl = ops.KLFunction.apply(pp_state.logit, pr_state.logit, self.dist.category_size, self.dist.class_size,
                                      float(self.kl_balancing_mix), float(self.kl_beta), tuple(sizes))
        else:
            kl = ops.GaussKLFunction.apply(pp_state.mean, pp_state.std, pr_state.mean, pr_state.std,
                                           float(self.kl_balancing_mix), float(self.kl_beta), tuple(sizes))
        output, act_pp = {}, []
        for i, (m, d, o, n) in enumerate(zip(names, mods, offs, sizes)):
            self.modality_scope = m
            met_pp, met_pr = ops.val_metrics(acts_pp[i], d["actions"]), ops.val_metrics(acts_pr[i], d["actions"])
            if "lang" in m and self.use_clip_auxiliary_loss:
                self.log("val/val_pred_clip_loss", self.clip_auxiliary_loss(seq_feat[o : o + n], goals[i], d.get("use_for_aux_lang_loss")), sync_dist=True)
            act_pp.append(loss_pp[i])
            self.log(f"val_total_mae/{m}_total_mae_pr", met_pr[0], sync_dist=True)
            self.log(f"val_total_mae/{m}_total_mae_pp", met_pp[0], sync_dist=True)
            self.log(f"val_pos_mae/{m}_pos_mae_pr", met_pr[1], sync_dist=True)
            self.log(f"val_pos_mae/{m}_pos_mae_pp", met_pp[1], sync_dist=True)
            self.log(f"val_orn_mae/{m}_orn_mae_pr", met_pr[2], sync_dist=True)
            self.log(f"val_orn_mae/{m}_orn_mae_pp", met_pp[2], sync_dist=True)
            self.log(f"val_kl/{m}_kl_loss", kl[i], sync_dist=True)
            self.log(f"val_act/{m}_act_loss_pp", loss_pp[i], sync_dist=True)
            self.log(f"val_act/{m}_act_loss_pr", loss_pr[i], sync_dist=True)
            self.log(f"val_grip/{m}_grip_sr_pr", met_pr[3], sync_dist=True)
            self.log(f"val_grip/{m}_grip_sr_pp", met_pp[3], sync_dist=True)
            self.log("val_act/action_loss_pp", ops.weighted_sum((1.0 / n_log,) * len(act_pp), [a.detach() for a in act_pp]), sync_dist=True)
            output[f"sampled_plan_pp_{m}"] = plan_pp[o : o + n]
            output[f"sampled_plan_pr_{m}"] = plan_pr[o : o + n]
            output[f"idx_{m}"] = d["idx"]
        return output

    def validation_step(self, batch: Dict[str, Dict], batch_idx: int) -> Dict[str, torch.Tensor]:  # type: ignore
        """hulc2.py:510-598.  Modalities with the same cameras run as ONE batch (``_validation_step_batched``)."""
        ops.invalidate_weight_mirrors()
        if self.batch_modalities and self._can_batch_modalities(batch) and hasattr(self.action_decoder, "loss_and_act_modalities"):
            return self._validation_step_batched(batch)
        output = {}
        act_pp = []
        n_mod = len(getattr(getattr(self.trainer, "datamodule", None), "modalities", None) or batch)
        for self.modality_scope, dataset_batch in batch.items():
            perceptual_emb = self.perceptual_encoder(dataset_batch["rgb_obs"], dataset_batch["depth_obs"], dataset_batch["robot_obs"])
            if "lang" in self.modality_scope:
                latent_goal = self.language_goal(dataset_batch["lang"])
            else:
                latent_goal = self.visual_goal(perceptual_emb[:, -1])
            (sampled_plan_pp, action_loss_pp, sampled_plan_pr, action_loss_pr, kl_loss, mae_pp, mae_pr,
             gripper_sr_pp, gripper_sr_pr, seq_feat) = self.lmp_val(
                perceptual_emb, latent_goal, dataset_batch["actions"], dataset_batch["state_info"]["robot_obs"]
            )
            m = self.modality_scope
            if "lang" in m and self.use_clip_auxiliary_loss:
                self.log("val/val_pred_clip_loss", self.clip_auxiliary_loss(seq_feat, latent_goal, dataset_batch.get("use_for_aux_lang_loss")), sync_dist=True)
            act_pp.append(action_loss_pp)
            self.log(f"val_total_mae/{m}_total_mae_pr", mae_pr.mean(), sync_dist=True)
            self.log(f"val_total_mae/{m}_total_mae_pp", mae_pp.mean(), sync_dist=True)
            self.log(f"val_pos_mae/{m}_pos_mae_pr", mae_pr[..., :3].mean(), sync_dist=True)
            self.log(f"val_pos_mae/{m}_pos_mae_pp", mae_pp[..., :3].mean(), sync_dist=True)
            self.log(f"val_orn_mae/{m}_orn_mae_pr", mae_pr[..., 3:6].mean(), sync_dist=True)
            self.log(f"val_orn_mae/{m}_orn_mae_pp", mae_pp[..., 3:6].mean(), sync_dist=True)
            self.log(f"val_kl/{m}_kl_loss", kl_loss, sync_dist=True)
            self.log(f"val_act/{m}_act_loss_pp", action_loss_pp, sync_dist=True)
            self.log(f"val_act/{m}_act_loss_pr", action_loss_pr, sync_dist=True)
            self.log(f"val_grip/{m}_grip_sr_pr", gripper_sr_pr, sync_dist=True)
            self.log(f"val_grip/{m}_grip_sr_pp", gripper_sr_pp, sync_dist=True)
            self.log("val_act/action_loss_pp", ops.weighted_sum((1.0 / n_mod,) * len(act_pp), [a.detach() for a in act_pp]), sync_dist=True)
            output[f"sampled_plan_pp_{m}"] = sampled_plan_pp
            output[f"sampled_plan_pr_{m}"] = sampled_plan_pr
            output[f"idx_{m}"] = dataset_batch["idx"]
        return output

    # ------------------------------------------------------------------ inference (hulc2.py:600-707)
    def reset(self):
        ops.invalidate_weight_mirrors()
        self.plan = None
        self.latent_goal = None
        self.rollout_step_counter = 0

    def step(self, obs, goal):
        if self.rollout_step_counter % self.replan_freq == 0:
            ops.invalidate_weight_mirrors()   # parameters may have been trained / loaded since the last plan
            if "lang" in goal:
                self.plan, self.latent_goal = self.get_pp_plan_lang(obs, goal)
            else:
                self.plan, self.latent_goal = self.get_pp_plan_vision(obs, goal)
        action = self.predict_with_plan(obs, self.latent_goal, self.plan)
        self.rollout_step_counter += 1
        return action

    def predict_with_plan(self, obs: Dict[str, Any], latent_goal: torch.Tensor, sampled_plan: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            perceptual_emb = self.perceptual_encoder(obs["rgb_obs"], obs["depth_obs"], obs["robot_obs"])
            action = self.action_decoder.act(sampled_plan, perceptual_emb, latent_goal, obs["robot_obs_raw"])
        return action

    def get_pp_plan_vision(self, obs: dict, goal: dict) -> Tuple[torch.Tensor, torch.Tensor]:
        assert len(obs["rgb_obs"]) == len(goal["rgb_obs"])
        imgs = {k: torch.cat([v, goal["rgb_obs"][k]], dim=1) for k, v in obs["rgb_obs"].items()}
        state = None
        depth_imgs: Dict[str, torch.Tensor] = {}
        for k in obs.keys():
            if "depth" in k:
                depth_imgs = {k2: torch.cat([v, goal["depth_obs"][k2]], dim=1) for k2, v in obs["depth_obs"].items()}
            if "robot_obs" in k and k in goal:
                state = torch.cat([obs["robot_obs"], goal["robot_obs"]], dim=1)
        with torch.no_grad():
            perceptual_emb = self.perceptual_encoder(imgs, depth_imgs, state)
            latent_goal = self.visual_goal(perceptual_emb[:, -1])
            pp_state = self.plan_proposal(perceptual_emb[:, 0], latent_goal)
            sampled_plan = self.dist.sample_latent_plan(self.dist.get_dist(pp_state))
        self.action_decoder.clear_hidden_state()
        return sampled_plan, latent_goal

    def get_pp_plan_lang(self, obs: dict, goal: dict) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.no_grad():
            perceptual_emb = self.perceptual_encoder(obs["rgb_obs"], obs["depth_obs"], obs["robot_obs"])
            latent_goal = self.language_goal(goal["lang"])
            pp_state = self.plan_proposal(perceptual_emb[:, 0], latent_goal)
            sampled_plan = self.dist.sample_latent_plan(self.dist.get_dist(pp_state))
        self.action_decoder.clear_hidden_state()
        return sampled_plan, latent_goal

    @rank_zero_only
    def on_train_epoch_start(self) -> None:
        logger.info(f"Start training epoch {self.current_epoch}")

    @rank_zero_only
    def on_train_epoch_end(self, unused: Optional[Any] = None) -> None:  # type: ignore
        logger.info(f"Finished training epoch {self.current_epoch}")

    @rank_zero_only
    def on_validation_epoch_end(self) -> None:
        logger.info(f"Finished validation epoch {self.current_epoch}")


def _val_metrics(sample_act: torch.Tensor, actions: torch.Tensor):
    """hulc2.py:292-302: per-dim MAE over the window [B,6] and discrete gripper success rate (validation
    metrics only -- host-side glue on device tensors, not part of the train/inference hot path)."""
    mae = torch.mean(torch.abs(sample_act[..., :-1] - actions[..., :-1]), 1)
    g = torch.where(sample_act[..., -1] > 0, 1.0, -1.0)
    return mae, torch.mean((actions[..., -1] == g).float())
