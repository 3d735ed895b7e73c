"""LogisticDecoderRNN (mirror of hulc2/models/decoders/logistic_decoder_rnn.py:27-284).

Same constructor kwargs, buffers and public methods (``forward``, ``loss``, ``loss_and_act``, ``act``,
``_loss``, ``_sample``, ``_logistic_loss``, ``clear_hidden_state``, ``_setup_action_bounds``).  Internally
the RNN runs time-major and the four heads land in one fused ``[rows, 184]`` buffer that the logistic
loss / sampling kernels read once; ``forward`` unpacks it into the reference's ``[B,S,A,M]`` tensors.
"""
import logging
from pathlib import Path
from typing import Sequence, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from ... import noise, ops
from ..._compat import ListConfig, OmegaConf
from .action_decoder import ActionDecoder
from .utils.gripper_control import tcp_to_world_frame, world_to_tcp_frame
from .utils.rnn import RNN_MODELS

logger = logging.getLogger(__name__)


class LogisticDecoderRNN(ActionDecoder):
    def __init__(
        self,
        perceptual_features: int,
        latent_goal_features: int,
        plan_features: int,
        n_mixtures: int,
        hidden_size: int,
        out_features: int,
        log_scale_min: float,
        act_max_bound: Union[List[float], ListConfig],
        act_min_bound: Union[List[float], ListConfig],
        dataset_dir: str,
        load_action_bounds: bool,
        num_classes: int,
        gripper_alpha: float,
        perceptual_emb_slice: tuple,
        policy_rnn_dropout_p: float,
        num_layers: int,
        rnn_model: str,
        gripper_control: bool,
        discrete_gripper: bool,
    ):
        super().__init__()
        if not discrete_gripper:
            raise NotImplementedError("continuous gripper head: every shipped config sets discrete_gripper: true")
        if num_layers != 2 or policy_rnn_dropout_p != 0.0:
            raise NotImplementedError("the recurrence kernels cover num_layers=2, policy_rnn_dropout_p=0 (conf defaults)")
        self.n_dist = n_mixtures
        self.gripper_control = gripper_control
        self.discrete_gripper = discrete_gripper
        self.log_scale_min = log_scale_min
        self.num_classes = num_classes
        self.plan_features = plan_features
        in_features = (perceptual_emb_slice[1] - perceptual_emb_slice[0]) + latent_goal_features + plan_features
        self.out_features = out_features - 1 if discrete_gripper else out_features
        self.gripper_alpha = gripper_alpha
        if rnn_model not in RNN_MODELS:
            raise ValueError(f"unknown rnn_model {rnn_model!r}")
        self.rnn = RNN_MODELS[rnn_model](in_features, hidden_size, num_layers, policy_rnn_dropout_p)
        self.mean_fc = nn.Linear(hidden_size, self.out_features * self.n_dist)
        self.log_scale_fc = nn.Linear(hidden_size, self.out_features * self.n_dist)
        self.prob_fc = nn.Linear(hidden_size, self.out_features * self.n_dist)
        self.register_buffer("one_hot_embedding_eye", torch.eye(self.n_dist))
        self.register_buffer("ones", torch.ones(1, 1, self.n_dist))
        self._setup_action_bounds(dataset_dir, act_max_bound, act_min_bound, load_action_bounds)
        self.gripper_fc = nn.Linear(hidden_size, 2)
        self.criterion = nn.CrossEntropyLoss()
        self.perceptual_emb_slice = perceptual_emb_slice
        self.hidden_state = None

    # ------------------------------------------------------------------ helpers
    def _loss_cfg(self):
        return (self.out_features, self.n_dist, self.num_classes, float(self.log_scale_min), float(self.gripper_alpha))

    def _head_params(self):
        return (self.prob_fc.weight, self.prob_fc.bias, self.mean_fc.weight, self.mean_fc.bias,
                self.log_scale_fc.weight, self.log_scale_fc.bias, self.gripper_fc.weight, self.gripper_fc.bias)

    def _run_rnn(self, latent_plan, perceptual_emb, latent_goal, h_0=None):
        """-> time-major hidden states [S,B,H] of the top layer and h_n [2,B,H]."""
        pe = perceptual_emb[..., slice(*self.perceptual_emb_slice)]
        r = self.rnn
        if isinstance(r, nn.Sequential):
            # mlp_decoder (logistic_decoder_rnn.py:270-272: ``x = self.rnn(x); h_n = None``): no state, every step independent
            B, S = pe.shape[:2]
            x = ops.DecoderInputFunction.apply(latent_plan, pe, latent_goal)
            Hs = ops.mlp(x, [(r[0].weight, r[0].bias), (r[2].weight, r[2].bias), (r[4].weight, r[4].bias)], [True, True, False])
            return Hs.view(S, B, -1), None
        if isinstance(r, (nn.GRU, nn.LSTM)):
            # h_0 follows torch: a [2,B,H] tensor for nn.GRU, an (h_0, c_0) pair for nn.LSTM; so does the returned h_n
            lstm = isinstance(r, nn.LSTM)
            h0, c0 = (h_0 if lstm else (h_0, None)) if h_0 is not None else (None, None)
            Hs, hn, cn = ops.GatedRNNDecoderFunction.apply(
                "lstm" if lstm else "gru", latent_plan, pe, latent_goal, h0, c0, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0,
                r.bias_hh_l0, r.weight_ih_l1, r.weight_hh_l1, r.bias_ih_l1, r.bias_hh_l1)
            return Hs, ((hn, cn) if lstm else hn)
        return ops.RNNDecoderFunction.apply(
            latent_plan, pe, latent_goal, h_0, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
            r.weight_ih_l1, r.weight_hh_l1, r.bias_ih_l1, r.bias_hh_l1,
        )

    def _fused_loss(self, Hs, actions):
        return ops.DecoderLossFunction.apply(Hs, actions, self.action_min_bound, self.action_max_bound, self._loss_cfg(),
                                             *self._head_params())

    def _fused_sample(self, heads, B, S, time_major, u1=None, u2=None):
        A, M = self.out_features, self.n_dist
        u1 = noise.uniform((B, S, A, M), heads.device) if u1 is None else u1
        u2 = noise.uniform((B, S, A), heads.device) if u2 is None else u2
        return ops.logistic_sample(heads, u1, u2, self.gripper_bounds, B, S, A, M, float(self.log_scale_min), time_major)

    # ------------------------------------------------------------------ reference API
    def clear_hidden_state(self) -> None:
        self.hidden_state = None

    def loss_and_act(self, latent_plan, perceptual_emb, latent_goal, actions, robot_obs) -> Tuple[torch.Tensor, torch.Tensor]:
        B, S = perceptual_emb.shape[:2]
        Hs, _ = self._run_rnn(latent_plan, perceptual_emb, latent_goal)
        with torch.no_grad():
            heads = ops.heads_forward(Hs.detach(), *self._head_params())
            pred_actions = self._fused_sample(heads, B, S, True)
        if self.gripper_control:
            loss = self._fused_loss(Hs, world_to_tcp_frame(actions, robot_obs))
            return loss, tcp_to_world_frame(pred_actions, robot_obs)
        return self._fused_loss(Hs, actions), pred_actions

    def loss_and_act_modalities(self, latent_plan, perceptual_emb, latent_goal, actions: Sequence[torch.Tensor],
                                robot_obs: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """``loss_and_act`` for windows of several modalities decoded by ONE recurrence call (validation, hulc2.py:268-291):
        returns the vector of per-modality mean losses and the per-modality sampled actions (world frame when
        ``gripper_control``), each exactly what ``loss_and_act`` returns for that modality."""
        B, S = perceptual_emb.shape[:2]
        Hs, _ = self._run_rnn(latent_plan, perceptual_emb, latent_goal)
        with torch.no_grad():
            heads = ops.heads_forward(Hs.detach(), *self._head_params())
            pred = self._fused_sample(heads, B, S, True)
        sizes = [a.shape[0] for a in actions]
        offs = [sum(sizes[:i]) for i in range(len(sizes))]
        preds = [pred[o : o + n] for o, n in zip(offs, sizes)]
        if self.gripper_control:
            losses = self._fused_loss(Hs, tuple(world_to_tcp_frame(a, r) for a, r in zip(actions, robot_obs)))
            return losses, [tcp_to_world_frame(p, r) for p, r in zip(preds, robot_obs)]
        return self._fused_loss(Hs, tuple(actions)), preds

    def act(self, latent_plan, perceptual_emb, latent_goal, robot_obs) -> torch.Tensor:
        B, S = perceptual_emb.shape[:2]
        with torch.no_grad():
            Hs, self.hidden_state = self._run_rnn(latent_plan, perceptual_emb, latent_goal, self.hidden_state)
            heads = ops.heads_forward(Hs, *self._head_params())
            pred_actions = self._fused_sample(heads, B, S, True)
            if self.gripper_control:
                return tcp_to_world_frame(pred_actions, robot_obs)
            return pred_actions

    def loss(self, latent_plan, perceptual_emb, latent_goal, actions, robot_obs) -> torch.Tensor:
        Hs, _ = self._run_rnn(latent_plan, perceptual_emb, latent_goal)
        if self.gripper_control:
            actions = world_to_tcp_frame(actions, robot_obs)
        return self._fused_loss(Hs, actions)

    def loss_modalities(self, latent_plan, perceptual_emb, latent_goal, actions: Sequence[torch.Tensor],
                        robot_obs: Sequence[torch.Tensor]) -> torch.Tensor:
        """``loss`` for windows of several modalities decoded by ONE recurrence call: ``latent_plan``, ``perceptual_emb``
        and ``latent_goal`` hold all windows in modality order, ``actions[i]`` / ``robot_obs[i]`` are modality i's own
        tensors.  Returns the vector of per-modality mean losses (each exactly what ``loss`` returns for that modality)."""
        Hs, _ = self._run_rnn(latent_plan, perceptual_emb, latent_goal)
        if self.gripper_control:
            actions = [world_to_tcp_frame(a, r) for a, r in zip(actions, robot_obs)]
        return self._fused_loss(Hs, tuple(actions))

    def _loss(self, logit_probs, log_scales, means, gripper_act, actions) -> torch.Tensor:
        return ops.LogisticLossFunction.apply(logit_probs, log_scales, means, gripper_act, actions,
                                              self.action_min_bound, self.action_max_bound, self._loss_cfg())

    def _setup_action_bounds(self, dataset_dir, act_max_bound, act_min_bound, load_action_bounds):
        if load_action_bounds:
            try:
                import hulc2_b200

                statistics_path = Path(hulc2_b200.__file__).parent / dataset_dir / "training/statistics.yaml"
                statistics = OmegaConf.load(statistics_path)
                act_max_bound = statistics.act_max_bound
                act_min_bound = statistics.act_min_bound
                logger.info(f"Loaded action bounds from {statistics_path}")
            except FileNotFoundError:
                logger.info("Could not load statistics.yaml, taking action bounds defined in hydra conf")
        act_max_bound, act_min_bound = list(act_max_bound), list(act_min_bound)
        dev = self.ones.device
        self.register_buffer("gripper_bounds", torch.tensor([act_min_bound[-1], act_max_bound[-1]], dtype=torch.float32, device=dev))
        action_max_bound = torch.tensor(act_max_bound[:-1], dtype=torch.float32, device=dev)
        action_min_bound = torch.tensor(act_min_bound[:-1], dtype=torch.float32, device=dev)
        assert action_max_bound.shape[0] == self.out_features
        assert action_min_bound.shape[0] == self.out_features
        action_max_bound = action_max_bound.view(1, 1, -1, 1) * self.ones  # [1, 1, action_space, N_DIST]
        action_min_bound = action_min_bound.view(1, 1, -1, 1) * self.ones
        self.register_buffer("action_max_bound", action_max_bound.contiguous())
        self.register_buffer("action_min_bound", action_min_bound.contiguous())

    def _logistic_loss(self, logit_probs, log_scales, means, actions) -> torch.Tensor:
        """Logistic NLL only (logistic_decoder_rnn.py:181-228); actions [B,S,A]."""
        B, S = actions.shape[:2]
        a7 = torch.cat([actions, actions.new_ones(B, S, 1)], -1)
        gr = actions.new_zeros(B, S, 2)
        cfg = (self.out_features, self.n_dist, self.num_classes, float(self.log_scale_min), 0.0)
        return ops.LogisticLossFunction.apply(logit_probs, log_scales, means, gr, a7, self.action_min_bound,
                                              self.action_max_bound, cfg)

    def _sample(self, logit_probs, log_scales, means, gripper_act, u1=None, u2=None) -> torch.Tensor:
        B, S = logit_probs.shape[:2]
        with torch.no_grad():
            heads = ops.heads_pack(logit_probs, log_scales, means, gripper_act)
            return self._fused_sample(heads, B, S, False, u1, u2)

    def forward(self, latent_plan, perceptual_emb, latent_goal, h_0: Optional[torch.Tensor] = None):
        """-> logit_probs, log_scales (clamped), means [B,S,A,M], gripper logits [B,S,2], h_n (inference/inspection
        API: the tensors are detached; training goes through ``loss`` which keeps everything fused)."""
        B, S = perceptual_emb.shape[:2]
        with torch.no_grad():
            Hs, h_n = self._run_rnn(latent_plan, perceptual_emb, latent_goal, h_0)
            heads = ops.heads_forward(Hs, *self._head_params())
            lp, ls, mu, gr = ops.heads_unpack(heads, B, S, self.out_features, self.n_dist, float(self.log_scale_min), True)
        return lp, ls, mu, gr, h_n
