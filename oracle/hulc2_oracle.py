"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the HULC++ low-level policy step.

This file restates, in plain functional torch (fp32 on the host), the algorithm of
the reference's hot path.  It is the parity oracle for the
CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
``hulc2_b200`` never imports anything under ``oracle/``.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself
run in the build container: ``tests/golden/make_golden.py`` imports the unmodified
reference modules from ``/root/reference`` (via ``oracle/ref_import.py``) and commits
their outputs as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against those fixtures everywhere, and
``tests/test_oracle_vs_reference.py`` checks it live against the reference when
``/root/reference`` is present.

All tensors are fp32.  ``P`` is a flat dict of parameters/buffers under the
reference's state_dict names.  All randomness (plan category indices, dropout keep
masks, sampling uniforms) is an explicit input.

Reference files restated (paths relative to /root/reference):
  hulc2/models/hulc2.py                                   (Hulc2)
  hulc2/models/perceptual_encoders/concat_encoders.py     (ConcatEncoders)
  hulc2/models/perceptual_encoders/vision_network.py      (VisionNetwork, SpatialSoftmax)
  hulc2/models/perceptual_encoders/vision_network_gripper.py
  hulc2/models/plan_encoders/plan_recognition_net.py
  hulc2/models/plan_encoders/plan_proposal_net.py
  hulc2/models/encoders/goal_encoders.py
  hulc2/models/auxiliary_loss_networks/proj_vis_lang.py
  hulc2/models/decoders/logistic_decoder_rnn.py
  hulc2/models/decoders/utils/gripper_control.py  (+ pytorch3d euler conversions, unpinned dep)
  hulc2/utils/distributions.py
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- small helpers
def linear(x: Tensor, P: Dict[str, Tensor], name: str) -> Tensor:
    return x @ P[name + ".weight"].t() + P[name + ".bias"]


def layer_norm(x: Tensor, P, name: str, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * P[name + ".weight"] + P[name + ".bias"]


def _drop(x: Tensor, keep: Optional[Tensor], p: float) -> Tensor:
    """Inverted dropout with an explicit boolean/0-1 keep mask (None = identity)."""
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype) / (1.0 - p)


# ----------------------------------------------------------------------------- a2/a3 static encoder
def spatial_softmax(x: Tensor, x_map: Tensor, y_map: Tensor, temperature: Tensor) -> Tensor:
    """vision_network.py:100-108.  x [N,C,H,W] -> [N, 2C] interleaved (x_c0, y_c0, x_c1, ...)."""
    n, c, h, w = x.shape
    flat = x.contiguous().view(-1, h * w)
    att = torch.softmax(flat / temperature, dim=1)
    ex = torch.sum(x_map * att, dim=1, keepdim=True)
    ey = torch.sum(y_map * att, dim=1, keepdim=True)
    return torch.cat((ex, ey), 1).view(-1, c * 2)


def static_encoder(x: Tensor, P, pre: str, l2_normalize: bool = False) -> Tensor:
    """vision_network.py:55-65.  x [N,C,H,W] -> [N,64]."""
    x = F.relu(F.conv2d(x, P[pre + "conv_model.0.weight"], P[pre + "conv_model.0.bias"], stride=4))
    x = F.relu(F.conv2d(x, P[pre + "conv_model.2.weight"], P[pre + "conv_model.2.bias"], stride=2))
    x = F.relu(F.conv2d(x, P[pre + "conv_model.4.weight"], P[pre + "conv_model.4.bias"], stride=1))
    x = spatial_softmax(
        x, P[pre + "spatial_softmax.x_map"], P[pre + "spatial_softmax.y_map"], P[pre + "spatial_softmax.temperature"]
    )
    x = F.relu(linear(x, P, pre + "fc1.0"))
    x = linear(x, P, pre + "fc2")
    if l2_normalize:
        x = F.normalize(x, p=2, dim=1)
    return layer_norm(x, P, pre + "ln")


# ----------------------------------------------------------------------------- a4 gripper encoder
def gripper_encoder(x: Tensor, P, pre: str, l2_normalize: bool = False) -> Tensor:
    """vision_network_gripper.py:82-89 with nature_cnn :11-26.  x [N,C,84,84] -> [N,64]."""
    x = F.relu(F.conv2d(x, P[pre + "conv_model.0.weight"], P[pre + "conv_model.0.bias"], stride=4))
    x = F.relu(F.conv2d(x, P[pre + "conv_model.2.weight"], P[pre + "conv_model.2.bias"], stride=2))
    x = F.relu(F.conv2d(x, P[pre + "conv_model.4.weight"], P[pre + "conv_model.4.bias"], stride=1))
    x = x.flatten(1)  # (C,H,W) order
    x = F.relu(linear(x, P, pre + "conv_model.7"))
    x = F.relu(linear(x, P, pre + "fc1.0"))
    x = linear(x, P, pre + "fc2")
    if l2_normalize:
        x = F.normalize(x, p=2, dim=1)
    return layer_norm(x, P, pre + "ln")


# ----------------------------------------------------------------------------- a1 concat encoders
def perceptual_encoder(rgb_obs: Dict[str, Tensor], depth_obs: Dict[str, Tensor], P) -> Tensor:
    """concat_encoders.py:59-109 (rgb_static [+depth_static] [+rgb_gripper]; proprio=none)."""
    rs = rgb_obs["rgb_static"]
    b, s, c, h, w = rs.shape
    enc = static_encoder(rs.reshape(-1, c, h, w), P, "perceptual_encoder.rgb_static_encoder.").reshape(b, s, -1)
    if depth_obs and "depth_static" in depth_obs:
        d = depth_obs["depth_static"].unsqueeze(2).reshape(-1, 1, h, w)
        ed = static_encoder(d, P, "perceptual_encoder.depth_static_encoder.").reshape(b, s, -1)
        enc = torch.cat([enc, ed], -1)
    if "rgb_gripper" in rgb_obs:
        rg = rgb_obs["rgb_gripper"]
        b, s, c, h, w = rg.shape
        eg = gripper_encoder(rg.reshape(-1, c, h, w), P, "perceptual_encoder.rgb_gripper_encoder.").reshape(b, s, -1)
        enc = torch.cat([enc, eg], -1)
    return enc


# ----------------------------------------------------------------------------- a5 goal encoders
def visual_goal(x: Tensor, P) -> Tensor:
    """goal_encoders.py:29-34."""
    x = F.relu(linear(x, P, "visual_goal.mlp.0"))
    x = F.relu(linear(x, P, "visual_goal.mlp.2"))
    x = linear(x, P, "visual_goal.mlp.4")
    return layer_norm(x, P, "visual_goal.ln")


def language_goal(x: Tensor, P) -> Tensor:
    """goal_encoders.py:62-71 (lang_net=None, word_dropout_p=0)."""
    x = F.relu(linear(x, P, "language_goal.mlp.1"))
    x = F.relu(linear(x, P, "language_goal.mlp.3"))
    x = linear(x, P, "language_goal.mlp.5")
    return layer_norm(x, P, "language_goal.ln")


# ----------------------------------------------------------------------------- a6 plan proposal
def plan_proposal(emb0: Tensor, goal: Tensor, P) -> Tensor:
    """plan_proposal_net.py:42-47 -> state logits [B, plan_features]."""
    x = torch.cat([emb0, goal], -1)
    for i in (0, 2, 4, 6):
        x = F.relu(linear(x, P, f"plan_proposal.fc_model.{i}"))
    return linear(x, P, "plan_proposal.fc_state.0")


# ----------------------------------------------------------------------------- a7 plan recognition
def mha(x: Tensor, P, pre: str, num_heads: int, attn_keep: Optional[Tensor], p: float) -> Tensor:
    """torch nn.MultiheadAttention self-attention on [B,S,E] (batch-major restatement).
    in_proj_weight packs [q;k;v] rows; scale 1/sqrt(head_dim) on q; dropout on probabilities."""
    B, S, E = x.shape
    dh = E // num_heads
    qkv = x @ P[pre + "in_proj_weight"].t() + P[pre + "in_proj_bias"]
    q, k, v = qkv.split(E, dim=-1)
    q = q.view(B, S, num_heads, dh).transpose(1, 2)
    k = k.view(B, S, num_heads, dh).transpose(1, 2)
    v = v.view(B, S, num_heads, dh).transpose(1, 2)
    att = torch.softmax((q / math.sqrt(dh)) @ k.transpose(-1, -2), dim=-1)  # [B,H,S,S]
    att = _drop(att, attn_keep, p)
    o = (att @ v).transpose(1, 2).reshape(B, S, E)
    return linear(o, P, pre + "out_proj")


def plan_recognition(
    emb: Tensor, P, num_heads: int = 8, num_layers: int = 2, dropout_p: float = 0.0, masks: Optional[dict] = None
) -> Tuple[Tensor, Tensor]:
    """plan_recognition_net.py:125-148 (position_embedding=true, no padding, post-LN layers).
    ``masks`` (all in batch-major [B,S,...] layout): 'emb', and per layer i 'attn{i}' [B,H,S,S],
    'sa{i}' [B,S,E], 'ff1{i}' [B,S,FF], 'ff2{i}' [B,S,E].  Returns (state logits, seq_feat)."""
    masks = masks or {}
    B, S, E = emb.shape
    x = emb + P["plan_recognition.position_embeddings.weight"][:S].unsqueeze(0)
    x = _drop(x, masks.get("emb"), dropout_p)
    for i in range(num_layers):
        pre = f"plan_recognition.transformer_encoder.layers.{i}."
        sa = mha(x, P, pre + "self_attn.", num_heads, masks.get(f"attn{i}"), dropout_p)
        x = layer_norm(x + _drop(sa, masks.get(f"sa{i}"), dropout_p), P, pre + "norm1")
        ff = _drop(F.relu(linear(x, P, pre + "linear1")), masks.get(f"ff1{i}"), dropout_p)
        ff = linear(ff, P, pre + "linear2")
        x = layer_norm(x + _drop(ff, masks.get(f"ff2{i}"), dropout_p), P, pre + "norm2")
    x = linear(x, P, "plan_recognition.fc")  # [B,S,4096]
    seq_feat = x.mean(dim=1)
    return linear(seq_feat, P, "plan_recognition.fc_state.0"), seq_feat


# ----------------------------------------------------------------------------- a8 distribution
def onehot_from_index(idx: Tensor, class_size: int) -> Tensor:
    """distributions.py:37-41: OneHotCategorical.sample() given the drawn indices [B,cat]."""
    return F.one_hot(idx.long(), class_size).to(torch.float32)


def rsample_straight_through(logits: Tensor, idx: Tensor, category_size: int, class_size: int) -> Tensor:
    """OneHotCategoricalStraightThrough.rsample (hulc2.py:235): one-hot + (p - sg(p)); flattened."""
    lg = logits.view(*logits.shape[:-1], category_size, class_size)
    probs = torch.softmax(lg, -1)
    st = onehot_from_index(idx, class_size) + (probs - probs.detach())
    return st.flatten(-2, -1)


def _kl_cat(p_logits: Tensor, q_logits: Tensor, category_size: int, class_size: int) -> Tensor:
    """torch.distributions kl(Independent(OneHotCategorical p) || Independent(... q)) -> [B]."""
    lp = torch.log_softmax(p_logits.view(-1, category_size, class_size), -1)
    lq = torch.log_softmax(q_logits.view(-1, category_size, class_size), -1)
    pp, qq = lp.exp(), lq.exp()
    t = pp * (lp - lq)
    t = torch.where(qq == 0, torch.full_like(t, float("inf")), t)
    t = torch.where(pp == 0, torch.zeros_like(t), t)
    return t.sum(-1).sum(-1)


def kl_loss(pp_logits, pr_logits, kl_beta, alpha, category_size=32, class_size=32) -> Tensor:
    """hulc2.py:444-466 (KL balancing)."""
    lhs = _kl_cat(pr_logits.detach(), pp_logits, category_size, class_size).mean()
    rhs = _kl_cat(pr_logits, pp_logits.detach(), category_size, class_size).mean()
    return (alpha * lhs + (1 - alpha) * rhs) * kl_beta


def gauss_state(x: Tensor) -> Tuple[Tensor, Tensor]:
    """distributions.py:55-59 (continuous): chunk -> mean, softplus(var) + 1e-4."""
    mean, var = torch.chunk(x, 2, dim=-1)
    return mean, F.softplus(var) + 0.0001


def _kl_normal(mp: Tensor, sp: Tensor, mq: Tensor, sq: Tensor) -> Tensor:
    """torch.distributions.kl._kl_normal_normal summed over the plan dims (Independent(Normal, 1)) -> [B]."""
    var_ratio = (sp / sq).pow(2)
    t1 = ((mp - mq) / sq).pow(2)
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(-1)


def kl_loss_gauss(pp: Tuple[Tensor, Tensor], pr: Tuple[Tensor, Tensor], kl_beta, alpha) -> Tensor:
    """hulc2.py:444-466 with the continuous plan (distributions.py:28-29): p = pr (posterior), q = pp (prior)."""
    lhs = _kl_normal(pr[0].detach(), pr[1].detach(), pp[0], pp[1]).mean()
    rhs = _kl_normal(pr[0], pr[1], pp[0].detach(), pp[1].detach()).mean()
    return (alpha * lhs + (1 - alpha) * rhs) * kl_beta


# ----------------------------------------------------------------------------- a16 tcp frames
def _rot(axis: str, a: Tensor) -> Tensor:
    c, s = torch.cos(a), torch.sin(a)
    o, z = torch.ones_like(a), torch.zeros_like(a)
    flat = {"X": (o, z, z, z, c, -s, z, s, c), "Y": (c, z, s, z, o, z, -s, z, c), "Z": (c, -s, z, s, c, z, z, z, o)}[axis]
    return torch.stack(flat, -1).reshape(a.shape + (3, 3))


def euler_xyz_to_matrix(e: Tensor) -> Tensor:
    """pytorch3d.transforms.euler_angles_to_matrix(e, "XYZ") = Rx(e0) Ry(e1) Rz(e2)."""
    return _rot("X", e[..., 0]) @ _rot("Y", e[..., 1]) @ _rot("Z", e[..., 2])


def matrix_to_euler_xyz(m: Tensor) -> Tensor:
    """pytorch3d.transforms.matrix_to_euler_angles(m, "XYZ")."""
    return torch.stack(
        (torch.atan2(-m[..., 1, 2], m[..., 2, 2]), torch.asin(m[..., 0, 2]), torch.atan2(-m[..., 0, 1], m[..., 0, 0])), -1
    )


def _wrap_pi(x: Tensor) -> Tensor:
    x = torch.where(x < -np.pi, x + 2 * np.pi, x)
    return torch.where(x > np.pi, x - 2 * np.pi, x)


def world_to_tcp_frame(action: Tensor, robot_obs: Tensor) -> Tensor:
    """gripper_control.py:16-36."""
    b, s, _ = action.shape
    w_T = euler_xyz_to_matrix(robot_obs[..., 3:6]).float().view(-1, 3, 3)
    t_T = torch.inverse(w_T)
    pos = t_T @ action[..., :3].reshape(-1, 3, 1)
    orn = action[..., 3:6] * 0.01
    w_T_new = euler_xyz_to_matrix(robot_obs[..., 3:6] + orn).float().view(-1, 3, 3)
    rel = torch.inverse(w_T_new) @ w_T
    o = _wrap_pi(matrix_to_euler_xyz(rel).float()) * 100
    return torch.cat([pos.view(b, s, -1), o.view(b, s, -1), action[..., -1:]], -1)


def matrix_to_quaternion(m: Tensor) -> Tensor:
    """pytorch3d.transforms.matrix_to_quaternion (unpinned third-party dependency, requirements.txt:20; restated from its
    published algorithm, rotation_conversions.py): the four candidate quaternions (r,i,j,k) built from the matrix, the one
    with the largest |component| selected, divided by 2 * max(that component, 0.1)."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(m.shape[:-2] + (9,)), -1)
    arg = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1)
    q_abs = torch.where(arg > 0, torch.sqrt(torch.clamp(arg, min=0)), torch.zeros_like(arg))      # _sqrt_positive_part
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
    ], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    pick = torch.nn.functional.one_hot(q_abs.argmax(-1), 4) > 0.5
    return cand[pick, :].reshape(m.shape[:-2] + (4,))


def quaternion_to_matrix(q: Tensor) -> Tensor:
    """pytorch3d.transforms.quaternion_to_matrix (real part first), restated from its published algorithm."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def tcp_to_world_frame(action: Tensor, robot_obs: Tensor) -> Tensor:
    """gripper_control.py:39-63, including the NaN fallback (:51-55): when asin leaves its domain by fp32 rounding near the
    gimbal pole (|M02| = 1 + ulp), the angles of the WHOLE batch are re-derived from the matrix re-normalised through a
    quaternion round trip."""
    b, s, _ = action.shape
    w_T = euler_xyz_to_matrix(robot_obs[..., 3:6]).float().view(-1, 3, 3)
    pos = w_T @ action[..., :3].reshape(-1, 3, 1)
    rel = euler_xyz_to_matrix(action[..., 3:6] * 0.01).float().view(-1, 3, 3)
    w_T_new = w_T @ torch.inverse(rel)
    e = matrix_to_euler_xyz(w_T_new).float()
    if bool(torch.any(e.isnan())):
        e = matrix_to_euler_xyz(quaternion_to_matrix(matrix_to_quaternion(w_T_new))).float()
    o = e - robot_obs[..., 3:6].reshape(-1, 3)
    o = _wrap_pi(o) * 100
    return torch.cat([pos.view(b, s, -1), o.view(b, s, -1), action[..., -1:]], -1)


# ----------------------------------------------------------------------------- a10-a15 decoder
def rnn_relu(x: Tensor, P, h0: Optional[Tensor], num_layers: int = 2) -> Tuple[Tensor, Tensor]:
    """nn.RNN(nonlinearity=relu, batch_first) (decoders/utils/rnn.py:5-14):
    h_t = relu(W_ih x_t + b_ih + W_hh h_{t-1} + b_hh); layer l input = layer l-1 output."""
    B, S, _ = x.shape
    hn = []
    inp = x
    for l in range(num_layers):
        Wi, Wh = P[f"action_decoder.rnn.weight_ih_l{l}"], P[f"action_decoder.rnn.weight_hh_l{l}"]
        bi, bh = P[f"action_decoder.rnn.bias_ih_l{l}"], P[f"action_decoder.rnn.bias_hh_l{l}"]
        h = h0[l] if h0 is not None else x.new_zeros(B, Wh.shape[0])
        outs = []
        for t in range(S):
            h = F.relu(inp[:, t] @ Wi.t() + bi + h @ Wh.t() + bh)
            outs.append(h)
        inp = torch.stack(outs, 1)
        hn.append(h)
    return inp, torch.stack(hn, 0)


def rnn_gru(x: Tensor, P, h0: Optional[Tensor], num_layers: int = 2) -> Tuple[Tensor, Tensor]:
    """nn.GRU (decoders/utils/rnn.py:28-36); gate order r,z,n."""
    B, S, _ = x.shape
    hn, inp = [], x
    for l in range(num_layers):
        Wi, Wh = P[f"action_decoder.rnn.weight_ih_l{l}"], P[f"action_decoder.rnn.weight_hh_l{l}"]
        bi, bh = P[f"action_decoder.rnn.bias_ih_l{l}"], P[f"action_decoder.rnn.bias_hh_l{l}"]
        H = Wh.shape[1]
        h = h0[l] if h0 is not None else x.new_zeros(B, H)
        outs = []
        for t in range(S):
            gi = inp[:, t] @ Wi.t() + bi
            gh = h @ Wh.t() + bh
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H : 2 * H] + gh[:, H : 2 * H])
            n = torch.tanh(gi[:, 2 * H :] + r * gh[:, 2 * H :])
            h = (1 - z) * n + z * h
            outs.append(h)
        inp = torch.stack(outs, 1)
        hn.append(h)
    return inp, torch.stack(hn, 0)


def rnn_lstm(x: Tensor, P, hc0, num_layers: int = 2):
    """nn.LSTM (decoders/utils/rnn.py:17-25); gate order i,f,g,o."""
    B, S, _ = x.shape
    hn, cn, inp = [], [], x
    for l in range(num_layers):
        Wi, Wh = P[f"action_decoder.rnn.weight_ih_l{l}"], P[f"action_decoder.rnn.weight_hh_l{l}"]
        bi, bh = P[f"action_decoder.rnn.bias_ih_l{l}"], P[f"action_decoder.rnn.bias_hh_l{l}"]
        H = Wh.shape[1]
        h = hc0[0][l] if hc0 is not None else x.new_zeros(B, H)
        c = hc0[1][l] if hc0 is not None else x.new_zeros(B, H)
        outs = []
        for t in range(S):
            g = inp[:, t] @ Wi.t() + bi + h @ Wh.t() + bh
            i, f, gg, o = g[:, :H], g[:, H : 2 * H], g[:, 2 * H : 3 * H], g[:, 3 * H :]
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, 1)
        hn.append(h)
        cn.append(c)
    return inp, (torch.stack(hn, 0), torch.stack(cn, 0))


def mlp_dec(x: Tensor, P, h0=None):
    """decoders/utils/rnn.py:39-46 (mlp_decoder): Linear-ReLU-Linear-ReLU-Linear per step, no state (h_n = None)."""
    x = F.relu(linear(x, P, "action_decoder.rnn.0"))
    x = F.relu(linear(x, P, "action_decoder.rnn.2"))
    return linear(x, P, "action_decoder.rnn.4"), None


def decoder_forward(plan, emb, goal, P, emb_slice=(64, 128), h0=None, n_dist=10, log_scale_min=-7.0, rnn="rnn_decoder"):
    """logistic_decoder_rnn.py:257-284."""
    pe = emb[..., emb_slice[0] : emb_slice[1]]
    B, S = pe.shape[:2]
    x = torch.cat([plan.unsqueeze(1).expand(-1, S, -1), pe, goal.unsqueeze(1).expand(-1, S, -1)], -1)
    fn = {"rnn_decoder": rnn_relu, "gru_decoder": rnn_gru, "lstm_decoder": rnn_lstm, "mlp_decoder": mlp_dec}[rnn]
    x, hn = fn(x, P, h0)
    probs = linear(x, P, "action_decoder.prob_fc")
    means = linear(x, P, "action_decoder.mean_fc")
    log_scales = torch.clamp(linear(x, P, "action_decoder.log_scale_fc"), min=log_scale_min)
    grip = linear(x, P, "action_decoder.gripper_fc")
    A = probs.shape[-1] // n_dist
    return probs.view(B, S, A, n_dist), log_scales.view(B, S, A, n_dist), means.view(B, S, A, n_dist), grip, hn


def logistic_loss(logit_probs, log_scales, means, actions, P, num_classes=10, log_scale_min=-7.0) -> Tensor:
    """logistic_decoder_rnn.py:181-228 (+ log_sum_exp :19-24)."""
    amax, amin = P["action_decoder.action_max_bound"], P["action_decoder.action_min_bound"]
    log_scales = torch.clamp(log_scales, min=log_scale_min)
    a = actions.unsqueeze(-1) * P["action_decoder.ones"]
    centered = a - means
    inv_std = torch.exp(-log_scales)
    act_range = (amax - amin) / 2.0
    plus_in = inv_std * (centered + act_range / (num_classes - 1))
    cdf_plus = torch.sigmoid(plus_in)
    min_in = inv_std * (centered - act_range / (num_classes - 1))
    cdf_min = torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    mid_in = inv_std * centered
    log_pdf_mid = mid_in - log_scales - 2.0 * F.softplus(mid_in)
    cdf_delta = cdf_plus - cdf_min
    log_probs = torch.where(
        a < amin + 1e-3,
        log_cdf_plus,
        torch.where(
            a > amax - 1e-3,
            log_one_minus_cdf_min,
            torch.where(
                cdf_delta > 1e-5,
                torch.log(torch.clamp(cdf_delta, min=1e-12)),
                log_pdf_mid - np.log((num_classes - 1) / 2),
            ),
        ),
    )
    log_probs = log_probs + torch.log_softmax(logit_probs, dim=-1)
    m = log_probs.max(-1, keepdim=True)[0]
    lse = m.squeeze(-1) + torch.log(torch.sum(torch.exp(log_probs - m), -1))
    return -torch.sum(lse, dim=-1).mean()


def decoder_loss(logit_probs, log_scales, means, grip, actions, P, gripper_alpha=1.0) -> Tensor:
    """logistic_decoder_rnn.py:133-152 (discrete_gripper=True)."""
    ll = logistic_loss(logit_probs, log_scales, means, actions[:, :, :-1], P)
    gt = actions[:, :, -1].clone()
    gt[gt == -1] = 0
    ce = F.cross_entropy(grip.reshape(-1, 2), gt.reshape(-1).long())
    return ll + gripper_alpha * ce


def decoder_sample(logit_probs, log_scales, means, grip, u1: Tensor, u2: Tensor, P) -> Tensor:
    """logistic_decoder_rnn.py:231-255 with torch.rand replaced by the supplied uniforms
    u1 [B,S,A,M] (first draw) and u2 [B,S,A] (second draw)."""
    r1, r2 = 1e-5, 1.0 - 1e-5
    temp = (r1 - r2) * u1 + r2
    temp = logit_probs - torch.log(-torch.log(temp))
    argmax = torch.argmax(temp, -1)
    dist = P["action_decoder.one_hot_embedding_eye"][argmax]
    ls = (dist * log_scales).sum(-1)
    mu = (dist * means).sum(-1)
    u = (r1 - r2) * u2 + r2
    act = mu + torch.exp(ls) * (torch.log(u) - torch.log(1.0 - u))
    g = P["action_decoder.gripper_bounds"][grip.argmax(-1)]
    return torch.cat([act, g.unsqueeze(-1)], 2)


# ----------------------------------------------------------------------------- a17/a18 InfoNCE
def clip_loss(seq_feat: Tensor, goal: Tensor, use: Optional[Tensor], P) -> Tensor:
    """hulc2.py:472-508 + proj_vis_lang.py:23-27."""
    if use is not None:
        if not bool(torch.any(use)):
            return torch.tensor(0.0)
        seq_feat, goal = seq_feat[use], goal[use]
    im = linear(F.relu(linear(seq_feat, P, "proj_vis_lang.mlp_im.0")), P, "proj_vis_lang.mlp_im.2")
    tx = linear(F.relu(linear(goal, P, "proj_vis_lang.mlp_lang.0")), P, "proj_vis_lang.mlp_lang.2")
    im = im / im.norm(dim=-1, keepdim=True)
    tx = tx / tx.norm(dim=-1, keepdim=True)
    logits = P["logit_scale"].exp() * im @ tx.t()
    labels = torch.arange(logits.shape[0])
    return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2


# ----------------------------------------------------------------------------- a19/a20 train step
def lmp_train(emb, goal, actions, robot_obs_raw, plan_idx, P, cfg, masks=None):
    """hulc2.py:200-245.  plan_idx [B,32] = the category indices pr_dist.rsample() drew (discrete plan), or the
    standard-normal eps [B,plan_features] of Normal.rsample() (continuous plan)."""
    dec = cfg["action_decoder"]
    pp_logits = plan_proposal(emb[:, 0], goal, P)
    pr_logits, seq_feat = plan_recognition(
        emb, P, cfg["plan_recognition"]["num_heads"], cfg["plan_recognition"]["num_layers"],
        cfg["plan_recognition"]["dropout_p"], masks,
    )
    if cfg["distribution"]["dist"] == "continuous":
        pp, pr = gauss_state(pp_logits), gauss_state(pr_logits)
        plan = pr[0] + pr[1] * plan_idx
        lp, ls, mu, grip, _ = decoder_forward(
            plan, emb, goal, P, tuple(dec["perceptual_emb_slice"]), None, dec["n_mixtures"], dec["log_scale_min"], dec["rnn_model"]
        )
        acts = world_to_tcp_frame(actions, robot_obs_raw) if dec["gripper_control"] else actions
        action_loss = decoder_loss(lp, ls, mu, grip, acts, P, dec["gripper_alpha"])
        kl = kl_loss_gauss(pp, pr, cfg["kl_beta"], cfg["kl_balancing_mix"])
        return kl, action_loss, action_loss + kl, pp_logits, pr_logits, seq_feat
    plan = rsample_straight_through(pr_logits, plan_idx, 32, 32)
    lp, ls, mu, grip, _ = decoder_forward(
        plan, emb, goal, P, tuple(dec["perceptual_emb_slice"]), None, dec["n_mixtures"], dec["log_scale_min"], dec["rnn_model"]
    )
    acts = world_to_tcp_frame(actions, robot_obs_raw) if dec["gripper_control"] else actions
    action_loss = decoder_loss(lp, ls, mu, grip, acts, P, dec["gripper_alpha"])
    kl = kl_loss(pp_logits, pr_logits, cfg["kl_beta"], cfg["kl_balancing_mix"])
    return kl, action_loss, action_loss + kl, pp_logits, pr_logits, seq_feat


def training_step(batch: dict, noise: dict, P, cfg) -> Dict[str, Tensor]:
    """hulc2.py:336-442.  ``noise[mod]`` = {"plan_idx": [B,32] int64, "masks": dict|None}.
    Returns the logged scalars under the reference's names; "loss" = returned total."""
    out: Dict[str, Tensor] = {}
    kl_sum = act_sum = tot_sum = 0.0
    clip = torch.tensor(0.0)
    for mod, db in batch.items():
        emb = perceptual_encoder(db["rgb_obs"], db["depth_obs"], P)
        goal = language_goal(db["lang"], P) if "lang" in mod else visual_goal(emb[:, -1], P)
        kl, act, tot, _, _, seq_feat = lmp_train(
            emb, goal, db["actions"], db["state_info"]["robot_obs"], noise[mod]["plan_idx"], P, cfg, noise[mod].get("masks")
        )
        if "lang" in mod and cfg["use_clip_auxiliary_loss"] and bool(torch.any(db["use_for_aux_lang_loss"])):
            clip = clip + clip_loss(seq_feat, goal, db["use_for_aux_lang_loss"], P)
        kl_sum, act_sum, tot_sum = kl_sum + kl, act_sum + act, tot_sum + tot
        out[f"train/kl_loss_scaled_{mod}"] = kl
        out[f"train/action_loss_{mod}"] = act
        out[f"train/total_loss_{mod}"] = tot
    n = len(batch)
    total = tot_sum / n
    if cfg["use_clip_auxiliary_loss"]:
        total = total + cfg["clip_auxiliary_loss_beta"] * clip
        out["train/lang_clip_loss"] = cfg["clip_auxiliary_loss_beta"] * clip
    out["train/kl_loss"] = kl_sum / n
    out["train/action_loss"] = act_sum / n
    out["train/total_loss"] = total
    out["loss"] = total
    return out


# ----------------------------------------------------------------------------- a21 validation
def lmp_val(emb, goal, actions, robot_obs_raw, noise, P, cfg):
    """hulc2.py:247-334.  noise: plan_idx_pp, plan_idx_pr [B,32]; u1_pp,u2_pp,u1_pr,u2_pr."""
    dec = cfg["action_decoder"]
    sl = tuple(dec["perceptual_emb_slice"])

    def loss_and_act(plan, u1, u2):  # logistic_decoder_rnn.py:82-99
        lp, ls, mu, grip, _ = decoder_forward(plan, emb, goal, P, sl, None, dec["n_mixtures"], dec["log_scale_min"], dec["rnn_model"])
        pred = decoder_sample(lp, ls, mu, grip, u1, u2, P)
        if dec["gripper_control"]:
            loss = decoder_loss(lp, ls, mu, grip, world_to_tcp_frame(actions, robot_obs_raw), P, dec["gripper_alpha"])
            return loss, tcp_to_world_frame(pred, robot_obs_raw)
        return decoder_loss(lp, ls, mu, grip, actions, P, dec["gripper_alpha"]), pred

    def metrics(sample):
        mae = torch.mean(torch.abs(sample[..., :-1] - actions[..., :-1]), 1)
        g = torch.where(sample[..., -1] > 0, 1.0, -1.0)
        return mae, torch.mean((actions[..., -1] == g).float())

    cont = cfg["distribution"]["dist"] == "continuous"      # plan_idx_* then hold the standard-normal eps of Normal.sample()
    pp_logits = plan_proposal(emb[:, 0], goal, P)
    if cont:
        pp = gauss_state(pp_logits)
        plan_pp = pp[0] + pp[1] * noise["plan_idx_pp"]
    else:
        plan_pp = onehot_from_index(noise["plan_idx_pp"], 32).flatten(-2, -1)
    loss_pp, act_pp = loss_and_act(plan_pp, noise["u1_pp"], noise["u2_pp"])
    mae_pp, sr_pp = metrics(act_pp)
    pr_logits, seq_feat = plan_recognition(emb, P, cfg["plan_recognition"]["num_heads"], cfg["plan_recognition"]["num_layers"], 0.0, None)
    if cont:
        pr = gauss_state(pr_logits)
        plan_pr = pr[0] + pr[1] * noise["plan_idx_pr"]
    else:
        plan_pr = onehot_from_index(noise["plan_idx_pr"], 32).flatten(-2, -1)
    loss_pr, act_pr = loss_and_act(plan_pr, noise["u1_pr"], noise["u2_pr"])
    mae_pr, sr_pr = metrics(act_pr)
    kl = (kl_loss_gauss(pp, pr, cfg["kl_beta"], cfg["kl_balancing_mix"]) if cont
          else kl_loss(pp_logits, pr_logits, cfg["kl_beta"], cfg["kl_balancing_mix"]))
    return plan_pp, loss_pp, plan_pr, loss_pr, kl, mae_pp, mae_pr, sr_pp, sr_pr, seq_feat


# ----------------------------------------------------------------------------- a22 inference
class OracleRollout:
    """hulc2.py:600-707: stateful step()/reset() (language goal path), noise supplied per call."""

    def __init__(self, P, cfg):
        self.P, self.cfg = P, cfg
        self.reset()

    def reset(self):
        self.plan = self.latent_goal = self.hidden = None
        self.counter = 0

    def step(self, obs, goal, plan_idx: Optional[Tensor], u1: Tensor, u2: Tensor) -> Tensor:
        P, cfg, dec = self.P, self.cfg, self.cfg["action_decoder"]
        with torch.no_grad():
            if self.counter % cfg["replan_freq"] == 0:
                emb = perceptual_encoder(obs["rgb_obs"], obs["depth_obs"], P)
                self.latent_goal = language_goal(goal["lang"], P)
                _ = plan_proposal(emb[:, 0], self.latent_goal, P)  # logits only matter through plan_idx
                self.plan = onehot_from_index(plan_idx, 32).flatten(-2, -1)
                self.hidden = None
            emb = perceptual_encoder(obs["rgb_obs"], obs["depth_obs"], P)
            lp, ls, mu, grip, self.hidden = decoder_forward(
                self.plan, emb, self.latent_goal, P, tuple(dec["perceptual_emb_slice"]), self.hidden,
                dec["n_mixtures"], dec["log_scale_min"], dec["rnn_model"],
            )
            act = decoder_sample(lp, ls, mu, grip, u1, u2, P)
            if dec["gripper_control"]:
                act = tcp_to_world_frame(act, obs["robot_obs_raw"])
        self.counter += 1
        return act
