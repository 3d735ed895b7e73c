"""ConcatEncoders (mirror of hulc2/models/perceptual_encoders/concat_encoders.py:10-109).

Runs the per-camera encoders on ``[B*S, C, H, W]`` frames; each encoder's final LayerNorm writes its
64 features straight into its column block of the ``[B, S, latent]`` embedding (no ``torch.cat``).
Tactile / proprio / state-decoder branches of the reference are not on the scoped path.
"""
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from ... import ops
from ..._compat import DictConfig, as_config, instantiate


class _ConcatLayerNorm(torch.autograd.Function):
    """LayerNorm of k feature blocks [F, D_i] into one [F, sum D_i] buffer."""

    @staticmethod
    def forward(ctx, eps, *args):
        from ..._lib import call

        k = len(args) // 3
        feats, gammas, betas = args[:k], args[k : 2 * k], args[2 * k :]
        rows = feats[0].shape[0]
        Ds = [f.shape[1] for f in feats]
        total = sum(Ds)
        out = torch.empty(rows, total, device=feats[0].device, dtype=torch.float32)
        saved, off = [], 0
        for f, g, b, D in zip(feats, gammas, betas, Ds):
            f = f.contiguous()
            mean = torch.empty(rows, device=f.device, dtype=torch.float32)
            rstd = torch.empty(rows, device=f.device, dtype=torch.float32)
            call("hulc2_layernorm_fwd", f.data_ptr(), D, None, 0, None, 1.0, g.data_ptr(), b.data_ptr(),
                 out.data_ptr() + 4 * off, total, None, mean.data_ptr(), rstd.data_ptr(), rows, D, eps)
            saved += [f, g, mean, rstd]
            off += D
        ctx.save_for_backward(*saved)
        ctx.Ds, ctx.total = Ds, total
        return out

    @staticmethod
    def backward(ctx, dout):
        from ..._lib import call

        dout = dout.contiguous()
        saved = ctx.saved_tensors
        k = len(ctx.Ds)
        rows = dout.shape[0]
        dfs, dgs, dbs, off = [], [], [], 0
        for i, D in enumerate(ctx.Ds):
            f, g, mean, rstd = saved[4 * i : 4 * i + 4]
            dx = torch.empty(rows, D, device=dout.device, dtype=torch.float32)
            dg = ops.grad_buffer(g, zero=True)
            db = torch.zeros(D, device=dout.device, dtype=torch.float32)
            call("hulc2_layernorm_bwd", dout.data_ptr() + 4 * off, ctx.total, f.data_ptr(), D, g.data_ptr(), mean.data_ptr(),
                 rstd.data_ptr(), dx.data_ptr(), D, None, None, 1.0, dg.data_ptr(), db.data_ptr(), rows, D)
            dfs.append(dx)
            dgs.append(dg)
            dbs.append(db)
            off += D
        return (None, *dfs, *dgs, *dbs)


class ConcatEncoders(nn.Module):
    def __init__(
        self,
        rgb_static: DictConfig,
        proprio: DictConfig,
        device: torch.device,
        depth_static: Optional[DictConfig] = None,
        rgb_gripper: Optional[DictConfig] = None,
        depth_gripper: Optional[DictConfig] = None,
        tactile: Optional[DictConfig] = None,
        state_decoder: Optional[DictConfig] = None,
    ):
        super().__init__()
        rgb_static, rgb_gripper = as_config(rgb_static), as_config(rgb_gripper)
        depth_static, depth_gripper = as_config(depth_static), as_config(depth_gripper)
        if tactile or proprio or state_decoder:
            raise NotImplementedError("tactile / proprio / state_decoder encoders are outside the scoped hot path")
        self._latent_size = rgb_static.visual_features
        if rgb_gripper:
            self._latent_size += rgb_gripper.visual_features
        if depth_static:
            self._latent_size += depth_static.visual_features
        if depth_gripper:
            self._latent_size += depth_gripper.visual_features
        self.rgb_static_encoder = instantiate(rgb_static)
        self.depth_static_encoder = instantiate(depth_static) if depth_static else None
        self.rgb_gripper_encoder = instantiate(rgb_gripper) if rgb_gripper else None
        self.depth_gripper_encoder = instantiate(depth_gripper) if depth_gripper else None
        self.tactile_encoder = None
        self.proprio_encoder = None
        self.state_decoder = None
        self.current_visual_embedding = None
        self.current_state_obs = None

    @property
    def latent_size(self):
        return self._latent_size

    def forward(self, imgs: Dict[str, torch.Tensor], depth_imgs: Dict[str, torch.Tensor], state_obs: torch.Tensor) -> torch.Tensor:
        return self.forward_modalities([imgs], [depth_imgs], state_obs)

    def forward_modalities(self, imgs: Sequence[Dict[str, torch.Tensor]], depth_imgs: Sequence[Dict[str, torch.Tensor]],
                           state_obs=None) -> torch.Tensor:
        """concat_encoders.py:59-109 for the windows of several modalities at once: every camera's frames go through
        ONE trunk call (frame groups are packed back to back, never concatenated in fp32), returning the embeddings of
        all windows ``[sum(B_i), S, latent]`` in modality order.  All modalities must carry the same cameras."""
        def cam(dicts, key):
            vals = [d.get(key) if d else None for d in dicts]
            if all(v is None for v in vals):
                return None
            if any(v is None for v in vals):
                raise ValueError(f"camera '{key}' is missing from some modalities; encode them separately")
            return vals

        def frames(dicts, key):
            """Per modality: fp32 [B,S,C,H,W] (the reference batch contract) -> [B*S,C,H,W]; uint8 [B,S,H,W,C] (frames as
            the dataset stores them, optional ``<key>_shift`` int32 [B,S,2] RandomShiftsAug draw) or a ready
            ``ops.U8Frames`` (window gather from a resident store) -> U8Frames, converted inside the trunk."""
            vals = cam(dicts, key)
            if vals is None:
                return None, None
            out, bs = [], []
            for d, t in zip(dicts, vals):
                if isinstance(t, ops.U8Frames):
                    out.append(t)
                    bs.append((t.F // t.S, t.S))
                elif t.dtype == torch.uint8:
                    if t.dim() == 4:                      # depth-like single channel [B,S,H,W] is not a uint8 format
                        raise ValueError(f"uint8 camera '{key}' must be [B,S,H,W,C]")
                    shift = d.get(key + "_shift")
                    out.append(ops.U8Frames(t.reshape(-1, *t.shape[2:]), None if shift is None else shift.reshape(-1, 2), S=t.shape[1]))
                    bs.append((t.shape[0], t.shape[1]))
                else:
                    chw = t.shape[2:] if t.dim() == 5 else (1, *t.shape[2:])
                    out.append(t.reshape(-1, *chw))
                    bs.append((t.shape[0], t.shape[1]))
            return tuple(out), bs

        rgb_static, bs = frames(imgs, "rgb_static")
        s = bs[0][1]
        b = sum(n for n, _ in bs)
        encs = [(self.rgb_static_encoder, rgb_static)]
        depth_static, _ = frames(depth_imgs, "depth_static")
        if depth_static is not None:
            encs.append((self.depth_static_encoder, depth_static))
        rgb_gripper, _ = frames(imgs, "rgb_gripper")
        if rgb_gripper is not None:
            encs.append((self.rgb_gripper_encoder, rgb_gripper))
            depth_gripper, _ = frames(depth_imgs, "depth_gripper")
            if depth_gripper is not None:
                encs.append((self.depth_gripper_encoder, depth_gripper))
        feats = [enc.features(x if len(x) > 1 else x[0]) for enc, x in encs]
        gammas = [enc.ln.weight for enc, _ in encs]
        betas = [enc.ln.bias for enc, _ in encs]
        out = _ConcatLayerNorm.apply(float(encs[0][0].ln.eps), *feats, *gammas, *betas)
        perceptual_emb = out.view(b, s, -1)
        self.current_visual_embedding = perceptual_emb
        self.current_state_obs = state_obs
        return perceptual_emb
