"""Batched rollout server around ``Hulc2.step`` (N parallel envs, S=1).

The reference evaluates one environment at a time (``rollout.py:342-346``, ``evaluate_policy.py``): per control step it
calls ``model.step(obs, goal)``, which re-plans every ``replan_freq`` steps (``hulc2.py:608-628``) and otherwise runs the
encoders on one frame, one recurrence step from the cached decoder hidden state and one draw from the logistic mixture
(``logistic_decoder_rnn.py:101-116``).  At B=1 that path is launch-latency bound (~150 module calls per step).  The
server keeps the same state machine -- plan, latent goal, step counter, decoder hidden state -- for N environments at
once in STATIC device buffers and replays two captured CUDA graphs over them:

* ``replan`` graph: ``get_pp_plan_lang`` / ``get_pp_plan_vision`` -> plan, latent goal; zeroes the hidden state
  (``clear_hidden_state``: ``h_0 = None`` == zeros);
* ``act`` graph: ``predict_with_plan`` from the static observation, plan, goal and hidden state -> action ``[N,1,7]``,
  hidden state updated in place.

Noise (plan categories / Gaussian eps, Gumbel + logistic uniforms) comes from the library's Philox kernels keyed by a
device-resident epoch counter that each replay bumps, so replays draw fresh noise.  ``use_graph=False`` drives the same
bodies eagerly (and accepts ``noise.supplied`` tensors: the parity-test mode).  Environments step in lock-step, like the
reference's vectorised evaluation would; ``reset()`` restarts all of them.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import noise
from ._lib import call, load_library
from .synthetic import tree_map


def _copy_tree(dst, src):
    if isinstance(dst, dict):
        for k in dst:
            _copy_tree(dst[k], src[k])
    elif isinstance(dst, torch.Tensor):
        if dst.data_ptr() != src.data_ptr():
            dst.copy_(src, non_blocking=True)


class RolloutServer:
    def __init__(self, model, use_graph: bool = True):
        self.model = model.eval()
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("RolloutServer needs the model on a CUDA device (there is no CPU path)")
        self.use_graph = use_graph
        self.stream = torch.cuda.Stream(device=self.device)
        self.obs: Optional[Dict[str, Any]] = None       # static observation buffers (what the graphs read)
        self.goal: Optional[Dict[str, Any]] = None
        self.plan = self.latent_goal = self.action = None
        self.hidden = None                              # tensor [L,N,H] (RNN/GRU) or (h, c) pair (LSTM)
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._keepalive: list = []
        self._act_offset = 0
        self.launches = {"replan": 0, "act": 0}
        self.rollout_step_counter = 0

    # ------------------------------------------------------------------ state machine (hulc2.py:600-628)
    def reset(self) -> None:
        self.rollout_step_counter = 0
        self.model.reset()

    def step(self, obs: Dict[str, Any], goal: Dict[str, Any]) -> torch.Tensor:
        """obs / goal as ``Hulc2.step`` takes them (device or pinned-host tensors, ``[N,1,...]``); returns the action
        ``[N,1,7]`` on the device (a static buffer, overwritten by the next call)."""
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        with torch.cuda.stream(self.stream), torch.no_grad():
            self._load(obs, goal)
            if self.rollout_step_counter % self.model.replan_freq == 0:
                self._run("replan")
            self._run("act")
        caller.wait_stream(self.stream)
        self.rollout_step_counter += 1
        self.model.rollout_step_counter = self.rollout_step_counter
        return self.action

    # ------------------------------------------------------------------ bodies
    def _load(self, obs, goal) -> None:
        if self.obs is None:
            dev = self.device
            self.obs = tree_map(lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev), obs)
            self.goal = tree_map(lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev), goal)
            rnn = self.model.action_decoder.rnn
            n = next(iter(obs["rgb_obs"].values())).shape[0]
            h = torch.zeros(rnn.num_layers, n, rnn.hidden_size, device=dev)
            self.hidden = (h, torch.zeros_like(h)) if isinstance(rnn, torch.nn.LSTM) else h
        _copy_tree(self.obs, obs)
        if self.rollout_step_counter % self.model.replan_freq == 0:
            _copy_tree(self.goal, goal)

    def _replan_body(self) -> None:
        m = self.model
        noise.begin_step()
        if "lang" in self.goal:
            plan, latent_goal = m.get_pp_plan_lang(self.obs, self.goal)
        else:
            plan, latent_goal = m.get_pp_plan_vision(self.obs, self.goal)
        if self.plan is None:
            self.plan, self.latent_goal = torch.empty_like(plan), torch.empty_like(latent_goal)
        self.plan.copy_(plan)
        self.latent_goal.copy_(latent_goal)
        m.plan, m.latent_goal = self.plan, self.latent_goal
        for h in (self.hidden if isinstance(self.hidden, tuple) else (self.hidden,)):
            h.zero_()                                    # clear_hidden_state(): h_0 = None == zeros
        self._act_offset = noise._state["counter"]
        call("hulc2_counter_add", noise.epoch_tensor(self.device).data_ptr(), 1)

    def _act_body(self) -> None:
        m = self.model
        noise._state["counter"] = self._act_offset      # the act draws follow the re-plan's in the Philox stream
        dec = m.action_decoder
        dec.hidden_state = self.hidden
        action = m.predict_with_plan(self.obs, self.latent_goal, self.plan)
        new = dec.hidden_state
        if self.action is None:
            self.action = torch.empty_like(action)
        for dst, src in zip(self.hidden if isinstance(self.hidden, tuple) else (self.hidden,),
                            new if isinstance(new, tuple) else (new,)):
            dst.copy_(src)
        dec.hidden_state = self.hidden
        self.action.copy_(action)
        call("hulc2_counter_add", noise.epoch_tensor(self.device).data_ptr(), 1)

    def _run(self, which: str) -> None:
        body = self._replan_body if which == "replan" else self._act_body
        if not self.use_graph:
            body()
            return
        g = self._graphs.get(which)
        if g is not None:
            g.replay()
            return
        # first call: run the body eagerly (this IS the step: it allocates the static outputs and fills the bf16 weight
        # mirrors), then record the same body for the following steps -- a capture executes nothing
        from . import ops

        body()
        torch.cuda.synchronize()
        n0 = load_library().hulc2_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            body()
        self.launches[which] = int(load_library().hulc2_launch_count() - n0)
        self._graphs[which] = g
        # the graph addresses the cached weight mirrors by pointer: keep them alive past ops.invalidate_weight_mirrors()
        self._keepalive.append((dict(ops._w16), [a[1] for a in ops._arenas]))

    def refresh(self) -> None:
        """Drop the captured graphs (parameters were trained / loaded since the capture: bf16 weight mirrors are stale)."""
        from . import ops

        ops.invalidate_weight_mirrors()
        self._graphs.clear()
        self._keepalive.clear()
