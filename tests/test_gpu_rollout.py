"""GPU: batched rollout server (hulc2_b200/rollout.py) against the reference's rollout fixtures
(hulc2.py:600-707 run step by step, tests/golden/make_golden.py) and graph replays against the eager state machine."""
import pytest
import torch

from hulc2_b200 import noise, ops
from hulc2_b200.rollout import RolloutServer
from hulc2_b200.synthetic import synthetic_obs

from helpers import assert_close, build_alt_model, build_model, gt, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32():
    ops.set_precision("fp32")
    yield
    ops.set_precision("fp32")


def test_eager_server_matches_reference_rollout_fixture():
    """Same fixture as test_gpu_step.test_rollout_golden (the unmodified reference's model.step over 4 steps, re-plan
    every 2), driven through the server's state machine with supplied noise."""
    m = build_model("calvin").to(DEV).eval()
    m.replan_freq = 2
    srv = RolloutServer(m, use_graph=False)
    obs, goal = synthetic_obs(4, seed=2)
    obs, goal = to_device(obs, DEV), to_device(goal, DEV)
    srv.reset()
    for s in range(4):
        cats = [gt(f"rollout_N4/step{s}/plan_idx")] if s % 2 == 0 else []
        with noise.supplied(categories=cats, uniforms=[gt(f"rollout_N4/step{s}/u1"), gt(f"rollout_N4/step{s}/u2")]):
            a = srv.step(obs, goal).clone()
        ref = gt(f"rollout_N4/step{s}/action")
        assert a.shape == (4, 1, 7)
        assert torch.equal(a[..., -1].cpu(), ref[..., -1]), f"gripper argmax differs at step {s}"
        assert_close(a, ref, 2e-5, f"action step {s}")


def _drive(srv, frames, goal, steps):
    noise.manual_seed(11)
    noise.epoch_tensor(DEV).zero_()
    srv.reset()
    out = []
    for s in range(steps):
        out.append(srv.step(frames[s % len(frames)], goal).clone())
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("precision,model", [("fp32", "rnn"), ("bf16", "rnn"), ("fp32", "gru"), ("fp32", "lstm")])
def test_graph_replays_equal_eager_state_machine(precision, model):
    """8 steps with re-plan every 3 (two re-plans replayed from the graph, hidden state carried in between): the captured
    graphs reproduce the eager bodies bit for bit (same kernels, same Philox (seed, offset, epoch) per draw)."""
    N = 6
    m = (build_model("calvin") if model == "rnn" else build_alt_model(model)).to(DEV).eval()
    m.replan_freq = 3
    ops.set_precision(precision)
    frames = [to_device(synthetic_obs(N, seed=20 + i)[0], DEV) for i in range(3)]
    goal = to_device(synthetic_obs(N, seed=20)[1], DEV)
    eager = _drive(RolloutServer(m, use_graph=False), frames, goal, 8)
    srv = RolloutServer(m, use_graph=True)
    graph = _drive(srv, frames, goal, 8)
    assert srv.launches["act"] > 0 and srv.launches["replan"] > 0 and len(srv._graphs) == 2
    for s, (a, b) in enumerate(zip(graph, eager)):
        assert torch.equal(a, b), f"step {s}: max diff {float((a - b).abs().max())}"
    # the noise really advances between replays, and the hidden state is carried (same frame, different action)
    assert not torch.equal(graph[1], graph[4])
    # a second episode through the same graphs
    again = _drive(srv, frames, goal, 8)
    for a, b in zip(again, eager):
        assert torch.equal(a, b)


def test_uint8_camera_frames_and_host_observations():
    """Observations as the environment produces them: uint8 HWC frames in pinned host memory."""
    N = 4
    m = build_model("calvin").to(DEV).eval()
    g = torch.Generator().manual_seed(5)
    obs_f, goal = synthetic_obs(N, seed=31)
    u8s = torch.randint(0, 256, (N, 1, 200, 200, 3), generator=g, dtype=torch.uint8)
    u8g = torch.randint(0, 256, (N, 1, 84, 84, 3), generator=g, dtype=torch.uint8)
    obs_u8 = dict(obs_f, rgb_obs={"rgb_static": u8s.pin_memory(), "rgb_gripper": u8g.pin_memory()})
    obs_f = dict(obs_f, rgb_obs={"rgb_static": (u8s.permute(0, 1, 4, 2, 3).float() / 255.0 - 0.5) / 0.5,
                                 "rgb_gripper": (u8g.permute(0, 1, 4, 2, 3).float() / 255.0 - 0.5) / 0.5})
    a = _drive(RolloutServer(m, use_graph=True), [obs_u8], goal, 3)
    b = _drive(RolloutServer(m, use_graph=False), [to_device(obs_f, DEV)], to_device(goal, DEV), 3)
    for s, (x, y) in enumerate(zip(a, b)):
        assert_close(x, y, 1e-5, f"step {s}")
