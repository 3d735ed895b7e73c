"""Diagnostic: per-parameter gradient agreement (norm ratio, cosine) of the CUDA path vs the CPU oracle."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
warnings.filterwarnings("ignore")
import torch
from hulc2_b200 import noise, ops
from hulc2_b200.config import hulc2_config
from hulc2_b200.synthetic import synthetic_batch
from oracle import hulc2_oracle as O
from helpers import build_model, oracle_params, to_device

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch_init = len(sys.argv) > 3 and sys.argv[3] == "torchinit"
if torch_init:
    from hulc2_b200._compat import instantiate
    torch.manual_seed(0)
    m = instantiate(hulc2_config(dropout_p=0.0))
else:
    m = build_model("calvin")
P = oracle_params(m)
cfg = hulc2_config(pkg="x", dropout_p=0.0)
batch = synthetic_batch(B, seed=1, aux="all")
idx = {mod: torch.randint(0, 32, (B, 32), generator=torch.Generator().manual_seed(5)) for mod in batch}
out = O.training_step(batch, {mod: {"plan_idx": idx[mod]} for mod in batch}, P, cfg)
out["loss"].backward()
m = m.to("cuda").train()
ops.set_precision(prec)
with noise.supplied(categories=[idx[mod] for mod in batch]):
    loss = m.training_step(to_device(batch, "cuda"), 0)
loss.backward()
print("loss", float(loss), float(out["loss"]))
for k, v in m.logged.items():
    print(f"  {k:40s} {float(v):.6f} {float(out[k]):.6f}")
rows = []
for n, p in m.named_parameters():
    r = P[n].grad
    if r is None: continue
    g = p.grad.cpu()
    cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
    rows.append((n, float(g.norm() / (r.norm() + 1e-30)), cos, float(r.norm())))
for n, ratio, cos, rn in rows:
    print(f"{n:75s} ratio {ratio:8.4f} cos {cos:8.5f} |ref| {rn:.3e}")
