"""Diagnostic: where does the bf16 gradient error of the static-encoder tail come from at the bench shape (B=64)?
Runs the CPU oracle and the CUDA path for loss-term subsets and prints, per parameter group, the worst norm ratio / cosine."""
import json, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import torch
from hulc2_b200 import noise, ops
from hulc2_b200._compat import instantiate
from hulc2_b200.config import hulc2_config
from hulc2_b200.synthetic import synthetic_batch
from oracle import hulc2_oracle as O
from helpers import oracle_params, to_device

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.set_num_threads(os.cpu_count() or 1)
res = {}
for tag, kw, aux, prec in [("clip_bf16", {"lvl": 0}, "half", "bf16"), ("clip_fp32_proj", {"lvl": 1}, "half", "bf16"),
                           ("clip_fp32_proj_fc", {"lvl": 2}, "half", "bf16"), ("clip_fp32_proj_fc_all", {"lvl": 2}, "all", "bf16"),
                           ("noclip", {"clip": 0.0, "lvl": 2}, "half", "bf16")]:
    ops.clip_fp32, ops.clip_fp32_level = kw["lvl"] > 0, kw["lvl"]
    torch.manual_seed(0)
    m = instantiate(hulc2_config(dropout_p=0.0))
    cfg = hulc2_config(pkg="x", dropout_p=0.0)
    if "clip" in kw:
        cfg["clip_auxiliary_loss_beta"] = kw["clip"]; m.clip_auxiliary_loss_beta = kw["clip"]
    if "kl" in kw:
        cfg["kl_beta"] = kw["kl"]; m.kl_beta = kw["kl"]
    P = oracle_params(m)
    batch = synthetic_batch(B, seed=1, aux=aux)
    idx = {mod: torch.randint(0, 32, (B, 32), generator=torch.Generator().manual_seed(5)) for mod in batch}
    out = O.training_step(batch, {mod: {"plan_idx": idx[mod]} for mod in batch}, P, cfg)
    out["loss"].backward()
    m = m.to("cuda").train()
    ops.set_precision(prec)
    with noise.supplied(categories=[idx[mod] for mod in batch]):
        loss = m.training_step(to_device(batch, "cuda"), 0)
    loss.backward()
    torch.cuda.synchronize()
    groups = {}
    for n, p in m.named_parameters():
        r = P[n].grad
        if r is None or p.grad is None:
            continue
        g = p.grad.cpu()
        ratio = float(g.norm() / (r.norm() + 1e-30)); cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        grp = ".".join(n.split(".")[:3]) if n.startswith(("perceptual", "plan_recognition.transformer")) else n.split(".")[0]
        w = groups.setdefault(grp, [0.0, 1.0])
        w[0] = max(w[0], abs(ratio - 1)); w[1] = min(w[1], cos)
        if abs(ratio - 1) > 0.02 or cos < 0.993:
            print(f"      ! {n:70s} ratio {ratio:.4f} cos {cos:.5f} |ref| {float(r.norm()):.3e}")
    res[tag] = {"loss": float(loss), "oracle": float(out["loss"]), "groups": groups}
    print(f"== {tag}: loss {float(loss):.5f} oracle {float(out['loss']):.5f}")
    for gname, (dr, c) in sorted(groups.items()):
        print(f"   {gname:60s} |ratio-1| {dr:.4f}  min cos {c:.5f}")
    del m
    ops.set_precision("fp32")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "diag_bf16_b64.json"), "w"), indent=1)
