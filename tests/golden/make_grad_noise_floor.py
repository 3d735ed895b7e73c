"""Generates tests/golden/grad_noise_floor.npz: the fp32 rounding-noise floor of the REFERENCE's own gradients.

Run in the build container only (needs /root/reference):
    python tests/golden/make_grad_noise_floor.py
The unmodified reference (oracle/ref_import.py shims) runs the calvin_B2 fixture case (same weights / batch / plan draw as
tests/golden/make_golden.py) twice: in fp32 (the arithmetic the 1e-5 parity target is quoted in) and in fp64.  Per
parameter the file holds the fp64 gradient norm, the fp64 gradient at a seeded sample of <= 4096 positions (full tensor
when small), and the reference's OWN fp32-vs-fp64 error on the norm and element-wise.  tests/test_gpu_step.py compares
the CUDA gradients with the fp64 values and reports both errors side by side (profiles/parity_r02_grad_noise_floor.md):
a CUDA error at the level of the reference's own fp32 error is summation-order noise, not a defect.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")

from helpers import grad_sample_index as sample_index  # noqa: E402
from make_golden import build  # noqa: E402
from hulc2_b200.synthetic import synthetic_batch  # noqa: E402
from oracle.ref_noise import supplied_categories  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grad_noise_floor.npz")


def run(dtype):
    m = build("calvin", (200, 200))
    batch = synthetic_batch(2, seed=1, static_hw=(200, 200), aux="half")
    g = torch.Generator().manual_seed(5)
    idx = {mod: torch.randint(0, 32, (2, 32), generator=g) for mod in batch}
    if dtype == torch.float64:
        m = m.double()

        def cast(x, key=""):
            # actions / state_info stay fp32: the tcp-frame transform forces fp32 itself (gripper_control.py:16-36) and carries no gradient
            if isinstance(x, dict):
                return {k: (v if k in ("actions", "state_info") else cast(v, k)) for k, v in x.items()}
            return x.double() if isinstance(x, torch.Tensor) and x.dtype == torch.float32 else x

        batch = cast(batch)
    with supplied_categories([idx[mod] for mod in batch]):
        loss = m.training_step(batch, 0)
    loss.backward()
    return float(loss), {n: p.grad.detach().double() for n, p in m.named_parameters() if p.grad is not None}


def main():
    l32, g32 = run(torch.float32)
    l64, g64 = run(torch.float64)
    out = {"loss64": np.float64(l64), "loss32_rel_err": np.float64(abs(l32 - l64) / abs(l64))}
    rows = []
    for n, r in g64.items():
        a = g32[n]
        ix = sample_index(n, r.numel())
        scale = float(r.abs().max()) + 1e-300
        out[f"gnorm64/{n}"] = np.float64(r.norm())
        out[f"gmax64/{n}"] = np.float64(scale)
        out[f"sample64/{n}"] = r.reshape(-1)[torch.from_numpy(ix)].numpy()
        e_norm = abs(float(a.norm()) - float(r.norm())) / (float(r.norm()) + 1e-300)
        e_elem = float((a - r).abs().max()) / scale
        e_samp = float((a.reshape(-1)[torch.from_numpy(ix)] - r.reshape(-1)[torch.from_numpy(ix)]).abs().max()) / scale
        out[f"ref32_err/{n}"] = np.asarray([e_norm, e_elem, e_samp], dtype=np.float64)
        rows.append((n, e_norm, e_elem))
    np.savez_compressed(OUT, **out)
    rows.sort(key=lambda t: -t[2])
    print(f"loss fp32 vs fp64: {out['loss32_rel_err']:.2e}; {len(rows)} parameters")
    for n, en, ee in rows[:12]:
        print(f"  {n:70s} norm {en:.2e}  elem {ee:.2e}")
    en = sorted(r[1] for r in rows)
    ee = sorted(r[2] for r in rows)
    print("norm err median/p90/max:", en[len(en) // 2], en[int(0.9 * len(en))], en[-1])
    print("elem err median/p90/max:", ee[len(ee) // 2], ee[int(0.9 * len(ee))], ee[-1])


if __name__ == "__main__":
    main()
