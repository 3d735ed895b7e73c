import sys, torch
sys.path.insert(0, "/root/repo")
from hulc2_b200 import ops
ops.set_precision("bf16")
g = torch.Generator(device="cuda").manual_seed(0)
u8 = torch.randint(0, 256, (2048, 200, 200, 3), generator=g, device="cuda", dtype=torch.uint8)
sh = torch.randint(-10, 11, (2048, 2), generator=g, device="cuda", dtype=torch.int32)
fr = ops.U8Frames(u8, sh, S=32)
for _ in range(3):
    xs = ops.pack_frames(fr)
torch.cuda.synchronize()
print(xs.shape)
