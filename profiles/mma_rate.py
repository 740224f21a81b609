"""tcgen05 TF32 MMA cost per instruction shape (128 x N x 8) in this library's no-swizzle operand layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crank_b200 import lib as L

buf = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
print("N   K   split grid | MMAs  cycles/MMA (exec)  cycles/MMA (issue loop)   [data: zeros, then random]")
for grid in (148,):
    for split in (1, 9, 0, 8):        # bit 0: 3xTF32 passes, bit 1: A from tensor memory, bit 2: commit per group, bit 3: warp-collective issue
        for N, K in ((256, 64), (192, 64), (128, 64), (64, 64), (64, 32), (16, 64)):
            reps = 40
            nmma = reps * (3 if split & 1 else 1) * (min(K, 64) if split & 2 else K) // 8
            out = []
            for sign in (1, -1):
                for _ in range(2):
                    buf.zero_()
                    L.call("crk_tc_mma_rate", N, K, sign * reps, split, grid, buf.data_ptr())
                    torch.cuda.synchronize()
                c = buf.view(-1, 2)[:grid].double()
                out.append(f"{c[:, 0].mean().item() / nmma:10.1f} {c[:, 1].mean().item() / nmma:10.1f}")
            print(f"{N:3d} {K:3d} {split:5d} {grid:4d} | {nmma:5d} " + "   |".join(out))
