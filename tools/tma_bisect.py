import os, sys, subprocess
sys.path.insert(0, "/root/repo")
case = os.environ.get("CASE")
if case is None:
    for c in ["3,224,224,64,7,2,3", "3,208,208,64,7,2,3", "3,224,112,64,7,2,3", "3,112,224,64,7,2,3", "64,56,56,64,3,1,1", "64,112,112,64,3,1,1", "64,224,224,64,3,1,1"]:
        r = subprocess.run([sys.executable, __file__], env=dict(os.environ, CASE=c), capture_output=True, text=True)
        print(c, "->", (r.stdout.strip().splitlines() or ["?"])[-1], "|", (r.stderr.strip().splitlines() or [""])[-1][:100], flush=True)
    sys.exit(0)
import numpy as np, torch
from oracle import bsr_oracle as O, c_oracle
from resnet_accel_b200 import ops
Cin, H, W, Cout, k, s, p = map(int, case.split(","))
rng = np.random.default_rng(1)
K = Cin * k * k
Wm = rng.integers(-128, 128, (Cout, K), dtype=np.int8)
nbr, nbc = -(-Cout // 14), -(-K // 14)
keep = rng.random((nbr, nbc)) < 0.5
Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :K].astype(np.int8)
bsr = O.build_bsr_14x14_int8_direct(Wm)
x = rng.integers(-128, 128, (2, Cin, H, W), dtype=np.int8)
plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
xd = ops.alloc_padded(x.shape); xd.copy_(torch.from_numpy(x).cuda())
out = plan.conv(xd, k, s, p, Cout, out_kind="i32").cpu().numpy()
ref, _ = O.conv2d_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, k, s, p)
print("match" if np.array_equal(out, ref) else "MISMATCH")
