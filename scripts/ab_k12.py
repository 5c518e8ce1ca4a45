"""A/B of K1 (linearize) / K2 (assemble) between two builds of the library (SRBD_LIB=... selects the build):
dumps the raw packed device records (BAbt, RSQrq, DCt, d, dmask, defect) of a seeded batch in both assemble modes to
an .npz and times the two kernels with CUDA events.  `python scripts/ab_k12.py cmp a.npz b.npz` compares two dumps
bit for bit."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if len(sys.argv) > 1 and sys.argv[1] == "cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        same = a[k].shape == b[k].shape and (a[k].view(np.uint64) == b[k].view(np.uint64)).all()
        if not same:
            d = a[k] != b[k]
            print("DIFF", k, int(d.sum()), "values differ; max abs", float(np.abs(a[k] - b[k]).max()))
            bad += 1
    print("bitwise identical" if not bad else "NOT identical", "(%d arrays)" % len(a.files))
    sys.exit(1 if bad else 0)

import torch
import srbd_pkg
pkg = srbd_pkg.load()
out = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = 20
w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait")
rng = np.random.default_rng(5)
w["x"] = w["x"] + 0.05 * rng.standard_normal(w["x"].shape)  # a trajectory with a defect (general linearization point)
w["u"] = w["u"] + 2.0 * rng.standard_normal(w["u"].shape)
ctx = pkg.Context(B)
ctx.set_model(pkg.default_model_params(N))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
C = pkg.capi
dump = {}
for mode, name in ((C.SRBD_HARD_INEQ, "hard"), (C.SRBD_BARRIER_SOFT, "soft")):
    ctx.linearize(); ctx.assemble(mode); ctx.sync()
    # buffer ids of include/srbd_b200.h (SRBD_BUF_BABT .. SRBD_BUF_DEFECT)
    for buf, bn in ((13, "babt"), (14, "rsq"), (15, "dct"), (16, "d"), (17, "dmask"), (18, "defect"), (19, "stage_rec")):
        n = 512  # QPs kept in the dump
        t = ctx.device_tensor(buf)
        dump[name + "_" + bn] = t.reshape(B, -1)[:n].cpu().numpy().copy()
np.savez(out, **dump)
# timing on the context's stream
st = torch.cuda.ExternalStream(ctx.stream)
res = {}
with torch.cuda.stream(st):
    for fn, nm in ((ctx.linearize, "K1"), (lambda: ctx.assemble(C.SRBD_HARD_INEQ), "K2")):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[nm] = float(np.median(ts))
gb1 = B * N * (336 + 12) * 8 / 1e9
gb2 = B * (N + 1) * (672 + 576 + 96 + 192) * 8 / 1e9  # RSQrq, DCt, d, dmask, compact stage record
print("B=%d  K1 %.3f ms (%.0f GB/s written)  K2 %.3f ms (%.0f GB/s written)  lib=%s" % (
    B, res["K1"], gb1 / res["K1"] * 1e3, res["K2"], gb2 / res["K2"] * 1e3, os.environ.get("SRBD_LIB", "in-tree")))
