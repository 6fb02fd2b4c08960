"""Timeline trace of CTA 0 of fused_tc2_kernel on the bench's layer shapes (development aid).

Build the instrumented library next to the product one and run on the GPU box:
    KAGNN_LIB_SUFFIX=_trace KAGNN_NVCC_EXTRA="-DKAGNN_TRACE=1" python -m kagnn_b200.build
    KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_trace.so python scripts/trace_tc2.py [layer0|layer1|layout]
Events (clock64 of SM of CTA 0): producers (role 0/1 = warp 0 of each warpgroup) per chunk: 0/1 wait xs_full begin/end,
2/3 wait empty begin/end, 4 a_full arrived; role 2 MMA per chunk: 0 begin, 1 a_full ok, 2 b_full ok, 3 issued+committed;
role 3 loader: 0/1 wait empty; roles 4/5 per layer: 0/1 acc wait (layer>0), 2/3 acc wait (epilogue), 4 epilogue done;
role 6 gather warp 0 per unit: 0/1 wait xs_empty, 2 unit done."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import kagnn_b200 as kb
from kagnn_b200 import _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "layer0"
torch.manual_seed(0)
n, e = 169_343, 1_166_243
dev = torch.device("cuda")
lib = L.lib()
lib.kagnn_debug_set_trace.restype = C.c_int
lib.kagnn_debug_set_trace.argtypes = [C.c_void_p]
buf = torch.zeros(8 * 512 * 8, dtype=torch.int64, device=dev)

if which == "layout":
    mod = kb.KANLinear(320, 40, grid_size=5, spline_order=3).to(dev)
    x = torch.randn(n, 320, device=dev) * 0.5
    run = lambda: mod(x)
    nch = [45]
else:
    f = 128 if which == "layer0" else 64
    mod = kb.GIKANLayer(f, 64, 5, 3, 64, 2).to(dev)
    x = torch.randn(n, f, device=dev) * 0.3
    ei = torch.randint(0, n, (2, e), device=dev)
    run = lambda: mod(x, ei)
    nch = [f // 8 + (f + 63) // 64, 9]
with torch.no_grad():
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    assert lib.kagnn_debug_set_trace(C.c_void_p(buf.data_ptr())) == 0
    run()
    torch.cuda.synchronize()
    lib.kagnn_debug_set_trace(None)
t = buf.cpu().numpy().reshape(8, 512, 8).astype(np.int64)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
per_tile = sum(nch)
ntiles = int((rel[2, :, 3] >= 0).sum()) // per_tile
print(f"{which}: chunks/tile {per_tile} (layers {nch}), tiles of CTA 0: {ntiles}, total cycles {rel.max()}")
for tile in range(min(ntiles, 3)):
    base = tile * per_tile
    print(f"--- tile {tile}: MMA-issued span {rel[2, base, 0]} .. {rel[2, base + per_tile - 1, 3]}")
    print(" chunk |   P0: xs_wait  empty_wait   work  |   P1: xs_wait empty_wait work  | MMA: wait_a wait_b issue | loader wait | t(MMA done issue)")
    for q in range(per_tile):
        k = base + q
        def d(r, a, b):
            return int(rel[r, k, b] - rel[r, k, a]) if rel[r, k, a] >= 0 and rel[r, k, b] >= 0 else 0
        print(f" {q:5d} | [cmp {d(0,3,5)} st {d(0,5,6)} stwait {d(0,6,7)} arr {d(0,7,4)}] {d(0,0,1):8d} {d(0,2,3):8d} {d(0,3,4):8d} | {d(1,0,1):8d} {d(1,2,3):8d} {d(1,3,4):8d} |"
              f" {d(2,0,2):8d} {0:6d} {d(2,2,3):6d} | {d(3,0,1):8d} | {int(rel[2,k,3])}"
              f" | A-ready {int(rel[0,k,4])} W-issued {int(rel[3,k,1])} MMA-go {int(rel[2,k,2])}  (go-A {int(rel[2,k,2]-rel[0,k,4])}, go-Wissue {int(rel[2,k,2]-rel[3,k,1])})")
    for l in range(len(nch)):
        lc = tile * len(nch) + l
        for r in (4, 5):
            ev = rel[r, lc]
            print(f"   layer {l} wg{r-4}: acc-wait(l>0) {int(ev[1]-ev[0]) if ev[0] >= 0 else 0}")
    lc = tile * len(nch) + len(nch)
    for r in (4, 5):
        ev = rel[r, lc]
        if ev[2] >= 0:
            print(f"   epilogue wg{r-4}: start {int(ev[2])} acc-wait {int(ev[3]-ev[2])} body {int(ev[4]-ev[3])}")
print("--- gather warp 0 per unit: wait_xs_empty, work, t_done")
for u in range(min(int((rel[6, :, 2] >= 0).sum()), 12)):
    print(f"   unit {u}: {int(rel[6,u,1]-rel[6,u,0])} {int(rel[6,u,2]-rel[6,u,1])} {int(rel[6,u,2])}")
# aggregate over all tiles
P = rel[0]
valid = P[:, 4] >= 0
print("producer wg0 totals over CTA 0: xs_wait", int((P[valid, 1] - P[valid, 0])[P[valid, 0] >= 0].sum()), "empty_wait",
      int((P[valid, 3] - P[valid, 2]).sum()), "work", int((P[valid, 4] - P[valid, 3]).sum()))
M = rel[2]
valid = M[:, 3] >= 0
print("MMA totals: wait", int((M[valid, 2] - M[valid, 0]).sum()), "issue",
      int((M[valid, 3] - M[valid, 2]).sum()))
# tile-level view (valid in the light trace build, KAGNN_TRACE=1)
print("--- tile-level (wg0 warp 0): t_start_wait_x, x_wait, [acc waits per later layer], epi_acc_wait, epi_body, t_end | gather unit: wait, work")
nl = len(nch)
prev_end = 0
upt = 3 if which == "layout" else 1
for tile in range(int((rel[4, :, 4] >= 0).sum())):
    k0 = tile * per_tile
    xs0, xs1 = int(rel[0, k0, 0]), int(rel[0, k0, 1])
    accw = [int(rel[4, tile * nl + l, 1] - rel[4, tile * nl + l, 0]) for l in range(1, nl)]
    e = rel[4, tile * nl + nl]
    g = rel[6, tile * upt]
    print(f" tile {tile}: start {xs0} (+{xs0 - prev_end} after prev end) x_wait {xs1 - xs0} acc_waits {accw} epi_wait {int(e[3]-e[2])} epi_body {int(e[4]-e[3])} end {int(e[4])}"
          f" | gather wait {int(g[1]-g[0])} work {int(g[2]-g[1])} done_at {int(g[2])}")
    prev_end = int(e[4])
print("--- gather warp 0, second tile of CTA 0, per sub-batch of U loads: t_start, addresses+loads issued (+), first data used (+), consumed (+)")
for b in range(12):
    r = rel[7, b]
    if r[0] < 0:
        break
    print(f"   sub-batch {b}: start {int(r[0])} issued +{int(r[1]-r[0])} first-data +{int(r[2]-r[1])} consumed +{int(r[3]-r[2])}")
