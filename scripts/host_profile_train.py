"""Where the HOST time of one training step goes (cProfile over 30 steps of the arxiv-shaped model, after warm-up).

    python scripts/host_profile_train.py > gpurun_out/host_profile.txt
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb


def main():
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(12345)
    n, e, f, c = 169_343, 1_166_243, 128, 40
    ei = torch.randint(0, n, (2, e), generator=gen).to(dev)
    x = (torch.randn(n, f, generator=gen) * 0.3).to(dev)
    y = torch.randint(0, c, (n,), generator=gen).to(dev)
    torch.manual_seed(0)
    model = kb.GKAN_Nodes("gin", 3, f, 64, c, skip=True, grid_size=5, spline_order=3, hidden_layers=2, dropout=0.0).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        model.train()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x, ei), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0) / 30:.3f} ms/step, with drain {1e3 * (t2 - t0) / 30:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(30):
        step()
    pr.disable()
    torch.cuda.synchronize()
    for key in ("tottime", "cumtime"):
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(45)
        print(buf.getvalue()[:9000])


if __name__ == "__main__":
    main()
