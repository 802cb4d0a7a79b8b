"""How fast can a B200 gather 8-byte values at uniformly random indices from an L2-resident vector?
(The bound of the power-law SpMV, config 5b: one such gather per stored entry.)  Uses torch's own
index kernels as an independent probe: out[i] = x[idx[i]], idx int32/int64, x of 2^21 / 2^22 doubles.
    python scripts/gather_probe.py"""
import torch

dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
for logn in (21, 22):
    n = 1 << logn
    m = 30 * n                                   # entries of the power-law matrix of that order
    x = torch.randn(n, dtype=torch.float64, device=dev)
    idx = torch.randint(0, n, (m,), device=dev, generator=g)
    srt = torch.sort(idx).values
    out = torch.empty(m, dtype=torch.float64, device=dev)
    for name, ii in (("random", idx), ("sorted", srt)):
        for _ in range(3):
            torch.index_select(x, 0, ii, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10):
            torch.index_select(x, 0, ii, out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        print(f"x: 2^{logn} doubles ({8 * n >> 20} MB), {m} gathers, {name} indices: {us:.0f} us per pass = "
              f"{m / us / 1e3:.1f} G gathers/s; streamed bytes (8 B index + 8 B out) {16 * m / us / 1e3:.0f} GB/s", flush=True)
