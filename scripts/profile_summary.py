import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
import torch
pkg = graft.load_package()
from bayes_js_b200.summary import CudaBlockReducer, RadixSelect, quantile_targets
rows, entries, chains = 100, 2, 1 << 20
dev = torch.device("cuda", 0)
block = torch.empty((rows, entries, chains), dtype=torch.float64, device=dev)
block[:, 0] = 184.3 + 0.14 * torch.randn((rows, chains), dtype=torch.float64, device=dev)
block[:, 1] = 4.5 + 0.1 * torch.randn((rows, chains), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
red = CudaBlockReducer(0)
def T(f, n=3):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
gb = rows * entries * chains * 8 / 1e9
ms = T(lambda: red.moments(block))[0]
print("moments: %.3f ms incl. launch/alloc/D2H; reads the block twice: %.0f GB/s" % (ms, 2 * gb / ms * 1e3))
ranks, plan = quantile_targets(rows * chains, (0.025, 0.25, 0.5, 0.75, 0.975))
sel = RadixSelect(entries, ranks)
for p in range(8):
    table, which = sel.prefixes()
    ms, counts = T(lambda: red.digit_counts(block, p, table))
    t = time.perf_counter(); sel.advance(counts.cpu().numpy(), which); host = (time.perf_counter() - t) * 1e3
    print("pass", p, "n_prefix", table.shape[1], "kernel+launch ms %.3f (%.0f GB/s)" % (ms, gb / ms * 1e3), "host advance ms %.3f" % host)
