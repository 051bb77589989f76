"""An unmodified PyTorch application whose steady state is a replayed CUDA graph
(torch.cuda.graphs): capture once, then cuGraphLaunch in a loop.  The reference does not gate
cuGraphLaunch (src/hook.c:545-577); here every replay waits for the GPU lock, and the memory the
graph's nodes point at is unmapped and re-mapped between replays (forced swaps).
usage: graph_app.py <n> <seconds>      prints RESULT PASS|FAIL replays=<k>"""
import sys
import time

import torch

n, seconds = int(sys.argv[1]), float(sys.argv[2])
dev = torch.device("cuda")
# integers only, and small enough that 20000 accumulations stay exact in fp32 (251 * 2 * 20000 < 2**24)
x = (torch.arange(n * n, device=dev, dtype=torch.int64) % 251).to(torch.float32).reshape(n, n)
acc = torch.zeros(n, n, device=dev)
w = torch.full((n, n), 2.0, device=dev)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):                      # warm-up on a side stream, as the PyTorch docs prescribe
    for _ in range(3):
        acc.add_(x * w)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
acc.zero_()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):                       # the hook must not synchronise inside the capture
    acc.add_(x * w)                             # exact: integers below 2**24
torch.cuda.synchronize()
acc.zero_()
t0, k = time.time(), 0
while time.time() - t0 < seconds:
    g.replay()
    k += 1
    if k % 8 == 0:
        torch.cuda.synchronize()
    if k >= 20000:
        break
torch.cuda.synchronize()
want = x * (2.0 * k)
bad = int((acc != want).sum().item()) + int((w != 2.0).sum().item())
print(f"RESULT {'PASS' if bad == 0 else 'FAIL'} replays={k} mismatches={bad}", flush=True)
sys.exit(0 if bad == 0 else 1)
