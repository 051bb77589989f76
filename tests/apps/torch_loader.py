"""A PyTorch process in its loading phase, then one computation.  Run under
libnvshare.so beside a client that holds the GPU: the uploads and the read-back
must complete without the lock (prints LOAD_DONE <seconds>), the computation
afterwards needs it (prints RESULT ...)."""
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20      # fp32 elements per tensor
torch.cuda.init()
t0 = time.time()
h1 = torch.arange(n, dtype=torch.float32)                      # pageable
h2 = (torch.arange(n, dtype=torch.float32) % 1024).pin_memory()
d1 = torch.empty(n, device="cuda")                             # cudaMalloc: no kernel
d2 = torch.empty(n, device="cuda")
d1.copy_(h1)                                                   # cudaMemcpyAsync H2D (pageable) + sync
d2.copy_(h2, non_blocking=True)                                # cudaMemcpyAsync H2D (pinned)
torch.cuda.current_stream().synchronize()
back = torch.empty(n, dtype=torch.float32).pin_memory()
back.copy_(d2, non_blocking=True)                              # D2H
torch.cuda.current_stream().synchronize()
ok_load = bool(torch.equal(back, h2)) and bool(torch.equal(d1.cpu(), h1))
print("LOAD_DONE %.3f %s" % (time.time() - t0, ok_load), flush=True)
t1 = time.time()
s = (d1 + d2).double().sum().item()                            # first kernel: waits for the GPU lock
want = h1.double().sum().item() + h2.double().sum().item()
print("RESULT %s waited=%.3f" % ("PASS" if ok_load and s == want else "FAIL", time.time() - t1), flush=True)
