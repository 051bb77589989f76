"""Small model workloads for BASELINE configs #4/#5 as PARITY cases (not bench
lines): the reference ships no ResNet / Llama tests, so these are authored here
(SURVEY 8d).  Each runs a fixed, seeded computation and prints one JSON line with
the numbers to compare between an un-hooked run and runs under libnvshare.so
(tolerance from north_star: 1e-5 relative).

  resnet   torchvision ResNet-50, synthetic 3x224x224 batch, SGD training steps -> losses
  llama    a small Llama-architecture decoder (transformers, random init), greedy decode -> logits checksum

Plain PyTorch on purpose: this is the unmodified-application side of the boundary.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time


def run(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["resnet", "llama"], required=True)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=0.0, help="keep repeating the computation for this long")
    args = ap.parse_args(argv)

    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")
    import torch
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.use_deterministic_algorithms(True, warn_only=True)
    dev = torch.device("cuda")
    out = {"kind": args.kind}
    t0 = time.time()
    rounds = 0
    while True:
        torch.manual_seed(1234)
        if args.kind == "resnet":
            import torchvision
            model = torchvision.models.resnet50(weights=None).to(dev).train()
            opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
            g = torch.Generator(device="cpu").manual_seed(7)
            x = torch.randn(args.batch, 3, 224, 224, generator=g).to(dev)
            y = torch.randint(0, 1000, (args.batch,), generator=g).to(dev)
            losses = []
            for _ in range(args.steps):
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.cross_entropy(model(x), y)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))
            out["losses"] = losses
            out["param_checksum"] = float(sum(p.double().sum().item() for p in model.parameters()))
        else:
            from transformers import LlamaConfig, LlamaForCausalLM
            cfg = LlamaConfig(vocab_size=4096, hidden_size=512, intermediate_size=1376, num_hidden_layers=6,
                              num_attention_heads=8, num_key_value_heads=8, max_position_embeddings=512)
            model = LlamaForCausalLM(cfg).to(dev).eval()
            g = torch.Generator(device="cpu").manual_seed(11)
            ids = torch.randint(0, 4096, (args.batch, 32), generator=g).to(dev)
            with torch.no_grad():
                for _ in range(args.steps):          # greedy decode, full re-forward each step (no cache: simple and exact)
                    logits = model(ids).logits[:, -1, :]
                    ids = torch.cat([ids, logits.argmax(-1, keepdim=True)], dim=1)
            out["last_logits_sum"] = float(logits.double().sum().item())
            out["last_logits_absmax"] = float(logits.abs().max().item())
            out["tokens"] = ids[:, -args.steps:].tolist()
        torch.cuda.synchronize()
        rounds += 1
        if args.seconds <= 0 or time.time() - t0 >= args.seconds:
            break
    out["rounds"] = rounds
    out["seconds"] = time.time() - t0
    print("RESULT " + json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(run())
