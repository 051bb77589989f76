"""Model workloads for BASELINE configs #4/#5.  The reference ships no ResNet /
Llama tests (its tests/ hold pytorch-add and tf-matmul only), so these are
authored here as plain PyTorch applications -- the unmodified side of
the LD_PRELOAD boundary; nothing in here knows about the swap engine.

  resnet   torchvision ResNet-50, synthetic 3x224x224 batch, SGD-momentum training
           steps (cuDNN convolutions, batch-norm, memsets)            -> per-step losses
  llama    a Llama-architecture decoder (transformers LlamaForCausalLM, random
           init; --size 7b = the 7B geometry: 32 layers, hidden 4096, 32 heads, FFN
           11008, vocab 32000), real prefill of a synthetic prompt into a static KV
           cache, then greedy decode                                   -> token ids, logit sums

One ROUND is a fixed, seeded computation (K steps from the same initial state), so
every round must reproduce the numbers of every other round -- and of the same
script run un-hooked: that is the parity check (north_star: 1e-5 relative).  A
golden file written by the un-hooked run (--write-golden) is compared against
after every round of the hooked clients (--golden); any deviation fails the run.

Footprint: --target-bytes pads the process with a position-dependent ballast tensor
(verified bit-exact at the end) up to the requested device footprint, which is how a
ResNet-50 step reaches ~1.0x HBM per client ("oversubscription reached with
batch/ballast"); the Llama footprint is weights + KV cache (batch x context).

Every step logs one {"event": "iter"} line with its completion time -- the same
format as nvshare_b200/workloads.py, so nvshare_b200/harness.py analyses hand-offs of
both in the same way.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time


def _rel_close(a, b, tol):
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-30)


def run(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["resnet", "llama"], required=True)
    ap.add_argument("--size", choices=["small", "7b"], default="small", help="llama geometry")
    ap.add_argument("--steps", type=int, default=6, help="steps per round")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--context", type=int, default=32, help="llama: prompt length that is prefilled into the KV cache")
    ap.add_argument("--tf32", type=int, default=0, help="allow TF32 matmuls/convolutions (the large configurations)")
    ap.add_argument("--seconds", type=float, default=0.0, help="keep repeating rounds for this long")
    ap.add_argument("--rounds", type=int, default=0, help="stop after this many rounds (0 = by time / stop file)")
    ap.add_argument("--target-bytes", type=int, default=0, help="pad the device footprint up to this with a verified ballast")
    ap.add_argument("--log", default="")
    ap.add_argument("--tag", default="client")
    ap.add_argument("--start-barrier", default="")
    ap.add_argument("--stop-file", default="")
    ap.add_argument("--write-golden", default="", help="write this round's numbers here (the un-hooked run)")
    ap.add_argument("--golden", default="", help="compare every round against this file")
    args = ap.parse_args(argv)

    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")
    import torch
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    torch.use_deterministic_algorithms(True, warn_only=True)
    dev = torch.device("cuda")
    t_start = time.time()
    log = open(args.log, "w", buffering=1) if args.log else None

    def emit(rec):
        rec["tag"] = args.tag
        if log:
            log.write(json.dumps(rec) + "\n")

    golden = json.load(open(args.golden)) if args.golden else None
    n_iter = 0

    # ------------------------------------------------------------- model state
    if args.kind == "resnet":
        import torchvision
        model = torchvision.models.resnet50(weights=None).to(dev).train()
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
        g = torch.Generator(device="cpu").manual_seed(7)
        x = torch.randn(args.batch, 3, 224, 224, generator=g).to(dev)
        y = torch.randint(0, 1000, (args.batch,), generator=g).to(dev)
        # a pristine copy of everything a step changes, so that each round starts from the same state
        init = [t.detach().clone() for t in list(model.parameters()) + list(model.buffers())]

        def reset():
            with torch.no_grad():
                for t, s in zip(list(model.parameters()) + list(model.buffers()), init):
                    t.copy_(s)
            opt.state.clear()

        def one_round():
            nonlocal n_iter
            reset()
            losses = []
            for _ in range(args.steps):
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.cross_entropy(model(x), y)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))        # synchronises
                n_iter += 1
                emit({"event": "iter", "i": n_iter, "t": time.time()})
            return {"losses": losses,
                    "param_checksum": float(sum(p.double().sum().item() for p in model.parameters()))}
    else:
        from transformers import LlamaConfig, LlamaForCausalLM, StaticCache
        if args.size == "7b":
            cfg = LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                              num_attention_heads=32, num_key_value_heads=32,
                              max_position_embeddings=max(4096, args.context + args.steps + 8))
        else:
            cfg = LlamaConfig(vocab_size=4096, hidden_size=512, intermediate_size=1376, num_hidden_layers=6,
                              num_attention_heads=8, num_key_value_heads=8,
                              max_position_embeddings=max(512, args.context + args.steps + 8))
        with torch.device(dev):                           # random init straight on the GPU (7B: 27 GB of fp32)
            model = LlamaForCausalLM(cfg)
        model.eval()
        g = torch.Generator(device="cpu").manual_seed(11)
        prompt = torch.randint(0, cfg.vocab_size, (args.batch, args.context), generator=g).to(dev)
        cache = StaticCache(config=cfg, max_cache_len=args.context + args.steps + 1)
        with torch.no_grad():                              # real prefill, in pieces (no [B, S, vocab] logits)
            piece = max(1, min(args.context, 8192 // max(1, args.batch)))
            for s0 in range(0, args.context, piece):
                s1 = min(args.context, s0 + piece)
                pos = torch.arange(s0, s1, device=dev).unsqueeze(0).expand(args.batch, -1)
                out = model(input_ids=prompt[:, s0:s1], past_key_values=cache, use_cache=True,
                            position_ids=pos, logits_to_keep=1)
            first = out.logits[:, -1, :].argmax(-1, keepdim=True)
        torch.cuda.synchronize()

        def one_round():
            nonlocal n_iter
            toks, sums = [], []
            nxt = first
            with torch.no_grad():
                for layer in cache.layers:                 # rewind the cache to the end of the prompt
                    layer.cumulative_length.fill_(args.context)
                for k in range(args.steps):                # decode: positions context .. context+steps-1, overwritten every round
                    pos = torch.full((args.batch, 1), args.context + k, device=dev)
                    out = model(input_ids=nxt, past_key_values=cache, use_cache=True, position_ids=pos)
                    logits = out.logits[:, -1, :]
                    nxt = logits.argmax(-1, keepdim=True)
                    sums.append(float(logits.double().sum().item()))   # synchronises
                    toks.append(nxt.flatten().tolist())
                    n_iter += 1
                    emit({"event": "iter", "i": n_iter, "t": time.time()})
            return {"tokens": toks, "logit_sums": sums}

    def check(res):
        """-> number of deviations from the golden (un-hooked) round, 1e-5 relative."""
        if golden is None:
            return 0
        bad = 0
        for k, v in golden["round"].items():
            got = res[k]
            if k == "tokens":
                bad += int(got != v)
            elif isinstance(v, list):
                bad += sum(0 if _rel_close(a, b, 1e-5) else 1 for a, b in zip(got, v)) + abs(len(got) - len(v))
            else:
                bad += 0 if _rel_close(got, v, 1e-5) else 1
        return bad

    # ----------------------------------------------------- warm round + ballast
    first_round = one_round()                              # also sizes the allocator's pools
    torch.cuda.synchronize()
    ballast = None
    if args.target_bytes > 0:
        held = torch.cuda.memory_reserved()
        room = args.target_bytes - held
        if room > (64 << 20):
            bn = room // 4
            ballast = torch.empty([bn], dtype=torch.float32, device=dev)
            blk = 1 << 26
            for o in range(0, bn, blk):
                m = min(blk, bn - o)
                ballast[o:o + m].copy_(((torch.arange(o, o + m, device=dev, dtype=torch.int64) * 2654435761)
                                        & ((1 << 22) - 1)).to(torch.float32))
    torch.cuda.synchronize()
    emit({"event": "setup_done", "t": time.time(), "setup_s": time.time() - t_start, "kind": args.kind,
          "torch_allocated": torch.cuda.memory_allocated(), "torch_reserved": torch.cuda.memory_reserved()})
    if args.start_barrier:
        while not os.path.exists(args.start_barrier):
            time.sleep(0.01)

    # ------------------------------------------------------------------ rounds
    bad = check(first_round)
    rounds = 1
    t_loop = time.time()
    res = first_round
    while True:
        if args.rounds and rounds >= args.rounds:
            break
        if not args.rounds and args.seconds <= 0 and not args.stop_file:
            break
        if args.seconds > 0 and time.time() - t_loop >= args.seconds:
            break
        if args.stop_file and os.path.exists(args.stop_file):
            break
        res = one_round()
        bad += check(res)
        rounds += 1
    t_end = time.time()

    if ballast is not None:                                # the padding is data too: it must have survived every hand-off
        bn = ballast.numel()
        blk = 1 << 26
        for o in range(0, bn, blk):
            m = min(blk, bn - o)
            exp = ((torch.arange(o, o + m, device=dev, dtype=torch.int64) * 2654435761) & ((1 << 22) - 1)).to(torch.float32)
            bad += int((ballast[o:o + m] != exp).sum().item())
    torch.cuda.synchronize()
    if args.write_golden:
        json.dump({"kind": args.kind, "round": res, "args": vars(args)}, open(args.write_golden, "w"))
    summary = {"event": "summary", "iters": n_iter, "rounds": rounds, "loop_s": t_end - t_loop, "total_s": time.time() - t_start,
               "mismatches": bad, "result": "PASS" if bad == 0 else "FAIL",
               "torch_reserved": torch.cuda.memory_reserved(), "ballast_bytes": 0 if ballast is None else ballast.numel() * 4}
    emit(summary)
    out = dict(res)
    out.update({"kind": args.kind, "rounds": rounds, "seconds": time.time() - t_start, "mismatches": bad})
    print("RESULT " + json.dumps(out), flush=True)
    print(("PASS" if bad == 0 else "FAIL") + " " + json.dumps(summary), flush=True)
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(run())
