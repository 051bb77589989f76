#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on its config:

    swap GB/s + iter/sec at 1.5x HBM oversubscription, 2 co-located clients
    (restated tests/pytorch-add.py), vs the reference's UVM page-fault path.

One STEP = one lock hand-off between the two clients (the leaving client's
working set goes out, the arriving client's comes in) together with the
quantum of client compute that follows it.  W warm-up hand-offs, then exactly K
timed hand-offs, delimited on the merged per-iteration timeline of the clients
(nvshare_b200/harness.py); both arms are analysed by the same code.

JSON line (one, from rank 0):
  value      swap GB/s measured ON THE DEVICE: payload bytes the sm_100a copy
             kernels moved in the timed hand-offs / CUDA-event time of those
             launches (both directions summed; inputs resident in HBM / pinned
             host memory when each launch starts)
  e2e.value  the same metric END TO END through the LD_PRELOAD boundary:
             algorithmic bytes per hand-off / measured stall per hand-off, as the
             unmodified PyTorch application experiences it (host<->device copies,
             map/unmap, protocol and scheduling all inside)
  iter_per_s oversubscribed-client iterations per second over the timed window
  roofline   per-direction GB/s of the dominant kernel vs the link peak measured
             live with the copy engines (the path is PCIe/NVLink-bound, not
             HBM- or tensor-bound: SURVEY 8d)
  cpu_baseline  oracle/nvshare_oracle.c's migration restatement (memcpy) on one
             host core, bounded sample -- reported, not a target

--impl reference runs the UNMODIFIED reference (oracle/_ref: its libnvshare.so
and nvshare-scheduler, i.e. cuMemAllocManaged + UVM faults) through the same
harness on the same workload.

Multi-GPU (--gpus N under torchrun): the path shards by slab, not by client:
rank 0 hosts the two clients on GPU 0 and stripes their backing slabs over the
HBM of GPUs 1..N-1 (peer tier, cuMemMap of peer physical memory, no NCCL);
the other ranks only take part in the barriers.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from nvshare_b200 import harness  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--tq", type=int, default=10, help="scheduler time quantum in seconds (reference default: 30)")
    ap.add_argument("--oversub", type=float, default=1.5, help="aggregate client footprint / HBM")
    ap.add_argument("--clients", type=int, default=2)
    ap.add_argument("--kind", choices=["add", "matmul"], default="add")
    ap.add_argument("--pattern", choices=["ones", "pos"], default="pos")
    ap.add_argument("--hbm-fraction", type=float, default=0.0,
                    help="share of the GPU's HBM the experiment may use; the rest is held by a ballast process. "
                         "1.0 = the configuration BASELINE.json names; 0 (default) = 1.0 if the host memory this "
                         "arm needs fits the box's limit, else the largest fraction that does")
    ap.add_argument("--keep", default="", help="directory to keep logs in")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measure_link_peak(torch, nbytes=2 * GiB):
    """Copy-engine H2D / D2H bandwidth with pinned memory, CUDA events: the
    roofline denominator for the host tier, measured on this box in this run."""
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    out = {}
    for name, (dst, src) in {"d2h": (host, dev), "h2d": (dev, host)}.items():
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            e1.synchronize()
            best = max(best, nbytes / 1e6 / e0.elapsed_time(e1))
        out[name] = best
    del dev, host
    torch.cuda.empty_cache()
    return out


def measure_hbm_roofline(torch, nbytes=4 * GiB):
    """The same sm_100a kernel used device-to-device, all 148 SMs: shows the kernel itself
    runs at the HBM copy peak, i.e. the link -- not the kernel -- bounds the swap path."""
    from nvshare_b200 import engine as E
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    src.random_(0, 255)
    descs = [(src.data_ptr() + o, dst.data_ptr() + o, 2 * MiB) for o in range(0, nbytes, 2 * MiB)]
    with E.Engine(prepin=0) as e:
        n_sms = torch.cuda.get_device_properties(0).multi_processor_count
        e.copy_slabs(descs, variant="tma", grid=n_sms)
        ms = min(e.copy_slabs(descs, variant="tma", grid=n_sms) for _ in range(3))
    ok = bool(torch.equal(src, dst))
    del src, dst
    torch.cuda.empty_cache()
    peak, source = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
        source = "MEASURED_PEAKS.json hbm_gbs (burst)"
    except Exception:
        pass
    traffic_gbps = 2 * nbytes / 1e6 / ms            # every payload byte is read once and written once
    return {"bound": "hbm", "kernel": "nvs_slab_copy_tma (device-to-device, %d CTAs)" % n_sms,
            "achieved": traffic_gbps, "peak": peak, "unit": "GB/s", "frac": traffic_gbps / peak,
            "peak_source": source, "payload_GBps": nbytes / 1e6 / ms, "verified": ok}


def ncu_traffic_ratio():
    """DRAM bytes (read + write) per payload byte of the eviction kernel, from the committed
    `ncu --set full` capture (profiles/r01_ncu_full_tma_selected.csv, first launch: 1 GiB HBM -> host)."""
    import csv
    try:
        rows = list(csv.reader(open(ROOT / "profiles" / "r01_ncu_full_tma_selected.csv")))
        hdr, first = rows[0], rows[2]
        rd = float(first[hdr.index("dram__bytes_read.sum")])
        wr = float(first[hdr.index("dram__bytes_write.sum")])
        return (rd + wr) * 1e9 / float(1 << 30)
    except Exception:
        return None


def cpu_baseline(sample_gib=4):
    """oracle_slab_move (memcpy restatement of page migration) on one core."""
    lib_path = ROOT / "oracle" / "_ref" / "liboracle.so"
    if not lib_path.exists():
        return {"value": None, "unit": "GB/s", "cores": 1, "kind": "port", "sample": "oracle not built"}
    import numpy as np
    from nvshare_b200.engine import CopyDesc
    lib = C.CDLL(str(lib_path))
    lib.oracle_slab_move.argtypes = [C.c_void_p, C.c_uint32]
    nbytes = sample_gib * GiB
    src = np.ones(nbytes, dtype=np.uint8)
    dst = np.zeros(nbytes, dtype=np.uint8)
    n = nbytes // (2 * MiB)
    arr = (CopyDesc * n)()
    for i in range(n):
        j = (i * 2654435761) % n if n & (n - 1) else (i * 2654435761) & (n - 1)
        arr[i].src, arr[i].dst, arr[i].bytes = src.ctypes.data + j * 2 * MiB, dst.ctypes.data + j * 2 * MiB, 2 * MiB
    lib.oracle_slab_move(arr, n)        # touch
    best = 0.0
    t_all = time.time()
    while time.time() - t_all < 10:
        t0 = time.time()
        lib.oracle_slab_move(arr, n)
        best = max(best, nbytes / 1e9 / (time.time() - t0))
    return {"value": round(best, 2), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"{sample_gib} GiB of 2 MiB slab descriptors, host->host memcpy (oracle_slab_move), best pass in 10 s"}


def host_memory_budget():
    """Bytes of host RAM this job may still take: the cgroup limit (pinned and UVM
    host pages are charged to it: profiles/r01_probe_h_pinned_accounting.txt) or MemAvailable."""
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except OSError:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        cur = int(open("/sys/fs/cgroup/memory.current").read())
        if lim != "max":
            room = int(lim) - cur
            avail = room if avail is None else min(avail, room)
    except (OSError, ValueError):
        pass
    return avail


def host_memory_needed(impl, clients, footprint, hbm_avail):
    """What each arm keeps in host RAM (measured, r01 call 3): the reference's UVM ends up
    holding every client's pages on the host; our shared pool holds what is swapped out plus
    the pinned windows ahead of it (peak 163 GB for 95.7 GB swapped at full scale)."""
    if impl == "reference":
        return clients * footprint * 1.03 + (8 << 30)
    swapped = max(clients * footprint - hbm_avail, 0)
    try:   # the scheduler-wide pool lives in /dev/shm; without room there every client pins its own arenas
        vfs = os.statvfs("/dev/shm")
        shared_ok = vfs.f_bavail * vfs.f_frsize >= hbm_avail
    except OSError:
        shared_ok = False
    return (swapped * 1.75 if shared_ok else swapped * clients * 1.15) + (12 << 30)


def pick_fraction(args, total_b):
    if args.hbm_fraction > 0:
        return args.hbm_fraction, None
    budget = host_memory_budget()
    frac = 1.0
    while frac > 0.1:
        hbm_avail = total_b * frac
        fp = args.oversub * hbm_avail / args.clients
        if budget is None or host_memory_needed(args.impl, args.clients, fp, hbm_avail) <= budget - (20 << 30):
            break
        frac = round(frac - 0.05, 2)
    note = None
    if frac < 0.999:
        note = (f"host RAM budget {budget / 1e9:.0f} GB cannot hold what the {args.impl} arm needs at full scale "
                f"({host_memory_needed(args.impl, args.clients, args.oversub * total_b / args.clients, total_b) / 1e9:.0f} GB); "
                f"scaled to {frac:.2f} of the HBM, the rest is held by a ballast process")
    return frac, note


def main():
    args = parse_args()
    rank, world, local = dist_env()
    import torch

    # reference arm: rank 0 alone runs and prints; the other ranks exit 0 without work
    if args.impl == "reference" and rank != 0:
        return 0
    use_dist = world > 1 and args.impl != "reference"
    if use_dist:
        import torch.distributed as dist
        import datetime
        torch.cuda.set_device(local)
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world,
                                timeout=datetime.timedelta(minutes=60))
        # create the communicator now, while GPU 0 is still empty; the data-path work then runs on rank 0
        # alone (the path shards by slab over the peers' HBM, not by rank) and everybody meets again below
        dist.barrier()

    result = None
    t_region = 0.0
    if rank == 0:
        result, t_region = run_rank0(args, torch, world)
    if use_dist:
        dist.barrier()
        t = torch.tensor([t_region], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # max over ranks (only rank 0 does data-path work)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result), flush=True)
        return 0 if result.get("verified", True) else 1
    return 0


def run_rank0(args, torch, world):
    paths = harness.impl_paths(args.impl)
    if not paths["lib"].exists():
        if args.impl == "reference":
            return {"impl": "reference", "unavailable": "oracle/_ref not built (reference sources absent at build time)"}, 0.0
        raise SystemExit("nvshare_b200/_build is missing: run __graft_entry__.build() first")

    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    free_b, total_b = torch.cuda.mem_get_info()
    ncpu = os.cpu_count()
    peak = measure_link_peak(torch)
    hbm_roof = measure_hbm_roofline(torch) if args.impl == "ours" else None
    cpu = cpu_baseline() if (args.impl == "ours" and world == 1) or args.impl == "reference" else None

    # -- geometry: clients x footprint = oversub x HBM the experiment may use
    args.hbm_fraction, scale_note = pick_fraction(args, total_b)
    hbm_avail = int(total_b * args.hbm_fraction)
    footprint = args.oversub * hbm_avail / args.clients
    # live n^2 fp32 blocks: x, y, the result and -- while `z = op(x, y)` rebinds -- the previous result
    blocks = 4
    n = int(math.floor(math.sqrt(footprint / (4 * blocks))))
    footprint = blocks * 4 * n * n
    algo_bytes_dir = max(args.clients * footprint - hbm_avail, 0.0)   # must come in (and go out) per hand-off
    # the clients stop as soon as warmup + steps + 2 hand-offs have been seen; this is only the safety limit
    seconds = (args.warmup + args.steps + 3) * (args.tq + (12 if args.impl == "ours" else 60)) + 30

    out_dir = Path(args.keep) if args.keep else Path(tempfile.mkdtemp(prefix="nvs_bench_"))
    out_dir.mkdir(parents=True, exist_ok=True)
    ballast = None
    if args.hbm_fraction < 0.999:
        ballast_bytes = int(total_b * (1 - args.hbm_fraction))
        code = ("import torch,time; b=torch.empty(%d,dtype=torch.uint8,device='cuda'); torch.cuda.synchronize();"
                "print('BALLAST',flush=True); time.sleep(100000)" % ballast_bytes)
        import subprocess
        ballast = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True)
        assert "BALLAST" in ballast.stdout.readline()
    # free the context-side memory this process holds before the clients start
    torch.cuda.empty_cache()

    extra = {}
    if world > 1 and args.impl == "ours":
        extra["NVSHARE_PEERS"] = ",".join(str(i) for i in range(1, world))
        # per peer; empty peer arenas are returned at once, so the clients share the peers' HBM over time
        extra["NVSHARE_PEER_CAPACITY_MIB"] = int(0.92 * total_b) >> 20
    sampler = harness.ClockSampler(out_dir / "clocks.csv")
    sampler.start()
    t0 = time.time()
    try:
        res = harness.run_clients(args.impl, out_dir, args.clients, args.kind, n, args.pattern, seconds, args.tq,
                                  extra_env=extra, stop_after_handoffs=args.warmup + args.steps + 2)
    finally:
        clocks = sampler.stop()
        if ballast:
            ballast.kill()
            ballast.wait()
    wall = time.time() - t0

    verified = all(r["rc"] == 0 and r["meta"].get("summary", {}).get("result") == "PASS" for r in res)
    clients = {f"client{i}": r["iters"] for i, r in enumerate(res)}
    try:
        a = harness.analyse(clients, args.warmup, args.steps)
    except Exception as ex:  # not enough hand-offs: report what happened, loudly
        tails = {f"client{i}": r["err_tail"][-600:] for i, r in enumerate(res)}
        last_ops = {}
        for i in range(args.clients):        # what the engines did last: the first thing one wants to know
            f = out_dir / f"engine{i}.jsonl"
            if f.exists():
                last_ops[f"client{i}"] = [l[:400] for l in f.read_text().splitlines() if '"op":"pin"' not in l][-6:]
        return {"metric": "swap_GBps_at_1.5x_hbm_oversub_2_clients", "error": str(ex), "impl": args.impl,
                "verified": False, "client_rc": [r["rc"] for r in res], "stderr_tails": tails,
                "engine_last_ops": last_ops, "n": n, "hbm_fraction_used": args.hbm_fraction, "scale_note": scale_note,
                "host_memory_budget": host_memory_budget(), "wall_s_total": wall}, wall

    stall = a["stall_per_handoff_s"]
    e2e_gbps = (2 * algo_bytes_dir / 1e9) / stall if stall > 0 else None

    line = {
        "metric": "swap_GBps_at_1.5x_hbm_oversub_2_clients",
        "unit": "GB/s",
        "impl": args.impl,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * a["window_s"] / args.steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "verified": verified,
        "iter_per_s": a["iter_per_s"],
        "iter_per_s_resident": a["iter_per_s_resident"],
        "stall_ms_per_handoff": 1e3 * stall,
        "first_iter_gap_ms": [1e3 * g for g in a["first_iter_gap_s"]],
        "gpu_busy_frac": a["gpu_busy_frac"],
        "config": {
            "workload": f"2x {args.kind} fp32 [n,n] n={n} ({args.pattern}), restated "
                        f"{'tests/pytorch-add.py' if args.kind == 'add' else 'tests/tf-matmul.py'}, "
                        f"footprint {footprint / 1e9:.1f} GB/client = {args.clients * footprint / hbm_avail:.2f}x of "
                        f"{hbm_avail / 1e9:.1f} GB HBM",
            "clients": args.clients, "oversubscription": args.clients * footprint / hbm_avail,
            "tq_s": args.tq, "hbm_bytes": total_b, "hbm_fraction_used": args.hbm_fraction, "scale_note": scale_note,
            "backing_tier": "peer-HBM over NVLink (GPUs 1..%d) + pinned host" % (world - 1) if world > 1 and args.impl == "ours" else "pinned host DRAM over PCIe Gen5 x16",
            "l2_policy": "inputs (>= tens of GB per hand-off) exceed the 126 MB L2",
            "algorithmic_bytes_per_handoff_per_direction": algo_bytes_dir,
            "host_cores": ncpu,
        },
        "clocks": clocks,
        "link_peak_GBps_measured": peak,
        "wall_s_total": wall,
    }

    if args.impl == "ours":
        recs = harness.engine_records([out_dir / f"engine{i}.jsonl" for i in range(args.clients)], a["t_start"], a["t_end"])
        ev = [r for r in recs if r["op"] == "evict" and (r["bytes"] or r.get("elided_bytes"))]
        fe = [r for r in recs if r["op"] == "fetch" and (r["bytes"] or r.get("elided_bytes"))]
        bytes_moved = sum(r["bytes"] for r in ev + fe)
        dev_ms = sum(r["copy_ms"] for r in ev + fe)
        launches = sum(r["launches"] for r in ev + fe)
        ev_gbps = sum(r["bytes"] for r in ev) / 1e6 / max(sum(r["copy_ms"] for r in ev), 1e-9) if ev else None
        fe_gbps = sum(r["bytes"] for r in fe) / 1e6 / max(sum(r["copy_ms"] for r in fe), 1e-9) if fe else None
        line["value"] = bytes_moved / 1e6 / dev_ms if dev_ms else None
        line["gpu_launches"] = launches
        line["device"] = {"evict_GBps": ev_gbps, "fetch_GBps": fe_gbps, "bytes_moved": bytes_moved,
                          "bytes_elided_same_filled": sum(r.get("elided_bytes", 0) for r in recs),
                          "evicts": len(ev), "fetches": len(fe),
                          "map_ms_mean": statistics_mean([r["map_ms"] for r in ev + fe]),
                          "wait_ms_mean": statistics_mean([r["wait_ms"] for r in fe]),
                          "wall_ms_mean": {"evict": statistics_mean([r["wall_ms"] for r in ev]),
                                           "fetch": statistics_mean([r["wall_ms"] for r in fe])}}
        pins = [r for r in harness.engine_records([out_dir / f"engine{i}.jsonl" for i in range(args.clients)], 0, 1e18)
                if r["op"] == "pin"]
        if pins:        # where the pinned pool's pages ended up (engine.c numa_init): the last report covers the whole pool
            last = max(pins, key=lambda r: r["t"])
            line["device"]["pool_pages_per_numa_node"] = last["pages_per_node"]
            line["device"]["pool_placed_from_cpus"] = last["near_cpus"]
        per_dir = [g for g in (ev_gbps, fe_gbps) if g]
        achieved = sum(per_dir) / len(per_dir) if per_dir else None
        link = "nvlink" if world > 1 else "pcie"
        pk = 770.0 if world > 1 else (peak["d2h"] + peak["h2d"]) / 2
        line["roofline"] = {"bound": link, "achieved": achieved, "peak": pk, "unit": "GB/s",
                            "frac": achieved / pk if achieved else None,
                            "traffic": (bytes_moved / launches * ncu_traffic_ratio()) if launches and ncu_traffic_ratio() else None,
                            "traffic_source": "DRAM read+write per launch = algorithmic bytes per launch x the ratio measured by "
                                              "ncu --set full (profiles/r01_ncu_summary.md: 1.0796 GB of DRAM traffic for a "
                                              "1.0737 GB launch)",
                            "peak_source": "770 GB/s measured peer copy (B200_PROFILING.md)" if world > 1 else
                                           "cuMemcpyAsync pinned<->HBM measured in this run (nominal PCIe Gen5 x16: 63.0 GB/s)",
                            "kernel": "nvs_slab_copy_tma",
                            "algorithmic_bytes_per_launch": bytes_moved / launches if launches else None}
        line["roofline_hbm"] = hbm_roof
        line["e2e"] = {"value": e2e_gbps, "unit": "GB/s", "iter_per_s": a["iter_per_s"],
                       "h2d_bytes_per_step": sum(r["bytes"] for r in fe) / max(args.steps, 1),
                       "d2h_bytes_per_step": sum(r["bytes"] for r in ev) / max(args.steps, 1),
                       "definition": "2 x algorithmic bytes per hand-off / stall per hand-off seen by the application"}
        if cpu:
            line["cpu_baseline"] = cpu
    else:
        line["value"] = e2e_gbps
        line["gpu_launches"] = 0
        line["e2e"] = {"value": e2e_gbps, "unit": "GB/s", "iter_per_s": a["iter_per_s"], "h2d_bytes_per_step": 0,
                       "d2h_bytes_per_step": 0}
        line["cpu_baseline"] = {"value": e2e_gbps, "unit": "GB/s", "cores": ncpu, "kind": "reference",
                                "sample": "the reference has no CPU compute path: this is its UVM page-fault path "
                                          "(cuMemAllocManaged) on the same box, same workload; host cores only service faults"}
    if not args.keep:
        shutil.rmtree(out_dir, ignore_errors=True)
    return line, wall


def statistics_mean(xs):
    xs = [x for x in xs if x is not None]
    return sum(xs) / len(xs) if xs else None


if __name__ == "__main__":
    sys.exit(main())
