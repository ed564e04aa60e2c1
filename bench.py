#!/usr/bin/env python
"""bench.py -- Gbp/s inserted by the batched multi-string insertion path (mr_insert_multi).

Workload (BASELINE.json configs[1]): 100 M x 101 bp synthetic uniform reads, forward strand,
RLO (`ropebwt2 -LRs`), one batch into an empty index, on one B200.  A "step" is one
mr_insert_multi call over that batch.  With N > 1 GPUs the ranks build ONE index of N x 100 M reads
together (sharded build: 36 sub-buckets spread over the ranks, one count all-gather and one
string-state exchange per column over NCCL): per-GPU work is fixed, so the
scaling is "weak".  `--layout replicas` runs N independent indexes instead (no exchange).

  value   device-resident input (rb2_insert_multi_dev), timed with CUDA events on the engine's
          stream around the whole call, max over ranks
  e2e     the reference-facing call (mr_insert_multi through the C-ABI) on a pinned HOST buffer:
          H2D copy of the batch and D2H of the symbol counts inside the timed region
  roofline  the dominant kernel (k_flat_merge in the dense regime this workload runs in): algorithmic
          bytes per launch / its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the unmodified reference binary (oracle/_ref/ropebwt2 -LRs) on a bounded sample

`--impl reference` times the reference's own CPU implementation (all the threads it can use:
4 workers + master) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbp/s inserted (101bp reads), bit-exact BWT"
UNIT = "Gbp/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU (BASELINE config 2: 100M)")
    ap.add_argument("--length", type=int, default=101)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-reads", type=int, default=4_000_000, help="reads in the bounded CPU-reference sample (about 15 s of reference time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1: one index sharded over all GPUs (default) or one independent index per GPU")
    return ap.parse_args()


def fill_batch(dst: np.ndarray, n: int, length: int, seed: int, chunk: int = 4_000_000) -> None:
    """Write the mr_insert_multi buffer (reversed read + NUL per read, main.c:200-225) for `n`
    seeded uniform reads into dst (n*(length+1) bytes), chunk by chunk to bound host memory."""
    view = dst.reshape(n, length + 1)
    for k, a in enumerate(range(0, n, chunk)):
        b = min(n, a + chunk)
        rng = np.random.default_rng([seed, k])
        view[a:b, :length] = rng.integers(1, 5, size=(b - a, length), dtype=np.uint8)[:, ::-1]
        view[a:b, length] = 0


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "ropebwt2")
    return p if os.path.exists(p) else None


def run_reference_sample(n: int, length: int, seed: int, path: str = None):
    """One timed run of the unmodified reference binary on n seeded reads; returns (Gbp/s over the
    reference's own hot-path timer (main.c:241,249), hot-path seconds, wall seconds)."""
    from oracle import oracle as orc  # the checker; allowed only in this baseline leg
    from ropebwt2_b200.synth import NT6
    own = path is None
    if own:
        buf = np.empty(n * (length + 1), dtype=np.uint8)
        fill_batch(buf, n, length, seed)
        lines = buf.reshape(n, length + 1)[:, :length][:, ::-1]  # forward reads
        txt = np.empty((n, length + 1), dtype=np.uint8)
        txt[:, :length] = NT6[lines]
        txt[:, length] = 10
        f = tempfile.NamedTemporaryFile(suffix=".txt", delete=False)
        f.write(txt.tobytes())
        f.close()
        path = f.name
    t = time.time()
    r = subprocess.run([ref_binary(), "-LRs", "-o", "/dev/null", path], capture_output=True, timeout=7200)
    wall = time.time() - t
    if r.returncode != 0:
        raise RuntimeError(r.stderr.decode()[-300:])
    hot = orc.ref_hot_path_seconds(r.stderr.decode())
    if own:
        os.unlink(path)
    return n * length / hot / 1e9, hot, wall, path


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if ref_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ropebwt2 was not prebuilt (needs /root/reference once)"}))
        return
    n, L = args.cpu_reads, args.length
    # write the sample once, time K runs after W warm-ups
    buf = np.empty(n * (L + 1), dtype=np.uint8)
    fill_batch(buf, n, L, args.seed)
    from ropebwt2_b200.synth import NT6
    txt = np.empty((n, L + 1), dtype=np.uint8)
    txt[:, :L] = NT6[buf.reshape(n, L + 1)[:, :L][:, ::-1]]
    txt[:, L] = 10
    f = tempfile.NamedTemporaryFile(suffix=".txt", delete=False)
    f.write(txt.tobytes())
    f.close()
    hots = []
    for it in range(args.warmup + args.steps):
        g, hot, wall, _ = run_reference_sample(n, L, args.seed, f.name)
        if it >= args.warmup:
            hots.append(hot)
    os.unlink(f.name)
    mean = sum(hots) / len(hots)
    val = n * L / mean / 1e9
    sample = f"{n} x {L} bp uniform reads (seed {args.seed}), ropebwt2 -LRs, hot-path timer main.c:241,249"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"bounded sample of BASELINE configs[1]: {sample}", "threads": "4 workers + master (reference maximum, mrope.h:53)",
                   "host_cores": os.cpu_count()},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 5, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        reference_arm(args)
        return
    # NCCL reads these once per process: the sharded build moves a few large point-to-point messages per column
    os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
    os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
    import torch
    import torch.distributed as dist
    from ropebwt2_b200.dist import Reducer, rank_info, shard_seed
    rank, world, local = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red = Reducer(world, torch.device("cuda", local))

    from ropebwt2_b200 import Engine, MRope, load
    L = load()
    n, ln = args.reads, args.length
    nbytes = n * (ln + 1)
    bp = n * ln

    # ---- inputs: pinned host batch (e2e leg) and a device-resident copy (value leg) ----------
    L.rb2_host_alloc.restype = C.c_void_p
    hptr = L.rb2_host_alloc(nbytes)
    host = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(hptr))
    fill_batch(host, n, ln, shard_seed(args.seed, rank))
    sharded = world > 1 and args.layout == "sharded"
    if sharded:
        # one index over all ranks: the library's own NCCL communicator, unique id handed out through torch.distributed
        from ropebwt2_b200.binding import ShardedEngine, nccl_unique_id
        from ropebwt2_b200.dist import broadcast_bytes
        eng = ShardedEngine(local, 1, rank, world, nccl_uid=broadcast_bytes(nccl_unique_id() if rank == 0 else None))
    else:
        eng = Engine(local, 1)
    dptr = eng.dev_alloc(nbytes)
    eng.dev_upload(dptr, host)
    os.environ["RB2_DEVICE"] = str(local)

    def barrier():
        torch.cuda.synchronize()
        red.barrier()
        torch.cuda.synchronize()

    max_over_ranks = red.max

    # ---- value leg: device-resident input ------------------------------------------------
    for _ in range(args.warmup):
        eng.reset()
        eng.insert_multi_dev(dptr, nbytes)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    eng.reset_stats()
    t0 = time.time()
    for _ in range(args.steps):
        eng.reset()
        eng.insert_multi_dev(dptr, nbytes)
    barrier()
    wall_value = time.time() - t0
    st = eng.stats()
    ms_value = max_over_ranks(st["ms_total"])
    counts = eng.counts()
    n_idx = world if sharded else 1  # strings in the index this rank sees
    ok = int(counts.sum()) == nbytes * n_idx and int(counts[:, 0].sum()) == n * n_idx
    if not ok:
        raise SystemExit("symbol conservation violated: the index does not hold the batch")
    eng.dev_free(dptr)

    # ---- e2e leg: the public host-buffer call on the pinned batch, counts read back ----------
    if sharded:
        def e2e_step():
            eng.reset()
            eng.insert_multi_ptr(hptr, nbytes)
            return int(eng.counts().sum())
        e2e_stats, e2e_reset_stats = eng.stats, eng.reset_stats
        e2e_api = "rb2_insert_multi_sharded (include/ropebwt2_b200.h) on a pinned host buffer per rank + rb2_counts"
    else:
        eng.close()
        mr = MRope(1)

        def e2e_step():
            L.rb2_reset(mr.engine_handle)
            mr.L.mr_insert_multi(mr.h, nbytes, C.cast(hptr, C.POINTER(C.c_uint8)), 1)
            return int(mr.counts().sum())
        e2e_stats, e2e_reset_stats = mr.stats, mr.reset_stats
        e2e_api = "mr_insert_multi (include/mrope.h) on a pinned host buffer + mr_get_c counts"

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e2e_reset_stats()
    t0 = time.time()
    for _ in range(args.steps):
        tot = e2e_step()
    barrier()
    wall_e2e = time.time() - t0
    st2 = e2e_stats()
    ms_e2e = max_over_ranks(st2["ms_total"])
    clocks = sampler.summary()
    assert tot == nbytes * n_idx
    ms_exch = max_over_ranks(st.get("ms_exchange", 0.0))
    if sharded:
        eng.close()
    else:
        mr.close()
    L.rb2_host_free(C.c_void_p(hptr))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured copy bandwidth (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # dominant kernel.  Dense regime (the headline workload): k_flat_merge, one streaming pass per column
    # over the flat symbol array -- algorithmic bytes = old array read + new array written + records
    # (20 B read, 8 B rank written each).  Sparse regime: k_merge_half (two leaf blocks per warp) --
    # every item it finishes reads one 512-byte leaf block and writes it back.
    dense = st.get("flat_batches", 0) > 0
    if dense:
        kernel = "k_flat_merge"
        fast_items = st["merge_blocks"]
        fast_bytes = st["merge_bytes_rw"]
    else:
        kernel = "k_merge_half"
        fast_items = st["merge_blocks"] - st["general_items"]
        fast_bytes = fast_items * 2 * 512
    merge_gbs = fast_bytes / (st["ms_merge"] * 1e-3) / 1e9 if st["ms_merge"] > 0 else 0.0
    traffic = None
    traffic_info = None
    ncu_json = os.path.join(ROOT, "profiles", "flat_merge_traffic.json" if dense else "merge_traffic.json")
    if os.path.exists(ncu_json):
        traffic_info = {k: v for k, v in json.load(open(ncu_json)).items() if k in ("capture", "items_in_launch", "algorithmic_bytes_same_launch", "note")}
        traffic = json.load(open(ncu_json)).get("dram_bytes_per_launch")  # one late launch (column 90) at this workload size
    value = world * bp * args.steps / (ms_value * 1e-3) / 1e9
    e2e_val = world * bp * args.steps / (ms_e2e * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: {n} x {ln} bp uniform reads per GPU, forward strand, RLO (-LRs), one batch into an empty index",
                   "reads_per_gpu": n, "read_length": ln, "sorting_order": "RLO",
                   "parallelism": (f"sharded x{world}: ONE index of {world * n} reads, 36 sub-buckets over {world} GPUs, "
                                   "per column one count all-gather + one string-state exchange (NCCL send/recv)") if sharded
                   else (f"replicas x{world}" if world > 1 else "one GPU"),
                   "l2": "inputs (%.1f GB batch, multi-GB leaf-block pool) far exceed the 126 MB L2" % (nbytes / 1e9),
                   "timing": "CUDA events on the engine stream around each call; wall-clock cross-check %.3f s/step" % (wall_value / args.steps),
                   "parity": "symbol conservation checked in-run; bit-exactness vs the reference is tests/test_parity_gpu.py"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 7 * 48,
                "ms_per_step": ms_e2e / args.steps, "wall_s_per_step": wall_e2e / args.steps,
                "api": e2e_api},
        "gpu_launches": int(st["n_launches"]),
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": merge_gbs, "peak": peak, "unit": "GB/s",
                     "frac": merge_gbs / peak, "traffic": traffic, "traffic_capture": traffic_info, "peak_source": peak_src,
                     "launches": int(st["n_merge_launches"]), "ms_in_kernel": st["ms_merge"],
                     "algorithmic_bytes": int(fast_bytes), "items": int(fast_items), "regime": "dense (flat symbol array)" if dense else "sparse (leaf blocks)",
                     "items_left_to_k_merge_fast_and_general": int(st["general_items"]), "ms_in_k_merge_fast_and_general": st["ms_merge_general"],
                     "share_of_step": st["ms_merge"] / st["ms_total"] if st["ms_total"] else None},
        "phases_ms_per_step": {k: st[k] / args.steps for k in ("ms_transpose", "ms_members", "ms_groups", "ms_merge", "ms_merge_general", "ms_directory", "ms_exchange", "ms_convert")},
    }
    if sharded:
        out["exchange"] = {"ms_per_step_max_over_ranks": ms_exch / args.steps, "bytes_received_per_step_rank0": int(st["exch_bytes"] / args.steps),
                           "collectives_per_column": "1 all-gather (1.8 KB/rank) + 1 grouped send/recv of the string state"}
    if not args.no_cpu_baseline and ref_binary() is not None:
        g, hot, wall, _ = run_reference_sample(args.cpu_reads, ln, args.seed)
        out["cpu_baseline"] = {"value": g, "unit": UNIT, "cores": 5, "kind": "reference",
                               "sample": f"{args.cpu_reads} x {ln} bp uniform reads, oracle/_ref/ropebwt2 -LRs (4 workers + master), "
                                         f"hot-path timer {hot:.2f} s, wall {wall:.2f} s, host has {os.cpu_count()} cores"}
    else:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/ropebwt2 not available"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
