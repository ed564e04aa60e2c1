#!/usr/bin/env python
"""bench.py -- Gbp/s inserted by the batched multi-string insertion path (mr_insert_multi).

Workloads (BASELINE.json configs; seeded counter-based generators, ropebwt2_b200/synth.py):

  --config cfg2 (default)  100 M x 101 bp uniform reads, forward strand, RLO (`ropebwt2 -LRs`), ONE batch
                           into an empty index, one B200.  A "step" = one mr_insert_multi call over that
                           batch.  With N > 1 GPUs the ranks build ONE index of N x 100 M reads together
                           (sharded build): per-GPU work is fixed -> "weak" scaling.
  --config cfg3            the north-star job: 1.2 B x 101 bp reads of a 30x genome, RLO, inserted the way
                           the reference driver does it: 12 successive mr_insert_multi calls of
                           102,110,743 reads (main.c:94,238-251) into ONE growing index.  A "step" = one of
                           those calls; the line reports the whole job and every batch.
  --config cfg4            1 M x 10 kbp uniform reads, input order (`-LR`), one batch (the long-read path).

  value   device time of the hot path with the batch already resident in HBM (CUDA events on the engine's
          stream around every call, max over ranks)
  e2e     the reference-facing call (mr_insert_multi through the C-ABI) on a pinned HOST buffer: H2D copy
          of the batch and D2H of the symbol counts inside the timed region
  parity  after the timed region the index is decoded block by block through mr_itr_next_block and its
          md5 compared with the md5 of the UNMODIFIED reference's output on the same reads
          (tests/golden/ref_full_runs.json, recorded by tools/ref_full_run.py); a mismatch is fatal
  roofline  the dominant kernel (k_flat_merge in the dense regime): algorithmic bytes per launch / its
          CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the unmodified reference binary (oracle/_ref/ropebwt2) on a bounded sample

`--impl reference` times the reference's own CPU implementation (all the threads it can use: 4 workers +
master) on the SAME workload, in full (cfg2: one run of ~6-13 minutes whatever --steps says).
"""
import argparse
import ctypes as C
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbp/s inserted (101bp reads), bit-exact BWT"
UNIT = "Gbp/s"
REF_BATCH_BYTES = 10415295693  # the reference driver's default -m (main.c:94)
FLAGS = {"cfg2": "-LRs", "cfg3": "-LRs", "cfg3u": "-LRs", "cfg4": "-LR"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(FLAGS))
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads (per GPU for cfg2); default: the config's own")
    ap.add_argument("--cpu-reads", type=int, default=4_000_000, help="reads in the bounded CPU-reference sample of the b200 arm (about 15 s of reference time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the md5 check of the final index")
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference: run a sample of this many reads instead of the full workload")
    ap.add_argument("--layout", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1: one index sharded over all GPUs (default) or one independent index per GPU")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled during the timed region.  NVML in-process (pynvml: what
    nvidia-smi itself reads) when it is importable -- spawning nvidia-smi five times a second next to a collective
    job delays the rank that does it -- else the nvidia-smi query of the profiling recipe."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip()]
        return int(ids[index]) if ids and all(x.strip().isdigit() for x in ids) and index < len(ids) else index

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        active = {nm for bit, nm in self.NVML_REASONS.items() if mask & bit}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        return [str(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), str(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                "%.1f" % (n.nvmlDeviceGetPowerUsage(h) / 1000.0)] + ["Active" if nm in active else "Not Active" for nm in names]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1 if self.nvml is not None else 0.2)

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
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "NVML in-process (pynvml), every 0.1 s" if self.nvml is not None else "nvidia-smi --query-gpu, every 0.2 s"}


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "ropebwt2")
    return p if os.path.exists(p) else None


def workload_of(args, world: int = 1):
    """(the workload all ranks build together, reads per rank, flags)."""
    from ropebwt2_b200 import synth
    w = synth.workload(args.config, 0)
    per_rank = args.reads or w["n"]
    if args.config == "cfg2":
        w = synth.workload(args.config, per_rank * world)  # rank r inserts reads [r*per_rank, (r+1)*per_rank)
    else:
        w = synth.workload(args.config, per_rank)
    return w, per_rank, FLAGS[args.config]


def mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


# ------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference binary on the box's host cores
# ------------------------------------------------------------------------------------------------------
def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if ref_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ropebwt2 was not prebuilt (needs /root/reference once)"}))
        return
    from oracle import oracle as orc  # the checker; allowed only in this baseline leg
    from ropebwt2_b200 import synth
    world = args.gpus
    w, per_rank, flags = workload_of(args, 1)   # the reference runs the one-GPU workload whatever N is (it does not scale with N)
    full_n = w["n"]
    # memory: ~25 GB RSS for cfg2 (10.2 GB buffer + 4.8 GB string state + the B+-trees)
    need_gb = 2.8e-7 * full_n if args.config != "cfg4" else 30.0
    sample = args.ref_sample
    note = ""
    if not sample and (args.config in ("cfg3", "cfg3u") or mem_available_gb() < need_gb):
        # cfg3 takes ~3-5 h on CPU: never inside a bench run.  Its recorded full run stands in (measured in the build container).
        rec = orc.ref_recorded(w, flags)
        if args.config in ("cfg3", "cfg3u") and rec:
            val = rec["gbp_per_s_hot_path"]
            print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": len(rec["hot_path_s_per_batch"]), "warmup": 0,
                              "ms_per_step": rec["hot_path_s"] * 1e3 / len(rec["hot_path_s_per_batch"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": "u8", "data": "synthetic", "recorded": True,
                              "config": {"workload": "RECORDED full run (tools/ref_full_run.py, %s, %d host cores): %s %s" % (rec["when"], rec["host_cores"], synth.workload_key(w, flags), rec["batch"]),
                                         "md5_text": rec["md5_text"]},
                              "cpu_baseline": {"value": val, "unit": UNIT, "cores": 5, "kind": "reference", "sample": "full workload, recorded"},
                              "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return
        sample = 4_000_000
        note = "host has %.0f GB available: a bounded sample instead of the full workload; " % mem_available_gb()
    if sample:
        w = synth.workload(args.config, sample)
    key = synth.workload_key(w, flags)
    cache = "/tmp/rb2_reference_arm_%s.json" % (socket.gethostname())
    rec = None
    if os.path.exists(cache):
        try:
            c = json.load(open(cache))
            if c.get("key") == key and time.time() - c.get("t", 0) < 6 * 3600:
                rec, note = c["rec"], note + "measured earlier in this session on this box (N does not change the CPU reference); "
        except Exception:
            rec = None
    if rec is None:
        rec = orc.ref_stream_run(w, flags, want_md5=True, gen_threads=max(2, (os.cpu_count() or 8) // 4))
        try:
            json.dump({"key": key, "t": time.time(), "rec": rec}, open(cache, "w"))
        except Exception:
            pass
    nb = max(1, len(rec["hot_path_s_per_batch"]))
    val = rec["gbp_per_s_hot_path"]
    recorded = orc.ref_recorded(w, flags)
    same = recorded["md5_text"] == rec["md5_text"] if recorded else None
    if same is False:
        raise SystemExit("the reference's md5 on this box differs from the recorded one: the generators disagree")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": nb, "warmup": 0,
        "ms_per_step": rec["hot_path_s"] * 1e3 / nb, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": note + "%s: %s, oracle/_ref/ropebwt2 %s (unmodified reference), ONE full run (%d mr_insert_multi call%s), hot-path timer main.c:241,249 = %.1f s, wall %.1f s"
                               % ("FULL workload of " + args.config if not sample else "sample of " + args.config, key, flags, nb, "" if nb == 1 else "s", rec["hot_path_s"], rec["wall_s"]),
                   "threads": "4 workers + master (reference maximum, mrope.h:53)", "host_cores": os.cpu_count(),
                   "steps_requested": args.steps, "md5_text": rec["md5_text"], "md5_equals_recorded_run": same},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 5, "kind": "reference", "sample": "the full workload" if not sample else "%d reads" % sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
# b200 arm
# ------------------------------------------------------------------------------------------------------
def fill_host_batch(host: np.ndarray, w: dict, a: int, b: int, device, slice_bytes: int = 400_000_000):
    """mr_insert_multi buffer (reversed read + NUL per read) of reads a..b-1 into the pinned host array,
    generated on the GPU slice by slice (bit-identical to the numpy / C generators, tests/test_synth.py)."""
    import torch
    from ropebwt2_b200 import synth
    L1 = w["L"] + 1
    slice_reads = max(1, slice_bytes // L1)
    ht = torch.from_numpy(host[:(b - a) * L1])
    for x in range(a, b, slice_reads):
        y = min(b, x + slice_reads)
        t = torch.empty((y - x) * L1, dtype=torch.uint8, device=device)
        synth.fill_batch_torch(t, w, x, y, chunk=slice_reads)
        ht[(x - a) * L1:(y - a) * L1].copy_(t)
        del t
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured copy bandwidth (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_of(st, steps):
    peak, peak_src = peaks()
    dense = st.get("flat_batches", 0) > 0
    if dense:
        kernel, items, nbytes = "k_flat_merge", st["merge_blocks"], st["merge_bytes_rw"]
    else:
        kernel = "k_merge_half"
        items = st["merge_blocks"] - st["general_items"]
        nbytes = items * 2 * 512
    gbs = nbytes / (st["ms_merge"] * 1e-3) / 1e9 if st["ms_merge"] > 0 else 0.0
    traffic = info = None
    pj = os.path.join(ROOT, "profiles", "flat_merge_traffic.json" if dense else "merge_traffic.json")
    if os.path.exists(pj):
        j = json.load(open(pj))
        info = {k: v for k, v in j.items() if k in ("capture", "items_in_launch", "algorithmic_bytes_same_launch", "note")}
        traffic = j.get("dram_bytes_per_launch")
    return {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
            "traffic_capture": info, "peak_source": peak_src, "launches": int(st["n_merge_launches"]), "ms_in_kernel": st["ms_merge"],
            "algorithmic_bytes": int(nbytes), "items": int(items),
            "regime": "dense (flat 3-bit plane array, TMA-staged)" if dense else "sparse (leaf blocks)",
            "items_left_to_k_merge_fast_and_general": int(st["general_items"]), "ms_in_k_merge_fast_and_general": st["ms_merge_general"],
            "share_of_step": st["ms_merge"] / st["ms_total"] if st["ms_total"] else None}


def cpu_sample(args, flags):
    from oracle import oracle as orc
    from ropebwt2_b200 import synth
    w = synth.workload(args.config, args.cpu_reads if args.config != "cfg4" else max(1, args.cpu_reads // 400))
    rec = orc.ref_stream_run(w, flags, want_md5=False)
    return {"value": rec["gbp_per_s_hot_path"], "unit": UNIT, "cores": 5, "kind": "reference",
            "sample": "%s, oracle/_ref/ropebwt2 %s (4 workers + master), hot-path timer %.2f s, wall %.2f s, host has %d cores"
                      % (synth.workload_key(w, flags), flags, rec["hot_path_s"], rec["wall_s"], os.cpu_count())}


def verify_md5(L, mr_handle, w, flags):
    """md5 of the index behind an mrope_t (decoded through mr_itr_next_block) vs the recorded reference run."""
    from oracle import oracle as orc  # the checker
    rec = orc.ref_recorded(w, flags)
    t = time.time()
    md5, total = orc.index_md5(L, mr_handle)
    out = {"md5": md5, "symbols": total, "seconds": round(time.time() - t, 1)}
    if rec is None:
        out["result"] = "unpinned: no recorded reference run for this workload"
        return out
    if md5 != rec["md5_text"] or total + 1 != rec["text_bytes"]:
        raise SystemExit("PARITY FAILURE: md5 of the GPU index %s (%d symbols) != reference %s (%d bytes)" % (md5, total, rec["md5_text"], rec["text_bytes"]))
    out["result"] = "md5 == reference (oracle/_ref/ropebwt2 %s, recorded %s)" % (flags, rec["when"])
    return out


def verify_sharded_md5(L, eng, args, torch, rank, world, local, hptr, host):
    """Bit-exactness of the sharded build on real NCCL, at the one-GPU bench size: the N ranks build ONE index of the
    config-2 read set together (rank r inserts reads [r*n/N, (r+1)*n/N)), the sub-buckets travel to rank 0 in
    sub-bucket order, and the md5 of the decoded text must equal the recorded reference run.  Outside the timed region."""
    import hashlib
    import torch.distributed as dist
    from oracle import oracle as orc  # the checker
    from ropebwt2_b200 import synth
    flags = FLAGS[args.config]
    w = synth.workload(args.config, args.reads or 0)   # the ONE-GPU workload
    rec = orc.ref_recorded(w, flags)
    if rec is None:
        return {"result": "unpinned: no recorded reference run for " + synth.workload_key(w, flags)}
    t0 = time.time()
    n, ln = w["n"], w["L"]
    share = n // world
    dev = torch.device("cuda", local)
    fill_host_batch(host, w, rank * share, (rank + 1) * share, dev)
    eng.reset()
    eng.insert_multi_ptr(hptr, share * (ln + 1))
    md5, total = hashlib.md5(), 0
    for s in range(36):
        owner = L.rb2_shard_owner(world, s)
        blocks = eng.fetch_subbucket(s) if rank == owner else None
        nb = torch.tensor([blocks.shape[0] if blocks is not None else 0], dtype=torch.int64, device=dev)
        dist.broadcast(nb, src=owner)
        k = int(nb.item())
        if k == 0:
            continue
        if rank == owner and rank == 0:
            got = blocks
        elif rank == owner or rank == 0:
            t = torch.empty((k, 512), dtype=torch.uint8, device=dev)
            if rank == owner:
                t.copy_(torch.from_numpy(blocks))
                dist.send(t, dst=0)
            else:
                dist.recv(t, src=owner)
            got = t.cpu().numpy() if rank == 0 else None
            del t
        else:
            got = None
        if rank == 0:
            txt = orc.blocks_ascii(got, n * (ln + 1))
            md5.update(memoryview(txt))
            total += txt.size
    if rank != 0:
        return {"result": "checked on rank 0"}
    md5.update(b"\n")
    if md5.hexdigest() != rec["md5_text"] or total + 1 != rec["text_bytes"]:
        raise SystemExit("PARITY FAILURE (sharded build, %d ranks): md5 %s (%d symbols) != reference %s" % (world, md5.hexdigest(), total, rec["md5_text"]))
    return {"md5": md5.hexdigest(), "symbols": total, "seconds": round(time.time() - t0, 1),
            "result": "md5 == reference for ONE index of %s built by %d ranks over NCCL (%d reads per rank; oracle/_ref/ropebwt2 %s, recorded %s); "
                      "the timed weak-scaling index (%d reads) is checked for symbol conservation" % (synth.workload_key(w, flags), world, share, flags, rec["when"], n * world)}


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
    from ropebwt2_b200.dist import Reducer, rank_info
    rank, world, local = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red = Reducer(world, torch.device("cuda", local))
    os.environ["RB2_DEVICE"] = str(local)
    if args.config in ("cfg3", "cfg3u"):
        out = run_multibatch(args, torch, red, rank, world, local)
    else:
        out = run_single_batch(args, torch, red, rank, world, local)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_single_batch(args, torch, red, rank, world, local):
    """cfg2 / cfg4: one batch into an empty index per step."""
    from ropebwt2_b200 import Engine, MRope, load
    L = load()
    w, n, flags = workload_of(args, world)
    ln = w["L"]
    so = 1 if "s" in flags else 0
    nbytes = n * (ln + 1)
    bp = n * ln
    dev = torch.device("cuda", local)

    # ---- inputs: pinned host batch (e2e leg) and a device-resident copy (value leg) ----------
    L.rb2_host_alloc.restype = C.c_void_p
    hptr = L.rb2_host_alloc(nbytes)
    host = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(hptr))
    fill_host_batch(host, w, rank * n, (rank + 1) * n, dev)
    sharded = world > 1 and args.layout == "sharded"
    if sharded:
        # one index over all ranks: the library's own NCCL communicator, unique id handed out through torch.distributed
        from ropebwt2_b200.binding import ShardedEngine, nccl_unique_id
        from ropebwt2_b200.dist import broadcast_bytes
        eng = ShardedEngine(local, so, rank, world, nccl_uid=broadcast_bytes(nccl_unique_id() if rank == 0 else None))
    else:
        eng = Engine(local, so)
    dptr = eng.dev_alloc(nbytes)
    eng.dev_upload(dptr, host)

    def barrier():
        torch.cuda.synchronize()
        red.barrier()
        torch.cuda.synchronize()

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
    ms_value = red.max(st["ms_total"])
    counts = eng.counts()
    n_idx = world if sharded else 1  # strings in the index this rank sees
    if not (int(counts.sum()) == nbytes * n_idx and int(counts[:, 0].sum()) == n * n_idx):
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
        mr = None
    else:
        eng.close()
        mr = MRope(so)

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
    if mr is not None:
        L.rb2_span_begin(mr.engine_handle)
    t0 = time.time()
    for _ in range(args.steps):
        tot = e2e_step()
    span_ms = L.rb2_span_ms(mr.engine_handle) if mr is not None else 0.0  # waits for the last insertion
    barrier()
    wall_e2e = time.time() - t0
    st2 = e2e_stats()
    # one GPU: the calls are pipelined (the copy of step k+1 overlaps the insertion of step k): the job's device time is
    # the CUDA-event span from the first copy to the end of the last insertion, not the sum of the per-call times
    ms_e2e = red.max(span_ms if span_ms > 0 else st2["ms_total"])
    clocks = sampler.summary()
    assert tot == nbytes * n_idx
    ms_exch = red.max(st.get("ms_exchange", 0.0))

    # ---- parity: md5 of the index the last e2e step built vs the unmodified reference (outside the timed region) ----
    parity = {"result": "symbol conservation only (--no-verify)"}
    if not args.no_verify:
        if mr is not None:
            parity = verify_md5(L, mr.h, w, flags)
        elif sharded and args.config == "cfg2" and n % world == 0:
            parity = verify_sharded_md5(L, eng, args, torch, rank, world, local, hptr, host)
        else:
            parity = {"result": "symbol conservation checked in-run at this N; bit-exactness of the sharded build: tests/test_sharded_gpu.py, tests/test_sharded_nccl.py"}
    if sharded:
        eng.quiesce()  # collective: the ranks close the mappings of each other's state buffers before anybody frees them
        eng.close()
    else:
        mr.close()
    L.rb2_host_free(C.c_void_p(hptr))
    if rank != 0:
        return None

    value = world * bp * args.steps / (ms_value * 1e-3) / 1e9
    e2e_val = world * bp * args.steps / (ms_e2e * 1e-3) / 1e9
    from ropebwt2_b200 import synth
    out = {
        "metric": METRIC if ln == 101 else "Gbp/s inserted (%d bp reads), bit-exact BWT" % ln, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE %s: %s (%d reads per GPU), forward strand, %s, one batch into an empty index" % (args.config, synth.workload_key(w, flags), n, "RLO" if so else "input order"),
                   "reads_per_gpu": n, "read_length": ln, "sorting_order": "RLO" if so else "IO",
                   "parallelism": (f"sharded x{world}: ONE index of {world * n} reads, 36 sub-buckets over {world} GPUs, "
                                   "per column one count all-gather + the string ids by NCCL send/recv under the merge + "
                                   + ("the new interval starts stored by the merge kernel straight into the peers' HBM (CUDA IPC mappings over NVLink)"
                                      if st.get("p2p_batches", 0) else "the interval starts by NCCL send/recv behind the merge")) if sharded
                   else (f"replicas x{world}" if world > 1 else "one GPU"),
                   "l2": "inputs (%.1f GB batch, multi-GB symbol array) far exceed the 126 MB L2" % (nbytes / 1e9),
                   "timing": "CUDA events on the engine stream around each call; wall-clock cross-check %.3f s/step" % (wall_value / args.steps),
                   "parity": parity["result"]},
        "parity": parity,
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 7 * 48,
                "ms_per_step": ms_e2e / args.steps, "wall_s_per_step": wall_e2e / args.steps, "api": e2e_api,
                "pipelining": ("mr_insert_multi returns when the batch is on the device (the caller's buffer is free, main.c:243); a worker thread inserts it "
                               "while the next call copies: K back-to-back steps are timed as one span (CUDA events, first copy -> end of last insertion); "
                               "sum of per-call copy + insertion times: %.1f ms per step" % (st2["ms_total"] / args.steps)) if mr is not None else "synchronous calls",
                "phases_ms_per_step": {k: st2[k] / args.steps for k in ("ms_h2d", "ms_transpose", "ms_members", "ms_groups", "ms_merge", "ms_directory", "ms_convert")}},
        "gpu_launches": int(st["n_launches"]),
        "roofline": roofline_of(st, args.steps),
        "phases_ms_per_step": {k: st[k] / args.steps for k in ("ms_transpose", "ms_members", "ms_groups", "ms_merge", "ms_merge_general", "ms_directory", "ms_exchange", "ms_convert")},
    }
    if sharded:
        out["exchange"] = {"ms_per_step_max_over_ranks": ms_exch / args.steps, "bytes_received_per_step_rank0": int(st["exch_bytes"] / args.steps),
                           "interval_starts": "direct delivery: peer stores from the k_flat_merge epilogue + a 16-byte stream barrier" if st.get("p2p_batches", 0)
                           else "ncclSend/ncclRecv behind the merge (RB2_P2P=0, or the peers' buffers could not be mapped)",
                           "collectives_per_column": "1 all-gather (1.8 KB/rank) + 1 grouped send/recv of the string ids"
                           + (" + 1 16-byte all-gather as barrier" if st.get("p2p_batches", 0) else " + 1 grouped send/recv of the interval starts")}
    if not args.no_cpu_baseline and ref_binary() is not None and world == 1:
        out["cpu_baseline"] = cpu_sample(args, flags)
    else:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "N > 1 or oracle/_ref/ropebwt2 not available"}
    return out


def run_multibatch(args, torch, red, rank, world, local):
    """cfg3: the reference driver's own batching -- successive mr_insert_multi calls into one growing index.
    The batches are generated up front into host memory (as the driver's parser would have them), then streamed
    through mr_insert_multi back to back: the calls are pipelined (copy of batch k+1 under the insertion of
    batch k), the job is timed as one CUDA-event span."""
    from ropebwt2_b200 import MRope, load, synth
    if world > 1:
        raise SystemExit("--config cfg3 runs on one GPU here (N > 1: --config cfg2, the sharded build)")
    L = load()
    w, n, flags = workload_of(args, 1)
    ln = w["L"]
    so = 1 if "s" in flags else 0
    per = (REF_BATCH_BYTES + ln) // (ln + 1)  # main.c:238: flush once the buffer holds >= m bytes
    if args.reads and args.reads < 2 * per:
        per = max(1, args.reads // 12)          # scaled-down runs keep the 12-batch shape
    cuts = list(range(0, n, per)) + [n]
    dev = torch.device("cuda", local)
    t_gen = time.time()
    bufs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        buf = np.empty((b - a) * (ln + 1), dtype=np.uint8)   # pageable host memory, like the reference driver's kstring buffer
        fill_host_batch(buf, w, a, b, dev)
        bufs.append(buf)
    t_gen = time.time() - t_gen
    torch.cuda.empty_cache()
    # the driver's batch buffer: ONE pinned buffer, refilled by the host before every call (a host memcpy stands in for
    # the reference driver's parser) and free again as soon as mr_insert_multi returns (main.c:243)
    L.rb2_host_alloc.restype = C.c_void_p
    cap = per * (ln + 1)
    hptr = L.rb2_host_alloc(cap)
    pinned = np.ctypeslib.as_array((C.c_uint8 * cap).from_address(hptr))
    mr = MRope(so)
    sampler = ClockSampler(local)
    sampler.start()
    mr.reset_stats()
    L.rb2_span_begin(mr.engine_handle)
    t0 = time.time()
    for buf in bufs:
        np.copyto(pinned[:buf.size], buf)
        mr.L.mr_insert_multi(mr.h, buf.size, C.cast(hptr, C.POINTER(C.c_uint8)), 1)
        tot = int(mr.counts().sum())   # mr_get_c: current on return (does not wait for the insertion)
    span_ms = L.rb2_span_ms(mr.engine_handle)  # waits for the last insertion
    wall = time.time() - t0
    clocks = sampler.summary()
    hist = mr.job_history(len(bufs))
    free, total_mem = torch.cuda.mem_get_info()
    if tot != n * (ln + 1):
        raise SystemExit("symbol conservation violated")
    batches = []
    idx = 0
    for (a, b), st in zip(zip(cuts[:-1], cuts[1:]), hist):
        idx += (b - a) * (ln + 1)
        ms_dev = st["ms_total"] - st["ms_h2d"]
        batches.append({"reads": b - a, "ms_insert": ms_dev, "ms_h2d": st["ms_h2d"], "Gbp/s": (b - a) * ln / (ms_dev * 1e-3) / 1e9,
                        "regime": "dense" if st["flat_batches"] else "sparse", "ms_merge": st["ms_merge"], "ms_groups": st["ms_groups"], "ms_members": st["ms_members"],
                        "ms_convert": st["ms_convert"], "merge_GB/s": st["merge_bytes_rw"] / max(st["ms_merge"], 1e-9) / 1e6,
                        "launches": int(st["n_launches"]), "index_symbols": idx})
    parity = {"result": "symbol conservation only (--no-verify)"}
    if not args.no_verify:
        parity = verify_md5(L, mr.h, w, flags)
    mr.close()
    L.rb2_host_free(C.c_void_p(hptr))
    bp = n * ln
    ms_ins = sum(x["ms_insert"] for x in batches)
    e2e_val = bp / (span_ms * 1e-3) / 1e9
    value = bp / (ms_ins * 1e-3) / 1e9
    peak, peak_src = peaks()
    mbytes = sum(x["merge_GB/s"] * x["ms_merge"] * 1e6 for x in batches)
    mms = sum(x["ms_merge"] for x in batches)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": len(batches), "warmup": 0, "ms_per_step": ms_ins / len(batches),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE %s (north star): %s, forward strand, RLO, %d successive mr_insert_multi calls of %d reads (the reference driver's default -m, main.c:94) into ONE growing index"
                               % (args.config, synth.workload_key(w, flags), len(batches), per),
                   "l2": "every column streams the whole symbol array (GBs): far beyond the 126 MB L2", "parity": parity["result"],
                   "timing": "value: sum of the insertions' device times (CUDA events around every batch on the engine stream); e2e: one CUDA-event span from the first "
                             "copy to the end of the last insertion; every batch is memcpy'd by the host into one pinned buffer (the stand-in for the driver's parser) and handed to "
                             "mr_insert_multi, whose copy runs under the previous batch's insertion; wall clock %.2f s" % wall,
                   "generation_s": round(t_gen, 1)},
        "parity": parity, "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": per * (ln + 1), "d2h_bytes_per_step": 7 * 48, "ms_per_step": span_ms / len(batches), "ms_job": span_ms,
                "wall_s_job": wall, "api": "mr_insert_multi (include/mrope.h) on host buffers + mr_get_c counts, once per batch, back to back"},
        "gpu_launches": sum(x["launches"] for x in batches),
        "roofline": {"bound": "hbm", "kernel": "k_flat_merge", "achieved": mbytes / max(mms, 1e-9) / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": mbytes / max(mms, 1e-9) / 1e6 / peak, "traffic": None, "peak_source": peak_src, "ms_in_kernel": mms, "algorithmic_bytes": int(mbytes),
                     "share_of_step": mms / ms_ins},
        "batches": batches,
        "hbm_high_water_GB": (total_mem - free) / 1e9,
        "cpu_baseline": cpu_sample(args, flags) if (not args.no_cpu_baseline and ref_binary() is not None) else {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "not run"},
    }


if __name__ == "__main__":
    main()
