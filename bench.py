#!/usr/bin/env python
"""bench.py -- spacer-frame queries/s of the SpacePHARER search hot path (prefilter + align) on B200.

Workload (BASELINE.json configs[1]): 10 000 synthetic spacers (25-45 nt, six-frame translated, ~38 k frames)
against a synthetic GenBank_phage_2018_09-shaped ORF set (8 000 genomes, ~0.5 M ORFs, ~53 M residues),
predictmatch parameters (VTML80 seeding, -s 5.7, k = 6 spaced 11011101, --max-seqs 300, VTML40 16/2, -e 200, -a 1).

A "step" is one pass of the search (prefilter -> align) over the whole query batch against the resident index:
  value   : device-resident passes (queries already in HBM) over `--lanes` contexts per GPU that share one index and run
            concurrently (one host thread each; the kernels of one lane fill the tails and host-sync bubbles of the other):
            K steps in total, CUDA events on the lane streams (latest end - earliest start), L2 flushed before every step,
            max over ranks.  `single_lane` repeats the figure for one context (the kernels alone on the device); the
            roofline numbers come from that segment
  e2e     : a stream of batches through the public API with HOST buffers (pinned), over the same lanes: every step uploads
            its frames (H2D), runs the kernels and reads its sorted result rows + CIGARs back (D2H); wall clock from the
            first upload to the last row landed
  roofline: the prefilter match kernel (dominant kernel), algorithmic bytes / its CUDA-event duration against the
            measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the reference's own CPU implementation -- the unmodified reference binary built by
            `make -C oracle ref` (oracle/_ref/spacepharer, AVX2 + OpenMP), `prefilter` + `align` with all host threads on DBs
            written from the same seeds; the oracle port only if that build is missing (said in the line).
N > 1 (torchrun): `--split-mode query` (default): every rank holds the whole index (1.7 GB of 180 GB) and searches its
own slice of the batch, no data-path collective; per-GPU work is fixed (10 000 spacers per GPU), so scaling is "weak".
`--split-mode target`: targets sharded across ranks like MMseqs2 --split, queries replicated, per-query hit lists
exchanged device to device over NCCL and merged to the unsplit result (strong scaling; the k-mer lookups per query do
not shrink with the shard, so this mode is about capacity, not speed -- DESIGN.md 6).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "spacer-frame queries/sec (prefilter+align)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spacers", type=int, default=10000)
    ap.add_argument("--genomes", type=int, default=8000)
    ap.add_argument("--split-mode", default="query", choices=["index", "target", "query"],
                    help="N > 1, mode of the main legs: query = full index per GPU, queries partitioned (default: the throughput mode while "
                         "the index fits one GPU); index = k-mer table sharded over groups of GPUs (--shard-group), gathered through NVLink "
                         "peer memory; target = MMseqs2-style target split with hit-list exchange + merge.  The line always carries the "
                         "index split (value, strong scaling on config 4, parity) next to the replicas.")
    ap.add_argument("--shard-group", type=int, default=0,
                    help="index split: GPUs sharing one copy of the k-mer table (0 = all N); N must be a multiple of it")
    ap.add_argument("--strong-spacers", type=int, default=1000000,
                    help="spacers of the strong-scaling leg (BASELINE config 4: the whole job split over the GPUs); 0 = skip")
    ap.add_argument("--long-steps", type=int, default=200,
                    help="steps of the additional long value leg (`steady_state`), 0 = skip")
    ap.add_argument("--no-control", action="store_true", help="skip the control-setDB pass of the strong-scaling leg")
    ap.add_argument("--no-disk-to-disk", action="store_true",
                    help="skip the product-level leg (the C++ host binary `spb200 search` on the config-4 DBs, from disk to disk)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_ok check (rows of the config-2 batch vs the reference digest)")
    ap.add_argument("--cpu-sample-spacers", type=int, default=0, help="spacers in the CPU baseline sample (0 = auto)")
    ap.add_argument("--lanes", type=int, default=3,
                    help="contexts per GPU sharing one index; batches of different lanes overlap on the device (query split / 1 GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mask", action="store_true", help="skip the tantan host pre-step (--mask 0)")
    ap.add_argument("--exhaustive-spacers", type=int, default=1000,
                    help="spacers per GPU in the exhaustive-search leg (BASELINE config 5: all frames x all ORFs, no k-mer stage); 0 = skip")
    return ap.parse_args()


def workload(args, world=1):
    """`world` > 1 with query split: the batch grows with the number of GPUs (weak scaling, 10 000 spacers per GPU)."""
    from spacepharer_b200 import synth
    tres, toff, _ = synth.make_targets(n_genomes=args.genomes)
    qres, qoff, _ = synth.make_query_frames(args.spacers * world, tres, toff)
    return tres, toff, qres, qoff


def workload_name(args, Q, T, nres, spacers=None):
    return ("%d synthetic spacers (25-45 nt, six-frame, %d frames) vs synthetic GenBank_phage_2018_09-shaped ORF setDB "
            "(%d genomes, %d ORFs, %d aa); predictmatch search parameters" % (spacers or args.spacers, Q, args.genomes, T, nres))


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# CPU baseline (oracle port)
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(args, tres, toff, qres, qoff, steps, warmup, sample_spacers, masked=None):
    """Times the CPU port (oracle/) of the same path: prefilter batch + alignment of its hit lists, OpenMP over
    queries with all host threads, on the first `sample_spacers`-worth of frames.  Index build is excluded (it is
    excluded on the GPU side too) and reported separately."""
    from oracle import pyoracle as O
    from spacepharer_b200 import matrices
    m80 = O.parse_matrix(matrices.matrix_text("VTML80.out"), 8.0, -0.2)
    m40u = O.parse_matrix(matrices.matrix_text("VTML40.out"), 2.0, -0.2)
    m40 = O.parse_matrix(matrices.matrix_text("VTML40.out"), 2.0, 0.0)
    cores = os.cpu_count() or 1
    Q = len(qoff) - 1
    frames_per_spacer = Q / float(args.spacers)
    nq = int(min(Q, max(1, round(sample_spacers * frames_per_spacer))))
    t0 = time.perf_counter()
    pf = O.Prefilter(masked if masked is not None else tres, toff, m80, m40u)
    build_s = time.perf_counter() - t0
    p = O.align_params(int(toff[-1]))
    sq, so = qres[: int(qoff[nq])], qoff[: nq + 1]
    times = []
    rows_acc = 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        off, hits = pf.batch(sq, so, threads=cores)
        rows = O.align_batch(m40, p, sq, so, tres, toff, off, hits["id"], threads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            rows_acc = int((rows["ok"] == 1).sum())
    pf.close()
    mean = float(np.mean(times))
    return {"value": nq / mean, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "first %d frames (%d spacers) of the workload, %d step(s); oracle index build %.1f s excluded"
                      % (nq, sample_spacers, len(times), build_s),
            "ms_per_step": mean * 1e3, "sample_queries": nq, "rows": rows_acc}


def reference_binary_run(args, tres, toff, qres, qoff, steps, warmup, single_thread=True):
    """Times the reference's own CPU implementation of the path: oracle/_ref/spacepharer (the unmodified reference, AVX2 +
    OpenMP build: `make -C oracle ref`) runs `prefilter` and `align` with the workflow's argv and all host threads on DBs written
    from the same seeds.  A step = the whole batch; its time = what the modules print themselves ('Time for processing' of
    prefilter minus 'Time for index table init' -- the GPU arm keeps its index resident too -- plus 'Time for processing' of
    align)."""
    from oracle import refbin
    cores = os.cpu_count() or 1
    Q = len(qoff) - 1
    wd = refbin.Workdir(tres, toff)
    try:
        wd.add_queries(qres, qoff)
        runs = []
        for it in range(warmup + steps):
            r = refbin.search(wd, threads=cores, tag="s%d" % it, mask=0 if args.no_mask else 1)
            if it >= warmup:
                runs.append(r)
        step = float(np.mean([r["step_s"] for r in runs]))
        out = {"value": Q / step, "unit": UNIT, "cores": cores, "kind": "reference", "ms_per_step": step * 1e3,
               "sample_queries": Q, "rows": runs[-1]["rows"],
               "prefilter_ms": float(np.mean([r["prefilter_s"] for r in runs])) * 1e3,
               "align_ms": float(np.mean([r["align_s"] for r in runs])) * 1e3,
               "index_build_s": float(np.mean([r["index_s"] for r in runs])),
               "sample": "the whole batch (%d frames, %d spacers), %d step(s) of `spacepharer prefilter` + `align --threads %d` "
                         "(reference binary, AVX2/OpenMP); module-reported processing time, index table init (%.1f s per run) excluded"
                         % (Q, args.spacers, len(runs), cores, float(np.mean([r["index_s"] for r in runs])))}
        if single_thread:
            nq = int(min(Q, max(1, round(500 * Q / float(args.spacers)))))
            wd.add_queries(qres[: int(qoff[nq])], qoff[: nq + 1], name="queries1")
            r1 = refbin.search(wd, queries="queries1", threads=1, tag="t1", mask=0 if args.no_mask else 1)
            out["single_thread"] = {"value": nq / r1["step_s"], "unit": UNIT, "sample_queries": nq,
                                    "note": "--threads 1 on the first %d frames, same target DB" % nq}
        return out
    finally:
        wd.close()


def auto_sample(cores):
    # the scalar port does roughly 350 frames/s/core; aim at ~10 s per CPU step
    return int(max(500, min(10000, 900 * cores)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refbin
    tres, toff, qres, qoff = workload(args)
    cores = os.cpu_count() or 1
    if refbin.available():
        # the reference itself, same --steps / --warmup as the GPU arm
        steps, warm = max(1, args.steps), max(0, args.warmup)
        r = reference_binary_run(args, tres, toff, qres, qoff, steps, warm)
        note = ("the unmodified reference binary (oracle/_ref/spacepharer, AVX2 + OpenMP) on this host's %d threads; each step = "
                "prefilter + align of the whole batch, module-reported processing time without the index table init" % cores)
    else:
        # fallback, said loudly: the reference build is missing from this snapshot -> scalar C port of the path
        sys.stderr.write("bench.py: oracle/_ref/spacepharer not found (make -C oracle ref); timing the oracle PORT instead\n")
        sample = args.cpu_sample_spacers or auto_sample(cores)
        sample = min(sample, args.spacers)
        masked = None
        if not args.no_mask:
            masked = oracle_mask(tres, toff)
        steps = max(1, min(args.steps, 3))
        warm = 1 if args.warmup > 0 else 0
        r = cpu_reference_run(args, tres, toff, qres, qoff, steps, warm, sample, masked)
        note = "FALLBACK: CPU port of the reference path (oracle/, scalar C + OpenMP) because the reference build is missing; each step = the bounded sample"
    Q = len(qoff) - 1
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args, Q, len(toff) - 1, int(toff[-1])), "note": note},
            "cpu_baseline": {k: r[k] for k in r if k in ("value", "unit", "cores", "kind", "sample", "prefilter_ms", "align_ms",
                                                        "index_build_s", "single_thread")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def oracle_mask(tres, toff):
    """tantan masking with the oracle's restatement (reference arm only)."""
    import ctypes as C
    from oracle import pyoracle as O
    from spacepharer_b200 import matrices
    m80 = O.parse_matrix(matrices.matrix_text("VTML80.out"), 8.0, -0.2)
    out = np.array(tres, dtype=np.uint8, copy=True)
    L = O.lib()
    base = out.ctypes.data
    for i in range(len(toff) - 1):
        a, b = int(toff[i]), int(toff[i + 1])
        L.spo_tantan_mask(C.byref(m80), C.cast(base + a, C.POINTER(C.c_ubyte)), b - a)
    return out


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from spacepharer_b200 import api, multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    weak = world > 1 and args.split_mode in ("query", "index")
    tres, toff, qres, qoff = workload(args, world if weak else 1)
    sq = so = None
    if args.strong_spacers > 0:
        from spacepharer_b200 import synth as _synth
        sq, so, _ = _synth.make_query_frames(args.strong_spacers, tres, toff)
    # product-level leg first: the host binary runs as its own process, before this one opens a CUDA context on the device
    # (two processes with live contexts on one GPU time-slice it, which would be charged to the binary)
    d2d = disk_to_disk(args, tres, toff, sq, so) if (world == 1 and args.strong_spacers > 0 and not args.no_disk_to_disk) else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world and rank == 0:
        print("bench.py: warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)

    T, Q = len(toff) - 1, len(qoff) - 1
    db_res = int(toff[-1])
    t0 = time.perf_counter()
    # one busy host thread per lane and rank: never more lanes than host cores per rank
    n_lanes = 1 if (world > 1 and args.split_mode == "target") else max(1, min(args.lanes, (os.cpu_count() or 1) // max(world, 1)))
    eng = multigpu.ShardedSearcher(device=local, rank=rank, world=world, split_mode=args.split_mode, lanes=n_lanes,
                                   shard_group=args.shard_group)
    masked, n_masked = (tres, 0) if args.no_mask else api.tantan_mask(eng.searcher.seed, tres, toff)
    mask_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ist = eng.load_targets(masked, tres, toff)
    build_s = time.perf_counter() - t0

    # pinned host copies of the query batch (what a caller hands to the public API)
    q_pin = torch.from_numpy(np.ascontiguousarray(qres)).pin_memory()
    o_pin = torch.from_numpy(np.ascontiguousarray(qoff).view(np.int64)).pin_memory()
    q_np, o_np = q_pin.numpy(), o_pin.numpy().view(np.uint64)
    stream = torch.cuda.ExternalStream(eng.searcher.stream_ptr(), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # three distinct batches of the same shape rotate through the lanes (value leg: lane l owns batch l mod 3; e2e leg: step
    # i searches batch i mod 3), so no timed step repeats the previous step's inputs
    from spacepharer_b200 import synth
    batches = [(q_np, o_np)]
    keep_pins = []
    for extra_seed in (1001, 1002):
        bq, bo, _ = synth.make_query_frames(args.spacers * (world if weak else 1), tres, toff, seed=extra_seed)
        bqp = torch.from_numpy(np.ascontiguousarray(bq)).pin_memory()
        bop = torch.from_numpy(np.ascontiguousarray(bo).view(np.int64)).pin_memory()
        keep_pins += [bqp, bop]
        batches.append((bqp.numpy(), bop.numpy().view(np.uint64)))
    batch_q = [eng.queries_per_step(len(o) - 1) for _, o in batches]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ("value") ----
    eng.set_queries(q_np, o_np)
    for _ in range(args.warmup):
        eng.step_resident(db_res)
    launches0 = eng.launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pk_ms, pk_bytes, sw_ms, sw_cells, pk_phase = [], [], [], [], []
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()          # L2 flush between timed iterations (256 MB > 126 MB L2), outside the events
            ev[k][0].record(stream)
        st = eng.step_resident(db_res)
        ev[k][1].record(stream)
        pk_ms.append(st["prefilter"]["kernel_ms"])
        pk_bytes.append(st["prefilter"]["algo_bytes"])
        pk_phase.append(st["prefilter"])
        sw_ms.append(st["align"]["kernel_ms"])
        sw_cells.append(st["align"]["n_cells"])
    barrier()
    wall_value = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = eng.launches() - launches0
    last = st
    single_ms = dev_ms
    single_launches = launches
    value_queries = eng.queries_per_step(Q) * args.steps
    phase = {k: float(np.mean([x[k] for x in pk_phase])) for k in ("expand_ms", "lookup_ms", "bookkeep_ms")}

    # ---- device-resident leg over the lanes: the whole-job throughput ("value") ----
    # K steps shared by n_lanes contexts (one host thread each, every lane holds the batch resident); one CUDA event per lane
    # stream before the first and after the last step, device time = latest end - earliest start; the per-step L2 flush runs
    # on the lane's stream INSIDE the timed span here
    def lanes_value_leg(engine, steps=None):
        """K steps shared by the lanes of `engine`, batches resident -> (device ms, wall s, own launches, queries searched)"""
        steps = steps or args.steps
        lanes = engine.lanes()
        dev = torch.device("cuda", local)
        ext = [torch.cuda.ExternalStream(ln.stream_ptr(), device=dev) for ln in lanes]
        fl = [flush] + [torch.empty(256 << 20, dtype=torch.uint8, device="cuda") for _ in lanes[1:]]

        def before_step(l, ln):
            with torch.cuda.stream(ext[l]):
                fl[l].zero_()

        engine.run_stream(batches, db_res, steps=max(args.warmup, 2) * len(lanes), resident=True, before_step=before_step)
        l0 = engine.launches()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        barrier()
        tw0 = time.perf_counter()
        for l in range(len(lanes)):
            e0[l].record(ext[l])
        vsts = engine.run_stream(batches, db_res, steps=steps, resident=True, before_step=before_step)
        nq_ = sum(batch_q[st["lane"] % len(batches)] for st in vsts)
        for l in range(len(lanes)):
            e1[l].record(ext[l])
        barrier()
        return max(a.elapsed_time(b) for a in e0 for b in e1), time.perf_counter() - tw0, engine.launches() - l0, nq_

    long_leg = None
    if n_lanes > 1:
        dev_ms, wall_value, launches, value_queries = lanes_value_leg(eng)
        if args.long_steps > args.steps:      # the same leg over a longer timed region (K steps are 30-40 ms of device time)
            lm, lw, _, lq = lanes_value_leg(eng, args.long_steps)
            long_leg = [lm, lq]
        eng.set_queries(q_np, o_np)

    # ---- end-to-end leg through the public API with HOST buffers ----
    # a stream of batches: every step uploads its queries from pinned host memory (H2D), searches, and reads its rows + CIGARs
    # back into pinned host buffers of the library (D2H); wall clock from the first upload to the last row landed
    d2h = 0
    if n_lanes > 1:
        eng.run_stream(batches, db_res, steps=2 * n_lanes)
        barrier()
        t0 = time.perf_counter()
        sts = eng.run_stream(batches, db_res, steps=args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_queries = sum(batch_q[i % len(batches)] for i in range(args.steps))
        d2h, n_rows_e2e = sts[-1]["d2h_bytes"], sts[-1]["rows"]
        e2e_how = ("%d lanes per GPU (contexts sharing one index, one host thread each): upload, kernels and read-back of different "
                   "batches overlap; all copies inside the timed region" % n_lanes)
    else:
        for _ in range(min(2, args.warmup)):
            eng.search_host(q_np, o_np, db_res, copy=False)["view"].close()
        warm = [eng.search_host_begin(q_np, o_np, db_res) for _ in range(max(2, min(3, args.warmup)))]
        for w in warm:
            w["view"].wait().close()
        barrier()
        t0 = time.perf_counter()
        pending = None

        def consume(p):
            v = p["view"].wait()
            n = (v.nbytes, len(v.rows))
            v.close()
            return n

        e2e_queries = eng.queries_per_step(Q) * args.steps
        for k in range(args.steps):
            cur = eng.search_host_begin(q_np, o_np, db_res)
            if pending is not None:
                d2h, n_rows_e2e = consume(pending)
            pending = cur
        d2h, n_rows_e2e = consume(pending)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_how = "D2H of step k on a second stream while step k+1 runs (spb_align_fetch_begin/_wait); all copies inside the timed region"
    clocks = sampler.stop() if rank == 0 else None

    # ---- exhaustive-search leg (config 5): a bounded sample of the batch against EVERY target, streaming SW kernel ----
    exh = None
    if args.exhaustive_spacers > 0:
        per = Q // world if weak else Q
        lo = rank * per if weak else 0
        nq = int(min(per, max(1, round(args.exhaustive_spacers * Q / float(args.spacers * (world if weak else 1))))))
        base = int(o_np[lo])
        eng.searcher.set_queries(q_np[base:int(o_np[lo + nq])], (o_np[lo:lo + nq + 1] - np.uint64(base)).astype(np.uint64))
        eng.searcher.pairs_exhaustive(db_res)
        barrier()
        runs = [eng.searcher.pairs_exhaustive(db_res) for _ in range(2)]
        barrier()
        exh = {"kernel_ms": float(np.mean([r["kernel_ms"] for r in runs])), "cells": float(runs[-1]["n_cells"]),
               "scored": float(runs[-1]["n_scored"]), "kept": float(runs[-1]["n_kept"]), "queries": nq}

    # ---- strong-scaling leg (BASELINE config 4): ONE job of 1 M spacers split over the GPUs ----
    # the whole job goes through the public API with host buffers (H2D of the frames, D2H of the sorted rows + CIGARs inside
    # the timed region), in chunks of ~125 k frames per GPU over the lanes; total work is fixed, so the N-GPU time against the
    # 1-GPU time is the strong-scaling curve of the split the engine uses
    # chunk size: at most 125 k frames, and small enough that every lane of every GPU sees two chunks (a job of 3.79 M frames
    # on 8 GPUs x 3 lanes in 125 k chunks would leave two lanes idle for the last third of the run)
    per_gpu = 125000
    if args.strong_spacers > 0:
        per_gpu = int(min(125000, max(32768, -(-(len(so) - 1) // (world * 2 * n_lanes)))))
    strong_chunks, SQ = None, 0
    if args.strong_spacers > 0:
        SQ = len(so) - 1
        sqp = torch.from_numpy(np.ascontiguousarray(sq)).pin_memory()
        sqn = sqp.numpy()
        so = np.ascontiguousarray(so, dtype=np.uint64)
        step_q = per_gpu * world
        strong_chunks = []
        for c0 in range(0, SQ, step_q):
            c1 = min(SQ, c0 + step_q)
            strong_chunks.append((sqn[int(so[c0]):int(so[c1])], (so[c0:c1 + 1] - so[c0]).astype(np.uint64)))

    def strong_leg(engine, mode, lanes_n):
        """-> [wall ms (max over ranks), rows, similar k-mers, expand, lookup, bookkeeping, prefilter kernels, align stage ms
        (summed over ranks)] or None"""
        if not strong_chunks:
            return None
        rows_total = [0]

        def run_chunks(lst):
            if world > 1 and mode == "target":
                pending, stats = None, []
                for cq, co in lst:
                    cur = engine.search_host_begin(cq, co, db_res)
                    stats.append(cur["stats"])
                    if pending is not None:
                        v = pending["view"].wait(); rows_total[0] += len(v.rows); v.close()
                    pending = cur
                v = pending["view"].wait(); rows_total[0] += len(v.rows); v.close()
                return stats
            sts_ = engine.run_stream(lst, db_res)
            rows_total[0] += sum(x["rows"] for x in sts_)
            return sts_

        run_chunks(strong_chunks[: max(2, 2 * lanes_n)])
        rows_total[0] = 0
        barrier()
        t0_ = time.perf_counter()
        ssts = run_chunks(strong_chunks)
        barrier()
        wall = time.perf_counter() - t0_
        vals = [float(rows_total[0]), float(sum(x["prefilter"]["n_kmers"] for x in ssts))]
        vals += [float(sum(x["prefilter"][k] for x in ssts)) for k in ("expand_ms", "lookup_ms", "bookkeep_ms", "kernel_ms")]
        vals.append(float(sum(x["align"]["total_ms"] for x in ssts)))
        tw = torch.tensor([wall * 1e3], dtype=torch.float64, device="cuda")
        ts = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        return [float(tw[0])] + [float(x) for x in ts]

    def strong_block(v, mode, group):
        if not v:
            return None
        return {"value": SQ / (v[0] * 1e-3), "unit": UNIT, "ms_total": v[0], "scaling": "strong",
                "workload": "BASELINE config 4: %d synthetic spacers = %d frames, ONE job split over %d GPU(s) (%s), chunks of %d frames "
                            "per GPU through the public API with pinned host buffers (H2D + D2H inside the timed region)"
                            % (args.strong_spacers, SQ, world, ("split_mode " + mode + (", shard groups of %d" % group if mode == "index" else ""))
                               if world > 1 else "single GPU", per_gpu),
                "rows": int(v[1]), "similar_kmers": int(v[2]),
                "stage_ms_summed_over_gpus": {"expand": v[3], "lookup": v[4], "bookkeeping": v[5], "prefilter_kernels": v[6], "align_stage": v[7]},
                "nvlink": ({"mode": "index split: table entries of the other GPUs' k-mer ranges are gathered with peer loads inside lookup_kernel",
                            "bytes_est": v[2] * 0.336 * (group - 1) / group * 32.0,
                            "how": "similar k-mers x 0.336 (share with an occupied table entry, measured on this index) x (G-1)/G remote (G = shard "
                                   "group) x one 32-byte sector per 16-byte entry; posting lists of >= 3 targets (5 % of the entries) add their own"}
                           if (world > 1 and mode == "index" and group > 1) else None)}

    # ---- parity of a configuration against the reference binary (config 2 digest) ----
    def parity_check(engine):
        import hashlib
        pq, po = (qres, qoff) if world == 1 else synth.make_query_frames(args.spacers, tres, toff)[:2]
        res = engine.search_host(np.ascontiguousarray(pq), np.ascontiguousarray(po, dtype=np.uint64), db_res)
        rows = res["rows"]
        lines = api.format_align_rows(rows, res["cigar"], np.diff(np.asarray(po, dtype=np.uint64)).astype(np.int64),
                                      np.diff(np.asarray(toff, dtype=np.uint64)).astype(np.int64))
        mine = (rows["query"].astype(np.int64).tolist(), lines)
        allparts = [None] * world
        if world > 1:
            dist.all_gather_object(allparts, mine)
        else:
            allparts = [mine]
        if rank != 0:
            return None
        ent = {}
        n_rows = 0
        for qs, ls in allparts:
            n_rows += len(ls)
            for q, l in zip(qs, ls):
                ent.setdefault(int(q), []).append(l)
        h = hashlib.sha256()
        for k in range(len(po) - 1):
            h.update(b"%d\n" % k)
            h.update("".join(x + "\n" for x in ent.get(k, [])).encode())
        out = {"aln_sha256": h.hexdigest(), "aln_rows": n_rows}
        try:
            want = json.load(open(os.path.join(ROOT, "tests", "golden", "config2_digest.json")))
            ok = (args.spacers == want.get("spacers", 10000) and args.genomes == 8000 and not args.no_mask and
                  out["aln_sha256"] == want["aln_sha256"] and n_rows == want["aln_rows"])
            out["reference"] = "tests/golden/config2_digest.json (SHA-256 of the reference binary's alignment DB, canonical form)"
            out["parity_ok"] = bool(ok)
        except Exception as e:  # noqa: BLE001
            out["parity_ok"] = None
            out["error"] = str(e)
        return out

    strong = strong_leg(eng, args.split_mode, n_lanes)
    parity = None if args.no_parity else parity_check(eng)
    # predictmatch searches the spacers twice: against the target setDB and against the reversed-ORF control setDB
    # (data/predictmatch.sh:78,96).  The same 1 M-spacer job against the control set, on a second engine (its own index, same split)
    strong_ctrl = None
    if strong and not args.no_control:
        ctres = synth.reverse_targets(tres, toff)
        cmasked = ctres if args.no_mask else api.tantan_mask(eng.searcher.seed, ctres, toff)[0]
        eng_c = multigpu.ShardedSearcher(device=local, rank=rank, world=world, split_mode=args.split_mode, lanes=n_lanes,
                                         shard_group=args.shard_group)
        eng_c.load_targets(cmasked, ctres, toff)
        strong_ctrl = strong_leg(eng_c, args.split_mode, n_lanes)
        eng_c.close()
        del ctres, cmasked

    # ---- the other way to spread the search over the GPUs, N > 1: value leg, strong-scaling leg and parity of that mode ----
    # main mode "query" (full index per GPU): the k-mer table sharded over groups of GPUs and gathered through NVLink peer memory
    # ("index"); main mode "index" / "target": full replicas ("query")
    other = None
    if world > 1 and n_lanes > 1:
        omode = "index" if args.split_mode == "query" else "query"
        ogroup = (args.shard_group or min(2, world)) if omode == "index" else 0
        eng_o = multigpu.ShardedSearcher(device=local, rank=rank, world=world, split_mode=omode, lanes=n_lanes, shard_group=ogroup)
        eng_o.load_targets(masked, tres, toff)
        o_ms, _, _, o_q = lanes_value_leg(eng_o)
        o_strong = strong_leg(eng_o, omode, n_lanes)
        o_parity = None if args.no_parity else parity_check(eng_o)
        to = torch.tensor([o_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(to, op=dist.ReduceOp.MAX)
        other = {"mode": omode, "group": ogroup, "ms": float(to[0]), "queries": o_q, "strong": o_strong, "parity": o_parity}
        eng_o.close()

    t = torch.tensor([dev_ms, e2e_s * 1e3, exh["kernel_ms"] if exh else 0.0, single_ms, long_leg[0] if long_leg else 0.0],
                     dtype=torch.float64, device="cuda")
    tsum = torch.tensor([exh["cells"], exh["scored"], exh["kept"], exh["queries"]] if exh else [0.0] * 4, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, exh_ms, single_ms, long_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4])
    q_total = eng.queries_per_step(Q)
    eng_group = eng.shard_group
    eng.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic = ncu_traffic()
        kms = float(np.mean(pk_ms))
        achieved = float(np.mean(pk_bytes)) / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value_queries / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, Q, T, db_res, args.spacers * (world if weak else 1)), "queries_per_step": q_total,
                       "split_mode": args.split_mode if world > 1 else "none",
                       "shard_group": (eng_group if (world > 1 and args.split_mode == "index") else None),
                       "lanes": n_lanes,
                       "l2": "256 MB memset on the lane's stream before every step (inside the timed span with lanes > 1, outside the "
                             "per-step events of the single-lane segment); index %.2f GB > L2" % (ist["bytes_device"] / 1e9),
                       "timing": ("CUDA events on the lane streams: latest end - earliest start over %d steps shared by %d lanes, max over ranks"
                                  % (args.steps, n_lanes)) if n_lanes > 1 else
                                 "sum of per-step CUDA-event intervals on the library stream, max over ranks",
                       "index": "resident (built once: %.1f ms device, %.2f s wall incl. H2D; tantan host pre-step %.2f s, %d residues masked)"
                                % (ist["build_ms"], build_s, mask_s, n_masked)},
            "e2e": {"value": e2e_queries / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(q_np.nbytes + o_np.nbytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "pipeline": e2e_how},
            "gpu_launches": int(launches),
            # the prefilter = three kernel classes (expand -> lookup -> bookkeeping); `achieved` follows the contract: ALGORITHMIC
            # bytes of the whole prefilter over the summed duration of its kernels.  The lookup kernel is the
            # HBM-bound one: its DRAM traffic over its own duration is reported beside it (dram_*), from the committed ncu capture.
            "roofline": dict({"bound": "hbm", "kernel": "expand_kernel + expand_heavy_kernel + lookup_kernel + bookkeep_kernel<1|8> (prefilter)",
                              "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak, "traffic": traffic.get("prefilter_dram_bytes_per_launch"), "peak_source": peak_src,
                              "algo_bytes_per_launch": float(np.mean(pk_bytes)), "kernel_ms": kms,
                              "share_of_step": kms / (single_ms / args.steps),
                              "kernels_ms": phase,
                              "measured": "single-lane segment (%d steps, the kernels alone on the device, CUDA events on their stream)" % args.steps},
                             **({"lookup": {"kernel_ms": phase["lookup_ms"],
                                            "dram_bytes_per_launch": traffic.get("lookup_kernel_dram_bytes_per_launch"),
                                            "dram_gbs": traffic["lookup_kernel_dram_bytes_per_launch"] / (phase["lookup_ms"] * 1e-3) / 1e9,
                                            "dram_frac_of_peak": traffic["lookup_kernel_dram_bytes_per_launch"] / (phase["lookup_ms"] * 1e-3) / 1e9 / peak,
                                            "note": "bytes from the committed ncu --set full capture of the same workload (profiles/), "
                                                    "duration measured live"}}
                                if traffic.get("lookup_kernel_dram_bytes_per_launch") and phase["lookup_ms"] > 0 else {})),
            "single_lane": {"value": q_total * args.steps / (single_ms * 1e-3), "ms_per_step": single_ms / args.steps,
                            "own_launches_per_step": single_launches / float(args.steps)},
            "sw": sw_block(float(np.mean(sw_cells)), float(np.mean(sw_ms)), last["align"], clocks),
            "clocks": clocks,
            "wall_s_value_leg": wall_value,
        }
        main_group = eng_group if (world > 1 and args.split_mode == "index") else 0
        if long_leg and long_ms > 0:
            line["steady_state"] = {"value": long_leg[1] / (long_ms * 1e-3), "unit": UNIT, "steps": args.long_steps,
                                    "ms_per_step": long_ms / args.long_steps,
                                    "note": "the value leg again over %d steps (%.0f ms of device time instead of %.0f)" % (args.long_steps, long_ms, dev_ms)}
        if strong:
            line["strong_scaling"] = strong_block(strong, args.split_mode, main_group)
            if strong_ctrl:
                line["strong_scaling"]["control"] = {"ms_total": strong_ctrl[0], "rows": int(strong_ctrl[1]), "value": SQ / (strong_ctrl[0] * 1e-3),
                                                     "workload": "the same spacers against the reversed-ORF control setDB (second search of predictmatch)"}
                both = strong[0] + strong_ctrl[0]
                line["strong_scaling"]["target_plus_control"] = {"ms_total": both, "value": SQ / (both * 1e-3), "unit": UNIT,
                                                                 "note": "both searches of predictmatch for the 1 M spacers; spacers/s = %.0f" % (args.strong_spacers / (both * 1e-3))}
        if parity:
            line["parity"] = parity
        if d2d:
            line["disk_to_disk"] = d2d
        if other:
            blk = {"split_mode": other["mode"], "shard_group": other["group"] or None,
                   "value": other["queries"] / (other["ms"] * 1e-3), "unit": UNIT, "ms_per_step": other["ms"] / args.steps,
                   "strong_scaling": strong_block(other["strong"], other["mode"], other["group"]), "parity": other["parity"],
                   "note": ("the k-mer table (direct table + posting lists) sharded over groups of %d GPUs, one k-mer range per GPU, the other "
                            "ranges gathered by lookup_kernel with loads on NVLink peer-mapped pointers; queries partitioned.  Measured on "
                            "8 x B200 (profiles/): groups of 2 / 4 / 8 reach 0.78 / 0.48 / 0.18 of the replica throughput -- remote 16-byte "
                            "reads through NVSwitch do not scale with the number of peers, so the default shards over pairs" % other["group"])
                           if other["mode"] == "index" else
                           "the whole index replicated on every GPU, queries partitioned: no NVLink traffic on the data path"}
            line["index_split" if other["mode"] == "index" else "query_split"] = blk
        if exh:
            line["sw_exhaustive"] = {
                "gcups": float(tsum[0]) / (exh_ms * 1e-3) / 1e9 if exh_ms > 0 else None, "kernel_ms": exh_ms,
                "roofline": {"bound": "integer pipe (DPX s16x2)", "alu_instr_per_cell_pair": 7,
                             "peak_gcups": world * 148 * 4 * 16 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6 * 2 / 7.0 / 1e9,
                             "frac": (float(tsum[0]) / (exh_ms * 1e-3) / 1e9) / (world * 148 * 4 * 16 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6 * 2 / 7.0 / 1e9)
                             if exh_ms > 0 else None},
                "cells": float(tsum[0]), "pairs_scored": int(tsum[1]), "pairs_kept": int(tsum[2]),
                "workload": "config 5 (--exhaustive-search, no k-mer stage): %d frames (%d spacers per GPU) x all %d ORFs, "
                            "sw_exhaustive_kernel; cells = sum qLen*tLen, time = max over ranks of the kernel's CUDA-event time, mean of 2 runs"
                            % (int(tsum[3]), args.exhaustive_spacers, T)}
        if not args.no_cpu_baseline and world == 1:
            from oracle import refbin
            cores = os.cpu_count() or 1
            if refbin.available():   # the reference binary itself, one pass over the whole batch (~10 s of CPU work incl. its index build)
                r = reference_binary_run(args, tres, toff, qres, qoff, 1, 0, single_thread=False)
            else:
                sys.stderr.write("bench.py: oracle/_ref/spacepharer not found (make -C oracle ref); cpu_baseline falls back to the oracle PORT\n")
                sample = min(args.spacers, args.cpu_sample_spacers or auto_sample(cores))
                r = cpu_reference_run(args, tres, toff, qres, qoff, 1, 0, sample, masked if not args.no_mask else None)
            line["cpu_baseline"] = {k: r[k] for k in r if k in ("value", "unit", "cores", "kind", "sample", "prefilter_ms", "align_ms",
                                                                "index_build_s")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def disk_to_disk(args, tres, toff, sq, so):
    """Product level: the C++ host binary on BASELINE config 4 from on-disk MMseqs2 DBs to the on-disk alignment DB -- DB read,
    tantan, index build, prefilter + align on the GPU, text formatting of every row, DB write (`spb200 search`, the fused
    prefilter + align modules with the workflow's argv).  Two runs (the first one also pays the cold page cache and driver); the
    phases are the binary's own wall clock (SPB_TIMING)."""
    import shutil
    import subprocess as sp
    import tempfile
    from spacepharer_b200 import dbio
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_host_cli import ALN_FLAGS, BIN, PREF_FLAGS
    if not os.path.exists(BIN):
        return {"error": "spacepharer_b200/lib/spb200 not built"}
    w = tempfile.mkdtemp(prefix="spb_d2d_")
    try:
        dbio.write_sequence_db_arrays(os.path.join(w, "targets"), tres, toff)
        dbio.write_sequence_db_arrays(os.path.join(w, "queries"), sq, so)
        threads = str(min(16, os.cpu_count() or 1))
        seen, merged = set(), []
        for lst in (PREF_FLAGS, ALN_FLAGS):
            for i in range(0, len(lst), 2):
                if lst[i] not in seen:
                    seen.add(lst[i])
                    merged += [lst[i], threads if lst[i] == "--threads" else lst[i + 1]]
        runs = []
        for it in range(2):
            t0 = time.perf_counter()
            r = sp.run([BIN, "search", w + "/queries", w + "/targets", w + "/aln"] + merged + ["--batch", "250000"], capture_output=True,
                       text=True, env=dict(os.environ, SPB_TIMING="1"))
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return {"error": (r.stderr or r.stdout)[-300:]}
            ph = {}
            for l in r.stderr.split("\n"):
                if "[spb200 timing]" in l:
                    name, val = l.split("]", 1)[1].rsplit(None, 2)[0].strip(), float(l.split()[-2])
                    ph[name] = val
            runs.append({"wall_s": dt, "phases_s": ph})
        best = min(runs, key=lambda x: x["wall_s"])
        n_rows = sum(1 for _ in open(w + "/aln.index"))
        return {"value": (len(so) - 1) / best["wall_s"], "unit": UNIT, "wall_s": best["wall_s"], "runs": runs, "host_threads": int(threads),
                "entries_written": n_rows, "aln_db_bytes": os.path.getsize(w + "/aln"),
                "workload": "BASELINE config 4 through the host binary: `spb200 search queries targets aln` (workflow argv, --batch 250000), "
                            "%d spacers = %d frames vs %d ORFs, on-disk DBs in, on-disk alignment DB out; wall clock of the whole process"
                            % (args.strong_spacers, len(so) - 1, len(toff) - 1)}
    finally:
        shutil.rmtree(w, ignore_errors=True)


def ncu_traffic():
    """DRAM bytes per launch of the prefilter kernels from the committed ncu capture (profiles/traffic.json), or {}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return {k: float(v) for k, v in json.load(open(p)).items() if isinstance(v, (int, float))}
    except Exception:
        return {}


def sw_block(cells, ms, aln, clocks):
    """Smith-Waterman forward kernel: GCUPS and its integer-pipe roofline.  One packed DPX / SIMD-video instruction updates two
    cells; per row of a column pair the kernel issues 7.5 of them on the ALU pipe (PRMT, VIADD.16x2, 3 x VIADDMNMX.S16x2, VIMNMX for H
    and half a VIMNMX3 for the column-maximum key; the key's IMAD runs on the FMA pipe beside them).  The ALU pipe takes one warp
    instruction every two cycles per scheduler (16 lanes / clock), so the ceiling is SMs x 4 schedulers x 16 lanes x f_clk x 2 cells /
    7.5 instructions.  (sw_exhaustive_kernel, 7 ALU instructions per cell pair, measures 4850 GCUPS = 0.91 of the same formula with 7,
    and ncu reports its ALU pipe 89-90 % busy: profiles/r1j_exhaustive_kernels.csv.)"""
    f = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * 4 * 16 * f * 1e6 * 2 / 7.5 / 1e9
    g = cells / (ms * 1e-3) / 1e9 if ms > 0 else None
    return {"gcups": g, "kernel_ms": ms, "cells_per_step": cells, "pairs": aln["n_pairs"], "rows": aln["n_rows"],
            "roofline": {"bound": "integer pipe (DPX s16x2)", "peak_gcups": peak, "frac": (g / peak) if g else None,
                         "alu_instr_per_cell_pair": 7.5, "alu_lanes_per_clock_per_sm": 64, "sm_mhz": f,
                         "note": "padded rows (queries shorter than the kernel's row count) are not counted as cells"}}


_JSON_FD = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was moved to stderr"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    # keep stdout clean for the JSON line: native libraries (NCCL prints its version banner there) write to fd 1 directly
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
