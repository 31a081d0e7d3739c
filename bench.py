#!/usr/bin/env python
"""bench.py -- ADCensus throughput on 1/2/4/8 B200 (BASELINE.json metric: Mpix*disp/s, 1920x1080, D=192).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE config C3 (the one the metric is quoted on): a batch of 64 synthetic 1080p RGB stereo
pairs (synth_v1), D = 0..192 (Dn = 193 cost planes), sharded over the ranks: rank r owns frames r, r+N, ...
One step = one pass of the whole ADCensus path over the batch.  No data-path collective (frames are independent;
SURVEY 8(e)); the batch is fixed, so scaling is "strong".  --config selects another BASELINE configuration
(one JSON line with the same keys, config.workload names it):
    C1  demo pair 0600, 1280x720, D=0..192        C2  Motorcycle, 1482x994, D=0..256
    C4  fused EpipolarRectify remap + ADCensus on a 2x1280x1024 side-by-side frame, D=0..128
    C5  synthetic gray 3840x2160, D=0..384 (12.8 GB of cost volume per view)

  value : Mpix*disp/s = frames*H*W*Dn / t / 1e6, inputs resident in HBM, device-timed (CUDA events on a stream
          ordered around the work streams, max over ranks).
  e2e   : same metric through the public operator (ADCensus.enqueue / wait over the C-ABI) with HOST buffers:
          pinned staging, H2D, kernels, D2H inside the timed region, over all --steps.
  roofline     : dominant kernel (the aggregation walk; algorithmic bytes = 8 B/cell per launch: every cell of both
                 views read once and written once) vs the measured HBM copy peak, timed live with CUDA events on the
                 stream the kernel runs on.
  cpu_baseline : the reference's own CPU implementation on this box's host cores (oracle/_ref when present, else
                 the oracle port), bounded sample: a full-width row STRIPE of frame 0, not a whole frame.
  outputs_sha256 : digest over the disparity maps of global frames 0..7 in frame order; identical at every --gpus N
                 (sharding does not change bits).

--impl reference times ONLY the CPU reference arm (rank 0; other ranks exit 0) on a >= 200-row full-width stripe
with all host threads (torchrun's OMP_NUM_THREADS=1 is overridden explicitly).
Only this file's cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

IN_FLIGHT = 6          # pairs in flight per GPU (measured at C3, same call: 2 -> 16.37 ms per pair, 4 -> 15.91, 6 -> 15.70) ...
IN_FLIGHT_BIG = 4      # ... unless one context's cost volumes exceed 12 GB (C5: 25 GB each)
DISTINCT = 8  # distinct frames per rank; the batch cycles through them
UNIT = "Mpix*disp/s"
HASH_FRAMES = 8  # global frames 0..7 are hashed into outputs_sha256


class Workload:
    """One BASELINE configuration: geometry, inputs, frames per step."""

    def __init__(self, name: str):
        self.name = name
        self.rectify = False
        self.gray = False
        if name == "C1":
            self.H, self.W, self.maxd, self.batch = 720, 1280, 192, 16
            self.desc = "C1: demo pair 0600 (1280x720 RGB, real), D=0..192 (Dn=193); the pair repeated"
            self.real = "pair_c1_0600_720p.npz"
        elif name == "C2":
            self.H, self.W, self.maxd, self.batch = 994, 1482, 256, 8
            self.desc = "C2: Motorcycle (1482x994 RGB, real), D=0..256 (Dn=257); the pair repeated"
            self.real = "pair_c2_motorcycle.npz"
        elif name == "C3":
            self.H, self.W, self.maxd, self.batch = 1080, 1920, 192, 64
            self.desc = "C3: synthetic 1920x1080 RGB stereo pairs (synth_v1, seeds 1000+i), D=0..192 (Dn=193)"
            self.real = None
        elif name == "C4":
            self.H, self.W, self.maxd, self.batch = 1024, 1280, 128, 16
            self.desc = ("C4: fused EpipolarRectify remap + ADCensus on synthetic side-by-side 2x1280x1024 frames "
                         "(synth_v1 seeds 2000+i, synthetic rectify maps), D=0..128 (Dn=129)")
            self.real = None
            self.rectify = True
        elif name == "C5":
            self.H, self.W, self.maxd, self.batch = 2160, 3840, 384, 4
            self.desc = "C5: synthetic gray 3840x2160 pairs (synth_v1 seeds 3000+i), D=0..384 (Dn=385)"
            self.real = None
            self.gray = True
        else:
            raise SystemExit(f"unknown --config {name}")
        self.Dn = self.maxd + 1
        self.metric = f"Mpix*disp/s ADCensus {self.W}x{self.H} D={self.maxd}"
        self.seed0 = {"C3": 1000, "C4": 2000, "C5": 3000}.get(name, 0)

    def frame(self, i: int):
        """Global frame i: (left, right) uint8 HxWx3, or for C4 the side-by-side frame Hx2Wx3."""
        from tea_stereo_matching_b200.synth import synth_v1

        if self.real:
            z = np.load(ROOT / "tests" / "golden" / self.real)
            return np.ascontiguousarray(z["left"]), np.ascontiguousarray(z["right"])
        l, r = synth_v1(self.H, self.W, self.maxd, seed=self.seed0 + i, gray=self.gray)
        if self.rectify:
            return (np.ascontiguousarray(np.concatenate([l, r], axis=1)),)
        return l, r

    def n_distinct(self, n_mine: int) -> int:
        if self.real:
            return 1
        return max(1, min(DISTINCT if self.name != "C5" else 2, n_mine))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def load_cpu_reference():
    """The reference's CPU implementation with ALL host threads: launchers such as torchrun export
    OMP_NUM_THREADS=1 to their workers, so the thread count is set explicitly through the library."""
    import oracle

    impl, kind = None, "port"
    try:
        if oracle.REF_SO.exists():
            impl, kind = oracle.Ref(), "reference"
    except Exception as e:  # pragma: no cover
        log("reference .so unusable:", e)
    if impl is None:
        impl = oracle.Port()
    impl.set_threads(host_cores())
    return impl, kind


def cpu_run(impl, kind, left, right, maxd):
    """One ADCensus::compute of the CPU implementation; returns seconds."""
    if kind == "reference":
        _, sec = impl.compute(left, right, maxd)  # the reference's public entry, as shipped (all OpenMP threads)
        return sec
    t0 = time.perf_counter()
    impl.compute(left, right, maxd)
    return time.perf_counter() - t0


def cpu_sample_rows(impl, kind, wl, frame, budget_s: float, runs: int) -> int:
    """Rows of the full-width stripe so that `runs` runs fit in about budget_s; never fewer than 200 rows."""
    left, right = frame
    probe = 48
    t = cpu_run(impl, kind, np.ascontiguousarray(left[:probe]), np.ascontiguousarray(right[:probe]), wl.maxd)
    per_row = t / probe
    rows = int(budget_s / max(runs, 1) / max(per_row, 1e-9))
    return int(min(wl.H, max(200, rows)))


def cpu_frame(wl):
    """Frame 0 as the CPU arm sees it (C4: the rectified pair is what ADCensus::compute gets; the remap is < 0.1 % of
    the CPU time and is left out of the CPU sample)."""
    f = wl.frame(0)
    if wl.rectify:
        return np.ascontiguousarray(f[0][:, : wl.W]), np.ascontiguousarray(f[0][:, wl.W:])
    return f


def reference_arm(args, wl):
    impl, kind = load_cpu_reference()
    threads, cores = impl.threads, host_cores()
    assert threads == cores, f"CPU arm would run on {threads} threads of {cores} cores"
    frame = cpu_frame(wl)
    runs = args.steps + args.warmup
    rows = cpu_sample_rows(impl, kind, wl, frame, budget_s=150.0, runs=runs)
    left, right = np.ascontiguousarray(frame[0][:rows]), np.ascontiguousarray(frame[1][:rows])
    for _ in range(args.warmup):
        cpu_run(impl, kind, left, right, wl.maxd)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_run(impl, kind, left, right, wl.maxd)
    cells = rows * wl.W * wl.Dn * args.steps
    value = cells / t / 1e6
    sample = (f"full-width STRIPE, rows 0-{rows - 1} of frame 0 ({wl.W}x{rows} of {wl.W}x{wl.H}, D=0..{wl.maxd}); "
              f"{args.steps} timed runs of ADCensus::compute on {threads} OpenMP threads")
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "real" if wl.real else "synthetic",
        "config": {"workload": wl.desc, "sample": sample, "sample_rows": rows, "cpu_threads": threads, "host_cores": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--batch", type=int, default=0, help="frames per step over all ranks (default: 64 for C3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=0,
                    help="stereo pairs in flight per GPU (contexts / streams); 0 = 6, or 4 when a context's volumes exceed 12 GB")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = Workload(args.config)
    if args.batch > 0:
        wl.batch = args.batch
    H, W, MAXD, DN = wl.H, wl.W, wl.maxd, wl.Dn

    if args.impl == "reference":
        if rank != 0:
            return 0
        reference_arm(args, wl)
        return 0

    import torch
    import torch.distributed as dist

    import tea_stereo_matching_b200 as tsm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the ADCensus path has no CPU fallback")
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    n_gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MAX)  # noqa: E731
    sum_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.SUM)  # noqa: E731

    # frames of this rank: global frame i -> rank i % N
    from tea_stereo_matching_b200.sharding import frames_for_rank

    my_frames = frames_for_rank(wl.batch, world, rank)
    n_distinct = wl.n_distinct(len(my_frames))
    t0 = time.time()
    frames = [wl.frame(my_frames[j]) for j in range(n_distinct)] if my_frames else []
    log(f"[rank {rank}] {wl.name}: {n_distinct} distinct frames ready in {time.time() - t0:.1f}s; "
        f"{len(my_frames)} frames per step on this rank")

    # IN_FLIGHT contexts (= streams, arenas) per GPU: frames alternate between them so the small serial refinement
    # kernels of one pair overlap with the bandwidth kernels of the others.
    NCTX = args.in_flight if args.in_flight > 0 else (IN_FLIGHT if 2 * 4 * wl.H * wl.W * (MAXD + 2) <= 12e9 else IN_FLIGHT_BIG)
    streams = [torch.cuda.Stream() for _ in range(NCTX)]

    def new_matcher(stream=None):
        m = tsm.ADCensus(device=local_rank, stream=stream)
        m.setMatchingStrategy(tsm.ColorModel.RGB, False, False)
        m.setMinMaxDisparity(0, MAXD)
        return m

    matchers = [new_matcher(st.cuda_stream) for st in streams]
    matcher, stream = matchers[0], streams[0]
    ctx = matcher.context
    rect = None
    if wl.rectify:
        from tea_stereo_matching_b200.synth import convert_maps_fixed, synth_rectify_maps

        (mx0, my0), (mx1, my1) = synth_rectify_maps(H, W)
        f0, f1 = convert_maps_fixed(mx0, my0), convert_maps_fixed(mx1, my1)
        rect = tsm.EpipolarRectify(tsm.EpipolarRectifyMap(map00=f0[0], map01=f0[1], map10=f1[0], map11=f1[1]), (W, H))

    d_frames = [tuple(torch.from_numpy(a).cuda() for a in f) for f in frames]
    d_out = [torch.empty((H, W), dtype=torch.float32, device="cuda") for _ in range(max(n_distinct, NCTX, 1))]
    torch.cuda.synchronize()

    def compute_device(m, j, out):
        f = d_frames[j % n_distinct]
        if wl.rectify:
            rect.rectify_adcensus_device(f[0].data_ptr(), 6 * W, m, out.data_ptr())
        else:
            m.compute_device(f[0].data_ptr(), f[1].data_ptr(), H, W, out.data_ptr())

    def step_device():
        for j in range(len(my_frames)):
            compute_device(matchers[j % NCTX], j, d_out[j % len(d_out)])

    def launches_total():
        return sum(m.context.launch_count for m in matchers)

    # ---- outputs_sha256: global frames 0..HASH_FRAMES-1, hashed where they are computed, ordered on rank 0 ----
    digests = {}
    for j, gi in enumerate(my_frames):
        if gi < HASH_FRAMES and j < n_distinct:
            with torch.cuda.stream(stream):
                compute_device(matcher, j, d_out[0])
            stream.synchronize()
            digests[gi] = hashlib.sha256(d_out[0].cpu().numpy().tobytes()).hexdigest()
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, digests)
        digests = {k: v for p in parts for k, v in p.items()}
    h = hashlib.sha256()
    for gi in sorted(digests):
        h.update(f"{gi}:{digests[gi]};".encode())
    outputs_sha256 = h.hexdigest()

    # ---- value: HBM-resident, device timed ----
    # events on a timing stream that is ordered after / before all work streams
    tstream = torch.cuda.Stream()

    def fence_all(on):  # make `on` wait for everything enqueued on the work streams
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            on.wait_event(ev)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = launches_total()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for st in streams:
        st.wait_event(e0)
    for _ in range(args.steps):
        step_device()
    fence_all(tstream)
    e1.record(tstream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(sum_over_ranks(float(launches_total() - launches0)))

    total_cells = float(wl.batch) * H * W * DN
    value = total_cells * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- per-stage and per-launch times (profiled frames, same stream, nothing else on the GPU) ----
    stage_acc: dict = {}
    if my_frames:
        ctx.set_profiling(True)
        with torch.cuda.stream(stream):
            nprof = 3
            for j in range(nprof):
                compute_device(matcher, j, d_out[0])
                for k, v in ctx.stage_times().items():
                    stage_acc[k] = stage_acc.get(k, 0.0) + v / nprof
        ctx.set_profiling(False)
    stages = {k: v for k, v in stage_acc.items() if "/" not in k}
    launches_ms = {k: stage_acc[k] / stage_acc[k + "#n"] for k in stage_acc if "/" in k and not k.endswith("#n")}
    launch_counts = {k: int(round(stage_acc[k + "#n"])) for k in launches_ms}
    peak, peak_src = measured_peak_gbs()
    cells_pair = 2.0 * H * W * DN  # both views
    pair_ms = sum(stages.values())

    def rl(name, bytes_per_cell):
        ms = launches_ms.get(name)
        if not ms:
            return None
        a = bytes_per_cell * cells_pair / (ms * 1e-3) / 1e9
        return {"achieved": a, "frac": a / peak, "avg_launch_ms": ms, "launches_per_pair": launch_counts[name],
                "algorithmic_bytes_per_cell": bytes_per_cell}

    # every aggregation launch reads and writes each cell of both views once (8 B/cell), whether it applies one
    # aggregation1D pass or two fused ones; the dominant kernel is the one with the largest share of the pair
    agg = {k: rl(k, 8.0) for k in launches_ms if k.startswith("aggregate/")}
    dom = max(agg, key=lambda k: agg[k]["avg_launch_ms"] * agg[k]["launches_per_pair"]) if agg else None
    # DRAM bytes per launch of the aggregation kernels from the committed `ncu --set full` capture (profiles/), if any
    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    if tfile.exists() and wl.name == "C3":
        tj = json.loads(tfile.read_text())
        traffic = tj.get("bytes_per_launch", {}).get(dom) if dom else None
        if traffic is None:
            traffic = tj.get("k_agg_persist_bytes_per_launch")
        traffic_src = tj.get("source")
    others = {k: v for k, v in agg.items() if k != dom}
    # scanline: the vertical launch (down + up) moves 16 B/cell; the horizontal one (right + left, WTA fused) 14 B/cell
    # because the right view's last store is skipped
    for name, bpc in (("scanline/vertical", 16.0), ("scanline/horizontal", 14.0)):
        if rl(name, bpc):
            others[name] = rl(name, bpc)
    if pair_ms > 0:
        others["pipeline (SURVEY 8(d): 120 B/cell end to end)"] = {
            "achieved": 60.0 * cells_pair / (pair_ms * 1e-3) / 1e9, "frac": 60.0 * cells_pair / (pair_ms * 1e-3) / 1e9 / peak,
            "ms_per_pair_one_stream": pair_ms}
    if "rectify" in stages:  # C4: both remaps of a frame, 12 B/pixel/view (3 B source + 6 B maps + 3 B output)
        a = 2 * 12.0 * H * W / (stages["rectify"] * 1e-3) / 1e9
        others["k_remap x2 (12 B/pixel/view)"] = {"achieved": a, "frac": a / peak, "avg_launch_ms": stages["rectify"] / 2}
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": agg[dom]["achieved"] if dom else None, "peak": peak, "unit": "GB/s",
        "frac": agg[dom]["frac"] if dom else None, "traffic": traffic, "traffic_kind": "static ncu capture (profiles/)",
        "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": 8.0 * cells_pair, "avg_launch_ms": agg[dom]["avg_launch_ms"] if dom else None,
        "launches_per_pair": agg[dom]["launches_per_pair"] if dom else None,
        "aggregate_stage": {"ms_per_pair": stages.get("aggregate"),
                            "algorithmic_GB_per_pair": sum(8.0 * cells_pair * launch_counts[k] for k in agg) / 1e9},
        "other_kernels": others,
    }

    # ---- e2e: public operator, host buffers, NCTX pairs in flight per GPU, all --steps ----
    for m in matchers:  # free the arenas of the device-resident leg (C5: 26 GB each)
        m.context.close()
    m2 = [new_matcher() for _ in range(NCTX)]
    h_out = [np.empty((H, W), np.float32) for _ in range(NCTX)]

    def run_e2e(nsteps):
        """nsteps steps as ONE stream of pairs: every pair is copied in from host memory and its map copied back;
        the pipeline (NCTX pairs in flight) is not drained between steps, only at the end of the timed region."""
        n = len(my_frames) * nsteps
        for j in range(n + NCTX):
            if NCTX <= j < n + NCTX:  # pair j - NCTX ran on the context pair j is about to use
                m2[j % NCTX].wait(h_out[j % NCTX])
            if j < n:
                f = frames[j % n_distinct]
                if wl.rectify:
                    rect.rectify_adcensus_enqueue(f[0], m2[j % NCTX])
                else:
                    m2[j % NCTX].enqueue(f[0], f[1])

    run_e2e(1)  # warm (arena + pinned allocations)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_cells * args.steps / e2e_s / 1e6
    in_bytes = H * 2 * W * 3
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(wl.batch * in_bytes),
           "d2h_bytes_per_step": int(wl.batch * H * W * 4), "ms_per_step": 1e3 * e2e_s / args.steps, "steps": args.steps,
           "in_flight_per_gpu": NCTX}

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            impl, kind = load_cpu_reference()
            cf = cpu_frame(wl)
            rows = cpu_sample_rows(impl, kind, wl, cf, budget_s=20.0, runs=1)
            sec = cpu_run(impl, kind, np.ascontiguousarray(cf[0][:rows]), np.ascontiguousarray(cf[1][:rows]), MAXD)
            cpu_baseline = {"value": rows * W * DN / sec / 1e6, "unit": UNIT, "cores": impl.threads, "kind": kind,
                            "sample": f"full-width STRIPE, rows 0-{rows - 1} of frame 0 ({W}x{rows} of {W}x{H}, D=0..{MAXD}), "
                                      f"1 run, {sec:.1f}s"}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}

    if rank == 0:
        line = {
            "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "ms_per_frame": dev_ms / args.steps / max(1, len(my_frames)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "real" if wl.real else "synthetic",
            "config": {"workload": wl.desc, "frames_per_step": wl.batch, "frames_per_gpu": len(my_frames),
                       "distinct_frames_per_gpu": n_distinct,
                       "parallelism": f"frame-sharded x{n_gpus}, no data-path collective; {NCTX} pairs in flight per GPU ({NCTX} streams)",
                       "l2": f"working set {8.0 * H * W * DN / 1e9:.1f} GB per frame >> 126 MB L2 (no flush needed)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "stages_ms": stages, "launches_ms": launches_ms, "outputs_sha256": outputs_sha256,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
