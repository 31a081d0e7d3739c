#!/usr/bin/env python
"""bench.py -- ADCensus 1920x1080 D=0..192 throughput on 1/2/4/8 B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config C3): a batch of 64 synthetic 1080p RGB stereo pairs
(synth_v1), D = 0..192 (Dn = 193 cost planes), sharded over the ranks: rank r owns
frames r, r+N, ...  One step = one pass of the whole ADCensus path over the batch.
No data-path collective (frames are independent); scaling is "strong".

  value : Mpix*disp/s = 64*H*W*Dn / t / 1e6, inputs resident in HBM, device-timed
          (CUDA events on the stream the kernels run on, max over ranks).
  e2e   : same metric through the public operator (ADCensus.enqueue/wait over the
          C-ABI) with HOST buffers: pinned staging, H2D, kernels, D2H inside the timed region.
  roofline     : dominant kernel (one aggregation walk pass, 8 B/cell) vs measured HBM peak.
  cpu_baseline : the reference's own CPU implementation on this box's host cores
                 (oracle/_ref when present, else the oracle port), bounded sample.

--impl reference times ONLY the CPU reference arm (rank 0; other ranks exit 0).
Only this file's cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
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

H, W, MAXD = 1080, 1920, 192
DN = MAXD + 1
BATCH = 64
IN_FLIGHT = 3
DISTINCT = 8  # distinct synthetic frames generated per rank; the batch cycles through them
METRIC = "Mpix*disp/s ADCensus 1920x1080 D=192"
UNIT = "Mpix*disp/s"


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
def load_cpu_reference():
    import oracle

    try:
        if oracle.REF_SO.exists():
            return oracle.Ref(), "reference"
    except Exception as e:  # pragma: no cover
        log("reference .so unusable:", e)
    return oracle.Port(), "port"


def cpu_run(impl, kind, left, right):
    """One ADCensus::compute of the CPU implementation; returns seconds."""
    if kind == "reference":
        _, sec = impl.compute(left, right, MAXD)  # the reference's public entry, as shipped (all OpenMP threads)
        return sec
    t0 = time.perf_counter()
    impl.compute(left, right, MAXD)
    return time.perf_counter() - t0


def cpu_sample_rows(impl, kind, frame, budget_s: float, runs: int) -> int:
    """Rows of the full-width stripe so that `runs` runs fit in about budget_s."""
    left, right = frame
    probe = 24
    t = cpu_run(impl, kind, np.ascontiguousarray(left[:probe]), np.ascontiguousarray(right[:probe]))
    per_row = t / probe
    rows = int(budget_s / max(runs, 1) / max(per_row, 1e-9))
    return int(min(H, max(32, rows)))


def reference_arm(args, frame):
    impl, kind = load_cpu_reference()
    runs = args.steps + args.warmup
    rows = cpu_sample_rows(impl, kind, frame, budget_s=150.0, runs=runs)
    left, right = np.ascontiguousarray(frame[0][:rows]), np.ascontiguousarray(frame[1][:rows])
    for _ in range(args.warmup):
        cpu_run(impl, kind, left, right)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_run(impl, kind, left, right)
    cells = rows * W * DN * args.steps
    value = cells / t / 1e6
    sample = f"rows 0-{rows - 1} of synthetic frame 0 (full width {W}x{rows}, D=0..{MAXD}); {args.steps} timed runs of ADCensus::compute"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3: synthetic 1920x1080 RGB stereo pairs, D=0..192 (Dn=193)", "sample": sample,
                   "cpu_threads": impl.threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": impl.threads, "kind": kind, "sample": sample},
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
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per step over all ranks (default 64 = config C3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=IN_FLIGHT, help="stereo pairs in flight per GPU (contexts / streams)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from tea_stereo_matching_b200.synth import synth_v1

    if args.impl == "reference":
        if rank != 0:
            return 0
        reference_arm(args, synth_v1(H, W, MAXD, seed=1000))
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

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # frames of this rank: global frame i -> rank i % N ; seeds 1000+i
    from tea_stereo_matching_b200.sharding import frames_for_rank

    my_frames = frames_for_rank(args.batch, world, rank)
    n_distinct = max(1, min(DISTINCT, len(my_frames)))
    t0 = time.time()
    frames = [synth_v1(H, W, MAXD, seed=1000 + my_frames[j]) for j in range(n_distinct)]
    log(f"[rank {rank}] generated {n_distinct} distinct synthetic frames in {time.time() - t0:.1f}s; "
        f"{len(my_frames)} frames per step on this rank")

    # IN_FLIGHT contexts (= streams, arenas) per GPU: frames alternate between them so the small
    # serial refinement kernels of one pair overlap with the bandwidth kernels of the other.
    NCTX = max(1, args.in_flight)
    streams = [torch.cuda.Stream() for _ in range(NCTX)]
    matchers = []
    for st in streams:
        m = tsm.ADCensus(device=local_rank, stream=st.cuda_stream)
        m.setMatchingStrategy(tsm.ColorModel.RGB, False, False)
        m.setMinMaxDisparity(0, MAXD)
        matchers.append(m)
    matcher, stream = matchers[0], streams[0]
    ctx = matcher.context

    d_frames = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in frames]
    d_out = [torch.empty((H, W), dtype=torch.float32, device="cuda") for _ in range(max(n_distinct, NCTX))]
    torch.cuda.synchronize()

    def step_device():
        for j in range(len(my_frames)):
            l, r = d_frames[j % n_distinct]
            matchers[j % NCTX].compute_device(l.data_ptr(), r.data_ptr(), H, W, d_out[j % len(d_out)].data_ptr())

    def launches_total():
        return sum(m.context.launch_count for m in matchers)

    # ---- value: HBM-resident, device timed ----
    # events on a timing stream that is ordered after / before both work streams
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

    total_cells = float(args.batch) * H * W * DN
    value = total_cells * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- per-stage times + roofline of the dominant kernel (one profiled frame, same stream) ----
    ctx.set_profiling(True)
    with torch.cuda.stream(stream):
        stage_acc: dict = {}
        nprof = 3
        for j in range(nprof):
            l, r = d_frames[j % n_distinct]
            matcher.compute_device(l.data_ptr(), r.data_ptr(), H, W, d_out[0].data_ptr())
            for k, v in ctx.stage_times().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v / nprof
    ctx.set_profiling(False)
    peak, peak_src = measured_peak_gbs()
    cells_pair = 2.0 * H * W * DN  # both views
    agg_launches = 8
    agg_ms = stage_acc.get("aggregate", float("nan")) / agg_launches
    agg_bytes = 8.0 * cells_pair  # one pass: read + write every cell of both views once
    achieved = agg_bytes / (agg_ms * 1e-3) / 1e9
    scan_ms = stage_acc.get("scanline", float("nan")) / 2
    scan_bytes = 16.0 * cells_pair  # one launch = forward + backward pass: 2 x (read + write)
    # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (profiles/), if any
    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    if tfile.exists():
        tj = json.loads(tfile.read_text())
        traffic, traffic_src = tj.get("k_agg_persist_bytes_per_launch"), tj.get("source")
    roofline = {
        "bound": "hbm", "kernel": "k_agg_persist (one aggregation1D pass over both views; stage time / 8 passes)",
        "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": agg_bytes, "avg_launch_ms": agg_ms,
        "other_kernels": {
            "k_scanline (fwd+bwd pass pair, both views)": {"achieved": scan_bytes / (scan_ms * 1e-3) / 1e9,
                                                           "frac": scan_bytes / (scan_ms * 1e-3) / 1e9 / peak,
                                                           "avg_launch_ms": scan_ms},
            "pipeline (120 B/cell end to end)": {"achieved": 60.0 * cells_pair / (sum(stage_acc.values()) * 1e-3) / 1e9,
                                                 "frac": 60.0 * cells_pair / (sum(stage_acc.values()) * 1e-3) / 1e9 / peak},
        },
    }

    # ---- e2e: public operator, host buffers, 2 pairs in flight per GPU ----
    m2 = [tsm.ADCensus(device=local_rank) for _ in range(NCTX)]
    for m in m2:
        m.setMatchingStrategy(tsm.ColorModel.RGB, False, False)
        m.setMinMaxDisparity(0, MAXD)
    h_out = [np.empty((H, W), np.float32) for _ in range(NCTX)]

    def run_e2e(nsteps):
        """nsteps steps as ONE stream of pairs: every pair is copied in from host memory and its map copied back;
        the pipeline (NCTX pairs in flight) is not drained between steps, only at the end of the timed region."""
        n = len(my_frames) * nsteps
        for j in range(n + NCTX):
            if j >= NCTX:  # pair j - NCTX ran on the context pair j is about to use: NCTX pairs stay in flight
                m2[j % NCTX].wait(h_out[j % NCTX])
            if j < n:
                l, r = frames[j % n_distinct]
                m2[j % NCTX].enqueue(l, r)

    e2e_steps = max(1, min(args.steps, 2))
    run_e2e(1)  # warm (arena + pinned allocations)
    barrier()
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_cells * e2e_steps / e2e_s / 1e6
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(args.batch * 2 * H * W * 3),
           "d2h_bytes_per_step": int(args.batch * H * W * 4), "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
           "in_flight_per_gpu": NCTX}

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            impl, kind = load_cpu_reference()
            rows = cpu_sample_rows(impl, kind, frames[0], budget_s=20.0, runs=1)
            sec = cpu_run(impl, kind, np.ascontiguousarray(frames[0][0][:rows]), np.ascontiguousarray(frames[0][1][:rows]))
            cpu_baseline = {"value": rows * W * DN / sec / 1e6, "unit": UNIT, "cores": impl.threads, "kind": kind,
                            "sample": f"rows 0-{rows - 1} of synthetic frame 0 (full width {W}x{rows}, D=0..{MAXD}), 1 run, {sec:.1f}s"}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "ms_per_frame": dev_ms / args.steps / len(my_frames),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: synthetic 1920x1080 RGB stereo pairs (synth_v1), D=0..192 (Dn=193)",
                       "frames_per_step": args.batch, "frames_per_gpu": len(my_frames), "distinct_frames_per_gpu": n_distinct,
                       "parallelism": f"frame-sharded x{n_gpus}, no data-path collective; {NCTX} pairs in flight per GPU ({NCTX} streams)",
                       "l2": "working set 3.2 GB per frame >> 126 MB L2 (no flush needed)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "stages_ms": stage_acc,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
