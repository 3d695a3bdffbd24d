#!/usr/bin/env python
"""bench.py — scans/sec & ms/IESKF-iter of the MA-LIO measurement hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N ...            # the reference's own CPU path on the host cores
  torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU, NCCL

A *step* is one scan: the full iterated update (esekfom.hpp:495-721; max_iteration 3 => up to 4 measurement
passes, k-NN on the first pass and after a converged step) of the 3-LiDAR 100k-point merged scan against the
1M-point map snapshot (BASELINE configs[1], "C2"; with --gpus N > 1 the same scan point-sharded over N ranks:
configs[2]).  Synthetic, seeded inputs (malio_b200/synth.py).

`value`   scans/sec with every input already resident in HBM when the timed region starts (the per-scan state is
          re-armed on the device; sort, k-NN, plane fit, gate, reduction, D2H of the 5 KB system and the host
          35x35 algebra are all inside).
`e2e`     the same metric through the public C-ABI calls with HOST (pinned) buffers: every step uploads the
          flattened map snapshot (compact form: 20 B per node, boxes rebuilt on the device) and the scan, runs the
          update, and reads back the per-point side outputs.
L2 is flushed (256 MiB write) between timed steps; each step is bracketed by CUDA events and the per-step
times are summed (max over ranks).  Only the cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))

METRIC = "scans/sec & ms/IESKF-iter, 100k-pt scan vs 1M-pt map, 1/2/4/8 GPU"
UNIT = "scans/s"
WORKLOAD = "C2: 3-LiDAR (Ouster+2xLivox) 100k-pt merged scan vs 1M-pt map snapshot, max_iteration=3"
# dram__bytes_read.sum + dram__bytes_write.sum of one knn_grid_kernel launch on C2
# (ncu --set full, profiles/r01_knn_grid_kernel_raw.csv / r01_knn_grid_kernel.md)
NCU_DRAM_BYTES_PER_KNN_LAUNCH = 8.23e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sort", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks sampler (recipe's nvidia-smi line)
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0 - 0.1) and (t1 is None or t <= t1 + 0.1)] or \
               [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ reference / CPU arm
class c_stdout_to_stderr:
    """The reference ikd-Tree printf()s thread start/stop notices on stdout; keep fd 1 clean for the JSON line."""

    def __enter__(self):
        import ctypes
        self.libc = ctypes.CDLL(None)
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        self.libc.fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(case, steps: int, warmup: int, threads: int):
    """The reference's CPU path: real ikd_Tree.cpp (oracle/_ref) for the k-NN when it was compiled here, else the
    restated search; restated h_share_model + update_iterated_dyn_share_modified (oracle/).  Returns
    (scans_per_s, ms_per_pass, kind, passes_per_scan)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    orc = po.Oracle(case.params)
    if po.ref_available():
        tree = po.RefTree(box_length=0.5)
        tree.build(case.map_xyz, case.map_normal_y)
        orc.set_knn_ref(tree)
        kind = "reference"
    else:
        from malio_b200 import plugin
        snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
        orc.set_map_snapshot(snap.nodes, snap.node_cov)
        kind = "port"
    times, passes = [], 0
    for i in range(warmup + steps):
        orc.set_scan(case.pts, case.table, case.table_off, case.temporal_comp)
        x, P = case.x_prop.copy(), case.P_prop.copy()
        t0 = time.perf_counter()
        rc, _, _, rep = orc.update_iterated(x, P, case.max_iter, nthreads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            passes += rep.passes
    tot = float(np.sum(times))
    orc.close()
    if kind == "reference":
        tree.close()
    return steps / tot, 1e3 * tot / max(passes, 1), kind, passes / max(steps, 1)


def run_reference(args, rank):
    if rank != 0:
        return
    from malio_b200 import synth
    case = synth.case_C2()
    threads = host_threads()
    steps = max(1, min(args.steps, 5))    # each step is one full scan on the CPU (~0.3-1 s with all cores): bounded
    warm = min(args.warmup, 1)
    with c_stdout_to_stderr():
        sps, ms_pass, kind, ppscan = cpu_reference_run(case, steps, warm, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 / sps, "ms_per_ieskf_iter": ms_pass, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic (seeded, malio_b200/synth.py)",
        "config": {"workload": WORKLOAD, "passes_per_scan": ppscan, "host_threads": threads},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{steps} full C2 scans (100k pts vs 1M-pt ikd-Tree), all host threads in the point loop"},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ CUDA arm
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from malio_b200 import capi, plugin, synth

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    case = synth.case_C2()
    N_total = case.pts.shape[0]
    # contiguous block of the merged scan per rank; map, tables and state are replicated
    from malio_b200 import dist as mdist
    lo, hi = mdist.shard_bounds(N_total, rank, world)
    snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)

    def pinned(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        return t, v

    keep = []
    t_nodes, h_nodes = pinned(snap.nodes); keep.append(t_nodes)
    t_cov, h_cov = pinned(snap.node_cov); keep.append(t_cov)
    t_pts, h_pts = pinned(np.ascontiguousarray(case.pts[lo:hi])); keep.append(t_pts)
    t_mp, h_mpts = pinned(plugin.compact_points(snap.nodes)); keep.append(t_mp)
    snap_p = plugin.MapSnapshot(h_nodes, h_cov, snap.node_ids, snap.max_depth)

    model = plugin.MeasurementModel(case.n_lidar, device=local_rank, sort_queries=not args.no_sort, params=case.params)
    if world > 1:
        mdist.init_comm(model, rank, world, device=torch.device("cuda", local_rank))
    model.upload_map(snap_p)
    model.upload_scan(h_pts, case.table, case.table_off, case.temporal_comp)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n_dof = case.n_dof

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        model.rearm_scan()
        x, P = case.x_prop.copy(), case.P_prop.copy()
        return model.update_iterated_dyn_share_modified(x, P, case.max_iter), x

    # the host side of a real integration knows the map's bounding box (ikd-Tree root node_range_*): node 0's point and
    # its two children's boxes
    n0 = snap.nodes[0]
    boxes = [n0["lbox"] if (n0["link"] & capi.LINK_HAS_LEFT) else None, n0["rbox"] if (n0["link"] & capi.LINK_HAS_RIGHT) else None]
    blo = np.array(n0["xyz"], np.float32); bhi = blo.copy()
    for b in boxes:
        if b is not None:
            blo = np.minimum(blo, b[0::2]); bhi = np.maximum(bhi, b[1::2])
    root_box = np.stack([blo, bhi], axis=1).reshape(6).astype(np.float32)
    t_ny, h_ny = pinned(np.zeros(h_pts.shape[0], np.float32)); keep.append(t_ny)
    t_sel, h_sel = pinned(np.zeros(h_pts.shape[0], np.uint8)); keep.append(t_sel)
    aux_out = {"normal_y": h_ny, "selected": h_sel}

    def step_e2e():
        model.upload_map_compact(snap_p, points=h_mpts, root_box=root_box)
        model.upload_scan(h_pts, case.table, case.table_off, case.temporal_comp)
        x, P = case.x_prop.copy(), case.P_prop.copy()
        rep = model.update_iterated_dyn_share_modified(x, P, case.max_iter)
        aux = model.aux(out=aux_out)
        return rep, x, aux

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
            flush.fill_(1)
        barrier()
        ms, reps = [], []
        c0 = model.counters()
        t_wall0 = time.time()
        for _ in range(steps):
            flush.fill_(1)          # L2 flush, outside the timed bracket
            torch.cuda.synchronize()
            if world > 1:           # start the ranks together: inside a pass they wait for each other (exchange), so any
                dist.barrier()      # start skew would be billed to the earlier rank's timed bracket
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            reps.append(out[0])
        t_wall1 = time.time()
        barrier()
        c1 = model.counters()
        total = torch.tensor([float(np.sum(ms))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), reps, c0, c1, out, (t_wall0, t_wall1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    model.set_timing(False)       # no per-pass event records inside the measured loops
    tot_ms, reps, c0, c1, out, (tw0, tw1) = timed(step_resident, args.steps, args.warmup)
    passes = sum(r.passes for r in reps)
    searches = sum(r.searches for r in reps)
    launches = int(c1.kernel_launches - c0.kernel_launches)
    host_ms = float(np.sum([r.ms_host_solve for r in reps]))
    value = args.steps / (tot_ms * 1e-3)

    e_steps = max(3, min(args.steps, 10))
    e_ms, e_reps, ec0, ec1, e_out, _ = timed(step_e2e, e_steps, min(args.warmup, 3))

    # separate short loop with per-kernel CUDA events on (same steps, L2 flushed): k-NN kernel time for the roofline
    model.set_timing(True)
    r_steps = max(3, min(args.steps, 10))
    _, r_reps, rc0, rc1, _, (_, tw_end) = timed(step_resident, r_steps, 1)
    # clocks / throttle reasons sampled (nvidia-smi, 20 ms) from the start of the resident loop to the end of the last
    # timed loop: all three loops keep the GPU under the same load
    clocks = sampler.stop(tw0, tw_end) if rank == 0 else None
    knn_launches = int(rc1.knn_launches - rc0.knn_launches)
    knn_ms = float(rc1.knn_ms - rc0.knn_ms)
    dev_ms = float(np.sum([r.ms_device_total for r in r_reps])) * args.steps / r_steps
    e_value = e_steps / (e_ms * 1e-3)
    h2d = int(ec1.h2d_bytes - ec0.h2d_bytes) // e_steps
    d2h = int(ec1.d2h_bytes - ec0.d2h_bytes) // e_steps

    if rank != 0:
        model.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant launch (the k-NN search: cell-list scan + the two list kernels behind it, timed together
    # with CUDA events on the library's stream): algorithmic bytes per search / event time.
    #   per query: 16 B scan point + 9 rows x 8 B cell ranges + C x 16 B candidates + 40 B neighbour list + 16 B world
    #   point + 1 B gate; C = candidates actually scanned (device counter, timing runs only)
    peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        peak = 6650.0
    q_per_launch = (hi - lo)
    cand = int(rc1.knn_candidates - rc0.knn_candidates)
    roof = None
    if knn_launches:
        cbar = cand / float(knn_launches * q_per_launch)
        bytes_per_launch = q_per_launch * (16 + 9 * 8 + cbar * 16 + 40 + 16 + 1)
        t_launch = knn_ms * 1e-3 / knn_launches
        achieved = bytes_per_launch / t_launch / 1e9
        roof = {"bound": "hbm", "kernel": "k-NN search: knn_grid_kernel (+ knn_list_kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_KNN_LAUNCH if world == 1 else None,
                "algorithmic_bytes_per_launch": bytes_per_launch, "us_per_launch": t_launch * 1e6,
                "launches_timed": knn_launches, "candidates_per_query": cbar,
                "fallback_queries_per_search": float(rc1.knn_fallback_queries - rc0.knn_fallback_queries) / knn_launches,
                "ring2_queries_per_search": float(rc1.knn_ring2_queries - rc0.knn_ring2_queries) / knn_launches,
                "peak_source": peak_src,
                "note": "bytes = Q*(16 + 72 + C*16 + 57), C = candidates scanned per query (device counter); the cell-sorted "
                        "point array (16 MB) and the scan are L2-resident, so DRAM traffic is far below the algorithmic "
                        "bytes: the kernel is bound by instruction issue of the top-6 insertion and by staging latency"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            th = host_threads()
            with c_stdout_to_stderr():
                sps, ms_pass, kind, _ = cpu_reference_run(case, 2, 1, th)
            cpu = {"value": sps, "unit": UNIT, "cores": th, "kind": kind, "ms_per_ieskf_iter": ms_pass,
                   "sample": "2 full C2 scans after 1 warm-up (100k pts vs 1M-pt ikd-Tree), all host threads"}
        except Exception as e:   # the checker is optional for the product line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "ms_per_ieskf_iter": tot_ms / max(passes, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (Jacobian, reduction, filter) / f32 (k-NN, plane fit)",
        "data": "synthetic (seeded, malio_b200/synth.py), random-free weights n/a",
        "config": {"workload": WORKLOAD, "parallelism": f"point-block x{world}", "passes_per_scan": passes / args.steps,
                   "knn_passes_per_scan": searches / args.steps, "l2": "flushed between timed steps (256 MiB write)",
                   "sort_queries": not args.no_sort, "points_per_rank": hi - lo, "map_nodes": snap.n_nodes},
        "device_ms_per_step": dev_ms / args.steps, "host_solve_ms_per_step": host_ms / args.steps,
        "clocks": clocks,
        "e2e": {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e_ms / e_steps, "steps": e_steps,
                "what": "upload_map_compact (20 B/node, boxes + cell index rebuilt on the device) + upload_scan (pinned host buffers) + IESKF update + download of normal_y/selected"},
        "gpu_launches": launches,
        "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    model.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}"}))
        sys.exit(2)
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
