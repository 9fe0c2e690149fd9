#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the OpenPose target-and-decode hot path.

Headline workload (BASELINE.json configs[1]): training-target generation, batch 256 per GPU
(augment warp + mask + 18 Gaussian maps + background + 38 PAF planes), 3 persons per sample,
synthetic COCO-shaped inputs.  One "step" = one pass of the path over one batch.

  python bench.py --gpus 1 --steps K --warmup W            -> one JSON line (this framework)
  python bench.py --impl reference --gpus 1 --steps K ...  -> one JSON line (reference CPU path)
  torchrun ... bench.py --gpus N ...                       -> weak scaling, samples sharded by rank

The other BASELINE.json configs ride on the same line as secondary legs under "configs" (each with its own
device-resident value, host-buffer e2e, per-kernel times and -- at N=1 -- cpu_baseline):
  configs2_single_scale_ski   single-scale decode, 674x712 frames, 8 per GPU (64 over 8 GPUs); also under "decode"
  configs3_multi_scale_1k     4-scale decode over the 1000 COCO2014-Val shapes, frames sharded i % N (strong scaling)
  configs4_crowded_gt         GT batch 64 per GPU (512 over 8), 20 persons per sample
  configs4_crowded_decode     single-scale decode of 20-person 674x712 frames, 8 per GPU

JSON keys beyond the base contract: roofline (dominant kernel vs measured HBM peak), cpu_baseline
(the reference's CPU arithmetic timed on this host), e2e (host buffers in/out through the C ABI),
kernels (device ms per launch of every kernel in the step).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 256
PERSONS = 3
SRC_HW = (368, 368)
DEC_FRAMES = 8          # per GPU: configs[2] is batch 64 over 8 GPUs
DEC_HW = (674, 712)     # sample_images/ski.jpg
CROWD_BATCH = 64        # per GPU: configs[4] is batch 512 over 8 GPUs
CROWD_PERSONS = 20
MS_TOTAL = 1000         # configs[3]: 1k COCO2014-Val-shaped images
WORKLOAD = "gt_batch256_3persons_368x368 (BASELINE.json configs[1])"


# --------------------------------------------------------------------------------------------
# algorithmic bytes (DESIGN.md "Roofline"): what one launch must move at minimum
# --------------------------------------------------------------------------------------------
def gt_bytes_per_sample(P, src_hw=SRC_HW):
    H, W = src_hw
    img_in, mask_in = H * W * 3, H * W
    img_out, mask_out, labels = 368 * 368 * 3, 46 * 46 * 4, 57 * 46 * 46 * 4
    joints = 432 * P
    return {
        "k_warp_fused": img_in + mask_in + img_out + mask_out + 48,
        "k_warp_simple": img_in + img_out + 48,
        "k_mask46": mask_in + mask_out + 48,
        "k_raster": mask_out + labels + 2 * joints + 48 + 1 + 4,
        "step": img_in + mask_in + joints + 48 + img_out + mask_out + labels + joints,
    }


def decode_bytes_per_frame(H, W, grids):
    """Staged dataflow of the reference: single scale 4*(57hw + 74HW); multi scale (f64 average map)
    4*57*sum(h_s w_s) + 8*74*HW."""
    cells = sum(h * w for h, w in grids)
    return 4 * 57 * cells + (4 if len(grids) == 1 else 8) * 74 * H * W


def blob_bytes_per_frame(grids):
    return 4 * 57 * sum(h * w for h, w in grids)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
_REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}


class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], 0
        self.sm_max = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001  (no NVML: report that, do not fail the bench)
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(v for k, v in _REASONS.items() if self.reasons & k),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Bind this process to the CPUs NVML reports as local to GPU `index` (its NUMA node / PCIe root), so that the pinned
    host buffers allocated afterwards and the thread that drives the copies sit next to the GPU.  With several ranks
    per box this keeps every rank's host<->device traffic on its own socket.  Returns the CPU count bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            try:
                phys = int(vis.split(",")[index])
            except (ValueError, IndexError):
                phys = index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001  (no NVML, no permission, unknown topology: run unbound)
        return None


# --------------------------------------------------------------------------------------------
# synthetic workloads.  Frames are made by a fork pool BEFORE CUDA is initialised.
# --------------------------------------------------------------------------------------------
def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()


def ms_shape_list():
    """The 1000 (H, W) of eval/val2014_1k.txt (tests/golden/val2014_1k_shapes.json), in a fixed shuffled order so that
    shards i % N are alike."""
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "val2014_1k_shapes.json")))["shapes"]
    pool = [(hh, ww) for hh, ww, c in shapes for _ in range(c)]
    order = np.random.RandomState(0).permutation(len(pool))
    return [pool[i] for i in order][:MS_TOTAL]


def _synth_frame(task):
    import rmpe_b200
    S = rmpe_b200.synth
    kind, seed, H, W, persons = task
    if kind == "ms":
        return S.multi_scale_frame(seed, H, W, persons)
    if kind == "dense":
        return S.dense_frame(seed, H, W)
    h, w = S.single_scale_grid(H, W)
    paf, heat, _ = S.decode_blobs(seed, (H, W), (h, w), persons)
    return dict(H=H, W=W, scales=[(paf, heat, 0, 0)])


def synth_frames(tasks, workers):
    if not tasks:
        return []
    if workers <= 1:
        return [_synth_frame(t) for t in tasks]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        return pool.map(_synth_frame, tasks, chunksize=max(1, len(tasks) // (workers * 4)))


def frame_grids(f):
    return [tuple(s[1].shape[:2]) for s in f["scales"]]


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own arithmetic (cv2.warpAffine / cv2.resize / numpy rasteriser, cv2.resize +
# scipy gaussian_filter + the Python limb / assembly loops) through oracle/cpu_port.py, one process per core
# like py_rmpe_server/rmpe_server.py:26 scales.  Decode frames are inherited from the parent through fork.
# --------------------------------------------------------------------------------------------
_W = {}
_FRAMES = {}      # kind -> list of frames, filled by the parent before the pool forks


def _cpu_init():
    import cv2
    cv2.setNumThreads(1)
    from oracle import cpu_port, gt_oracle as go
    _W["port"], _W["go"] = cpu_port, go


def _cpu_gt(task):
    persons, n = task
    port, go = _W["port"], _W["go"]
    key = ("gt", persons)
    if key not in _W:
        import rmpe_b200
        _W[key] = [rmpe_b200.synth.gt_sample(i, persons, SRC_HW, True) for i in range(16)]
    samples = _W[key]
    for i in range(n):
        s = samples[i % len(samples)]
        flip, deg, crop, scale = s["aug"]
        # T1 (AugmentSelection.affine) is part of the per-sample work, as in RawDataIterator.transform_data
        M = go.affine_closed_form(flip, deg, crop, scale, s["objpos"][0], s["scale_provided"][0])
        port.gt_sample(s["img"], s["mask"], s["joints"], M, flip)
    return n


def _cpu_decode(task):
    kind, idx = task
    f = _FRAMES[kind][idx]
    _W["port"].decode_frame(f["scales"], f["H"], f["W"])
    return 1


class CpuArm:
    def __init__(self, cores=None):
        self.cores = cores or host_cores()
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init)

    def gt(self, n_samples, persons):
        """Process n_samples spread over the pool; returns wall seconds."""
        per = max(1, n_samples // (self.cores * 4))
        chunks = [per] * (n_samples // per)
        if n_samples - per * len(chunks) > 0:
            chunks.append(n_samples - per * len(chunks))
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_gt, [(persons, c) for c in chunks], chunksize=1))
        dt = time.perf_counter() - t0
        assert done == n_samples
        return dt

    def gt_rate(self, batch, persons, target_seconds):
        self.pool.map(_cpu_gt, [(persons, 1)] * self.cores)      # page everything in, build the samples
        dt = self.gt(batch, persons)
        reps = int(max(1, min(40, round(target_seconds * (batch / dt) / batch))))
        dt = self.gt(batch * reps, persons)
        return batch * reps / dt, batch * reps

    def decode_rate(self, kind, per_core):
        n = min(len(_FRAMES[kind]), self.cores * per_core)
        tasks = [(kind, i) for i in range(n)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_decode, tasks, chunksize=1)
        return n / (time.perf_counter() - t0), n

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baselines(arm, which):
    c = arm.cores
    out = {}
    if "gt" in which:
        v, n = arm.gt_rate(BATCH, PERSONS, 8.0)
        out["gt"] = {"value": v, "unit": "samples/s", "cores": c, "kind": "port",
                     "sample": "%d samples (3 persons, random aug) through AugmentSelection.affine + cv2.warpAffine + cv2.resize "
                               "+ the NumPy rasteriser, one process per core" % n}
    if "crowd_gt" in which:
        v, n = arm.gt_rate(CROWD_BATCH, CROWD_PERSONS, 4.0)
        out["crowd_gt"] = {"value": v, "unit": "samples/s", "cores": c, "kind": "port",
                           "sample": "%d samples (20 persons) through the same path" % n}
    for kind, per_core, what in (("ss3", 3, "ski-shaped single-scale frames, 3 persons"),
                                 ("ss20", 2, "ski-shaped single-scale frames, 20 persons"),
                                 ("ms", 2, "COCO-val-shaped 4-scale frames (first of the shuffled 1k list), 3 persons")):
        if kind in which and _FRAMES.get(kind):
            v, n = arm.decode_rate(kind, per_core)
            out[kind] = {"value": v, "unit": "frames/s", "cores": c, "kind": "port",
                         "sample": "%d %s through cv2.resize + scipy gaussian_filter + the reference's Python limb and "
                                   "assembly loops (oracle/cpu_port.decode_frame; no drawing, no model.predict), one process "
                                   "per core" % (n, what)}
    return out


def prepare_frames(rank, world, want_secondary, workers):
    """Synthetic decode frames of this rank: configs[2] / [4] ski frames and this rank's shard of the configs[3] list."""
    H, W = DEC_HW
    tasks = [("ss", 9000 + 100 * rank + i, H, W, PERSONS) for i in range(64 if want_secondary else DEC_FRAMES)]
    n_ss3 = len(tasks)
    n_ss20 = n_ms = n_dense = 0
    ms_idx = []
    if want_secondary:
        tasks += [("ss", 19000 + 100 * rank + i, H, W, CROWD_PERSONS) for i in range(32)]
        n_ss20 = 32
        tasks += [("dense", 29000 + 100 * rank + i, H, W, 0) for i in range(DEC_FRAMES)]
        n_dense = DEC_FRAMES
        shapes = ms_shape_list()
        ms_idx = [i for i in range(len(shapes)) if i % world == rank]
        tasks += [("ms", 700 + i, shapes[i][0], shapes[i][1], PERSONS) for i in ms_idx]
        n_ms = len(ms_idx)
    frames = synth_frames(tasks, workers)
    _FRAMES["ss3"] = frames[:n_ss3]
    _FRAMES["ss20"] = frames[n_ss3:n_ss3 + n_ss20]
    _FRAMES["dense"] = frames[n_ss3 + n_ss20:n_ss3 + n_ss20 + n_dense]
    _FRAMES["ms"] = frames[n_ss3 + n_ss20 + n_dense:n_ss3 + n_ss20 + n_dense + n_ms]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    want_secondary = not args.no_configs
    prepare_frames(0, 1, want_secondary, host_cores())
    # bounded: the decode baselines see the first frames of each list only
    arm = CpuArm()
    arm.pool.map(_cpu_gt, [(PERSONS, 1)] * arm.cores)
    for _ in range(args.warmup):
        arm.gt(BATCH, PERSONS)
    t = 0.0
    for _ in range(args.steps):
        t += arm.gt(BATCH, PERSONS)
    ms = t / args.steps * 1e3
    v = BATCH / (ms * 1e-3)
    secondary = cpu_baselines(arm, ("crowd_gt", "ss3", "ss20", "ms")) if want_secondary else cpu_baselines(arm, ("ss3",))
    arm.close()
    sample = "each step = the 256-sample GT batch over %d processes (cv2 + NumPy, oracle/cpu_port.py)" % arm.cores
    print(json.dumps({
        "impl": "reference", "metric": "gt_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "persons": PERSONS,
                   "labels": "f64 (57,46,46) (reference dtype)", "image": "u8 HWC 368x368x3"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "decode": secondary.get("ss3"),
        "configs": {"configs2_single_scale_ski": secondary.get("ss3"), "configs3_multi_scale_1k": secondary.get("ms"),
                    "configs4_crowded_gt": secondary.get("crowd_gt"), "configs4_crowded_decode": secondary.get("ss20")},
    }), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def make_gt_inputs(rmpe, seed0, batch=BATCH, persons=PERSONS):
    b = rmpe.synth.gt_batch(batch, n_persons=persons, seed0=seed0)
    b["flip"] = np.array([a[0] for a in b["augs"]], np.uint8)
    b["degree"] = np.array([a[1] for a in b["augs"]], np.float64)
    b["crop"] = np.array([a[2] for a in b["augs"]], np.int32)
    b["scale"] = np.array([a[3] for a in b["augs"]], np.float64)
    b["M"] = rmpe.batch.aug_affine(b["flip"], b["degree"], b["crop"], b["scale"], b["centers"], b["scale_self"])
    return b


def pinned(shape, dtype):
    import torch
    t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
    return t, t.numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-decode", action="store_true", help="skip every decode leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[3] / configs[4] legs and the sweeps")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    want_secondary = not args.no_configs and not args.no_decode

    # ---- host-side preparation and the CPU legs first: forked workers must not inherit a CUDA context ----
    under_profiler = "NV_NSIGHT_INJECTION_TRANSPORT_TYPE" in os.environ   # set by ncu for the processes it launches
    cores = host_cores()
    if not args.no_decode:
        prepare_frames(rank, world, want_secondary, 1 if under_profiler else max(1, cores // world))
    cpu = {}
    if under_profiler and not args.no_cpu:
        # ncu injects into every child process; the forked CPU workers crash it (SIGSEGV).  A number printed under a
        # profiler is never a bench value anyway.
        print("bench.py: running under Nsight Compute, skipping the cpu_baseline legs", file=sys.stderr)
    if rank == 0 and world == 1 and not args.no_cpu and not under_profiler:
        arm = CpuArm()
        which = ["gt"] + ([] if args.no_decode else ["ss3"]) + (["crowd_gt", "ss20", "ms"] if want_secondary else [])
        if args.no_configs and not args.no_decode:
            which = ["gt", "ss3"]
        cpu = cpu_baselines(arm, which)
        arm.close()

    # every rank next to its GPU before anything is pinned (the CPU legs above ran on all cores)
    numa_cpus = None if os.environ.get("RMPE_BENCH_NO_AFFINITY") else bind_to_gpu_numa_node(local)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import rmpe_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = rmpe_b200.lib
    L.ensure_init(local)
    lib = L.load()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def max_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"

    def load_json(name):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:  # noqa: BLE001
            return {}

    traffic_db, limiters = load_json("traffic.json"), load_json("limiters.json")

    sampler = ClockSampler(local)
    sampler.start()

    def timed(fn, steps, warmup):
        """K steps back to back, CUDA events on the launching (current torch) stream, max over ranks."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.rmpe_launch_count()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        return max_ranks(e0.elapsed_time(e1) / steps), lib.rmpe_launch_count() - n0

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_flushed(fn, steps, warmup):
        """Steps timed one by one with L2 flushed (256 MB write) before each: for steps whose inputs fit in L2."""
        for i in range(warmup):
            fn(i)
        barrier()
        n0 = lib.rmpe_launch_count()
        tot = 0.0
        for i in range(steps):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(warmup + i)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        barrier()
        return max_ranks(tot / steps), lib.rmpe_launch_count() - n0

    def host_timed(fn, steps, warmup=2):
        """Wall clock around synchronous host-buffer calls (each returns with its results on the host)."""
        for i in range(warmup):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        torch.cuda.synchronize(dev)
        return max_ranks((time.perf_counter() - t0) / steps)

    def kernel_profile(fn, steps):
        L.profile_enable(True)
        timed(fn, steps, 1)
        L.profile_enable(False, reset=False)
        prof = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in L.profile_read().items() if v[1]}
        L.profile_enable(False, reset=True)
        return prof

    sec_steps = max(3, min(args.steps, 10))

    # =====================================================================================================
    # GT legs: headline (configs[1]) and crowded (configs[4])
    # =====================================================================================================
    def gt_leg(batch, persons, seed_base, steps, warmup, headline):
        NPOOL = 3      # rotating buffer sets: in+out of one set is 369 MB (93 MB crowded), three of them > 126 MB L2
        hosts = [make_gt_inputs(rmpe_b200, seed_base + 100000 * rank + 1000 * k, batch, persons) for k in range(NPOOL)]
        plans = []
        for hb in hosts:
            p = rmpe_b200.batch.GtDevicePlan(batch, persons, SRC_HW)
            p.upload(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"])
            plans.append(p)
        torch.cuda.synchronize(dev)
        run = lambda i: plans[i % NPOOL].run()
        ms, launches = timed(run, steps, warmup)
        status_bad = int(sum(int((p.status != 0).sum().item()) for p in plans))
        prof = kernel_profile(run, steps)
        per = gt_bytes_per_sample(persons)
        kernels = {}
        for name, kv in prof.items():
            kb = per.get(name, 0) * batch
            kernels[name] = dict(kv, algorithmic_bytes=kb,
                                 achieved_gbs=kb / (kv["ms_per_launch"] * 1e-3) / 1e9 if kb else None)
        roofline = None
        if kernels:
            top = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches"])
            kt = kernels[top]
            tot = sum(k["ms_per_launch"] * k["launches"] for k in kernels.values())
            roofline = {"bound": "hbm", "kernel": top, "achieved": kt["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": (kt["achieved_gbs"] or 0.0) / peak, "traffic": traffic_db.get(top) if headline else None,
                        "peak_source": peak_src, "kernel_ms": kt["ms_per_launch"],
                        "kernel_share_of_step": kt["ms_per_launch"] * kt["launches"] / tot if tot else None,
                        "step_achieved": per["step"] * batch / (ms * 1e-3) / 1e9,
                        "step_frac": per["step"] * batch / (ms * 1e-3) / 1e9 / peak,
                        # what ncu says holds the kernel below the HBM roofline (from the committed capture, not live)
                        "limiter": limiters.get(top) if headline else None}
        # ---- e2e: host (pinned) buffers in, host buffers out, through rmpe_aug_affine + rmpe_gt_batch_host ----
        e2e = None
        if not args.no_e2e:
            hb = hosts[0]
            keep, pin_in = [], {}
            for k, dt in (("imgs", torch.uint8), ("masks", torch.uint8), ("joints", torch.float64)):
                t, a = pinned(hb[k].shape, dt)
                a[...] = hb[k]
                keep.append(t)
                pin_in[k] = a

            def outputs(ft):
                out = {}
                for k, shape, dt in (("img", (batch, 368, 368, 3), torch.uint8), ("mask", (batch, 46, 46), ft),
                                     ("labels", (batch, 57, 46, 46), ft), ("joints", (batch, persons, 18, 3), torch.float64)):
                    t, a = pinned(shape, dt)
                    keep.append(t)
                    out[k] = a
                return out

            h2d = pin_in["imgs"].nbytes + pin_in["masks"].nbytes + pin_in["joints"].nbytes + batch * (4 + 48 + 1 + 32)
            n_e2e = max(3, min(steps, 10))

            def e2e_variant(ft, f64):
                out = outputs(ft)

                def step(i):
                    # T1 inside the timed region: the affine matrices are made from the augmentation draws every step
                    M = rmpe_b200.batch.aug_affine(hb["flip"], hb["degree"], hb["crop"], hb["scale"], hb["centers"],
                                                   hb["scale_self"])
                    rmpe_b200.batch.gt_batch_host(pin_in["imgs"], pin_in["masks"], pin_in["joints"], hb["n_persons"],
                                                  M, hb["flip"], out=out, f64=f64)

                dt = host_timed(step, n_e2e)
                return dt, int(sum(a.nbytes for a in out.values()) + batch * 4)

            dt, d2h = e2e_variant(torch.float32, False)
            e2e = {"value": world * batch / dt, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": n_e2e,
                   "api": "rmpe_aug_affine + rmpe_gt_batch_host (pinned host buffers; copies inside the call)"}
            if headline:
                # what a caller of the drop-in classes sees: labels and mask in the reference's dtype (f64)
                dt64, d2h64 = e2e_variant(torch.float64, True)
                e2e["f64_labels"] = {"value": world * batch / dt64, "unit": "samples/s", "ms_per_step": dt64 * 1e3,
                                     "d2h_bytes_per_step": d2h64}
                # the same f32 call with ordinary (pageable) numpy buffers, as a caller that does not pin anything has them:
                # the driver stages every copy through its own pinned buffers and the chunk pipeline serialises
                pg_out = {"img": np.empty((batch, 368, 368, 3), np.uint8), "mask": np.empty((batch, 46, 46), np.float32),
                          "labels": np.empty((batch, 57, 46, 46), np.float32), "joints": np.empty((batch, persons, 18, 3))}

                def pageable_step(i):
                    M = rmpe_b200.batch.aug_affine(hb["flip"], hb["degree"], hb["crop"], hb["scale"], hb["centers"],
                                                   hb["scale_self"])
                    rmpe_b200.batch.gt_batch_host(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], M, hb["flip"],
                                                  out=pg_out)

                dtp = host_timed(pageable_step, 3, 1)
                e2e["pageable_buffers"] = {"value": world * batch / dtp, "unit": "samples/s", "ms_per_step": dtp * 1e3}
                # the reference's own calling pattern, one sample at a time through the drop-in class
                # (RawDataIterator.transform_data = AugmentSelection.random + one fused C call; py_rmpe_data_iterator.py:68-74)
                import random as _random
                it = rmpe_b200.data_iterator.RawDataIterator(None, shuffle=False, augment=True)
                _random.seed(1234)

                def drop_in_step(i):
                    k = i % batch
                    meta = {"joints": hb["joints"][k].copy(), "objpos": [list(hb["centers"][k])],
                            "scale_provided": [float(hb["scale_self"][k])]}
                    it.transform_data(hb["imgs"][k], hb["masks"][k], meta)

                dts = host_timed(drop_in_step, 64, 4)
                e2e["drop_in_per_sample"] = {"value": world / dts, "unit": "samples/s", "ms_per_sample": dts * 1e3,
                                             "api": "RawDataIterator.transform_data (reference signature, f64 labels, one call per sample)"}
                # what the interconnect alone takes for the f32 step's bytes: the same pinned buffers copied in and out
                # on two streams at once, no kernels (explains e2e against the device-resident value)
                big_in = [keep[0], keep[1]]                       # imgs, masks
                big_out = [keep[3], keep[5]]                      # img, labels
                dev_in = [torch.empty_like(t, device=dev) for t in big_in]
                dev_out = [torch.empty_like(t, device=dev) for t in big_out]
                s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

                def copies(_i):
                    with torch.cuda.stream(s_in):
                        for d_, h_ in zip(dev_in, big_in):
                            d_.copy_(h_, non_blocking=True)
                    with torch.cuda.stream(s_out):
                        for d_, h_ in zip(dev_out, big_out):
                            h_.copy_(d_, non_blocking=True)

                floor = host_timed(copies, 5, 1)
                e2e["copy_floor_ms"] = floor * 1e3
                e2e["frac_of_copy_floor"] = floor / dt
        del plans
        return dict(ms=ms, launches=launches, status_bad=status_bad, kernels=kernels, roofline=roofline, e2e=e2e, per=per)

    gt = gt_leg(BATCH, PERSONS, 0, args.steps, args.warmup, True)
    value = world * BATCH / (gt["ms"] * 1e-3)

    # =====================================================================================================
    # decode legs
    # =====================================================================================================
    def decode_leg(frames, steps, *, flushed, strong_total=None, caps=None, with_e2e=True, with_kernels=True, what=""):
        caps = caps or {}
        n_local = len(frames)
        dplans = [rmpe_b200.batch.DecodeDevicePlan(frames, **caps) for _ in range(2 if flushed else 1)]
        run = lambda i: dplans[i % len(dplans)].run()
        dms, dl = (timed_flushed if flushed else timed)(run, steps, 3)
        res = dplans[0].results()
        n_total = sum_ranks(n_local)
        leg = {"metric": "decoded_frames_per_s", "value": n_total / (dms * 1e-3), "unit": "frames/s", "ms_per_step": dms,
               "steps": steps, "gpu_launches": int(dl), "frames_per_step_all_ranks": int(n_total),
               "us_per_frame": dms * 1e3 / max(n_local, 1),
               "persons_found_mean": float(np.mean([len(r["subset"]) for r in res])) if res else 0.0,
               "status_nonzero": int(sum(1 for r in res if r["status"]))}
        grids = [frame_grids(f) for f in frames]
        staged = float(sum(decode_bytes_per_frame(f["H"], f["W"], gr) for f, gr in zip(frames, grids)))
        blob = float(sum(blob_bytes_per_frame(gr) for gr in grids))
        if with_kernels:
            dprof = kernel_profile(run, steps)
            leg["kernels"] = dprof
            dtop = max(dprof, key=lambda k: dprof[k]["ms_per_launch"] * dprof[k]["launches"]) if dprof else None
            # Decode is latency-bound: the figure to judge is us_per_frame and `frac` = the blob bytes
            # the kernels really have to read over the step time.  `nominal_staged_frac` divides the reference's STAGED
            # dataflow (up-sampled maps written and read back) by the same time; the kernels never
            # materialise those maps, so it is not a bandwidth and may exceed 1.
            ach = blob / (dms * 1e-3) / 1e9
            leg["roofline"] = {"bound": "hbm", "kernel": "decode step (all kernels; dominant: %s)" % dtop,
                               "kernel_ms": dprof[dtop]["ms_per_launch"] if dtop else None,
                               "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                               "traffic": traffic_db.get(dtop) if dtop else None,
                               "algorithmic": "blobs in (4*57*sum h_s*w_s B per frame): what the fused kernels must read",
                               "nominal_staged_bytes_per_step": staged,
                               "nominal_staged_frac": staged / (dms * 1e-3) / 1e9 / peak}
        del dplans
        if with_e2e and not args.no_e2e:
            hp = rmpe_b200.batch.DecodeHostPlan(frames, **caps)
            dt = host_timed(lambda i: hp.run(), max(3, min(steps, 5)))
            leg["e2e"] = {"value": n_total / dt, "unit": "frames/s", "h2d_bytes_per_step": hp.h2d_bytes,
                          "d2h_bytes_per_step": hp.d2h_bytes(), "ms_per_step": dt * 1e3,
                          "api": "rmpe_decode_batch_host (pinned blobs in, filled prefix of every result table out)"}
            r2 = hp.results()
            leg["e2e"]["matches_device_resident_run"] = bool(all(
                np.array_equal(a["candidate"], b["candidate"]) and np.array_equal(a["subset"], b["subset"])
                for a, b in zip(res, r2)))
            del hp
        return leg

    decode = None
    configs = {}
    if not args.no_decode:
        ss3, ss20, msf = _FRAMES["ss3"], _FRAMES.get("ss20", []), _FRAMES.get("ms", [])
        decode = decode_leg(ss3[:DEC_FRAMES], sec_steps, flushed=True)
        decode["config"] = {"workload": "single_scale_decode_674x712_84x89_blobs_3persons (configs[2] per-GPU share)",
                            "frames_per_gpu": DEC_FRAMES, "l2": "flushed (256 MB write) before every timed step",
                            "scaling": "weak"}
        decode["cpu_baseline"] = cpu.get("ss3")
        configs["configs2_single_scale_ski"] = decode
        if want_secondary:
            # frames-per-call sweep (device-resident): where the per-GPU plateau is
            sweep = {str(DEC_FRAMES): decode["value"] / world}
            for n in (64, 512):
                fr = [ss3[i % len(ss3)] for i in range(n)]
                lg = decode_leg(fr, max(3, sec_steps // 2), flushed=False, with_e2e=(n == 64), with_kernels=(n == 64))
                sweep[str(n)] = lg["value"] / world
                if n == 64:
                    lg["config"] = {"workload": "single-scale decode, 64 ski-shaped frames per call per GPU",
                                    "l2": "109 MB of blobs per step, not flushed"}
                    configs["configs2_single_scale_ski_64_per_gpu"] = lg
            decode["frames_per_call_sweep_frames_per_s_per_gpu"] = sweep

            # configs[3]: the whole 1000-image list, sharded i % N -> strong scaling
            lg = decode_leg(msf, max(3, sec_steps // 2), flushed=False)
            lg["config"] = {"workload": "multi-scale (4 scales) decode over the 1000 COCO2014-Val shapes (configs[3])",
                            "frames_total": MS_TOTAL, "frames_this_rank": len(msf), "shard": "i % n_gpus",
                            "scaling": "strong", "l2": "4.8 GB of blobs per pass over the list, no flush needed"}
            lg["cpu_baseline"] = cpu.get("ms")
            if world == 1:
                sw = {str(len(msf)): lg["value"]}
                for n in (32, 256):
                    sw[str(n)] = decode_leg(msf[:n], 3, flushed=False, with_e2e=False, with_kernels=False)["value"]
                lg["frames_per_call_sweep_frames_per_s_per_gpu"] = sw
            configs["configs3_multi_scale_1k"] = lg

            # configs[4]: crowded scenes
            cg = gt_leg(CROWD_BATCH, CROWD_PERSONS, 5000, sec_steps * 2, 3, False)
            configs["configs4_crowded_gt"] = {
                "metric": "gt_samples_per_s", "value": world * CROWD_BATCH / (cg["ms"] * 1e-3), "unit": "samples/s",
                "ms_per_step": cg["ms"], "steps": sec_steps * 2, "gpu_launches": int(cg["launches"]),
                "config": {"workload": "GT batch 64 per GPU, 20 persons per sample (configs[4], 512 over 8 GPUs)",
                           "l2": "rotating 3 buffer sets of %.0f MB" % (cg["per"]["step"] * CROWD_BATCH / 1e6),
                           "scaling": "weak"},
                "kernels": cg["kernels"], "roofline": cg["roofline"], "e2e": cg["e2e"], "status_nonzero": cg["status_bad"],
                "cpu_baseline": cpu.get("crowd_gt")}
            lg = decode_leg(ss20[:DEC_FRAMES], sec_steps, flushed=True, caps=dict(max_cand=2048))
            lg["config"] = {"workload": "single-scale decode of 20-person ski-shaped frames (configs[4])",
                            "frames_per_gpu": DEC_FRAMES, "l2": "flushed before every timed step", "scaling": "weak"}
            lg["cpu_baseline"] = cpu.get("ss20")
            configs["configs4_crowded_decode"] = lg
            # not a BASELINE config: the screening's worst case (nothing culled), to bracket real network output
            lg = decode_leg(_FRAMES["dense"], sec_steps, flushed=True, caps=dict(max_persons=128), with_e2e=False)
            lg["config"] = {"workload": "single-scale decode of DENSE ski-shaped blobs: smooth random fields, every (tile, part) "
                                        "pair active, ~25 peaks per part (worst case of the screening; not a BASELINE config)",
                            "frames_per_gpu": DEC_FRAMES, "l2": "flushed before every timed step", "scaling": "weak"}
            configs["extra_dense_blobs_worst_case"] = lg

    clocks = sampler.finish()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    per = gt["per"]
    line = {
        "metric": "gt_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": gt["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH,
                   "persons": PERSONS, "labels": "f32 (57,46,46)", "image": "u8 HWC 368x368x3",
                   "l2": "rotating 3 buffer sets of %.0f MB (in+out) > 126 MB L2" % (per["step"] * BATCH / 1e6),
                   "parallelism": "samples sharded by rank, no collective"},
        "clocks": clocks, "e2e": gt["e2e"], "gpu_launches": int(gt["launches"]), "roofline": gt["roofline"],
        "cpu_baseline": cpu.get("gt"), "kernels": gt["kernels"], "decode": decode, "configs": configs,
        "status_nonzero": gt["status_bad"], "host_cores": cores, "numa_bound_cpus": numa_cpus,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
