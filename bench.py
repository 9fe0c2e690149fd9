#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the OpenPose target-and-decode hot path.

Headline workload (BASELINE.json configs[1]): training-target generation, batch 256 per GPU
(augment warp + mask + 18 Gaussian maps + background + 38 PAF planes), 3 persons per sample,
synthetic COCO-shaped inputs.  One "step" = one pass of the path over one batch.

  python bench.py --gpus 1 --steps K --warmup W            -> one JSON line (this framework)
  python bench.py --impl reference --gpus 1 --steps K ...  -> one JSON line (reference CPU path)
  torchrun ... bench.py --gpus N ...                       -> weak scaling, samples sharded by rank

JSON keys beyond the base contract: roofline (dominant kernel vs measured HBM peak), cpu_baseline
(the reference's CPU arithmetic timed on this host), e2e (host buffers in/out through the C ABI),
kernels (device ms per launch of every kernel in the step), decode (secondary metric: decoded
frames/s on ski.jpg-shaped frames, BASELINE.json configs[2] per-GPU share).
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
        "k_gt_fused": img_in + mask_in + joints + 48 + img_out + mask_out + labels + joints,
        "step": img_in + mask_in + joints + 48 + img_out + mask_out + labels + joints,
    }


def decode_bytes_per_frame(H, W, h, w):
    """Staged dataflow of the reference, single scale: 4*(57*h*w + 74*H*W)."""
    return 4 * (57 * h * w + 74 * H * W)


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


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own arithmetic (cv2.warpAffine / cv2.resize / numpy rasteriser) through
# oracle/cpu_port.py, one process per core like py_rmpe_server/rmpe_server.py:26 scales.
# --------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(seed0, n_local, persons):
    import cv2
    cv2.setNumThreads(1)
    from oracle import cpu_port, gt_oracle as go
    import rmpe_b200
    samples = []
    for i in range(n_local):
        s = rmpe_b200.synth.gt_sample(seed0 + i, persons, SRC_HW, True)
        flip, deg, crop, scale = s["aug"]
        M = go.affine_closed_form(flip, deg, crop, scale, s["objpos"][0], s["scale_provided"][0])
        samples.append((s["img"], s["mask"], s["joints"], M, flip))
    _W["samples"], _W["port"] = samples, cpu_port


def _cpu_work(n):
    samples, port = _W["samples"], _W["port"]
    for i in range(n):
        img, mask, joints, M, flip = samples[i % len(samples)]
        port.gt_sample(img, mask, joints, M, flip)
    return n


class CpuArm:
    def __init__(self, cores=None, persons=PERSONS):
        self.cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init, initargs=(0, 16, persons))
        self.pool.map(_cpu_work, [1] * self.cores)  # page everything in

    def run(self, n_samples):
        """Process n_samples spread over the pool; returns wall seconds."""
        per = max(1, n_samples // (self.cores * 4))
        chunks = [per] * (n_samples // per)
        if n_samples - per * len(chunks) > 0:
            chunks.append(n_samples - per * len(chunks))
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_work, chunks, chunksize=1))
        dt = time.perf_counter() - t0
        assert done == n_samples
        return dt

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(target_seconds=8.0):
    arm = CpuArm()
    dt = arm.run(BATCH)
    rate = BATCH / dt
    reps = int(max(1, min(40, round(target_seconds * rate / BATCH))))
    dt = arm.run(BATCH * reps)
    arm.close()
    return {"value": BATCH * reps / dt, "unit": "samples/s", "cores": arm.cores, "kind": "port",
            "sample": "%d x the 256-sample GT batch (%d samples, 3 persons, random aug) through cv2.warpAffine + "
                      "cv2.resize + the NumPy rasteriser, one process per core" % (reps, BATCH * reps)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    for _ in range(args.warmup):
        arm.run(BATCH)
    t = 0.0
    for _ in range(args.steps):
        t += arm.run(BATCH)
    arm.close()
    ms = t / args.steps * 1e3
    v = BATCH / (ms * 1e-3)
    sample = "each step = the 256-sample GT batch over %d processes (cv2 + NumPy, oracle/cpu_port.py)" % arm.cores
    print(json.dumps({
        "impl": "reference", "metric": "gt_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "persons": PERSONS,
                   "labels": "f64 (57,46,46) (reference dtype)", "image": "u8 HWC 368x368x3"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def make_gt_inputs(rmpe, seed0, batch=BATCH, persons=PERSONS):
    b = rmpe.synth.gt_batch(batch, n_persons=persons, seed0=seed0)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    M = rmpe.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]],
                              [a[3] for a in b["augs"]], b["centers"], b["scale_self"])
    b["flip"], b["M"] = flip, M
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-decode", action="store_true", help="skip the secondary decode metric")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU leg first: forked workers must not inherit a CUDA context
    cpu = None
    under_profiler = "NV_NSIGHT_INJECTION_TRANSPORT_TYPE" in os.environ   # set by ncu for the processes it launches
    if under_profiler and not args.no_cpu:
        # ncu injects into every child process; the forked CPU workers crash it (SIGSEGV).  A number printed under a
        # profiler is never a bench value anyway.
        print("bench.py: running under Nsight Compute, skipping the cpu_baseline leg", file=sys.stderr)
    if rank == 0 and world == 1 and not args.no_cpu and not under_profiler:
        cpu = cpu_baseline()

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

    # ---- inputs: NPOOL rotating buffer sets (each ~370 MB in+out, > L2 126 MB) ----
    NPOOL = 3
    hosts = [make_gt_inputs(rmpe_b200, 100000 * rank + 1000 * k) for k in range(NPOOL)]
    plans = []
    for hb in hosts:
        p = rmpe_b200.batch.GtDevicePlan(BATCH, PERSONS, SRC_HW)
        p.upload(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"])
        plans.append(p)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()

    def timed(fn, steps, warmup):
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

    # ---- value: device-resident, kernels only ----
    ms, launches = timed(lambda i: plans[i % NPOOL].run(), args.steps, args.warmup)
    value = world * BATCH / (ms * 1e-3)
    status_bad = int(sum(int((p.status != 0).sum().item()) for p in plans))

    # ---- per-kernel device time over the same K steps (events around every launch) ----
    L.profile_enable(True)
    timed(lambda i: plans[i % NPOOL].run(), args.steps, 1)
    L.profile_enable(False, reset=False)
    prof = L.profile_read()
    L.profile_enable(False, reset=True)
    per = gt_bytes_per_sample(PERSONS)
    kernels = {}
    for name, (tms, n) in prof.items():
        if n == 0:
            continue
        avg = tms / n
        kb = per.get(name, 0) * BATCH
        kernels[name] = {"ms_per_launch": avg, "launches": n, "algorithmic_bytes": kb,
                         "achieved_gbs": kb / (avg * 1e-3) / 1e9 if kb else None}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    traffic_db = {}
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:  # noqa: BLE001
        pass
    limiters = {}
    try:
        limiters = json.load(open(os.path.join(ROOT, "profiles", "limiters.json")))
    except Exception:  # noqa: BLE001
        pass
    roofline = None
    if kernels:
        top = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches"])
        kt = kernels[top]
        tot = sum(k["ms_per_launch"] * k["launches"] for k in kernels.values())
        roofline = {"bound": "hbm", "kernel": top, "achieved": kt["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": (kt["achieved_gbs"] or 0.0) / peak, "traffic": traffic_db.get(top),
                    "peak_source": peak_src, "kernel_ms": kt["ms_per_launch"],
                    "kernel_share_of_step": kt["ms_per_launch"] * kt["launches"] / tot if tot else None,
                    "step_achieved": per["step"] * BATCH / (ms * 1e-3) / 1e9,
                    "step_frac": per["step"] * BATCH / (ms * 1e-3) / 1e9 / peak,
                    # what ncu says holds the kernel below the HBM roofline (from the committed capture, not live)
                    "limiter": limiters.get(top)}

    # ---- e2e: host (pinned) buffers in, host buffers out, through rmpe_gt_batch_host ----
    e2e = None
    if not args.no_e2e:
        hb = hosts[0]
        keep = []
        pin_in = {}
        for k, dt in (("imgs", torch.uint8), ("masks", torch.uint8), ("joints", torch.float64)):
            t, a = pinned(hb[k].shape, dt)
            a[...] = hb[k]
            keep.append(t)
            pin_in[k] = a
        out = {}
        for k, shape, dt in (("img", (BATCH, 368, 368, 3), torch.uint8), ("mask", (BATCH, 46, 46), torch.float32),
                             ("labels", (BATCH, 57, 46, 46), torch.float32),
                             ("joints", (BATCH, PERSONS, 18, 3), torch.float64)):
            t, a = pinned(shape, dt)
            keep.append(t)
            out[k] = a
        h2d = pin_in["imgs"].nbytes + pin_in["masks"].nbytes + pin_in["joints"].nbytes + BATCH * (4 + 48 + 1 + 32)
        d2h = sum(a.nbytes for a in out.values()) + BATCH * 4

        def e2e_step(i):
            rmpe_b200.batch.gt_batch_host(pin_in["imgs"], pin_in["masks"], pin_in["joints"], hb["n_persons"],
                                          hb["M"], hb["flip"], out=out)

        n_e2e = max(3, min(args.steps, 10))
        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_step(i)
        torch.cuda.synchronize(dev)
        dt = max_ranks((time.perf_counter() - t0) / n_e2e)
        e2e = {"value": world * BATCH / dt, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": n_e2e,
               "api": "rmpe_gt_batch_host (pinned host buffers; copies inside the call)"}
        # what the interconnect alone takes for these bytes: the same pinned buffers copied in and out on two streams
        # at once, no kernels (explains e2e against the device-resident value; not part of any timed figure above)
        big_in = [keep[0], keep[1]]                       # imgs, masks
        big_out = [keep[3], keep[5]]                      # img, labels
        dev_in = [torch.empty_like(t, device=dev) for t in big_in]
        dev_out = [torch.empty_like(t, device=dev) for t in big_out]
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def copies():
            with torch.cuda.stream(s_in):
                for d_, h_ in zip(dev_in, big_in):
                    d_.copy_(h_, non_blocking=True)
            with torch.cuda.stream(s_out):
                for d_, h_ in zip(dev_out, big_out):
                    h_.copy_(d_, non_blocking=True)

        copies()
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            copies()
        torch.cuda.synchronize(dev)
        floor = max_ranks((time.perf_counter() - t0) / 5)
        e2e["copy_floor_ms"] = floor * 1e3
        e2e["frac_of_copy_floor"] = floor / dt
        del dev_in, dev_out

    # ---- secondary metric: single-scale decode of ski.jpg-shaped frames ----
    decode = None
    if not args.no_decode:
        H, W = DEC_HW
        h, w = rmpe_b200.synth.single_scale_grid(H, W)
        frames = []
        for i in range(DEC_FRAMES):
            paf, heat, _ = rmpe_b200.synth.decode_blobs(9000 + 100 * rank + i, (H, W), (h, w), PERSONS)
            frames.append(dict(H=H, W=W, scales=[(paf, heat, 0, 0)]))
        dplans = [rmpe_b200.batch.DecodeDevicePlan(frames) for _ in range(2)]
        dsteps = max(3, min(args.steps, 10))
        # the blobs of a step (14 MB) fit in L2: flush it (write 256 MB) before every timed step and time the
        # steps one by one, so that every step reads its blobs from HBM
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed_flushed(fn, steps, warmup):
            for i in range(warmup):
                fn(i)
            barrier()
            n0 = lib.rmpe_launch_count()
            tot = 0.0
            for i in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(warmup + i)
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
            barrier()
            return max_ranks(tot / steps), lib.rmpe_launch_count() - n0

        dms, dl = timed_flushed(lambda i: dplans[i % 2].run(), dsteps, 3)
        L.profile_enable(True)
        timed(lambda i: dplans[i % 2].run(), dsteps, 1)
        L.profile_enable(False, reset=False)
        dprof = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in L.profile_read().items() if v[1]}
        L.profile_enable(False, reset=True)
        fb = decode_bytes_per_frame(H, W, h, w)
        # heat half of the staged dataflow (what k_heat_screen + k_peak_verify replace): blob in, 18 maps
        # written by the resize and read back by the smoothing; the PAF half (38 up-sampled planes) is never
        # produced at all -- k_limbs samples it from the blob
        heat_b = 4 * (19 * h * w + 36 * H * W)
        dtop = max(dprof, key=lambda k: dprof[k]["ms_per_launch"] * dprof[k]["launches"]) if dprof else None
        droof = None
        if dtop:
            # decode is accounted with the reference's STAGED dataflow, B1 = 4(57hw + 74HW) bytes per
            # frame ("fusing stages is allowed and simply scores higher"); the heat kernels here read only the blob
            # (4*19hw bytes) because the up-sampled maps are never materialised, so `fused_frac` is reported beside it
            heat_ms = sum(dprof[k]["ms_per_launch"] * dprof[k]["launches"] for k in dprof
                          if k in ("k_screen_plan", "k_screen_pairs", "k_peak_verify", "k_peaks_finalize")) / dsteps
            ach = fb * DEC_FRAMES / (dms * 1e-3) / 1e9
            fused = 4 * 57 * h * w * DEC_FRAMES / (dms * 1e-3) / 1e9
            droof = {"bound": "hbm", "kernel": "decode step (all kernels; dominant: %s)" % dtop,
                     "kernel_ms": dprof[dtop]["ms_per_launch"], "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic_db.get(dtop),
                     "algorithmic": "staged reference dataflow 4*(57hw + 74HW) B/frame (SURVEY.md 8d)",
                     "fused_bytes_per_frame": 4 * 57 * h * w, "fused_achieved": fused, "fused_frac": fused / peak,
                     "heat_stage_ms": heat_ms, "heat_stage_staged_bytes_per_frame": heat_b}
        decode = {"metric": "decoded_frames_per_s", "value": world * DEC_FRAMES / (dms * 1e-3), "unit": "frames/s",
                  "ms_per_step": dms, "steps": dsteps, "gpu_launches": int(dl),
                  "config": {"workload": "single_scale_decode_674x712_84x89_blobs_3persons (configs[2] per-GPU share)",
                             "frames_per_gpu": DEC_FRAMES, "l2": "flushed (256 MB write) before every timed step"},
                  "staged_bytes_per_frame": fb,
                  "staged_frac_of_hbm": fb * DEC_FRAMES / (dms * 1e-3) / 1e9 / peak,
                  "roofline": droof, "kernels": dprof}

    clocks = sampler.finish()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "gt_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH,
                   "persons": PERSONS, "labels": "f32 (57,46,46)", "image": "u8 HWC 368x368x3",
                   "l2": "rotating %d buffer sets of %.0f MB (in+out) > 126 MB L2" % (
                       NPOOL, per["step"] * BATCH / 1e6),
                   "parallelism": "samples sharded by rank, no collective"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": cpu, "kernels": kernels, "decode": decode, "status_nonzero": status_bad,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
