"""A/B aid for the decode kernels: device-side time (CUDA events), per-kernel split and a digest of the results of five
decode workloads (8 / 64 ski-shaped single-scale frames, 8 crowded, 8 dense, N COCO-val-shaped multi-scale frames).
The build switches are environment variables read once per process (RMPE_SCREEN_CULL, RMPE_SS_GROUP, RMPE_MS_GROUP), so
one variant = one process; equal digests = bit-identical candidate / connection / subset tables.

   RMPE_SCREEN_CULL=0 python tools/decode_ab.py [n_multi_scale=256]
"""
import hashlib
import json
import multiprocessing as mp
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (frame synthesis helpers)


def digest(res):
    h = hashlib.sha256()
    for r in res:
        for k in ("candidate", "subset"):
            h.update(np.ascontiguousarray(r[k]).tobytes())
        for c in r["connections"]:
            h.update(b"-" if c is None else np.ascontiguousarray(c).tobytes())
        h.update(str(int(r["status"])).encode())
    return h.hexdigest()[:16]


def main():
    n_ms = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    H, W = 674, 712
    tasks = {
        "ss8": [("ss", 900 + i, H, W, 3) for i in range(8)],
        "ss64": [("ss", 900 + i, H, W, 3) for i in range(64)],
        "crowd8": [("ss", 1900 + i, H, W, 20) for i in range(8)],
        "dense8": [("dense", 2900 + i, H, W, 0) for i in range(8)],
        "ms%d" % n_ms: [("ms", 700 + i, hh, ww, 3) for i, (hh, ww) in enumerate(bench.ms_shape_list()[:n_ms])],
    }
    workers = min(32, os.cpu_count() or 8)
    cache = "/tmp/decode_ab_frames_%d.pkl" % n_ms                               # the variants of one session share the frames
    if os.path.exists(cache):
        frames = pickle.load(open(cache, "rb"))
    else:
        frames = {k: bench.synth_frames(t, workers) for k, t in tasks.items()}     # before CUDA is initialised (fork)
        pickle.dump(frames, open(cache, "wb"), protocol=4)
    import torch
    import rmpe_b200
    L = rmpe_b200.lib
    L.ensure_init(0)
    out = {"env": {k: os.environ.get(k) for k in ("RMPE_SCREEN_CULL", "RMPE_SS_GROUP", "RMPE_MS_GROUP")}}
    only = os.environ.get("DECODE_AB_ONLY")            # e.g. "ss64" (the short command ncu is wrapped around)
    for name, fr in frames.items():
        if only and not name.startswith(only):
            continue
        dp = rmpe_b200.batch.DecodeDevicePlan(fr, max_persons=64)
        for _ in range(3):
            dp.run()
        torch.cuda.synchronize()
        iters = 10 if len(fr) <= 64 else 4
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            dp.run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        L.profile_enable(True)
        for _ in range(3):
            dp.run()
        torch.cuda.synchronize()
        L.profile_enable(False, reset=False)
        kern = {k: round(t / 3 * 1e3, 1) for k, (t, n) in L.profile_read().items()}      # us per pass
        L.profile_enable(False, reset=True)
        res = dp.results()
        out[name] = {"ms": round(ms, 4), "frames_per_s": round(len(fr) / ms * 1e3), "kernels_us_per_pass": kern,
                     "digest": digest(res), "persons": int(sum(len(r["subset"]) for r in res))}
        del dp
    print(json.dumps(out))


if __name__ == "__main__":
    main()
