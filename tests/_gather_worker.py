"""Worker of tests/test_decode_gpu.py::test_decode_records_gather_across_ranks (one process per rank, launched by
torch.distributed.run): decode this rank's frames on its GPU, turn them into COCO keypoint records, gather them on every
rank, and let rank 0 compare the merged list with a single-process decode of all frames and with the oracle's persons."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import rmpe_b200
    from cases import DECODE_CASES, frames_of, decode_case_inputs
    from oracle import decode_oracle as do
    backend = os.environ["RMPE_TEST_BACKEND"]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    rmpe_b200.lib.ensure_init(local)
    cases = [DECODE_CASES[2], DECODE_CASES[3], DECODE_CASES[6], DECODE_CASES[2], DECODE_CASES[3]]
    image_ids = [101, 202, 303, 404, 505]
    shard = rmpe_b200.sub("shard")
    mine = shard.shard_indices(len(cases), rank, world)
    dec = rmpe_b200.decode
    local_records = []
    if len(mine):
        res = rmpe_b200.batch.decode_batch_host([frames_of(cases[i]) for i in mine])
        for i, r in zip(mine, res):
            assert r["status"] == 0
            local_records.append((int(i), dec.coco_keypoint_records([r["candidate"]], [r["subset"]], [image_ids[i]])))
    merged = dec.gather_records(local_records, len(cases))
    if rank == 0:
        full = rmpe_b200.batch.decode_batch_host([frames_of(c) for c in cases])
        want = dec.coco_keypoint_records([r["candidate"] for r in full], [r["subset"] for r in full], image_ids)
        plain = lambda recs: json.loads(json.dumps(recs, default=lambda v: v.item()))
        assert plain(merged) == plain(want), "gathered records differ from the single-process decode"
        n_oracle = 0
        for c in cases:
            name, H, W, P, seed, multi = c
            b = decode_case_inputs(c)
            _, sub = do.single_scale(b[0][0], b[0][1], H, W)
            n_oracle += len(sub)
        assert len(merged) == n_oracle and n_oracle > 0, (len(merged), n_oracle)
        print("GATHER_OK %d records over %d ranks (%s)" % (len(merged), world, backend), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
