"""1 rank vs N ranks give the identical gathered score table.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multirank_check.py [n_total]

Every rank enhances its contiguous shard of the utterance index in fixed sub-batches of 16 (so an utterance's result
does not depend on how many ranks there are), scores it, and the score rows are all-gathered with NCCL.  Rank 0 also
computes every utterance itself and compares bit for bit.  Prints one JSON line on rank 0; exit code 1 on a mismatch."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import avzoom  # noqa: E402
from avzoom import ops, parallel, pipeline, synth  # noqa: E402

SUB = 16


def scores_of(eng, lo, hi, dev):
    rows = []
    for b0 in range(lo, hi, SUB):
        mix, tgt, itf = synth.make_batch(2, SUB, 1.0, 3, start=b0)
        mix, tgt, itf = (torch.from_numpy(a).to(dev) for a in (mix, tgt, itf))
        out = eng.run(mix, tgt, itf)
        rows.append(ops.sir_scores(out, tgt, itf).clone())
    return torch.cat(rows)


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert n_total % (SUB * world) == 0, "n_total must be a multiple of 16 * world"
    eng = pipeline.OracleMvdr(avzoom.PRESETS["baseline_oracle"], SUB, 16000, dev)
    lo, hi = parallel.shard_range(n_total, rank, world)
    gathered = parallel.gather_scores(scores_of(eng, lo, hi, dev), n_total)
    ok = True
    if rank == 0:
        alone = scores_of(eng, 0, n_total, dev)
        ok = bool(torch.equal(gathered, alone)) and bool(torch.isfinite(alone).all())
        print(json.dumps({"world": world, "utterances": n_total, "gathered_equals_single_rank": ok,
                          "osir_mean_dB": float(alone[:, 1].mean())}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
