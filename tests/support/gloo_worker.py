"""Worker process of the world_size-2 gloo test: runs the sharded driver over the host stand-in."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))

import torch.distributed as dist  # noqa: E402

from host_shard_engine import HostShardEngine  # noqa: E402
from kafka_specification_b200.sharded import ShardedChecker  # noqa: E402


def main():
    name, out_path, chunk = sys.argv[1], sys.argv[2], int(sys.argv[3])
    cont = len(sys.argv) > 4 and sys.argv[4] == "cont"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = HostShardEngine(name, rank, world, chunk_states=chunk)
    res = ShardedChecker(eng, cont=cont).run()
    trace_ok = None
    if res.trace:
        steps = [t["words"] for t in res.trace]
        trace_ok = bool(eng.is_init(steps[0])) and all(eng.is_successor(a, b) >= 0 for a, b in zip(steps, steps[1:]))
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"distinct": res.distinct, "generated": res.generated, "depth": res.depth,
                       "deadlocks": res.deadlocks, "levels": res.levels, "complete": res.complete,
                       "violation": res.violation, "per_rank": res.per_rank_distinct, "trace_len": len(res.trace),
                       "trace_ok": trace_ok, "trace_ranks": sorted({t["rank"] for t in res.trace}),
                       "exchanged_rows": res.exchanged_rows}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
