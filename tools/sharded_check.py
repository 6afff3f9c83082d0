"""Multi-rank run of one model through the sharded driver (torchrun); rank 0 prints one JSON line.
Used by the 2-GPU test to check verdict, counts and the cross-rank error trace."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kafka_specification_b200.sharded import CudaShardEngine, ShardedChecker  # noqa: E402


def main():
    model = sys.argv[1]
    cont = "cont" in sys.argv[2:]
    p2p = "nccl" not in sys.argv[2:]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    eng = CudaShardEngine(model, rank, world, local, table_log2=22, max_states=2_000_000, p2p=p2p)
    res = ShardedChecker(eng, cont=cont).run()
    if rank == 0:
        print(json.dumps({"distinct": res.distinct, "generated": res.generated, "depth": res.depth,
                          "levels": res.levels, "complete": res.complete, "violation": res.violation,
                          "trace": [{"words": t["words"], "rank": t["rank"], "action": t.get("action_name")} for t in res.trace],
                          "p2p": eng.p2p}), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
