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


class BoardStandIn(HostShardEngine):
    """Stand-in for the device-synchronised entry points (kmc_shard_round_p2p / kmc_shard_level_sync): a round =
    expand + exchange + insert, a level end = every rank's summary row gathered into one board that all ranks read.
    The exchange goes through gloo instead of peer memory; what is under test is the DRIVER's side of that protocol
    (ShardedChecker._run_device_sync: chunk loops on idle ranks, termination, violation stop, totals from the board)."""
    p2p = True
    device_sync = True
    driver = None

    def round_p2p(self, first, count, seed=False):
        if seed:
            self.seed_init()
        else:
            self.reset_cand()
            if count:
                self.expand(first, count)
        rows = self.driver._exchange(self.counts())
        if rows:
            self.insert_received(rows)

    def level_sync(self):
        self._level = getattr(self, "_level", 0) + 1
        _, count = self.level_done()
        st = self.stats()
        row = [self._level, count, 1 if self.violation() else 0, st["distinct"], st["generated"], st["fail"], st["deadlocks"], 0]
        return self.driver._all_gather_counts(row)


def main():
    name, out_path, chunk = sys.argv[1], sys.argv[2], int(sys.argv[3])
    cont = "cont" in sys.argv[4:]
    board = "board" in sys.argv[4:]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = (BoardStandIn if board else HostShardEngine)(name, rank, world, chunk_states=chunk)
    drv = ShardedChecker(eng, cont=cont)
    eng.driver = drv
    res = drv.run()
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
