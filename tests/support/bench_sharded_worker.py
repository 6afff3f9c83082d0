"""Worker of tests/test_bench_contract.py::test_sharded_line_*: runs bench.run_sharded under gloo with the host
stand-in engine in place of the CUDA engine (tests only -- the control flow and the JSON line, not a measurement)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from host_shard_engine import HostShardEngine  # noqa: E402
import kafka_specification_b200.sharded as sharded  # noqa: E402


class StandIn(HostShardEngine):
    """Claims the fused + device-synchronised path when INJECT=1 and fails its first round on every rank, so that
    bench.py's acceptance run must fall back collectively; otherwise the plain exchange path."""

    def __init__(self, model, rank, world, local, **opts):
        super().__init__(model, rank, world, chunk_states=700)
        self.ck = types.SimpleNamespace(words=self.row_words - 1)
        self.device_sync = os.environ.get("INJECT") == "1"

    @property
    def p2p(self):
        return self.device_sync

    def round_p2p(self, first, count, seed=False):
        raise RuntimeError("injected device-sync failure")

    def level_sync(self):
        raise RuntimeError("unreachable")

    def stats(self):
        st = dict(super().stats())
        for k, v in (("gpu_ms_total", 1.0), ("gpu_ms_expand", 0.5), ("gpu_ms_insert", 0.4), ("launches_expand", 3),
                     ("launches_insert", 3), ("launches_other", 3)):
            st.setdefault(k, v)                   # (the host stand-in has no CUDA events)
        return st

    def close(self):
        pass


def main():
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, rank, world_size, device_id=None: real_init("gloo", rank=rank, world_size=world_size)
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    sharded.CudaShardEngine = StandIn
    import bench
    bench.ClockSampler.start = lambda self: None
    sys.argv = ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--model", sys.argv[1]]
    sys.exit(bench.main())


if __name__ == "__main__":
    main()
