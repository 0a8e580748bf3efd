"""Grid-merge micro-benchmark (run under torchrun, one rank per GPU): the 512^3 float32 grid merged by NCCL reduce / all_reduce
and by pvp_grid_reduce_peers (NVSwitch multicast ld_reduce or plain peer loads; to rank 0 or to every rank).
CUDA-event time on every rank, max over ranks, best of `reps`."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pixeltovoxelprojector_b200 as pvp  # noqa: E402
from pixeltovoxelprojector_b200 import distributed as pd  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = 8
    g = pd.symmetric_grid(n, 2.0, (0.0, 0.0, 500.0))
    plain = torch.zeros_like(g.data)
    ctx = pvp.default_context()

    def timed(fn):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best
    res = {
        "nccl reduce -> rank 0": timed(lambda: dist.reduce(plain, dst=0)),
        "nccl all_reduce": timed(lambda: dist.all_reduce(plain)),
        "peers multicast -> rank 0 (every rank reduces a slice)": timed(lambda: pd.reduce_grid_peers(g, dst=0, ctx=ctx, multicast=True)),
        "peers multicast -> rank 0 (rank 0 pulls everything)": (ctx.set_option("reduce_pull", 1), timed(lambda: pd.reduce_grid_peers(g, dst=0, ctx=ctx, multicast=True)),
                                                                 ctx.set_option("reduce_pull", 0))[1],
        "peers multicast -> all": timed(lambda: pd.reduce_grid_peers(g, dst=None, ctx=ctx, multicast=True)),
        "peers plain loads -> rank 0": timed(lambda: pd.reduce_grid_peers(g, dst=0, ctx=ctx, multicast=False)),
        "two symmetric-memory barriers alone": timed(lambda: (g.symm.barrier(channel=0), g.symm.barrier(channel=1))),
    }
    for unroll in (4, 8):                 # bytes in flight of the merge kernel: loads per thread x CTAs per SM
        for per_sm in (4, 8, 16):
            ctx.set_option("reduce_unroll", unroll)
            ctx.set_option("reduce_blocks_per_sm", per_sm)
            res[f"peers multicast -> all, unroll {unroll}, {per_sm} CTAs/SM"] = timed(lambda: pd.reduce_grid_peers(g, dst=None, ctx=ctx, multicast=True))
            res[f"peers multicast -> rank 0, unroll {unroll}, {per_sm} CTAs/SM"] = timed(lambda: pd.reduce_grid_peers(g, dst=0, ctx=ctx, multicast=True))
    ctx.set_option("reduce_unroll", 0)
    ctx.set_option("reduce_blocks_per_sm", 0)
    if rank == 0:
        mib = n ** 3 * 4 / 2 ** 20
        for k, v in res.items():
            print(f"world={world} grid {n}^3 f32 ({mib:.0f} MiB): {k}: {v:.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
