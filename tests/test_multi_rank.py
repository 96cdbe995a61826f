"""World-sharding logic of bench.py across ranks (config 5), exercised with a world_size-2 gloo group on CPU: the
shards partition the batch exactly, there is no data-path collective, and the timing reduction is a MAX over ranks."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_batch():
    sys.path.insert(0, ROOT)
    import bench
    for total in (1, 7, 148, 8192):
        for world_size in (1, 2, 3, 4, 8):
            covered = []
            for rank in range(world_size):
                lo, hi = bench.shard(total, rank, world_size)
                assert 0 <= lo <= hi <= total
                covered += list(range(lo, hi))
            assert covered == list(range(total))
            sizes = [bench.shard(total, r, world_size)[1] - bench.shard(total, r, world_size)[0] for r in range(world_size)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_max_over_ranks_and_sharded_stepping(tmp_path):
    """Two processes, gloo: each steps ITS shard of a batch of (emulated) worlds independently, then the per-rank times
    are reduced with MAX and the per-rank world hashes gathered: both shards evolved identically, no collective was
    needed for the step itself."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, hashlib
        sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
        import torch, torch.distributed as dist
        import bench, harness as H
        from forge2d_b200 import scenes
        dist.init_process_group("gloo")
        rank, world_size = dist.get_rank(), dist.get_world_size()
        lo, hi = bench.shard(6, rank, world_size)
        emu = H.load("emu")
        t = scenes.bench2d(emu, rows=5)
        batch = emu.f2dBatch_Create(t.world, hi - lo)
        dist.barrier()
        emu.f2dBatch_StepN(batch, scenes.TIME_STEP, scenes.SUB_STEPS, 20)
        dist.barrier()
        emu.f2dBatch_DownloadWorld(batch, hi - lo - 1, t.world)
        digest = H.state_hash(H.snapshot(emu, t.world, trees=False))
        ms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out = [None] * world_size
        dist.all_gather_object(out, (lo, hi, digest))
        if rank == 0:
            assert ms.item() == 10.0 + world_size - 1
            assert [o[:2] for o in out] == [(0, 3), (3, 6)]
            assert out[0][2] == out[1][2]
            print("OK")
        dist.destroy_process_group()
    """ % (ROOT, ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout
