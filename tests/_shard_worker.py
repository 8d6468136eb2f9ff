"""Worker of tests/test_host.py::test_sharding_gloo_world2 (one process per rank, gloo).
Each rank answers its slice of a seeded query batch with the host build of the product's query
code (tests/hostemu; the CUDA path needs a GPU), the slices are gathered over gloo, and every
rank checks the result against the unsharded answer."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import gather_shards, shard_slice
    import conftest
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = conftest.hostemu()
    img = conftest.navmesh_image("c2_apartment")
    h = C.c_void_p(emu.emu_create(img, len(img)))
    f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint)
    P = lambda a, t: a.ctypes.data_as(t)  # noqa: E731

    def find(st, en):
        d = np.zeros(len(st), np.float32)
        emu.emu_find_path(h, P(st, f32p), P(en, f32p), C.c_long(len(st)), 2048, 1, P(d, f32p), None, None, 0,
                          None, None, None)
        return d

    def rand(n, q0):
        out = np.zeros((n, 3), np.float32)
        emu.emu_random_points(h, C.c_long(n), 10, None, C.c_ulonglong(7), C.c_ulonglong(q0), P(out, f32p), None)
        return out

    n = 1001  # not divisible by world: ragged slices
    pts = conftest.query_points("c2_apartment", 2 * n, 77)
    st, en = np.ascontiguousarray(pts[:n]), np.ascontiguousarray(pts[n:])
    b, e = shard_slice(n, rank, world)
    d = gather_shards(find(st[b:e], en[b:e]), n, rank, world)
    r = gather_shards(rand(e - b, b), n, rank, world)
    assert np.array_equal(d.view(np.uint32), find(st, en).view(np.uint32)), "sharded distances differ"
    assert np.array_equal(r.view(np.uint32), rand(n, 0).view(np.uint32)), "random points depend on the sharding"
    # granule-aligned slices (multi-goal: G goals of a start stay together), equal sizes -> all_gather
    g = 4
    b, e = shard_slice(n // g * g * world, rank, world, granule=g)
    assert b % g == 0 and e % g == 0
    x = gather_shards(np.arange(b, e, dtype=np.int64), n // g * g * world, rank, world, granule=g)
    assert np.array_equal(x, np.arange(n // g * g * world))
    # the timing reduction bench.py uses: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()
    emu.emu_destroy(h)
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
