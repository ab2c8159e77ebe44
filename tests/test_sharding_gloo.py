"""world_size-2 gloo run of the host-side sharding logic (the N>1 path of bench.py / the CLI): every
rank computes its contiguous share (here with the CPU emulation of the device algorithm), the shares
tile the catalog exactly, and the all-reduced checksum equals the single-process one."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from egg_b200.sharding import flux_checksum, shard_galaxies, shard_range


def test_shard_ranges_tile_the_catalog():
    for n in (0, 1, 7, 16, 195_000):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import emul
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = cases.make("b4_defaults", ngal=41, seed=5)
    mine = shard_galaxies(c["gal"], rank, world)
    out = emul.compute(c["opt"], c["ir"], c["bands"], c["rfbands"], mine, line_lambda=c["lines"], **c["opts"])
    chk = torch.from_numpy(flux_checksum(out["flux_disk"] + out["flux_bulge"]))
    dist.all_reduce(chk)
    a, b = shard_range(41, rank, world)
    q.put((rank, a, b, chk.numpy(), out["flux_disk"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_the_catalog_and_agree_with_one():
    import emul
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=300) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = cases.make("b4_defaults", ngal=41, seed=5)
    full = emul.compute(c["opt"], c["ir"], c["bands"], c["rfbands"], c["gal"], line_lambda=c["lines"], **c["opts"])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 41
    np.testing.assert_array_equal(np.concatenate([res[0][4], res[1][4]]), full["flux_disk"])
    np.testing.assert_allclose(res[0][3], flux_checksum(full["flux_disk"] + full["flux_bulge"]), rtol=1e-12)
    np.testing.assert_array_equal(res[0][3], res[1][3])


def _props_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import props_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    z, m, pas = rng.uniform(0.1, 7, 57), rng.uniform(8, 11.5, 57), rng.random(57) < 0.3
    a, b = shard_range(57, rank, world)
    # a rank passes the catalog index of its first galaxy: that is what selects the random streams (eggprops.h)
    out = props_oracle.compute(z[a:b], m[a:b], pas[a:b], opt_lib=cases.libs()["opt"], seed=21, first_index=a)
    q.put((rank, out["sfr"], out["opt_sed_disk"], out["line_lum_disk"]))
    dist.barrier()
    dist.destroy_process_group()


def test_property_draws_do_not_depend_on_the_number_of_ranks():
    from oracle import props_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_props_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=300) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    z, m, pas = rng.uniform(0.1, 7, 57), rng.uniform(8, 11.5, 57), rng.random(57) < 0.3
    one = props_oracle.compute(z, m, pas, opt_lib=cases.libs()["opt"], seed=21)
    for k, i in (("sfr", 1), ("opt_sed_disk", 2), ("line_lum_disk", 3)):
        np.testing.assert_array_equal(np.concatenate([res[0][i], res[1][i]]), one[k])
