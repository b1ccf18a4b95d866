"""N>1 path on CPU: world_size-2 gloo run of the sharding and gather plumbing used by bench.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import jobgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, js_bytes, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import jobgen as jg
    from oracle import oracle_py
    par = jg.pkg.__class__  # noqa: F841
    from __graft_entry__ import load_package
    pkg = load_package()
    import importlib
    parallel = importlib.import_module("airs_compression_b200.parallel")
    abi = pkg.abi
    jobs = np.frombuffer(js_bytes["jobs"], dtype=abi.JOB_DTYPE).copy()
    b, e = parallel.shard_range(len(jobs), rank, world)
    mine = jobs[b:e].copy()
    base = int(mine["first_result"][0]) if len(mine) else 0
    mine["first_result"] -= base
    js = dict(src=np.frombuffer(js_bytes["src"], dtype=np.uint8).copy(), jobs=mine, dst_size=js_bytes["dst_size"],
              work_size=js_bytes["work_size"], n_results=int(mine["n_frames"].sum()), layout=1)
    dst, res, init, offs, _ = jg.run_cpu(oracle_py.load_oracle(), js)   # each rank compresses its shard
    sizes, counts = parallel.allgather_sizes(torch.from_numpy(res.astype(np.int64)))
    stream, offsets = parallel.allgather_streams(torch.from_numpy(dst[:int(offs[-1])].copy()))
    if rank == 0:
        ret["sizes"] = sizes.numpy().astype(np.uint32)
        ret["stream"] = stream.numpy()
        ret["counts"] = counts
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("method", ["sendrecv", "allgather"])
def test_two_rank_shard_and_gather(oracle, method, monkeypatch):
    """Both gathers of parallel.py: grouped send / receive into place (the default) and the padded all-gather."""
    monkeypatch.setenv("AIRS_GATHER", method)   # (inherited by the spawned ranks)
    rng = np.random.default_rng(77)
    js = jobgen.build_jobs(rng, 41, sizes=[5, 64, 257, 2048], max_frames=3, layout=1)
    want_dst, want_res, _, want_offs, _ = jobgen.run_cpu(oracle, js)
    payload = dict(jobs=js["jobs"].tobytes(), src=js["src"].tobytes(), dst_size=js["dst_size"], work_size=js["work_size"])
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, payload, ret), nprocs=2, join=True)
    assert np.array_equal(ret["sizes"], want_res)
    total = int(want_offs[-1])
    assert len(ret["stream"]) == total and np.array_equal(ret["stream"], want_dst[:total])


def test_shard_range_partitions(pkg):
    import importlib
    parallel = importlib.import_module("airs_compression_b200.parallel")
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 4, 8):
            cuts = [parallel.shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in cuts) - min(e - b for b, e in cuts) <= 1
