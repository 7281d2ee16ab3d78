"""N>1 host logic on CPU: two gloo ranks shard one batch by bytes, each encodes its shard (the
oracle stands in for the device here -- this test is about the sharding/stitching arithmetic, the
GPU parity tests cover the kernels), exchange only their token counts, and the stitched result
must equal the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tekken_rs_b200 import corpus, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tekken_json, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tekken_oracle as TO
    orc = TO.OracleTekkenizer.from_file(tekken_json)
    data, off = corpus.mixed_script_docs(600, seed=5)
    b, e = sharding.local_shard(off, rank, world)
    loc_off = sharding.rebase_offsets(off, b, e)
    loc = data[int(off[b]):int(off[e])]
    ids, toff = orc.encode_batch_np(loc, loc_off, True, True)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(ids)], dtype=torch.int64))
    counts = [int(c.item()) for c in counts]
    g_off = sharding.stitch_token_offsets(toff, counts, rank)
    np.save(os.path.join(out_dir, "ids_%d.npy" % rank), ids)
    np.save(os.path.join(out_dir, "off_%d.npy" % rank), g_off)
    np.save(os.path.join(out_dir, "rng_%d.npy" % rank), np.array([b, e]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path, tekken_json, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), tekken_json, str(tmp_path)), nprocs=world, join=True)
    data, off = corpus.mixed_script_docs(600, seed=5)
    ref_ids, ref_off = oracle.encode_batch_np(data, off, True, True)
    ids = np.concatenate([np.load(tmp_path / ("ids_%d.npy" % r)) for r in range(world)])
    assert np.array_equal(ids, ref_ids)
    rngs = [np.load(tmp_path / ("rng_%d.npy" % r)) for r in range(world)]
    assert rngs[0][0] == 0 and rngs[0][1] == rngs[1][0] and rngs[1][1] == 600
    sizes = [int(off[r[1]] - off[r[0]]) for r in rngs]
    assert abs(sizes[0] - sizes[1]) <= 2048                       # byte-balanced
    stitched = np.concatenate([np.load(tmp_path / "off_0.npy")[:-1], np.load(tmp_path / "off_1.npy")])
    assert np.array_equal(stitched, ref_off)
