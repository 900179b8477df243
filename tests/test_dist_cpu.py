"""N>1 host logic on CPU: gloo backend, world_size 2 (spawned processes, rendezvous on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trajsde_b200.dist import FlatGradBucket, shard_row_offsets, shard_scenes


def test_shard_scenes_partitions_exactly():
    for n, w in [(8192, 8), (1024, 1), (10, 4), (7, 8), (0, 2)]:
        spans = [shard_scenes(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - s for s, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_scenes(4, 2, 2)
    o = shard_row_offsets(8192, 20, 8, 3)
    assert o == {'scene_start': 3072, 'enc_agent_row': 61440, 'dec_row': 614400}


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, pack=False):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(66, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1))
    bucket = FlatGradBucket(net.parameters(), pack=pack)
    assert bucket.numel == 66 * 64 + 64 + 64 + 1
    # scene-sharded batch: each rank sees its own rows of one global batch
    g = torch.Generator().manual_seed(123)
    x_all = torch.randn(40, 66, generator=g)
    s, e = shard_scenes(40, world, rank)
    bucket.zero_()
    net(x_all[s:e]).sum().div(e - s).backward()
    if pack:                                                        # autograd handed over its own tensors: nothing in the bucket yet
        assert all(p.grad is not None and p.grad.data_ptr() != v.data_ptr() for p, v in zip(bucket.params, bucket._views))
    else:
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())   # grads accumulated into the bucket
    if rank == 0:
        bucket.all_reduce_mean()
    else:
        bucket.all_reduce_mean_async().wait()                      # same collective through the async entry point
    # reference: gradient of the mean over per-rank means, computed in one process
    ref = torch.nn.Sequential(torch.nn.Linear(66, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1))
    ref.load_state_dict(net.state_dict())
    loss = sum(ref(x_all[slice(*shard_scenes(40, world, r))]).sum() / (shard_scenes(40, world, r)[1] - shard_scenes(40, world, r)[0])
               for r in range(world)) / world
    loss.backward()
    flat_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    ok = torch.allclose(bucket.flat, flat_ref, atol=1e-6, rtol=1e-5)
    ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket._views))   # the optimizer reads the reduced values
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(t.item() == 1.0))
    dist.destroy_process_group()


@pytest.mark.parametrize('pack', [False, True])
def test_flat_bucket_allreduce_gloo_world2(pack):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, pack)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_bucket_without_process_group_is_a_noop():
    net = torch.nn.Linear(4, 2)
    b = FlatGradBucket(net.parameters())
    net(torch.ones(3, 4)).sum().backward()
    before = b.flat.clone()
    assert b.all_reduce_mean() is None
    assert torch.equal(before, b.flat)
