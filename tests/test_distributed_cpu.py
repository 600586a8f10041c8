"""CPU, world_size 2, gloo: the data-parallel plumbing -- batch sharding, DDP wrapping (gradient averaging ==
single-process gradient on the concatenated batch / world), scalar reductions and the max-over-ranks timer."""
import os
import socket

import torch
import torch.multiprocessing as mp

from swin_v2_weather_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    ddp = D.wrap_ddp(model)
    g = torch.Generator().manual_seed(1)
    x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    idx = list(D.shard_indices(8, rank, world))
    assert len(idx) == D.local_batch_size(8, world)
    loss = ((ddp(x_all[idx]) - y_all[idx]) ** 2).sum()          # sum over the local batch, like the reference loss
    loss.backward()
    grads = [p.grad.clone() for p in model.parameters()]
    logged = D.all_reduce_mean_scalar(loss.detach().clone())
    tmax = D.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        out.put(([g.tolist() for g in grads], float(logged), tmax))   # plain lists: no shared-memory handles
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_ddp_gloo_world2_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads, logged, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(1)
    x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    loss = ((model(x_all) - y_all) ** 2).sum()
    loss.backward()
    for got, p in zip(grads, model.parameters()):
        assert torch.allclose(torch.tensor(got), p.grad / world, rtol=1e-5, atol=1e-6)   # DDP averages the per-rank sums
    assert abs(logged - float(loss) / world) < 1e-4
    assert tmax == 2.0


def test_sharding_helpers():
    assert D.local_batch_size(64, 8) == 8
    import pytest
    with pytest.raises(ValueError):
        D.local_batch_size(10, 4)
    seen = []
    for r in range(4):
        seen += list(D.shard_indices(16, r, 4))
    assert seen == list(range(16))
    m = torch.nn.Linear(2, 2)
    assert D.wrap_ddp(m) is m      # single process: no wrapping


def test_numa_binding_never_raises():
    """bind_to_gpu_numa_node is best effort: without NVML / a GPU it reports False instead of failing the launch."""
    from swin_v2_weather_b200 import distributed as D
    assert D.bind_to_gpu_numa_node(0) in (True, False)
    assert D.bind_to_gpu_numa_node(99) is False
