"""world_size-2 `gloo` test of the data-parallel host logic (no GPU): the flat gradient arena is exchanged with ONE
all-reduce and averaged by 1/world inside the optimizer's grad_scale, which must equal training on the concatenated
batch (train_text_to_image_control_lora.py: DDP average of the ControlLoRA gradients)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from controllora_b200.arena import ParamArena
    import controllora_b200 as cb
    from controllora_b200.configs import NAMED

    torch.manual_seed(0)                       # identical initial parameters on every rank (accelerate set_seed)
    cl = cb.ControlLoRA.from_config(NAMED["diffusiondb-canny-v2"])
    arena = ParamArena(list(cl.parameters()), torch.device("cpu"))
    # parameters were re-homed into the arena without changing identity or values
    assert sum(p.numel() for p in cl.parameters()) == arena.numel
    first = next(cl.parameters())
    assert first.data.data_ptr() == arena.flat_p.data_ptr()
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(arena.flat_g.shape, generator=g)
    arena.flat_g.copy_(local)
    arena.all_reduce()                         # ONE collective over the whole arena
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered)
    ok = torch.allclose(arena.flat_g, want, atol=1e-6)
    # the per-parameter grad views alias the arena
    off_ok = torch.equal(arena.grad_of(first).flatten(), arena.flat_g[: first.numel()])
    scale_ok = abs(arena.grad_scale - 1.0 / world) < 1e-12
    q.put((rank, bool(ok), bool(off_ok), bool(scale_ok)))
    dist.destroy_process_group()


def test_flat_arena_allreduce_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def _seed_worker(rank, world, port, q):
    """Ranks seeded DIFFERENTLY: the arena must broadcast rank 0's parameters (accelerate/DDP does, train_...:513)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from controllora_b200.arena import ParamArena
    import controllora_b200 as cb
    from controllora_b200.configs import NAMED

    torch.manual_seed(1234 + rank)
    cl = cb.ControlLoRA.from_config(NAMED["diffusiondb-canny-v2"])
    arena = ParamArena(list(cl.parameters()), torch.device("cpu"))
    gathered = [torch.zeros_like(arena.flat_p) for _ in range(world)]
    dist.all_gather(gathered, arena.flat_p)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    nonzero = float(arena.flat_p.abs().sum()) > 0
    # the parameters still alias the arena after the broadcast
    alias = next(cl.parameters()).data.data_ptr() == arena.flat_p.data_ptr()
    q.put((rank, bool(same), bool(nonzero), bool(alias)))
    dist.destroy_process_group()


def test_arena_broadcasts_initial_parameters_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_seed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] and r[2] and r[3] for r in res), res
