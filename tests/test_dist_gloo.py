"""CPU, world_size 2, gloo: the host-side logic of the data-parallel step (nerf2mesh_b200.parallel):
gradient buffers are averaged, found_inf is OR-ed, ranks stay bit-identical."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeTrainer:
    def __init__(self, rank):
        g = torch.Generator().manual_seed(rank)
        self.gtable = torch.randn(1000, 4, generator=g)
        self.g_mlp = torch.randn(7648, generator=g)
        self.opt_state = torch.zeros(8)
        self.opt_state[3] = 1.0 if rank == 1 else 0.0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf2mesh_b200.parallel import GradSync
    t = FakeTrainer(rank)
    sync = GradSync(t)
    assert sync.bytes_per_step() == 1000 * 16 + 7648 * 4 + 4
    sync()
    q.put((rank, t.gtable.clone(), t.g_mlp.clone(), float(t.opt_state[3])))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    exp_t = (FakeTrainer(0).gtable + FakeTrainer(1).gtable) / 2
    exp_m = (FakeTrainer(0).g_mlp + FakeTrainer(1).g_mlp) / 2
    for _, gt, gm, inf in res:
        assert torch.allclose(gt, exp_t) and torch.allclose(gm, exp_m) and inf == 1.0
    assert torch.equal(res[0][1], res[1][1])
