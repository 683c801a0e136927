"""Multi-GPU data parallelism for the fused stage-0 step: rays are sharded across ranks (one process
per GPU, each rank draws its own batch), every rank holds a full replica of the tables / MLPs, and the
only exchange per step is one all-reduce of the flat gradient buffers (hash-table gradients + MLP
gradients) plus the 4-byte found_inf flag, NCCL over NVLink 5 / NVSwitch.  The reference has no working
multi-GPU path (its DDP scaffolding is unreachable)."""
import torch
import torch.distributed as dist


class GradSync:
    """Averages the loss-scaled gradients of a Stage0Trainer-like object (attributes gtable, g_mlp,
    opt_state) over the default process group; found_inf is OR-ed so every rank skips the same steps."""

    def __init__(self, trainer, group=None):
        self.t = trainer
        self.group = group
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def buffers(self):
        return [self.t.gtable, self.t.g_mlp]

    def __call__(self):
        if self.world == 1:
            return
        for buf in self.buffers():
            if self.backend == "nccl":
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
            else:       # gloo (CPU tests) has no AVG
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                buf.div_(self.world)
        dist.all_reduce(self.t.opt_state[3:4], op=dist.ReduceOp.MAX, group=self.group)

    def bytes_per_step(self):
        return sum(b.numel() * b.element_size() for b in self.buffers()) + 4
