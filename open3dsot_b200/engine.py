"""Training-step engine: flat parameter bucket + one-launch Adam + optional whole-step CUDA graph.

The reference trains through Lightning (`main.py:82-86`): per step DDP-wrapped `training_step`, backward, bucketed
all-reduce, `Adam(betas=(0.5, 0.999), eps=1e-6)` (`models/base_model.py:28-36`).  A BAT step is a few milliseconds of GPU
work spread over ~10^3 kernel launches, so the step is launch-bound unless it is captured: `TrainStep` records
zero-grad -> training_step -> backward -> gradient all-reduce -> Adam into ONE CUDA graph over static batch buffers
and replays it; the loss stays on the device (the reference's twelve `.item()` syncs per step, `bat.py:146-163`, are
not reproduced).
"""
import torch

from . import _lib, ddp, ops


class FlatAdam:
    """Adam over `FlatParams` with the step counter and learning rate in device memory (graph-replayable)."""

    def __init__(self, flat, lr=1e-3, betas=(0.5, 0.999), eps=1e-6, weight_decay=0.0):
        self.flat = flat
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.state = torch.tensor([0.0, lr], dtype=torch.float32, device=flat.flat.device)

    def set_lr(self, lr):
        self.state[1] = lr

    def step(self):
        ops._call("o3d_adam_step", self.flat.flat.data_ptr(), self.flat.grad.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), self.flat.numel, self.state.data_ptr(), self.betas[0], self.betas[1],
                  self.eps, self.wd, ops._stream())


class TrainStep:
    """`step(batch) -> loss` for a model exposing `training_step(batch, idx)`; `batch` is a dict of device tensors."""

    def __init__(self, model, lr=1e-3, weight_decay=0.0, use_graph=True, warmup=3):
        self.model = model
        self.flat = ddp.FlatParams(model)
        ddp.broadcast_parameters(self.flat, model)
        self.opt = FlatAdam(self.flat, lr=lr, weight_decay=weight_decay)
        self.use_graph = use_graph
        self.warmup = warmup
        self.graph = None
        self.static_batch = None
        self.static_loss = None
        self.calls = 0

    def _fwd_bwd(self, batch):
        self.flat.zero_grad()
        loss = self.model.training_step(dict(batch), 0)
        loss.backward()
        return loss.detach()

    def _finish(self):
        ddp.allreduce_gradients(self.flat)      # one NCCL all-reduce of the flat bucket (no-op on a single rank)
        self.opt.step()

    def _eager(self, batch):
        loss = self._fwd_bwd(batch)
        self._finish()
        return loss

    def _capture(self, batch):
        self.static_batch = {k: v.clone() for k, v in batch.items()}
        # settle allocator / lazy initialisation on a side stream, WITHOUT advancing training: parameters, optimizer
        # state and BatchNorm buffers are snapshotted and restored around the two throw-away steps
        buffers = list(self.model.buffers())
        snap = (self.flat.flat.clone(), self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.state.clone(),
                [b.clone() for b in buffers])
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._eager(self.static_batch)
        torch.cuda.current_stream().wait_stream(s)
        with torch.no_grad():
            self.flat.flat.copy_(snap[0]); self.opt.exp_avg.copy_(snap[1]); self.opt.exp_avg_sq.copy_(snap[2])
            self.opt.state.copy_(snap[3])
            for b, old in zip(buffers, snap[4]):
                b.copy_(old)
        # single rank: the whole step (incl. Adam) is one graph; multi-rank: forward+backward are captured and the
        # gradient all-reduce + Adam are enqueued right behind the replay (NCCL stays outside the capture)
        self.graph = torch.cuda.CUDAGraph()
        self.graph_has_update = not ddp.is_distributed()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._fwd_bwd(self.static_batch)
            if self.graph_has_update:
                self._finish()

    def step(self, batch):
        self.calls += 1
        if not self.use_graph or self.calls <= self.warmup:
            return self._eager(batch)
        if self.graph is None:
            self._capture(batch)
        for k, v in batch.items():
            self.static_batch[k].copy_(v, non_blocking=True)
        self.graph.replay()
        if not self.graph_has_update:
            self._finish()
        return self.static_loss
