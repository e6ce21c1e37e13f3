"""N>1 host logic on CPU with 2 gloo ranks: flat parameter/gradient bucket, rank-0 broadcast, mean all-reduce,
and equivalence of the 2-rank step with a single-process step on the concatenated batch (BN-free model, since the
reference keeps BatchNorm statistics per rank: main.py:82-85 has no sync_batchnorm)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open3dsot_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = ddp.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    net = _model(seed=100 + rank)                 # ranks start different; broadcast must align them
    flat = ddp.FlatParams(net)
    ddp.broadcast_parameters(flat, net)
    opt = torch.optim.Adam([flat.flat], lr=1e-2, betas=(0.5, 0.999), eps=1e-6)
    flat.flat.grad = flat.grad
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(3):
        flat.zero_grad()
        ((net(xs) - ys) ** 2).mean().backward()
        ddp.allreduce_gradients(flat)
        opt.step()
    torch.save(flat.flat.clone(), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    p1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(p0, p1)                    # replicas stay bit-identical
    # single process on the full batch (mean loss over 8 == mean of the two rank means over 4)
    net = _model(seed=100)
    flat = ddp.FlatParams(net)
    opt = torch.optim.Adam([flat.flat], lr=1e-2, betas=(0.5, 0.999), eps=1e-6)
    flat.flat.grad = flat.grad
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    for _ in range(3):
        flat.zero_grad()
        ((net(x) - y) ** 2).mean().backward()
        opt.step()
    assert torch.allclose(p0, flat.flat, rtol=1e-5, atol=1e-6)


def test_flat_params_views_alias_module():
    net = _model(seed=1)
    flat = ddp.FlatParams(net)
    assert flat.numel == sum(p.numel() for p in net.parameters())
    net[0].weight.data.fill_(2.0)
    assert float(flat.flat[:96].min()) == 2.0
    (net(torch.ones(2, 6)).sum()).backward()
    assert flat.grad.abs().sum() > 0 and net[0].weight.grad.data_ptr() == flat.grad.data_ptr()
