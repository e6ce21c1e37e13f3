"""Training-step engine on the GPU: the one-launch Adam against torch.optim.Adam, the CUDA-graph replay against the eager
step, and the fused path against the reference's op composition at BASELINE.json's full config-2 size (48 pairs,
512 / 1024 points) through size-independent properties (finite loss / gradients, agreement of the two execution modes)."""
import os

import pytest
import torch

from open3dsot_b200 import ddp, runtime
from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.synthetic import synthetic_siamese_batch
from open3dsot_b200.engine import FlatAdam, TrainStep
from open3dsot_b200.models import get_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adam_kernel_matches_torch_adam():
    torch.manual_seed(0)
    net = torch.nn.Linear(37, 53).cuda()                     # 2014 parameters: exercises the non-multiple-of-4 tail
    ref = torch.nn.Linear(37, 53).cuda()
    ref.load_state_dict(net.state_dict())
    flat = ddp.FlatParams(net)
    opt = FlatAdam(flat, lr=1e-3, betas=(0.5, 0.999), eps=1e-6)
    topt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.5, 0.999), eps=1e-6)
    for i in range(6):
        x = torch.randn(16, 37, device="cuda")
        flat.zero_grad()
        net(x).square().sum().backward()
        opt.step()
        topt.zero_grad()
        ref(x).square().sum().backward()
        topt.step()
    for a, b in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7)
    assert float(opt.state[0]) == 6.0


def _bat(seed=0):
    cfg = load_config(os.path.join(ROOT, "cfgs", "BAT_Car.yaml"))
    torch.manual_seed(seed)
    return cfg, get_model(cfg.net_model)(cfg).cuda().train()


def _snapshot(eng):
    return (eng.flat.flat.clone(), eng.opt.exp_avg.clone(), eng.opt.exp_avg_sq.clone(), eng.opt.state.clone(),
            [b.clone() for b in eng.model.buffers()])


def _restore(eng, snap):
    with torch.no_grad():
        eng.flat.flat.copy_(snap[0]); eng.opt.exp_avg.copy_(snap[1]); eng.opt.exp_avg_sq.copy_(snap[2]); eng.opt.state.copy_(snap[3])
        for b, old in zip(eng.model.buffers(), snap[4]):
            b.copy_(old)


def test_graph_replay_equals_eager_step_from_the_same_state():
    """Five optimisation steps at B = 4; at EVERY step the captured graph and the eager path are run from the same state
    (parameters, Adam moments, BatchNorm buffers restored in between), so the comparison is one step deep and no trajectory
    chaos enters: loss, the full gradient, and the Adam-updated parameters / moments must agree.  The two executions differ only
    by the summation order of the fp32 / fp64 atomics (BatchNorm sums, the 128x128 weight-gradient tile, the scatter of the lifted
    layer), measured at <= 2e-6 of the gradient norm; a wrong Adam step, a stale graph input or a missed BN update would show as
    an O(1) difference."""
    cfg, net = _bat()
    batches = [synthetic_siamese_batch(4, 256, 512, seed=100 + i, device="cuda") for i in range(6)]
    eng = TrainStep(net, lr=cfg.lr, use_graph=True, warmup=1)
    eng.step(batches[0])                                      # eager warm-up step
    worst = 0.0
    for b in batches[1:]:
        snap = _snapshot(eng)
        lg = eng.step(b).clone()                              # step 2 captures the graph, later steps replay it
        assert eng.graph is not None
        gg, pg, mg = eng.flat.grad.clone(), eng.flat.flat.clone(), eng.opt.exp_avg.clone()
        step_g = float(eng.opt.state[0])
        bn_g = [x.clone() for x in net.buffers()]
        _restore(eng, snap)
        le = eng._eager(b).clone()
        ge, pe, me = eng.flat.grad, eng.flat.flat, eng.opt.exp_avg
        assert float(eng.opt.state[0]) == step_g             # both advanced the step counter once
        assert abs(float(lg) - float(le)) <= 1e-6 * abs(float(le)), (float(lg), float(le))
        rg = float((gg - ge).norm() / ge.norm())
        worst = max(worst, rg)
        assert rg < 1e-4, rg
        assert float((mg - me).norm() / me.norm()) < 1e-4
        # Adam's update is lr * m / (sqrt(v) + eps): bounded by lr per element, and equal wherever the gradients are
        assert float((pg - pe).abs().max()) <= 2.0 * cfg.lr
        assert float((pg - pe).norm() / (pe - snap[0]).norm().clamp_min(1e-12)) < 2e-2
        for x, y in zip(bn_g, net.buffers()):                  # running statistics / num_batches_tracked
            assert torch.allclose(x.float(), y.float(), rtol=1e-5, atol=1e-7)
    print(f"\n[graph vs eager, same state] worst gradient difference over 5 steps: {worst:.1e}")


def test_full_size_step_fused_vs_composed():
    """BASELINE.json configs[1] at full size (48 pairs, 512/1024 points)."""
    cfg, net = _bat(seed=1)
    batch = synthetic_siamese_batch(48, cfg.template_size, cfg.search_size, seed=20260924, device="cuda")
    base = {k: v.clone() for k, v in net.state_dict().items()}
    res = {}
    for fused_mode in (True, False):
        runtime.set_fused(fused_mode)
        try:
            net.load_state_dict(base)
            net.zero_grad()
            loss = net.training_step({k: v.clone() for k, v in batch.items()}, 0)
            loss.backward()
            g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
            res[fused_mode] = (float(loss.detach()), g)
            assert torch.isfinite(loss) and torch.isfinite(g).all()
        finally:
            runtime.set_fused(True)
    (lf, gf), (lc, gc) = res[True], res[False]
    assert abs(lf - lc) <= 5e-3 * abs(lc), (lf, lc)
    cos = torch.dot(gf, gc) / (gf.norm() * gc.norm())
    assert float(cos) > 0.99                                # neighbour / top-k flips on computed coordinates aside


def test_inplace_gradient_accumulation_matches_autograd_accumulation():
    """The engine lets every stack add its parameter gradients straight into the flat bucket (runtime.grad_inplace_scope);
    the result must equal what autograd's own accumulation produces from the same state and batch (shared backbone weights
    receive two contributions per step either way)."""
    cfg, net = _bat(seed=3)
    batch = synthetic_siamese_batch(4, 256, 512, seed=7, device="cuda")
    eng = TrainStep(net, lr=cfg.lr, use_graph=False)
    eng._fwd_bwd(batch)                                        # in-place path
    g_in = eng.flat.grad.clone()
    bn = [b.clone() for b in net.buffers()]
    eng.flat.zero_grad()
    loss = net.training_step({k: v.clone() for k, v in batch.items()}, 0)
    loss.backward()                                            # plain autograd accumulation into the same .grad views
    g_ag = eng.flat.grad
    assert float(g_ag.norm()) > 0
    assert float((g_in - g_ag).norm() / g_ag.norm()) < 1e-5
    assert len(bn) > 0
