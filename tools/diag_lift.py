"""Diagnostic: one SA layer, lifted / materialised x tc0 / tc3, against the float64 oracle (gradients: error and scale)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from open3dsot_b200 import runtime
from open3dsot_b200.datasets.synthetic import synthetic_siamese_batch
from open3dsot_b200.pointnet2.utils.pointnet2_modules import PointnetSAModule
from oracle import modules as om
from _params import det_state_dict

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import itertools
SHAPES = [(512, 128, 256, 32, 0.5)] if len(sys.argv) < 2 else [(64, 256, 32, 32, 0.7), (128, 128, 64, 32, 0.5), (256, 0, 128, 32, 0.3),
                                                                (128, 256, 64, 32, 0.7), (512, 0, 256, 32, 0.3)]
for B, shape in itertools.product((2, 8, 48) if len(sys.argv) < 2 else (16,), SHAPES):
    for kind in ("synthetic",) if len(sys.argv) > 1 else ("synthetic", "uniform"):
        N, C, npoint, S, r = shape
        g = torch.Generator().manual_seed(B)
        if kind == "synthetic":
            xyz = synthetic_siamese_batch(B, 512, 1024, seed=3)["search_points"][:, :N].contiguous()
        else:
            xyz = torch.rand(B, N, 3, generator=g) * 2
        feats = torch.randn(B, max(C, 1), N, generator=g)
        mlp = [C, 128, 128, 256] if C == 128 else ([C, 256, 256, 256] if C == 256 else [0, 64, 64, 128])
        sa = PointnetSAModule(mlp=list(mlp), radius=r, nsample=S, use_fps=False)
        base = det_state_dict(sa.state_dict(), seed=1)
        pn = [k for k, _ in sa.named_parameters()]

        def oracle(dt):
            sd = {"sa." + k: (v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in base.items()}
            for k in pn:
                sd["sa." + k].requires_grad_(True)
            f = feats.clone().to(dt).requires_grad_(True)
            _, y, _ = om.sa_module(sd, "sa", xyz.to(dt), f if C else None, npoint, r, S, False, True)
            w = torch.linspace(0.5, 1.5, y.shape[1], dtype=dt)[None, :, None]
            (y * w).square().sum().backward()
            return {k: sd["sa." + k].grad for k in pn}, (f.grad if C else torch.ones(1, dtype=dt))
        g64, f64 = oracle(torch.float64)
        g32, f32 = oracle(torch.float32)
        den = sum(float(g64[k].norm()) ** 2 for k in pn) ** 0.5

        def rep(tag, gr, fg):
            num = sum(float((gr[k].double().cpu() - g64[k]).norm()) ** 2 for k in pn) ** 0.5
            dot = sum(float((gr[k].double().cpu() * g64[k]).sum()) for k in pn)
            fe = float((fg.double().cpu() - f64).norm() / f64.norm())
            print(f"B={B:2d} N={N} C={C} np={npoint} {kind:9s} {tag:18s} param err {num / den:.2e} scale {dot / den ** 2:.6f}  feat-grad err {fe:.2e}")
        rep("oracle32", g32, f32)
        sa = sa.cuda().train()
        for tag, lift, tc in (("lift tc0", True, 0), ("lift tc3", True, 3), ("nolift tc0", False, 0), ("nolift tc3", False, 3)):
            sa.load_state_dict(base)
            sa.zero_grad(set_to_none=True)
            runtime.set_lift(lift); runtime.set_tc(tc)
            f = feats.clone().cuda().requires_grad_(True)
            _, y = sa(xyz.cuda(), f if C else None, npoint)
            w = torch.linspace(0.5, 1.5, y.shape[1], device="cuda")[None, :, None]
            (y * w).square().sum().backward()
            rep(tag, {k: p.grad.detach() for k, p in sa.named_parameters()}, f.grad if C else torch.ones(1))
