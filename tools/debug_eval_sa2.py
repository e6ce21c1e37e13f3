import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_fused import reference_stack, randomise
from open3dsot_b200 import fused
from open3dsot_b200.pointnet2.utils import pytorch_utils as pt

for training in (True, False):
    import zlib
    torch.manual_seed(zlib.crc32(b"sa2") % 1000)
    chans, P, S = [132, 128, 128, 256], 1184, 32
    mod = pt.SharedMLP(list(chans), bn=True)
    randomise(mod, 7)
    mod = mod.cuda(); mod.train(training)
    specs = fused.parse_stack(mod)
    x = torch.randn(P, chans[0], device="cuda"); x[:, -1] = 0
    x1 = x.clone().requires_grad_(True); x2 = x.clone().requires_grad_(True)
    out = fused.mlp_stack(x1, specs, S, training)
    want = reference_stack(x2, specs, S, training)
    print("training", training, "fwd rel", float((out.double() - want).norm() / want.norm()))
    go = torch.randn_like(want)
    params = list(mod.parameters())
    names = [n for n, _ in mod.named_parameters()]
    g_ref = torch.autograd.grad(want, [x2] + params, go)
    g_out = torch.autograd.grad(out, [x1] + params, go.float())
    for n, a, b in zip(["x"] + names, g_out, g_ref):
        print(f"  {n:28s} rel {float((a.double()-b).norm()/b.norm()):.3e}  norm {float(b.norm()):.3e}")
    e = (g_out[0].double() - g_ref[0]).abs()
    print("  dx err by row block of 32:", [round(float(v), 3) for v in e.view(37, 32, 132).amax(dim=(1, 2))])
    bad = (e.view(37, 32, 132).amax(dim=(1, 2)) > 1e-3).nonzero().flatten().tolist()
    print("  bad groups", bad)
    print("  dx err by col block of 4 :", [round(float(v), 3) for v in e.view(1184, 33, 4).amax(dim=(0, 2))])
