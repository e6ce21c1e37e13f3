"""Diagnostic: xcorr heads / RPN, lifted vs materialised x tc0 / tc3, against the float64 oracle (gradients: error and scale)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from open3dsot_b200 import runtime
from open3dsot_b200.models.head.xcorr import BoxAwareXCorr, P2B_XCorr
from open3dsot_b200.models.head.rpn import P2BVoteNetRPN
from oracle import modules as om
from _params import det_state_dict
from test_gpu_parity_full import Choices

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for kind, B, Mt, Ns in (("p2b", 8, 64, 128), ("bat", 16, 32, 64), ("bat", 48, 64, 128), ("rpn", 8, 64, 128), ("rpn", 16, 32, 64)):
    f = 256
    g = torch.Generator().manual_seed(B)
    tf, sf = torch.randn(B, f, Mt, generator=g), torch.randn(B, f, Ns, generator=g)
    txyz, sxyz = torch.rand(B, Mt, 3, generator=g) * 2, torch.rand(B, Ns, 3, generator=g) * 2
    tbc, sbc = torch.rand(B, Mt, 9, generator=g), torch.rand(B, Ns, 9, generator=g)
    if kind == "p2b":
        m = P2B_XCorr(f, 256, f)
    elif kind == "bat":
        m = BoxAwareXCorr(f, 256, f, k=4, bc_channel=9)
    else:
        m = P2BVoteNetRPN(f, vote_channel=256, num_proposal=64)
    base = det_state_dict(m.state_dict(), seed=9)
    pn = [k for k, _ in m.named_parameters()]

    def oracle(dt, force=None):
        c = lambda v: v.clone().to(dt)
        sd = {"m." + k: (c(v) if v.is_floating_point() else v.clone()) for k, v in base.items()}
        for k in pn:
            sd["m." + k].requires_grad_(True)
        a, b = c(tf).requires_grad_(True), c(sf).requires_grad_(True)
        om.TAPS = {}
        om.set_force(force)
        if kind == "p2b":
            outs = [om.p2b_xcorr(sd, "m", a, b, c(txyz), True)]
        elif kind == "bat":
            outs = [om.boxaware_xcorr(sd, "m", a, b, c(txyz), c(sxyz), c(tbc), c(sbc), 4, True)[0]]
        else:
            outs = list(om.rpn(sd, "m", c(sxyz), b, 64, True))
        taps = om.TAPS
        om.TAPS = None
        om.set_force(None)
        loss = sum((o * torch.linspace(0.5, 1.5, o.shape[1], dtype=dt).view(1, -1, *([1] * (o.dim() - 2)))).square().sum() for o in outs)
        loss.backward()
        return {k: sd["m." + k].grad for k in pn}, (a.grad, b.grad), taps
    g32, i32, taps = oracle(torch.float32)
    force = {"topk": [taps["m:topk"][0]] if kind == "bat" else [], "ball_query": [taps["m.vote_aggregation:bq_idx"][0]] if kind == "rpn" else []}
    g64, i64, _ = oracle(torch.float64, force)
    den = sum(float(g64[k].norm()) ** 2 for k in pn) ** 0.5

    def rep(tag, gr, ig):
        num = sum(float((gr[k].double().cpu() - g64[k]).norm()) ** 2 for k in pn) ** 0.5
        dot = sum(float((gr[k].double().cpu() * g64[k]).sum()) for k in pn)
        ie = [float((a.double().cpu() - b).norm() / b.norm()) if (a is not None and b is not None) else -1 for a, b in zip(ig, i64)]
        print(f"{kind} B={B:2d} {tag:12s} param err {num / den:.2e} scale {dot / den ** 2:.6f}  input-grad err {ie[0]:.2e} {ie[1]:.2e}")
    rep("oracle32", g32, i32)
    m = m.cuda().train()
    inject = {("boxaware_topk", 0): taps["m:topk"][0]} if kind == "bat" else ({("ball_query", 0): taps["m.vote_aggregation:bq_idx"][0]} if kind == "rpn" else {})
    for tag, lift, tc in (("lift tc0", True, 0), ("lift tc3", True, 3), ("nolift tc0", False, 0), ("nolift tc3", False, 3)):
        m.load_state_dict(base)
        m.zero_grad(set_to_none=True)
        runtime.set_lift(lift); runtime.set_tc(tc)
        runtime.CHOICE_HOOK = Choices(inject) if lift else None
        a, b = tf.clone().cuda().requires_grad_(True), sf.clone().cuda().requires_grad_(True)
        try:
            if kind == "p2b":
                outs = [m(a, b, txyz.cuda())]
            elif kind == "bat":
                outs = [m(a, b, txyz.cuda(), sxyz.cuda(), tbc.cuda(), sbc.cuda())]
            else:
                outs = list(m(sxyz.cuda(), b))
            loss = sum((o * torch.linspace(0.5, 1.5, o.shape[1], device="cuda").view(1, -1, *([1] * (o.dim() - 2)))).square().sum() for o in outs)
            loss.backward()
        finally:
            runtime.CHOICE_HOOK = None
        rep(tag, {k: p.grad.detach() for k, p in m.named_parameters()}, (a.grad, b.grad))
