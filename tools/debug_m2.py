import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from _params import det_state_dict
from open3dsot_b200 import runtime
from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.synthetic import synthetic_motion_batch
from open3dsot_b200.models import get_model

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg = load_config("cfgs/M2_track_kitti.yaml")
net = get_model(cfg.net_model)(cfg)
base = det_state_dict(net.state_dict(), seed=31)
net.load_state_dict(base)
net = net.cuda().train()
batch = synthetic_motion_batch(4, 256, seed=77, device="cuda")
x = torch.cat([batch["points"].transpose(1, 2), batch["candidate_bc"].transpose(1, 2)], dim=1).contiguous()


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


outs = {}
for mode in (False, True):
    runtime.set_fused(mode)
    net.load_state_dict(base)
    with torch.no_grad():
        seg = net.seg_pointnet(x)
        mp_in = torch.randn(4, 13, 512, generator=torch.Generator().manual_seed(1)).cuda()
        mini = net.mini_pointnet(mp_in)
        head = net._mlp(net.motion_mlp, mini)
    outs[mode] = (seg, mini, head)
for name, a, b in zip(("seg_pointnet", "mini_pointnet", "motion_mlp"), outs[True], outs[False]):
    print(name, "fused vs composed rel", rel(a, b), tuple(a.shape))
for lv in (0, 3):
    runtime.set_tc(lv); runtime.set_fused(True)
    net.load_state_dict(base)
    with torch.no_grad():
        seg = net.seg_pointnet(x)
    print("tc level", lv, "seg rel vs composed", rel(seg, outs[False][0]))
