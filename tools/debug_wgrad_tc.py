import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open3dsot_b200 import _lib, ops

L = _lib.lib()
torch.manual_seed(0)
P, Cout, Cin = 4096, 128, 128
g = torch.randn(P, Cout, device="cuda")
x = torch.randn(P, Cin, device="cuda")
want = g.double().t() @ x.double()
for name, dbg in (("o3d_pw_wgrad", 0), ("o3d_pw_wgrad_tc", 0), ("o3d_pw_wgrad_tc", 16), ("o3d_pw_wgrad_tc", 32)):
    L.o3d_debug_set(dbg, 0)
    dw = torch.zeros(Cout, Cin, device="cuda")
    rc = getattr(L, name)(g.data_ptr(), Cout, None, 0, None, None, None, None, None, 0, 0, x.data_ptr(), Cin, None, None, 0,
                          P, Cout, Cin, dw.data_ptr(), Cin, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(name, "dbg", dbg, "rc", rc, "rel err", float((dw.double() - want).norm() / want.norm()), "dw[0,:4]", dw[0, :4].tolist(), "want", want[0, :4].tolist())
    if name.endswith("tc"):
        # structure of the error: correlate rows / cols
        c = (dw.double().T @ want) / (want.norm(dim=0)[None, :] * dw.double().norm(dim=0)[:, None] + 1e-9)
        print(" col match", c.argmax(dim=1)[:48].tolist())
        r = (dw.double() @ want.T) / (want.norm(dim=1)[None, :] * dw.double().norm(dim=1)[:, None] + 1e-9)
        print(" row match", r.argmax(dim=1)[:48].tolist())
        print(" nonzero frac", float((dw != 0).float().mean()))
