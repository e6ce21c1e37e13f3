"""Fused execution of the reference's MLP stacks on the sm_100a point-wise kernels (csrc/pwmlp.cu).

A *stack* is what the reference builds with `pt_utils.SharedMLP` / `pt_utils.Seq` / a bare `nn.Conv1d`:
a chain of 1x1 convolutions, each optionally followed by BatchNorm and ReLU, optionally ending in a max over
groups of S consecutive positions (SA max-pool over nsample — pointnet2_modules.py:70-73; BoxAware max over k —
xcorr.py:100; P2B max over template points — xcorr.py:49).  `mlp_stack` runs such a chain as

    per layer:  o3d_pw_fwd (GEMM + BN statistics [+ pooling] in the epilogue)  ->  o3d_bn_fwd_finalize
    backward :  o3d_pool_bwd_prep | o3d_dense_bwd_prep  ->  per layer  o3d_bn_bwd_finalize, o3d_pw_wgrad, o3d_pw_dgrad

on channels-last activations; normalised activations, the BN backward and the ReLU masks live only inside the
operand loaders / epilogues of the GEMMs.  Parameters stay in the reference's modules (same names, same
state-dict), torch is used for buffers and the autograd graph only.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .pointnet2.utils import pointnet2_utils

_call = ops._call
_stream = ops._stream


def _r4(n):
    return (n + 3) & ~3


def _ptr(t):
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------- layer parsing
class _LayerSpec:
    __slots__ = ("weight", "bias", "bn", "relu")

    def __init__(self, weight, bias, bn, relu):
        self.weight, self.bias, self.bn, self.relu = weight, bias, bn, relu


def _spec_from_unit(unit):
    """unit: pt_utils._ConvBase (children conv / bn / activation), nn.ConvNd, or nn.Linear."""
    if isinstance(unit, (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.Linear)):
        return _LayerSpec(unit.weight, unit.bias, None, False)
    conv = getattr(unit, "conv", None)
    if conv is None:
        raise TypeError(f"cannot fuse {type(unit).__name__}")
    bn = getattr(unit, "bn", None)
    if bn is not None and not isinstance(bn, nn.modules.batchnorm._BatchNorm):
        bn = bn[0]  # _BNBase wrapper -> the torch BatchNorm inside
    return _LayerSpec(conv.weight, conv.bias, bn, hasattr(unit, "activation"))


def parse_stack(module):
    """SharedMLP / Seq (children are _ConvBase units), bare conv, or nn.Sequential(conv, BN, ReLU, ...)."""
    if isinstance(module, (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.Linear)):
        return [_spec_from_unit(module)]
    kids = list(module.children())
    if kids and all(hasattr(k, "conv") for k in kids):
        return [_spec_from_unit(k) for k in kids]
    specs = []
    for k in kids:  # flat torch Sequential: conv / linear, then optional BatchNorm, then optional ReLU
        if isinstance(k, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            specs.append(_LayerSpec(k.weight, k.bias, None, False))
        elif isinstance(k, nn.modules.batchnorm._BatchNorm):
            specs[-1].bn = k
        elif isinstance(k, nn.ReLU):
            specs[-1].relu = True
        else:
            raise TypeError(f"cannot fuse {type(k).__name__} inside {type(module).__name__}")
    return specs


# ---------------------------------------------------------------------------------------------- the autograd op
class _Meta:
    """Static description of one stack invocation (python objects only)."""

    def __init__(self, specs, S, training):
        self.n = len(specs)
        self.S = int(S)
        self.training = bool(training)
        self.has_bias = [s.bias is not None for s in specs]
        self.has_bn = [s.bn is not None for s in specs]
        self.relu = [bool(s.relu) for s in specs]
        self.bns = [s.bn for s in specs]
        self.cout = [s.weight.shape[0] for s in specs]


class _MLPStackFn(torch.autograd.Function):
    """x (P, K0p) channels-last fp32; per layer W (Cout, Cin) 2-D, bias|None, gamma|None, beta|None."""

    @staticmethod
    def forward(ctx, meta, x, *params):
        dev = x.device
        P, K0 = x.shape
        n = meta.n
        Ws, bs, gs, betas = params[0::4], params[1::4], params[2::4], params[3::4]
        S = meta.S
        need_grad = any(ctx.needs_input_grad)
        ctx.wshapes = meta.wshapes
        st = _stream()
        nws = [_r4(c) for c in meta.cout]
        kin = [K0] + nws[:-1]
        # one zeroed fp64 workspace for all batch statistics
        stat = torch.zeros(2 * sum(nws), dtype=torch.float64, device=dev) if meta.training and any(meta.has_bn) else None
        vec = torch.zeros(4 * sum(nws), dtype=torch.float32, device=dev)  # scale | shift | mean | invstd per layer
        Wp, ys, scales, shifts, means, invstds = [], [], [], [], [], []
        cur, cur_ld, in_scale, in_shift, in_relu = x, K0, None, None, 0
        so = vo = 0
        pooled = None
        for l in range(n):
            Nw, K, Cout = nws[l], kin[l], meta.cout[l]
            W = Ws[l]
            wp = torch.zeros(Nw, K, dtype=torch.float32, device=dev)
            wp[:Cout, :W.shape[1]] = W
            wt = wp.t().contiguous()
            bias = None
            if bs[l] is not None:
                bias = torch.zeros(Nw, dtype=torch.float32, device=dev)
                bias[:Cout] = bs[l]
            last = l == n - 1
            pool_here = last and S > 0
            keep_y = (not last) or need_grad or (not pool_here)   # inference skips the last raw tensor when pooling
            y = torch.empty(P, Nw, dtype=torch.float32, device=dev) if keep_y else None
            use_stats = meta.training and meta.has_bn[l]
            sm = stat[so:so + Nw] if use_stats else None
            sq = stat[so + Nw:so + 2 * Nw] if use_stats else None
            if pool_here:
                G = P // S
                ymax = torch.empty(G, Nw, dtype=torch.float32, device=dev)
                ymin = torch.empty(G, Nw, dtype=torch.float32, device=dev)
                arg = torch.empty(G, Nw, dtype=torch.int32, device=dev)
            else:
                G, ymax, ymin, arg = 0, None, None, None
            _call("o3d_pw_fwd", cur.data_ptr(), cur_ld, _ptr(in_scale), _ptr(in_shift), int(in_relu), wt.data_ptr(), Nw,
                  _ptr(bias), P, K, Cout, _ptr(y), Nw, _ptr(sm), _ptr(sq), S if pool_here else 0, _ptr(ymax), _ptr(ymin),
                  _ptr(arg), Nw, st)
            sc = sh = mu = istd = None
            if meta.has_bn[l]:
                bn = meta.bns[l]
                sc, sh = vec[vo:vo + Nw], vec[vo + Nw:vo + 2 * Nw]
                mu, istd = vec[vo + 2 * Nw:vo + 3 * Nw], vec[vo + 3 * Nw:vo + 4 * Nw]
                mom = bn.momentum if bn.momentum is not None else 0.1
                track = bn.track_running_stats and bn.running_mean is not None
                _call("o3d_bn_fwd_finalize", _ptr(sm), _ptr(sq), float(P), _ptr(gs[l]), _ptr(betas[l]),
                      _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None,
                      _ptr(bn.num_batches_tracked) if (track and meta.training) else None, float(mom), float(bn.eps),
                      int(meta.training), Cout, sc.data_ptr(), sh.data_ptr(), mu.data_ptr(), istd.data_ptr(), st)
            Wp.append(wp); ys.append(y); scales.append(sc); shifts.append(sh); means.append(mu); invstds.append(istd)
            cur, cur_ld, in_scale, in_shift, in_relu = y, Nw, sc, sh, int(meta.relu[l])
            so += 2 * Nw
            vo += 4 * Nw
        # ---- output of the last layer
        Nw, Cout = nws[-1], meta.cout[-1]
        sel = ysel = None
        if S > 0:
            out = torch.empty(G, Nw, dtype=torch.float32, device=dev)
            sel = torch.empty(G, Nw, dtype=torch.int32, device=dev) if need_grad else None
            ysel = torch.empty(G, Nw, dtype=torch.float32, device=dev) if need_grad else None
            _call("o3d_pool_finalize", ymax.data_ptr(), ymin.data_ptr(), arg.data_ptr(), _ptr(scales[-1]),
                  _ptr(shifts[-1]), int(meta.relu[-1]), G, Nw, Nw, out.data_ptr(), Nw, _ptr(sel), _ptr(ysel), st)
        elif meta.has_bn[-1] or meta.relu[-1]:
            out = torch.empty(P, Nw, dtype=torch.float32, device=dev)
            _call("o3d_act_apply", ys[-1].data_ptr(), Nw, _ptr(scales[-1]), _ptr(shifts[-1]), int(meta.relu[-1]), P, Nw,
                  out.data_ptr(), Nw, st)
        else:
            out = ys[-1]
        ctx.meta = meta
        ctx.dims = (P, K0, nws, kin)
        ctx.saved = (x, Wp, ys, scales, shifts, means, invstds, sel, ysel, out, gs, vec)
        return out if Nw == Cout else out[:, :Cout]

    @staticmethod
    def backward(ctx, dout):
        meta = ctx.meta
        P, K0, nws, kin = ctx.dims
        x, Wp, ys, scales, shifts, means, invstds, sel, ysel, out, gs, _vec = ctx.saved
        n, S = meta.n, meta.S
        dev = x.device
        st = _stream()
        Nw = nws[-1]
        rows = P // S if S > 0 else P
        # padded, contiguous upstream gradient
        if dout.shape[1] != Nw or not dout.is_contiguous():
            d = torch.zeros(rows, Nw, dtype=torch.float32, device=dev)
            d[:, :dout.shape[1]] = dout
            dout = d
        stat = torch.zeros(2 * sum(nws), dtype=torch.float64, device=dev)     # s1 | s2y per layer
        coef = torch.zeros(5 * sum(nws), dtype=torch.float32, device=dev)     # a | b | cc | dgamma | dbeta per layer
        offs, o = [], 0
        for w in nws:
            offs.append(o)
            o += w
        s1 = [stat[2 * offs[l]:2 * offs[l] + nws[l]] for l in range(n)]
        s2 = [stat[2 * offs[l] + nws[l]:2 * offs[l] + 2 * nws[l]] for l in range(n)]

        def cf(l, j):
            return coef[5 * offs[l] + j * nws[l]:5 * offs[l] + (j + 1) * nws[l]]

        # ---- gradient entering the last layer's BN/ReLU (or the plain output)
        L = n - 1
        g = dpool = None
        if S > 0:
            dpool = torch.empty(rows, Nw, dtype=torch.float32, device=dev)
            _call("o3d_pool_bwd_prep", dout.data_ptr(), Nw, out.data_ptr(), Nw, ysel.data_ptr(), int(meta.relu[L]), rows,
                  Nw, Nw, dpool.data_ptr(), s1[L].data_ptr(), s2[L].data_ptr(), st)
        elif meta.has_bn[L] or meta.relu[L]:
            g = torch.empty(P, Nw, dtype=torch.float32, device=dev)
            _call("o3d_dense_bwd_prep", dout.data_ptr(), Nw, out.data_ptr(), Nw, ys[L].data_ptr(), Nw, int(meta.relu[L]), P,
                  Nw, g.data_ptr(), Nw, s1[L].data_ptr(), s2[L].data_ptr(), st)
        else:
            g = dout
            if meta.has_bias[L]:
                _call("o3d_dense_bwd_prep", dout.data_ptr(), Nw, None, 0, None, 0, 0, P, Nw, None, 0, s1[L].data_ptr(),
                      None, st)
        grads = [None] * (4 * n)
        dx = None
        need_dx = ctx.needs_input_grad[1]
        for l in range(L, -1, -1):
            Nl, K = nws[l], kin[l]
            a = b = cc = None
            if meta.has_bn[l]:
                a, b, cc = cf(l, 0), cf(l, 1), cf(l, 2)
                _call("o3d_bn_bwd_finalize", s1[l].data_ptr(), s2[l].data_ptr(), float(P), _ptr(gs[l]), means[l].data_ptr(),
                      invstds[l].data_ptr(), int(meta.training), meta.cout[l], a.data_ptr(), b.data_ptr(), cc.data_ptr(),
                      cf(l, 3).data_ptr(), cf(l, 4).data_ptr(), st)
                grads[4 * l + 2] = cf(l, 3)[:meta.cout[l]]
                grads[4 * l + 3] = cf(l, 4)[:meta.cout[l]]
                if meta.has_bias[l]:
                    grads[4 * l + 1] = torch.zeros(meta.cout[l], dtype=torch.float32, device=dev)  # BN removes the mean
            elif meta.has_bias[l]:
                grads[4 * l + 1] = s1[l][:meta.cout[l]].float()
            pooled_mode = (l == L and S > 0)
            dy_args = (_ptr(None if pooled_mode else g), Nl, _ptr(ys[l]) if a is not None else None, Nl, _ptr(a), _ptr(b),
                       _ptr(cc), _ptr(dpool) if pooled_mode else None, _ptr(sel) if pooled_mode else None,
                       S if pooled_mode else 0, Nl)
            # wgrad: dW = dY^T * A(x_l)
            xin = x if l == 0 else ys[l - 1]
            psc = None if l == 0 else scales[l - 1]
            psh = None if l == 0 else shifts[l - 1]
            prelu = 0 if l == 0 else int(meta.relu[l - 1])
            dw = torch.zeros(Nl, K, dtype=torch.float32, device=dev)
            _call("o3d_pw_wgrad", *dy_args, xin.data_ptr(), K, _ptr(psc), _ptr(psh), prelu, P, Nl, K, dw.data_ptr(), K, st)
            grads[4 * l] = dw[:meta.cout[l], :ctx.wshapes[l]]
            # dgrad: gradient w.r.t. this layer's input, masked by the previous layer's ReLU
            if l > 0 or need_dx:
                gout = torch.empty(P, K, dtype=torch.float32, device=dev)
                if l > 0:
                    want_stats = meta.has_bn[l - 1] or meta.has_bias[l - 1]
                    mask = meta.has_bn[l - 1] or meta.relu[l - 1]
                    _call("o3d_pw_dgrad", *dy_args, Wp[l].data_ptr(), K, P, Nl, K, gout.data_ptr(), K,
                          ys[l - 1].data_ptr() if mask else None, K, _ptr(psc), _ptr(psh), prelu,
                          s1[l - 1].data_ptr() if want_stats else None, s2[l - 1].data_ptr() if want_stats else None, st)
                    g = gout
                else:
                    _call("o3d_pw_dgrad", *dy_args, Wp[l].data_ptr(), K, P, Nl, K, gout.data_ptr(), K, None, 0, None, None,
                          0, None, None, st)
                    dx = gout
        return (None, dx, *grads)


def mlp_stack(x2d, specs, S=0, training=True, weights=None):
    """Run a stack on a channels-last matrix.

    x2d      (P, K0) fp32, contiguous, K0 % 4 == 0 (zero-padded input channels)
    specs    list of _LayerSpec (parameters of the reference modules)
    S        pooling group size over consecutive positions (0 = dense output)
    weights  optional list of 2-D (Cout, Cin<=K) tensors replacing spec.weight (re-ordered / padded input columns)
    returns  (P or P//S, Cout_last)
    """
    if not x2d.is_cuda:
        raise RuntimeError("open3dsot_b200.fused: CUDA tensors required (there is no CPU path)")
    assert x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype == torch.float32 and x2d.shape[1] % 4 == 0
    meta = _Meta(specs, S, training)
    params = []
    wshapes = []
    for i, s in enumerate(specs):
        W = weights[i] if weights is not None and weights[i] is not None else s.weight.reshape(s.weight.shape[0], -1)
        wshapes.append(W.shape[1])
        params += [W, s.bias, s.bn.weight if s.bn is not None else None, s.bn.bias if s.bn is not None else None]
    meta.wshapes = wshapes
    return _MLPStackFn.apply(meta, x2d, *params)


# ---------------------------------------------------------------------------------------------- layout helpers
def to_channels_last(feat, pad_to4=True):
    """(B, C, N) [any strides] -> (B, N, Cp) contiguous with Cp = round_up(C, 4) (zero padded)."""
    cl = feat.transpose(1, 2)
    C = cl.shape[2]
    Cp = _r4(C) if pad_to4 else C
    if Cp != C:
        cl = F.pad(cl, (0, Cp - C))
    return cl.contiguous()


def from_channels_last(cl):
    """(B, N, C) contiguous -> (B, C, N) view (no copy); consumers that want channels-last undo it for free."""
    return cl.transpose(1, 2)


# ---------------------------------------------------------------------------------------------- modules
def seq_forward(module, x):
    """Seq / SharedMLP-on-1D / bare Conv1d applied to x (B, C, L) -> (B, Cout, L) (channels-last view)."""
    B, C, L = x.shape
    cl = to_channels_last(x)
    specs = parse_stack(module)
    out = mlp_stack(cl.view(B * L, cl.shape[2]), specs, 0, module.training)
    return from_channels_last(out.reshape(B, L, out.shape[1]))


def sa_forward(sa, xyz, features, sample_idxs):
    """Fused body of _PointnetSAModuleBase.forward (pointnet2_modules.py:58-76) for every (grouper, mlp) scale."""
    B, N, _ = xyz.shape
    npoint = sample_idxs.shape[1]
    if sa.use_fps:
        new_xyz = torch.gather(xyz, 1, sample_idxs.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    else:
        new_xyz = xyz[:, :npoint].contiguous()
    feat_cl, C = None, 0
    if features is not None:
        C = features.shape[1]
        feat_cl = to_channels_last(features)
    Cp = 0 if feat_cl is None else feat_cl.shape[2]
    outs = []
    for grouper, mlp in zip(sa.groupers, sa.mlps):
        S = grouper.nsample
        if 128 % S != 0:
            raise RuntimeError(f"fused SA layer: nsample={S} must divide 128")
        grouped, _idx = pointnet2_utils.query_and_group_cl(xyz, new_xyz, feat_cl, grouper.radius, S,
                                                           grouper.normalize_xyz)
        specs = parse_stack(mlp)
        W0 = specs[0].weight.reshape(specs[0].weight.shape[0], -1)
        # reference channel order is [xyz(3), features(C)]; kernel rows are [features(Cp) | dx dy dz 0]
        if grouper.use_xyz:
            wx, wf = W0[:, :3], W0[:, 3:]
        else:
            wx, wf = W0.new_zeros(W0.shape[0], 3), W0
        parts = [wf]
        if Cp > C:
            parts.append(W0.new_zeros(W0.shape[0], Cp - C))
        parts.append(wx)
        packed = torch.cat(parts, dim=1) if len(parts) > 1 else wx
        pooled = mlp_stack(grouped.view(B * npoint * S, Cp + 4), specs, S, sa.training, weights=[packed] + [None] * (len(specs) - 1))
        outs.append(from_channels_last(pooled.reshape(B, npoint, pooled.shape[1])))
    return new_xyz, outs


class _ThreeNNInterpCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unknown, known, known_feat_cl):
        out, idx, w = ops.three_nn_interpolate(unknown, known, known_feat_cl)
        ctx.save_for_backward(idx, w)
        ctx.m = known.shape[1]
        return out

    @staticmethod
    def backward(ctx, g):
        idx, w = ctx.saved_tensors
        return None, None, ops.three_nn_interpolate_grad(g.contiguous(), idx, w, ctx.m)


def fp_forward(fp, unknown, known, unknow_feats, known_feats):
    """Fused PointnetFPModule.forward (pointnet2_modules.py:187-212)."""
    B, n, _ = unknown.shape
    kf = to_channels_last(known_feats)
    C2 = known_feats.shape[1]
    interp = _ThreeNNInterpCL.apply(unknown.contiguous(), known.contiguous(), kf)[:, :, :C2]
    if unknow_feats is not None:
        cl = torch.cat([interp, unknow_feats.transpose(1, 2)], dim=2)
    else:
        cl = interp
    C = cl.shape[2]
    if C % 4:
        cl = F.pad(cl, (0, _r4(C) - C))
    cl = cl.contiguous()
    out = mlp_stack(cl.view(B * n, cl.shape[2]), parse_stack(fp.mlp), 0, fp.training)
    return from_channels_last(out.reshape(B, n, out.shape[1]))


class _GroupRowsCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_cl, idx):
        B, N, C = feat_cl.shape
        L = idx.shape[1]
        out = torch.empty(B, L, C, dtype=torch.float32, device=feat_cl.device)
        _call("o3d_group_rows", feat_cl.data_ptr(), idx.data_ptr(), B, N, L, C, out.data_ptr(), _stream())
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        B, L, C = g.shape
        gf = torch.zeros(B, ctx.N, C, dtype=torch.float32, device=g.device)
        _call("o3d_group_rows_grad", g.data_ptr(), idx.data_ptr(), B, ctx.N, L, C, gf.data_ptr(), _stream())
        return gf, None


def boxaware_xcorr_forward(xc, template_feature, search_feature, template_xyz, template_bc, search_bc):
    """Fused BoxAwareXCorr.forward (models/head/xcorr.py:81-103): box-cloud top-k, row gather, MLP + max over k."""
    B, f, M = template_feature.shape
    N = search_bc.shape[1]
    k = xc.k
    if 128 % k != 0:
        raise RuntimeError(f"fused BoxAwareXCorr: k={k} must divide 128")
    dist = torch.cdist(template_bc, search_bc)                                        # same formulation as the reference
    topk = torch.argsort(dist, dim=1, stable=True)[:, :k, :].transpose(1, 2).contiguous().int()   # (B,N,k)
    # channel order [xyz(3), bc(9), feat(f)] == the reference's cat order (xcorr.py:82-84)
    tmpl = torch.cat([template_xyz, template_bc, template_feature.transpose(1, 2)], dim=2)
    C = tmpl.shape[2]
    if C % 4:
        tmpl = F.pad(tmpl, (0, _r4(C) - C))
    rows = _GroupRowsCL.apply(tmpl.contiguous(), topk.view(B, N * k))                 # (B, N*k, Cp)
    pooled = mlp_stack(rows.view(B * N * k, rows.shape[2]), parse_stack(xc.mlp), k, xc.training)
    fusion = from_channels_last(pooled.reshape(B, N, pooled.shape[1]))
    return seq_forward(xc.fea_layer, fusion)


def p2b_xcorr_forward(xc, template_feature, search_feature, template_xyz):
    """Fused P2B_XCorr.forward (models/head/xcorr.py:33-53); positions ordered (b, search j, template i) so that the
    max over the template dimension is a max over groups of n1 consecutive positions."""
    B, f, n1 = template_feature.shape
    n2 = search_feature.shape[2]
    if 128 % n1 != 0:
        raise RuntimeError(f"fused P2B_XCorr: number of template points {n1} must divide 128")
    sim = F.cosine_similarity(template_feature.unsqueeze(-1), search_feature.unsqueeze(2), dim=1)   # (B,n1,n2), eps 1e-8
    t_cl = template_feature.transpose(1, 2)                                                        # (B,n1,f)
    fusion = torch.cat([sim.transpose(1, 2).unsqueeze(-1),                                          # (B,n2,n1,1)
                        template_xyz.unsqueeze(1).expand(B, n2, n1, 3),
                        t_cl.unsqueeze(1).expand(B, n2, n1, f)], dim=3)
    C = fusion.shape[3]
    if C % 4:
        fusion = F.pad(fusion, (0, _r4(C) - C))
    pooled = mlp_stack(fusion.reshape(B * n2 * n1, fusion.shape[3]), parse_stack(xc.mlp), n1, xc.training)
    out = from_channels_last(pooled.reshape(B, n2, pooled.shape[1]))
    return seq_forward(xc.fea_layer, out)
