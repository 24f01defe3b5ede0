"""Torch definitions of the ``b200med_tcn_*`` entry points (test infrastructure, like oracle/).

Each function restates what ONE kernel of csrc/tcn.cu must compute, straight from the C-ABI contract in
include/b200med.h (pack / gradient record layouts included).  Two uses:
* CPU: the autograd node ``tcn.TcnStageFunction`` is run with ``ops.tcn_*`` replaced by these definitions and checked
  against the oracle TeCNo (oracle/nets.py, pinned to the reference by tests/golden) -- host logic, gradient routing,
  layouts (tests/test_tcn_host.py);
* GPU: every kernel is compared with its definition on the same inputs (tests/test_gpu_tcn.py).
"""
import torch

MAPS, WD, W1 = 64, 3 * 64 * 64, 64 * 64
PACK, GRAD = 32896, 16512
OFF_WDF, OFF_W1F, OFF_WDB, OFF_W1B, OFF_BIAS = 0, WD, WD + W1, 2 * WD + W1, 2 * WD + 2 * W1


def pack_layer(wd, bd, w1, b1):
    """WdF[k][ci][co] | W1F[ci][co] | WdB[k][co][ci] | W1B[co][ci] | b_d | b_1."""
    wd, w1 = wd.reshape(64, 64, 3), w1.reshape(64, 64)
    return torch.cat([wd.permute(2, 1, 0).reshape(-1), w1.t().reshape(-1), wd.permute(2, 0, 1).reshape(-1), w1.reshape(-1),
                      bd.reshape(-1), b1.reshape(-1)])


def offsets(dilation, causal):
    return (-2 * dilation, -dilation, 0) if causal else (-dilation, 0, dilation)


def shift_rows(x, off, tloc=None, trem=None):
    """rows[t] = x[t + off] when frame t + off lies in the same video as frame t, else 0."""
    T = x.shape[0]
    t = torch.arange(T, device=x.device)
    tl = t if tloc is None else tloc.long()
    tr = (T - 1 - t) if trem is None else trem.long()
    ok = (tl + off >= 0) & (off <= tr)
    src = (t + off).clamp(0, T - 1)
    return x[src] * ok.unsqueeze(1).to(x.dtype)


def layer_fwd(x, pack, dilation, causal, tloc=None, trem=None):
    wdf = pack[OFF_WDF:OFF_WDF + WD].view(3, 64, 64)       # [k][ci][co]
    w1f = pack[OFF_W1F:OFF_W1F + W1].view(64, 64)          # [ci][co]
    bd, b1 = pack[OFF_BIAS:OFF_BIAS + 64], pack[OFF_BIAS + 64:OFF_BIAS + 128]
    pre = bd + sum(shift_rows(x, o, tloc, trem) @ wdf[k] for k, o in enumerate(offsets(dilation, causal)))
    y = torch.relu(pre)
    return x + (y @ w1f + b1), y


def layer_bwd_hidden(dout, x, y, pack, dilation, causal, tloc=None, trem=None):
    """-> dpre, gradient record [GRAD] = dWd[co][ci][k] | dW1[co][ci] | db_d | db_1 (dropout off)."""
    w1b = pack[OFF_W1B:OFF_W1B + W1].view(64, 64)          # [co][ci]
    dz = dout
    dpre = (dz @ w1b) * (y > 0).to(dout.dtype)
    taps = torch.stack([shift_rows(x, o, tloc, trem) for o in offsets(dilation, causal)], dim=2)   # [T, ci, k]
    dwd = torch.einsum("to,tik->oik", dpre, taps)
    dw1 = dz.t() @ y
    return dpre, torch.cat([dwd.reshape(-1), dw1.reshape(-1), dpre.sum(0), dz.sum(0)])


def layer_bwd_input(dpre, dout, pack, dilation, causal, tloc=None, trem=None):
    wdb = pack[OFF_WDB:OFF_WDB + WD].view(3, 64, 64)       # [k][co][ci]
    return dout + sum(shift_rows(dpre, -o, tloc, trem) @ wdb[k] for k, o in enumerate(offsets(dilation, causal)))


def out_fwd(x, w, b):
    return (x @ w.t() + b).t().contiguous()                # [C, T]


def out_bwd(dlogits, w):
    return dlogits.t() @ w, dlogits.t().contiguous()       # dx [T,64], dl_t [T,C]


def softmax_fwd(logits):
    return torch.softmax(logits.t(), dim=1).contiguous()   # [T, C]


def softmax_bwd(p, dp):
    return (p * (dp - (p * dp).sum(1, keepdim=True))).t().contiguous()


class EmulatedOps:
    """Drop-in for the ``ops.tcn_*`` / ``ops.linear_*_f32`` calls of tcn.TcnStageFunction, on any device."""

    TCN_PACK_FLOATS, TCN_GRAD_FLOATS, TCN_MAPS = PACK, GRAD, MAPS

    @staticmethod
    def tcn_slots(T):
        return max(1, min(64, (T + 15) // 16))

    @staticmethod
    def tcn_pack(params, n_layers, out=None):               # `params`: the tensors themselves (no addresses on the host)
        return torch.stack([pack_layer(*[p.detach() for p in params[4 * l:4 * l + 4]]) for l in range(n_layers)])

    @staticmethod
    def tcn_layer_fwd(x, pack, out, y_save, dilation, causal, drop_p=0.0, seed=0, drop_base=0, tloc=None, trem=None):
        assert drop_p == 0.0
        o, y = layer_fwd(x, pack, dilation, causal, tloc, trem)
        out.copy_(o)
        if y_save is not None:
            y_save.copy_(y)
        return out

    @staticmethod
    def tcn_layer_bwd_hidden(dout, x, y, pack, dpre, partials, n_slots, dilation, causal, drop_p=0.0, seed=0, drop_base=0,
                             tloc=None, trem=None):
        d, g = layer_bwd_hidden(dout, x, y, pack, dilation, causal, tloc, trem)
        dpre.copy_(d)
        partials.zero_()
        partials[0].copy_(g)
        return dpre

    @staticmethod
    def tcn_layer_bwd_input(dpre, dout, pack, dx, dilation, causal, tloc=None, trem=None):
        dx.copy_(layer_bwd_input(dpre, dout, pack, dilation, causal, tloc, trem))
        return dx

    @staticmethod
    def tcn_reduce_grads(partials, n_layers, n_slots):
        return partials.sum(1)

    tcn_out_fwd = staticmethod(out_fwd)
    tcn_out_bwd = staticmethod(out_bwd)
    tcn_softmax_fwd = staticmethod(softmax_fwd)
    tcn_softmax_bwd = staticmethod(softmax_bwd)

    @staticmethod
    def linear_fwd_f32(x, w, b, relu, out=None):
        y = x @ w.t() + b
        y = torch.relu(y) if relu else y
        if out is not None:
            out.copy_(y)
            return out
        return y

    @staticmethod
    def linear_bwd_data_f32(dy, w, relu_out=None):
        return dy @ w

    @staticmethod
    def linear_bwd_weight_f32(dy, x, want_bias=True):
        return dy.t() @ x, dy.sum(0)

    # ---- whole-stage calls: the sequencing of b200med_tcn_stage_fwd / _bwd (csrc/tcn.cu), kernel by kernel
    @classmethod
    def tcn_stage_fwd(cls, x, softmax_in, in_w, in_b, params, n_layers, out_w, out_b, causal, drop_p=None, seed=0,
                      layer_base=0, keep=True, tloc=None, trem=None, seed_dev=None):
        assert not drop_p or max(drop_p) == 0.0
        xin = softmax_fwd(x) if softmax_in else x
        T = xin.shape[0]
        acts = torch.empty(n_layers + 1 if keep else 2, T, MAPS, dtype=x.dtype)
        ys = torch.empty(n_layers, T, MAPS, dtype=x.dtype) if keep else None
        cls.linear_fwd_f32(xin, in_w, in_b, False, out=acts[0])
        pack = cls.tcn_pack(params, n_layers)
        for l in range(n_layers):
            src, dst = (acts[l], acts[l + 1]) if keep else (acts[l & 1], acts[(l + 1) & 1])
            cls.tcn_layer_fwd(src, pack[l], dst, ys[l] if keep else None, 2 ** l, causal, 0.0, seed, 0, tloc, trem)
        last = acts[n_layers] if keep else acts[n_layers & 1]
        return dict(logits=out_fwd(last, out_w, out_b), xin=xin, acts=acts, ys=ys, pack=pack)

    @classmethod
    def tcn_stage_bwd(cls, dlogits, xin, softmax_in, in_w, out_w, n_layers, causal, acts, ys, pack, want_dx, drop_p=None,
                      seed=0, layer_base=0, tloc=None, trem=None, seed_dev=None):
        T = xin.shape[0]
        dA, dl_t = out_bwd(dlogits, out_w)
        d_out_w, d_out_b = cls.linear_bwd_weight_f32(dl_t, acts[n_layers])
        n_slots = cls.tcn_slots(T)
        partials = torch.empty(n_layers, n_slots, GRAD, dtype=dlogits.dtype)
        dpre, spare = torch.empty(T, MAPS, dtype=dlogits.dtype), torch.empty(T, MAPS, dtype=dlogits.dtype)
        for l in reversed(range(n_layers)):
            cls.tcn_layer_bwd_hidden(dA, acts[l], ys[l], pack[l], dpre, partials[l], n_slots, 2 ** l, causal, 0.0, seed, 0, tloc, trem)
            cls.tcn_layer_bwd_input(dpre, dA, pack[l], spare, 2 ** l, causal, tloc, trem)
            dA, spare = spare, dA
        d_in_w, d_in_b = cls.linear_bwd_weight_f32(dA, xin)
        dx = None
        if want_dx:
            dx = cls.linear_bwd_data_f32(dA, in_w)
            if softmax_in:
                dx = softmax_bwd(xin, dx)
        return dict(dx=dx, d_in_w=d_in_w, d_in_b=d_in_b, layer_grads=cls.tcn_reduce_grads(partials, n_layers, n_slots),
                    d_out_w=d_out_w, d_out_b=d_out_b)
