"""nn.LSTM(batch_first, multi-layer, inter-layer dropout) forward + backward on the b200med kernels
(throughput mode of the LSTM head, reference MED/modeling/models.py:161, 204-206).

hidden_size 128 (every reference config): ``LSTMRecFunction`` -- per layer ONE batched tcgen05 GEMM for the x-part
of the gates over all ``W*B`` rows + ONE persistent recurrence kernel that keeps W_hh in shared memory, the gates in
TMEM and walks all W steps (csrc/lstm_rec.cu); the backward mirrors it (persistent kernel -> dG, then batched
dX = dG W_ih and dW = dG^T [x|h] GEMMs).  Other hidden sizes: ``LSTMStackFunction`` -- one GEMM over
``[x_t | h_{t-1}]`` + one fused cell kernel per time step (csrc/lstm.cu).
Only ``h_{W-1}`` of the top layer is returned: that is all the reference head consumes
(``F.relu(out)[:, -1, :]``, models.py:205-206).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import ops
from ._lib import call


def _raw(ptr: int) -> C.c_void_p:
    return C.c_void_p(ptr)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class LSTMStackFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, drop_p, seed_dev, *params):
        """x [B, F, W] f32 (the head input); params = (w_ih, w_hh, b_ih, b_hh) per layer; returns h_{W-1} [B, H]."""
        B, F, W = x.shape
        L = len(params) // 4
        H = params[1].shape[1]
        dev = x.device
        st = _stream()
        A, G, Cs, Wcat, Kp, ins = [], [], [], [], [], []
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh = params[4 * l:4 * l + 4]
            in_l = F if l == 0 else H
            kp = (in_l + H + 7) // 8 * 8
            wc = torch.zeros(4 * H, kp, dtype=torch.float32, device=dev)
            wc[:, :in_l] = w_ih.detach()
            wc[:, in_l:in_l + H] = w_hh.detach()
            Wcat.append((ops.to_bf16(wc), (b_ih.detach() + b_hh.detach()).contiguous()))
            A.append(torch.empty(W, B, kp, dtype=torch.bfloat16, device=dev))
            G.append(torch.empty(W, B, 4 * H, dtype=torch.float32, device=dev))
            Cs.append(torch.empty(W, B, H, dtype=torch.float32, device=dev))
            Kp.append(kp); ins.append(in_l)
        call("b200med_lstm_pack_inputs", _raw(x.contiguous().data_ptr()), _raw(A[0].data_ptr()), B, B, F, W, H, Kp[0], F, 0, st)
        for l in range(1, L):
            call("b200med_zero_cols_bf16", _raw(A[l][0].data_ptr()), B, Kp[l], ins[l], H, st)
            if Kp[l] > ins[l] + H:
                call("b200med_zero_cols_bf16", _raw(A[l].data_ptr()), W * B, Kp[l], ins[l] + H, Kp[l] - ins[l] - H, st)
        out = torch.empty(B, H, dtype=torch.float32, device=dev)
        seed_ptr = 0 if seed_dev is None else seed_dev.data_ptr()
        for l in range(L):
            wb, bias = Wcat[l]
            p = drop_p if l < L - 1 else 0.0
            for t in range(W):
                ops.gemm_bf16(A[l][t], wb, B, 4 * H, Kp[l], True, True, bias=bias, out=G[l][t])
                call("b200med_lstm_cell_fwd", _raw(G[l][t].data_ptr()), _raw(Cs[l][t - 1].data_ptr() if t else 0),
                     _raw(Cs[l][t].data_ptr()),
                     _raw(A[l][t + 1].data_ptr() + 2 * ins[l] if t + 1 < W else 0), Kp[l],
                     _raw(A[l + 1][t].data_ptr() if l < L - 1 else 0), Kp[l + 1] if l < L - 1 else 0,
                     _raw(out.data_ptr() if (l == L - 1 and t == W - 1) else 0), B, H, float(p), _raw(seed_ptr),
                     (l * W + t) * B * H, st)
        ctx.save_for_backward(*A, *G, *Cs, *[w for w, _ in Wcat])
        ctx.meta = (B, F, W, L, H, Kp, ins, float(drop_p), seed_dev)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, F, W, L, H, Kp, ins, drop_p, seed_dev = ctx.meta
        saved = ctx.saved_tensors
        A, G, Cs, Wb = saved[:L], saved[L:2 * L], saved[2 * L:3 * L], saved[3 * L:4 * L]
        dev = dout.device
        st = _stream()
        seed_ptr = 0 if seed_dev is None else seed_dev.data_ptr()
        dout = dout.contiguous().float()
        grads = [None] * (4 * L)
        dA_up, kp_up = None, 0
        for l in reversed(range(L)):
            dG = torch.empty(W, B, 4 * H, dtype=torch.bfloat16, device=dev)
            dA = torch.empty(W, B, Kp[l], dtype=torch.float32, device=dev)
            dc = torch.empty(B, H, dtype=torch.float32, device=dev)
            for t in reversed(range(W)):
                if l == L - 1:
                    up_ptr, up_ld, p = (dout.data_ptr() if t == W - 1 else 0), H, 0.0
                else:
                    up_ptr, up_ld, p = dA_up[t].data_ptr(), kp_up, drop_p
                rec_ptr = dA[t + 1].data_ptr() + 4 * ins[l] if t + 1 < W else 0
                call("b200med_lstm_cell_bwd", _raw(G[l][t].data_ptr()), _raw(Cs[l][t].data_ptr()),
                     _raw(Cs[l][t - 1].data_ptr() if t else 0), _raw(up_ptr), up_ld, _raw(rec_ptr), Kp[l],
                     _raw(dc.data_ptr()), int(t == W - 1), _raw(dG[t].data_ptr()), B, H, float(p), _raw(seed_ptr),
                     (l * W + t) * B * H, st)
                # [dx_t | dh_{t-1}] [B, Kp] = dG_t [B, 4H] * Wcat [4H, Kp]   (B operand MN-major)
                ops.gemm_bf16(dG[t], Wb[l], B, Kp[l], 4 * H, True, False, out_dtype=torch.float32, out=dA[t])
            # dWcat [4H, Kp] = dG^T A over all W*B rows (both operands MN-major), deterministic split-K
            dW = ops.gemm_bf16(dG.view(W * B, 4 * H), A[l].view(W * B, Kp[l]), 4 * H, Kp[l], W * B, False, False,
                               out_dtype=torch.float32, split_k=ops.gemm_split_k(4 * H, Kp[l], W * B))
            db = ops.colsum(dG.view(W * B, 4 * H))
            grads[4 * l] = dW[:, :ins[l]].contiguous()
            grads[4 * l + 1] = dW[:, ins[l]:ins[l] + H].contiguous()
            grads[4 * l + 2] = db
            grads[4 * l + 3] = db.clone()
            dA_up, kp_up = dA, Kp[l]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, F, W, dtype=torch.float32, device=dev)
            call("b200med_lstm_unpack_dx", _raw(dA_up.data_ptr()), _raw(dx.data_ptr()), B, B, F, W, Kp[0], 0, st)
        return (dx, None, None, *grads)


_PERM = {}
_SIDE = {}
# Generation of the forward recurrence kernel: 2 = 2-CTA clusters with the x-part fused in (csrc/lstm_rec2.cu), 1 = the
# first-generation kernel behind a separate x-part GEMM (csrc/lstm_rec.cu).  Both write the same buffers; the tests hold
# one against the other.  Not a user switch: the product runs generation 2.
REC_GEN = 2
# Deferred join of the side stream.  The weight / bias gradients of the LSTM layers are computed on a side stream; with
# DEFER_JOIN the backward returns them WITHOUT making the main stream wait (autograd only stores the tensors when
# `.grad is None`), so the side work of layer 0 overlaps the FeatureExtractor backward.  Whoever consumes the gradients
# (FusedAdam._refresh_active -> all-reduce / Adam) calls join_pending() first.  Off by default: a caller that reads
# `.grad` right after backward() gets the join inside the backward.
DEFER_JOIN = False
_PENDING = []
SIDE_SMS = 148       # SMs the side-stream weight-gradient GEMMs may fill.  All of them: the step is bound by its total work, not by
# the recurrence chain (scripts/step_timeline.py) -- with 20 / 40 / 60 / 80 SMs the gradients pile up in front of the join at the
# end of the backward (2.03 / 1.90 / 1.86 / 1.83 ms per step against 1.81 ms with 148, B = 8192)


def join_pending():
    """Make the current stream wait for gradient work still running on the side stream."""
    while _PENDING:
        side, keep = _PENDING.pop()
        torch.cuda.current_stream().wait_stream(side)
        keep.clear()


def _side_stream(device) -> torch.cuda.Stream:
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]



def _dg_perm(H: int, device):
    """Column permutation of the recurrence kernel's dG output (csrc/lstm_rec.cu): returns (orig_of_col', col'_of_orig)."""
    key = (H, str(device))
    if key not in _PERM:
        cp = torch.arange(4 * H)
        c, uq, g, i = cp // 128, (cp % 128) // 32, (cp % 32) // 8, cp % 8
        orig = g * H + uq * 32 + c * 8 + i
        inv = torch.empty_like(orig)
        inv[orig] = cp
        _PERM[key] = (orig.to(device), inv.to(device))
    return _PERM[key]


def _dg_perm2(device):
    """Column order of generation 2's dG (csrc/lstm_rec2.cu): column' = chunk*256 + unit_half*128 + warp_column*32 + gate*8 + i
    stands for gate row gate*128 + unit_half*64 + warp_column*16 + chunk*8 + i.  Returns (orig_of_col', col'_of_orig)."""
    key = ("gen2", str(device))
    if key not in _PERM:
        kp = torch.arange(512)
        c, uh, cq, g, i = kp >> 8, (kp >> 7) & 1, (kp >> 5) & 3, (kp >> 3) & 3, kp & 7
        orig = g * 128 + uh * 64 + cq * 16 + c * 8 + i
        inv = torch.empty_like(orig)
        inv[orig] = kp
        _PERM[key] = (orig.to(device), inv.to(device))
    return _PERM[key]


class WindowParts:
    """The kinematics half of the window path's head input, still in the frame table: (table [N, Cb] f32, mean, std
    ([1 | W, Cb] or None), starts [B] int32).  With it the recurrence's first operand is built from the FeatureExtractor output
    [B, W, Ca] and the table rows in ONE kernel (b200med_lstm_pack_parts) instead of gather + torch.cat + pack."""

    def __init__(self, table, mean, std, starts):
        self.table, self.mean, self.std, self.starts = table, mean, std, starts


_PREPACKED = {}      # id(weight_ih_l0) -> what prepack() left for the next forward of that LSTM


def _prepack(params, B, W, F, need_grad, dev):
    """Generation-2 recurrence: packed weight operands (wp, bias_p, wt) per layer and the upper layers' operand buffers with
    their h_{-1} columns zeroed, all on the side stream (forked from the current one here)."""
    L = len(params) // 4
    H = params[1].shape[1]
    Bp = (B + 31) // 32 * 32
    ins = [F if l == 0 else H for l in range(L)]
    inp = [(i + 63) // 64 * 64 for i in ins]
    Kp = [ip + H for ip in inp]
    alloc = torch.empty if Bp == B else torch.zeros
    main, side = torch.cuda.current_stream(), _side_stream(dev)
    side.wait_stream(main)
    packed, upper = [], []
    with torch.cuda.stream(side):
        sst = _stream()
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh = (t.detach().contiguous() for t in params[4 * l:4 * l + 4])
            wp = torch.empty(4 * H, Kp[l], dtype=torch.bfloat16, device=dev)
            bias_p = torch.empty(4 * H, dtype=torch.float32, device=dev)
            call("b200med_lstm_pack_weights2", _raw(w_ih.data_ptr()), _raw(w_hh.data_ptr()), _raw(b_ih.data_ptr()),
                 _raw(b_hh.data_ptr()), ins[l], inp[l], _raw(wp.data_ptr()), _raw(bias_p.data_ptr()), sst)
            wt = None
            if need_grad:
                wt = torch.empty(Kp[l], 4 * H, dtype=torch.bfloat16, device=dev)
                call("b200med_lstm_pack_weights2_bwd", _raw(w_ih.data_ptr()), _raw(w_hh.data_ptr()), ins[l], inp[l],
                     _raw(wt.data_ptr()), _raw(0), sst)
            packed.append((wp, bias_p, wt))
        for l in range(1, L):
            a = alloc(W, Bp, Kp[l], dtype=torch.bfloat16, device=dev)
            call("b200med_zero_cols_bf16", _raw(a[0].data_ptr()), Bp, Kp[l], inp[l], H, sst)
            upper.append(a)
    return {"key": (B, W, F, L, H, bool(need_grad)), "side": side, "packed": packed, "upper": upper}


def prepack(lstm: torch.nn.LSTM, B: int, W: int, F: int, need_grad: bool = True):
    """Start the parameter-only preparation of the next ``lstm_last_hidden(..)`` call of this LSTM (bf16 mode, hidden size 128)
    NOW, on the side stream: a train step calls it first thing, so that these ~8 small kernels run under the FeatureExtractor
    instead of between it and the recurrence.  Harmless when the forward then runs with other shapes (it redoes the work)."""
    if REC_GEN != 2 or lstm.hidden_size != 128 or lstm.bidirectional or lstm.proj_size or not lstm.bias:
        return False
    params = []
    for l in range(lstm.num_layers):
        params += [getattr(lstm, f"weight_ih_l{l}"), getattr(lstm, f"weight_hh_l{l}"),
                   getattr(lstm, f"bias_ih_l{l}"), getattr(lstm, f"bias_hh_l{l}")]
    if (F + 63) // 64 * 64 not in (64, 128) or not params[0].is_cuda:
        return False
    _PREPACKED[id(params[0])] = _prepack(params, B, W, F, need_grad, params[0].device)
    return True


class LSTMRecFunction(torch.autograd.Function):
    """Persistent-recurrence path (hidden_size 128).  Layer buffers, time-major with the batch padded to Bp (multiple
    of 32) rows per step:
    A_l [W, Bp, inp_l + H] bf16 row-major = [x_t (padded to a multiple of 64) | h_{t-1}] (operand of the x-part GEMM and of dW),
    dG_l [W, Bp, 4H] bf16 row-major; row-block-interleaved (csrc/lstm_rec.cu): xg / gact_l [W*Bp, 4H] fp16, c_l [W*Bp, H] f32,
    dX_l [W*Bp, inp_l] f32."""

    @staticmethod
    def forward(ctx, x, drop_p, seed_dev, parts, *params):
        """x: the head input [B, F, W]; or, with ``parts`` (a WindowParts), the FeatureExtractor output [B, W, Ca]."""
        if parts is None:
            B, F, W = x.shape
        else:
            B, W, Ca = x.shape
            F = Ca + parts.table.shape[1]
        L = len(params) // 4
        H = params[1].shape[1]
        dev = x.device
        st = _stream()
        Bp = (B + 31) // 32 * 32
        ins = [F if l == 0 else H for l in range(L)]
        inp = [(i + 63) // 64 * 64 for i in ins]
        Kp = [ip + H for ip in inp]
        need_grad = any(ctx.needs_input_grad)
        gen2 = REC_GEN == 2 and all(i in (64, 128) for i in inp)      # generation 2 stages [x | h] tiles of 64 / 128 + 128 columns
        alloc = torch.empty if Bp == B else torch.zeros          # pad rows must hold finite values
        packed, side = None, None
        pre = _PREPACKED.pop(id(params[0]), None)
        if gen2:
            # The weight operands of ALL layers (forward layout, and the transposed one of the backward) and the zeroed h_{-1}
            # columns of the upper layers' operands depend on the parameters only: they are made on the side stream -- at the
            # start of the step when the caller asked for it (prepack(): under the FeatureExtractor's kernels), else here.
            if pre is None or pre["key"] != (B, W, F, L, H, bool(need_grad)):
                pre = _prepack(params, B, W, F, need_grad, dev)
            side, packed, upper = pre["side"], pre["packed"], pre["upper"]
            A = [alloc(W, Bp, Kp[0], dtype=torch.bfloat16, device=dev)] + upper
        else:
            A = [alloc(W, Bp, Kp[l], dtype=torch.bfloat16, device=dev) for l in range(L)]
        # the reference hands the head cat(...).permute(0, 2, 1): a view of a contiguous [B, W, F] tensor -- read it as such
        feat_bf16 = False
        if parts is not None:
            bwf = True
            kt, km, ks = parts.table, parts.mean, parts.std
            Cb = kt.shape[1]
            feat_bf16 = (x.dtype == torch.bfloat16 and Ca % 2 == 0 and Cb % 2 == 0 and Ca + Cb <= 64 and inp[0] == 64
                         and (km is None or km.shape[0] == 1))
            if feat_bf16:      # bf16 features (TableWindows.out_bf16): copied as they are; their gradient leaves in bf16 too
                feats = x.contiguous()
                call("b200med_lstm_pack_parts_bf16", _raw(feats.data_ptr()), Ca, _raw(kt.data_ptr()), kt.shape[0], Cb,
                     _raw(km.data_ptr() if km is not None else 0), _raw(ks.data_ptr() if ks is not None else 0),
                     _raw(parts.starts.data_ptr()), _raw(A[0].data_ptr()), B, Bp, W, H, Kp[0], st)
            else:
                feats = x.contiguous().float()
                call("b200med_lstm_pack_parts", _raw(feats.data_ptr()), Ca, _raw(kt.data_ptr()), kt.shape[0], kt.shape[1],
                     _raw(km.data_ptr() if km is not None else 0), _raw(ks.data_ptr() if ks is not None else 0),
                     km.shape[0] if km is not None else 1, _raw(parts.starts.data_ptr()), _raw(A[0].data_ptr()), B, Bp, W, H, Kp[0], inp[0], st)
        else:
            bwf = x.transpose(1, 2).is_contiguous() and not x.is_contiguous()
            xsrc = x.transpose(1, 2) if bwf else x.contiguous()
            call("b200med_lstm_pack_inputs", _raw(xsrc.data_ptr()), _raw(A[0].data_ptr()), B, Bp, F, W, H, Kp[0], inp[0], int(bwf), st)
        for l in range(1, L if not gen2 else 0):
            call("b200med_zero_cols_bf16", _raw(A[l][0].data_ptr()), Bp, Kp[l], inp[l], H, st)
        out = torch.empty(B, H, dtype=torch.float32, device=dev)
        seed_ptr = 0 if seed_dev is None else seed_dev.data_ptr()
        Wih, Whh, Gact, Cs = [], [], [], []
        if gen2:
            torch.cuda.current_stream().wait_stream(side)
            for t_ in [t_ for tup in packed for t_ in tup] + A[1:]:
                if t_ is not None:
                    t_.record_stream(torch.cuda.current_stream())     # allocated on the side stream, read on this one
        xg = None if gen2 else torch.empty(W * Bp, 4 * H, dtype=torch.float16, device=dev)
        # rows of the sigmoid gates (i, f, o) are halved for the forward kernels: sigmoid(z) = 0.5 * tanh(z / 2) + 0.5
        half = None
        if not gen2:
            half = torch.ones(4 * H, 1, dtype=torch.float32, device=dev)
            half[:2 * H] = 0.5
            half[3 * H:] = 0.5
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh = params[4 * l:4 * l + 4]
            if need_grad and not gen2:                                                 # unscaled copies: backward operands
                wpad = torch.zeros(4 * H, inp[l], dtype=torch.float32, device=dev)
                wpad[:, :ins[l]] = w_ih.detach()
                Wih.append(ops.to_bf16(wpad))
                Whh.append(ops.to_bf16(w_hh.detach().contiguous()))
            elif need_grad:                                                            # generation 2: the packed backward operand
                Wih.append(packed[l][2]); Whh.append(None)
            else:
                Wih.append(None); Whh.append(None)
            gact = torch.empty(W * Bp, 4 * H, dtype=torch.float16, device=dev) if need_grad else None
            cs = torch.empty(W * Bp, H, dtype=torch.float32, device=dev) if need_grad else None
            top = l == L - 1
            if gen2:
                wp, bias_p, _ = packed[l]
                call("b200med_lstm_rec2_fwd", _raw(A[l].data_ptr()), Kp[l], inp[l], _raw(wp.data_ptr()), _raw(bias_p.data_ptr()),
                     _raw(gact.data_ptr() if need_grad else 0), _raw(cs.data_ptr() if need_grad else 0),
                     _raw(0 if top else A[l + 1].data_ptr()), 0 if top else Kp[l + 1], _raw(out.data_ptr() if top else 0),
                     B, Bp, W, 0.0 if top else float(drop_p), _raw(seed_ptr), l * W * Bp * H, st)
            else:
                wpad = torch.zeros(4 * H, inp[l], dtype=torch.float32, device=dev)
                wpad[:, :ins[l]] = w_ih.detach()
                wih_f = ops.to_bf16(wpad * half)
                whh_f = ops.to_bf16(w_hh.detach() * half)
                bias = ((b_ih.detach() + b_hh.detach()) * half[:, 0]).contiguous()
                # x-part of the gates for every step at once: [W*Bp, inp] x [4H, inp]^T + bias
                ops.gemm_bf16(A[l].view(W * Bp, Kp[l]), wih_f, W * Bp, 4 * H, inp[l], True, True, bias=bias, out=xg, rbi=True)
                call("b200med_lstm_rec_fwd", _raw(xg.data_ptr()), _raw(whh_f.data_ptr()),
                     _raw(gact.data_ptr() if need_grad else 0), _raw(cs.data_ptr() if need_grad else 0),
                     _raw(A[l].data_ptr() if need_grad else 0), Kp[l], inp[l],
                     _raw(0 if top else A[l + 1].data_ptr()), 0 if top else Kp[l + 1],
                     _raw(out.data_ptr() if top else 0), B, Bp, W, H, 0.0 if top else float(drop_p), _raw(seed_ptr),
                     l * W * Bp * H, st)
            Gact.append(gact); Cs.append(cs)
        if need_grad:
            ctx.save_for_backward(*A, *Gact, *Cs, *Wih, *Whh)
        ctx.meta = (B, Bp, F, W, L, H, Kp, ins, inp, float(drop_p), seed_dev, bwf)
        ctx.gen2 = gen2
        ctx.n_feat = None if parts is None else Ca      # with parts only the FeatureExtractor columns carry a gradient
        ctx.feat_bf16 = feat_bf16
        return out

    @staticmethod
    def backward(ctx, dout):
        B, Bp, F, W, L, H, Kp, ins, inp, drop_p, seed_dev, bwf = ctx.meta
        gen2 = ctx.gen2
        saved = ctx.saved_tensors
        A, Gact, Cs, Wih, Whh = (saved[i * L:(i + 1) * L] for i in range(5))
        dev = dout.device
        st = _stream()
        seed_ptr = 0 if seed_dev is None else seed_dev.data_ptr()
        dout = dout.contiguous().float()
        grads = [None] * (4 * L)
        dX_up = None
        orig_of, colp_of = _dg_perm2(dev) if gen2 else _dg_perm(H, dev)
        main, side = torch.cuda.current_stream(), _side_stream(dev)
        keep = []        # tensors read on the side stream stay referenced until the join (no early reuse of their memory)
        for l in reversed(range(L)):
            top = l == L - 1
            dG = torch.empty(W * Bp, 4 * H, dtype=torch.bfloat16, device=dev)   # gate columns permuted (see _dg_perm)
            keep.append(dG)
            keep.append(A[l])       # read by the side-stream weight-gradient GEMM after autograd has released the saved tensors
            if gen2:
                # one launch: dG_t, dh_{t-1} AND dX_t = dG_t W_ih (csrc/lstm_rec2.cu); dX leaves in the layout its consumer reads
                # (row-block-interleaved for the recurrence of the layer below, row-major for the unpack into [B, F, W])
                wt = Wih[l]           # [Kp, 4H] bf16, packed in the forward (lstm_pack_weights2_bwd)
                dX = torch.empty(W * Bp, inp[l], dtype=torch.float32, device=dev)
                call("b200med_lstm_rec2_bwd", _raw(Gact[l].data_ptr()), _raw(Cs[l].data_ptr()), _raw(wt.data_ptr()), inp[l],
                     _raw(dout.data_ptr() if top else 0), _raw(0 if top else dX_up.data_ptr()), _raw(dG.data_ptr()),
                     _raw(dX.data_ptr()), B, Bp, W, 0.0 if top else drop_p, _raw(seed_ptr), l * W * Bp * H, st)
            else:
                call("b200med_lstm_rec_bwd", _raw(Gact[l].data_ptr()), _raw(Cs[l].data_ptr()), _raw(Whh[l].data_ptr()),
                     _raw(dout.data_ptr() if top else 0), _raw(0 if top else dX_up.data_ptr()), 0 if top else inp[l + 1],
                     _raw(dG.data_ptr()), B, Bp, W, H, 0.0 if top else drop_p, _raw(seed_ptr), l * W * Bp * H, st)
            # The weight / bias gradients of this layer are off the critical path (dG_l -> dX_l -> recurrence of layer l-1):
            # they run on a side stream, next to the next recurrence kernel which only fills 64 of the 148 SMs.
            side.wait_stream(main)
            with torch.cuda.stream(side), ops.sm_limit(SIDE_SMS):
                # dWcat [4H, Kp] = dG^T [x | h_prev] over all W*Bp rows (both operands MN-major), deterministic split-K
                dW = ops.gemm_bf16(dG, A[l].view(W * Bp, Kp[l]), 4 * H, Kp[l], W * Bp, False, False,
                                   out_dtype=torch.float32, split_k=ops.gemm_split_k(4 * H, Kp[l], W * Bp))
                dbp = ops.colsum(dG)
                if gen2:      # one kernel: un-permute the gate rows, split [x | h] columns, both bias gradients
                    gw = [torch.empty(4 * H, ins[l], dtype=torch.float32, device=dev), torch.empty(4 * H, H, dtype=torch.float32, device=dev),
                          torch.empty(4 * H, dtype=torch.float32, device=dev), torch.empty(4 * H, dtype=torch.float32, device=dev)]
                    call("b200med_lstm_unpack_grads2", _raw(dW.data_ptr()), _raw(dbp.data_ptr()), ins[l], inp[l], _raw(gw[0].data_ptr()),
                         _raw(gw[1].data_ptr()), _raw(gw[2].data_ptr()), _raw(gw[3].data_ptr()), _stream())
                    grads[4 * l:4 * l + 4] = gw
                else:
                    db = dbp.index_select(0, colp_of)
                    dW = dW.index_select(0, colp_of)
                    grads[4 * l] = dW[:, :ins[l]].contiguous()
                    grads[4 * l + 1] = dW[:, inp[l]:inp[l] + H].contiguous()
                    grads[4 * l + 2] = db
                    grads[4 * l + 3] = db.clone()
            if gen2:
                dX_up = dX
            elif l > 0 or ctx.needs_input_grad[0]:
                # dX [W*Bp, inp] = dG [W*Bp, 4H] * W_ih [4H, inp]   (B operand MN-major); interleaved rows for the layer
                # below's recurrence kernel, row-major for the unpack into [B, F, W]
                dX_up = ops.gemm_bf16(dG, Wih[l].index_select(0, orig_of), W * Bp, inp[l], 4 * H, True, False,
                                      out_dtype=torch.float32, rbi=l > 0)
        if DEFER_JOIN:
            _PENDING.append((side, keep))      # joined by the gradient consumer (join_pending)
        else:
            main.wait_stream(side)
            keep.clear()
        for g in grads:
            g.record_stream(main)
        dx = None
        if ctx.needs_input_grad[0] and ctx.n_feat is not None:
            # columns [0, Ca) of dX_0, window-major: the gradient of the FeatureExtractor output (no cat to slice back out of)
            if ctx.feat_bf16:
                dx = torch.empty(B, W, ctx.n_feat, dtype=torch.bfloat16, device=dev)
                call("b200med_lstm_unpack_dx_bf16", _raw(dX_up.data_ptr()), _raw(dx.data_ptr()), B, Bp, ctx.n_feat, W, inp[0], st)
            else:
                dx = torch.empty(B, W, ctx.n_feat, dtype=torch.float32, device=dev)
                call("b200med_lstm_unpack_dx", _raw(dX_up.data_ptr()), _raw(dx.data_ptr()), B, Bp, ctx.n_feat, W, inp[0], 1, st)
        elif ctx.needs_input_grad[0]:
            dx = torch.empty((B, W, F) if bwf else (B, F, W), dtype=torch.float32, device=dev)
            call("b200med_lstm_unpack_dx", _raw(dX_up.data_ptr()), _raw(dx.data_ptr()), B, Bp, F, W, inp[0], int(bwf), st)
            if bwf:
                dx = dx.transpose(1, 2)          # gradient of the permuted view, in the layout of its base tensor
        return (dx, None, None, None, *grads)


class LSTMF32Function(torch.autograd.Function):
    """fp32 parity mode (1e-5 bar): the same recurrence with fp32 operands on the SIMT GEMM (csrc/gemm_f32.cu) and exact-math
    cell kernels (csrc/lstm.cu).  Time-major fp32 buffers X_l [W, B, in_l], G_l [W, B, 4H], C_l / Hs_l [W, B, H]; per layer one
    GEMM for the x-part of all steps, then per step one accumulating GEMM (h_{t-1} W_hh^T) + one cell kernel."""

    @staticmethod
    def forward(ctx, x, drop_p, seed_dev, *params):
        B, F, W = x.shape
        L = len(params) // 4
        H = params[1].shape[1]
        dev = x.device
        st = _stream()
        bwf = x.transpose(1, 2).is_contiguous() and not x.is_contiguous()
        xsrc = (x.transpose(1, 2) if bwf else x.contiguous()).float()
        X = [torch.empty(W, B, F, dtype=torch.float32, device=dev)]
        call("b200med_lstm_pack_f32", _raw(xsrc.data_ptr()), _raw(X[0].data_ptr()), B, F, W, int(bwf), st)
        seed_ptr = 0 if seed_dev is None else seed_dev.data_ptr()
        Gs, Cs, Hs = [], [], []
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh = (q.detach() for q in params[4 * l:4 * l + 4])
            in_l = F if l == 0 else H
            G = torch.empty(W, B, 4 * H, dtype=torch.float32, device=dev)
            ops.linear_f32(X[l].view(W * B, in_l), w_ih, b_ih + b_hh, out=G.view(W * B, 4 * H))
            C_ = torch.empty(W, B, H, dtype=torch.float32, device=dev)
            H_ = torch.empty(W, B, H, dtype=torch.float32, device=dev)
            top = l == L - 1
            Xn = None if top else torch.empty(W, B, H, dtype=torch.float32, device=dev)
            for t in range(W):
                if t > 0:
                    ops.linear_f32(H_[t - 1], w_hh, None, ops.GEMM_ACCUM, out=G[t])
                call("b200med_lstm_cell_fwd_f32", _raw(G[t].data_ptr()), _raw(C_[t - 1].data_ptr() if t else 0), _raw(C_[t].data_ptr()),
                     _raw(H_[t].data_ptr()), _raw(0 if top else Xn[t].data_ptr()), B, H, 0.0 if top else float(drop_p),
                     _raw(seed_ptr), (l * W + t) * B * H, st)
            Gs.append(G); Cs.append(C_); Hs.append(H_)
            if not top:
                X.append(Xn)
        out = Hs[-1][W - 1].clone()
        ctx.save_for_backward(*X, *Gs, *Cs, *Hs, *[params[4 * l].detach() for l in range(L)], *[params[4 * l + 1].detach() for l in range(L)])
        ctx.meta = (B, F, W, L, H, float(drop_p), seed_dev, bwf)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, F, W, L, H, drop_p, seed_dev, bwf = ctx.meta
        sv = ctx.saved_tensors
        X, Gs, Cs, Hs, Wih, Whh = (sv[i * L:(i + 1) * L] for i in range(6))
        dev = dout.device
        st = _stream()
        seed_ptr = 0 if seed_dev is None else seed_dev.data_ptr()
        dout = dout.contiguous().float()
        grads = [None] * (4 * L)
        dX_up = None
        for l in reversed(range(L)):
            top = l == L - 1
            in_l = F if l == 0 else H
            dG = torch.empty(W, B, 4 * H, dtype=torch.float32, device=dev)
            dc = torch.empty(B, H, dtype=torch.float32, device=dev)
            dh_rec = None
            for t in reversed(range(W)):
                if top:
                    up_ptr, p = (dout.data_ptr() if t == W - 1 else 0), 0.0
                else:
                    up_ptr, p = dX_up[t].data_ptr(), drop_p
                call("b200med_lstm_cell_bwd_f32", _raw(Gs[l][t].data_ptr()), _raw(Cs[l][t].data_ptr()),
                     _raw(Cs[l][t - 1].data_ptr() if t else 0), _raw(up_ptr), H, _raw(0 if dh_rec is None else dh_rec.data_ptr()),
                     _raw(dc.data_ptr()), int(t == W - 1), _raw(dG[t].data_ptr()), B, H, float(p), _raw(seed_ptr),
                     (l * W + t) * B * H, st)
                if t > 0:
                    dh_rec = ops.linear_dgrad_f32(dG[t], Whh[l])
            dG2 = dG.view(W * B, 4 * H)
            grads[4 * l] = ops.linear_wgrad_f32(dG2, X[l].view(W * B, in_l))
            if W > 1:
                grads[4 * l + 1] = ops.linear_wgrad_f32(dG[1:].view((W - 1) * B, 4 * H), Hs[l][:W - 1].view((W - 1) * B, H))
            else:
                grads[4 * l + 1] = torch.zeros_like(Whh[l])
            db = ops.colsum(dG2)
            grads[4 * l + 2] = db
            grads[4 * l + 3] = db.clone()
            if l > 0 or ctx.needs_input_grad[0]:
                dX_up = ops.linear_dgrad_f32(dG2, Wih[l]).view(W, B, in_l)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((B, W, F) if bwf else (B, F, W), dtype=torch.float32, device=dev)
            call("b200med_lstm_unpack_f32", _raw(dX_up.data_ptr()), _raw(dx.data_ptr()), B, F, W, F, int(bwf), st)
            if bwf:
                dx = dx.transpose(1, 2)
        return (dx, None, None, *grads)


def lstm_last_hidden(x: torch.Tensor, lstm: torch.nn.LSTM, training: bool, seed_dev=None, impl: str = "auto",
                     precision: str = "bf16", parts: WindowParts = None) -> torch.Tensor:
    """h_{W-1} of the top layer for head input x [B, F, W].  precision "fp32": exact-math fp32 kernels (1e-5 mode);
    "bf16": persistent tcgen05 recurrence (hidden_size 128) or per-step tcgen05 gate GEMMs + cell kernels (other sizes;
    ``impl="per_step"`` forces that path -- used by the tests that hold the two implementations against each other)."""
    if not x.is_cuda:
        raise RuntimeError("b200med LSTM head runs on CUDA tensors only (no CPU fallback)")
    if lstm.bidirectional or lstm.proj_size or not lstm.bias:
        raise ValueError("b200med LSTM head: unidirectional nn.LSTM with biases and no projection (what the reference builds)")
    params = []
    for l in range(lstm.num_layers):
        params += [getattr(lstm, f"weight_ih_l{l}"), getattr(lstm, f"weight_hh_l{l}"),
                   getattr(lstm, f"bias_ih_l{l}"), getattr(lstm, f"bias_hh_l{l}")]
    p = float(lstm.dropout) if training else 0.0
    if precision == "fp32":
        if parts is not None:
            raise ValueError("WindowParts input is served by the bf16 mode only")
        return LSTMF32Function.apply(x, p, seed_dev, *params)
    if precision != "bf16":
        raise ValueError(f"precision {precision!r} is not supported (fp32 | bf16)")
    if lstm.hidden_size == 128 and impl != "per_step":
        return LSTMRecFunction.apply(x, p, seed_dev, parts, *params)
    if parts is not None:
        raise ValueError("WindowParts input is served by the persistent recurrence path only (hidden_size 128)")
    return LSTMStackFunction.apply(x, p, seed_dev, *params)
