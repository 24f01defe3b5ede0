"""Drop-in for the reference ``MED/modeling/models.py`` (FeatureExtractor, CNN, LSTM).

Same constructor signatures, ``forward`` contracts, ``.name`` attributes, ``state_dict`` keys and
seed-42 initialisation order as the reference (models.py:6-220), so checkpoints interchange.

* ``FeatureExtractor`` -- ~98 % (CNN head) / ~75 % (LSTM head) of the model FLOPs -- runs on the
  hand-written kernels of libb200med.so: fp32 SIMT GEMMs in ``precision="fp32"`` (1e-5 parity mode)
  or bf16 tcgen05/TMEM GEMMs in ``precision="bf16"`` (2e-2 throughput mode), forward and backward.
* ``CNN`` / ``LSTM`` heads run on the kernels of csrc/head.cu, csrc/gemm_f32.cu, csrc/lstm.cu and
  csrc/lstm_rec.cu through ``heads.py`` / ``lstm_stack.py``; their ``nn.Module`` members only hold the parameters
  (same ``state_dict`` keys as the reference).  There is no torch-layer path: a CPU tensor raises.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


# Data-parallel overlap hook (set by modeling_utils._backward when several ranks train): called inside the backward of the
# FeatureExtractor right BEFORE the gradients of its FIRST layer -- the last and longest GEMM of the whole backward (dW1 is a
# [512 x 2048] product over all B*W rows) -- with every other gradient of the step already computed.  The hook moves those into
# the flat gradient buffer and starts their all-reduce, which then runs UNDER the dW1 GEMM; only the first layer's own gradient
# is exchanged after the backward.  Signature: hook(params, grads) -> grads (entries may be replaced by flat-buffer views).
EARLY_EXCHANGE_HOOK = None


class TableWindows:
    """The FeatureExtractor's input as (device-resident fp32 table, per-column mean / std, window start rows, window length)
    instead of a gathered batch: the first layer's kernel gathers and standardises the windows itself (csrc/gather_gemm.cu)."""

    def __init__(self, table, mean, std, starts, W, events=None, out_bf16=False):
        self.table, self.mean, self.std, self.starts, self.W, self.events = table, mean, std, starts, W, events
        # features leave in bf16 (for a consumer that rounds them to bf16 anyway: the LSTM's first operand): same operand bits,
        # and the gradient comes back in bf16 -- the rounding this layer's backward applies first -- without a cast pass
        self.out_bf16 = out_bf16


def _bf16_copies(weights):
    """bf16 copies of all layers' weights in ONE launch (b200med_multi_copy_f32; was one cast kernel per layer in front of the
    first GEMM of the step)."""
    src = [w.detach().contiguous() for w in weights]
    wb = [torch.empty(w.shape, dtype=torch.bfloat16, device=w.device) for w in src]
    ops.multi_copy([b.view(-1) for b in wb], [w.view(-1) for w in src])
    return wb


class _MLPFunction(torch.autograd.Function):
    """y = L_n(...relu(L_1(x))) with every product on the b200med GEMM kernels (K2)."""

    @staticmethod
    def forward(ctx, x, precision, *params):
        weights, biases = params[0::2], params[1::2]
        n = len(weights)
        ctx.precision, ctx.n = precision, n
        ctx.params = params            # the Parameter objects: the data-parallel hook looks their flat-buffer slots up
        if isinstance(x, TableWindows):
            # K1 fused into the first layer: table rows -> standardised bf16 operand tile -> tcgen05.mma; the bf16 batch is
            # written once (the backward's weight-gradient operand) and never read back by this layer
            if precision != "bf16" or n < 2:
                raise ValueError("the fused gather serves the bf16 mode of an MLP with at least two layers")
            wb = _bf16_copies(weights)
            # inference (no gradient will be asked for): the bf16 batch is not written at all -- 0.54 GB per 8192 windows
            keep = any(ctx.needs_input_grad)      # all False under torch.no_grad() / for frozen parameters
            xb, y1 = ops.gather_linear_bf16(x.table, x.mean, x.std, x.starts, x.W, wb[0], biases[0].detach(), relu=True, events=x.events,
                                            want_xb=keep)
            acts = [xb, y1]
            M = y1.shape[0]
            for i in range(1, n):
                N, K = weights[i].shape
                last = i == n - 1
                acts.append(ops.gemm_bf16(acts[-1], wb[i], M, N, K, True, True, bias=biases[i], relu=not last,
                                          out_dtype=torch.float32 if (last and not x.out_bf16) else torch.bfloat16))
            if not keep:
                return acts[-1]
            ctx.save_for_backward(*acts[:-1], *wb)
            return acts[-1]
        M = x.shape[0]
        if precision == "fp32":
            h = x if x.dtype == torch.float32 else x.float()
            acts = [h.contiguous()]
            for i in range(n):
                acts.append(ops.linear_fwd_f32(acts[-1], weights[i], biases[i], relu=i < n - 1))
            ctx.save_for_backward(*acts[:-1], *weights)
            return acts[-1]
        # bf16: activations and weight copies in bf16, fp32 accumulation in TMEM, fp32 final output
        h = x if x.dtype == torch.bfloat16 else ops.to_bf16(x.contiguous())
        wb = _bf16_copies(weights)
        acts = [h.contiguous()]
        for i in range(n):
            N, K = weights[i].shape
            last = i == n - 1
            acts.append(ops.gemm_bf16(acts[-1], wb[i], M, N, K, True, True, bias=biases[i], relu=not last,
                                      out_dtype=torch.float32 if last else torch.bfloat16))
        ctx.save_for_backward(*acts[:-1], *wb)
        return acts[-1]

    @staticmethod
    def backward(ctx, dy):
        n = ctx.n
        saved = ctx.saved_tensors
        acts, weights = saved[:n], saved[n:]
        grads = [None] * (2 * n)
        need_dx = ctx.needs_input_grad[0]
        M = dy.shape[0]
        if ctx.precision == "fp32":
            g = dy.contiguous().float()
            for i in reversed(range(n)):
                if i == 0 and n > 1 and EARLY_EXCHANGE_HOOK is not None:
                    grads = EARLY_EXCHANGE_HOOK(ctx.params, grads)
                grads[2 * i], grads[2 * i + 1] = ops.linear_bwd_weight_f32(g, acts[i])
                if i > 0:
                    g = ops.linear_bwd_data_f32(g, weights[i], relu_out=acts[i])
                elif need_dx:
                    g = ops.linear_bwd_data_f32(g, weights[i], relu_out=None)
            return (g if need_dx else None, None, *grads)
        g = dy.contiguous() if dy.dtype == torch.bfloat16 else ops.to_bf16(dy.contiguous().float())
        # The bias gradients (column sums: a second pass over dY, HBM-bound) run on a side stream NEXT TO the tensor-bound
        # weight-gradient GEMMs instead of between them; `keep` holds every dY they read until the join at the end.
        main = torch.cuda.current_stream()
        side = _bias_stream(dy.device)
        keep = []
        for i in reversed(range(n)):
            if i == 0 and n > 1 and EARLY_EXCHANGE_HOOK is not None:
                main.wait_stream(side)
                grads = EARLY_EXCHANGE_HOOK(ctx.params, grads)
            N, K = weights[i].shape
            side.wait_stream(main)
            with torch.cuda.stream(side):
                grads[2 * i + 1] = ops.colsum(g)
            keep.append(g)
            # dW[N,K] = g[M,N]^T acts_i[M,K]: both operands reduce over their ROW index -> MN-major
            grads[2 * i] = ops.gemm_bf16(g, acts[i], N, K, M, a_kmajor=False, b_kmajor=False, out_dtype=torch.float32,
                                         split_k=ops.gemm_split_k(N, K, M))
            if i > 0:    # dh[M,K] = g[M,N] W[N,K], masked by relu'(acts_i)
                g = ops.gemm_bf16(g, weights[i], M, K, N, a_kmajor=True, b_kmajor=False, mask=acts[i])
            elif need_dx:
                g = ops.to_f32(ops.gemm_bf16(g, weights[i], M, K, N, a_kmajor=True, b_kmajor=False))
        main.wait_stream(side)
        keep.clear()
        for i in range(n):
            grads[2 * i + 1].record_stream(main)      # allocated on the side stream, consumed on the main one
        return (g if need_dx else None, None, *grads)


_BIAS_STREAMS = {}


def _bias_stream(device) -> torch.cuda.Stream:
    key = str(device)
    if key not in _BIAS_STREAMS:
        _BIAS_STREAMS[key] = torch.cuda.Stream(device=device)
    return _BIAS_STREAMS[key]


class FeatureExtractor(nn.Module):
    """Per-frame MLP input_dim -> hidden_dims... -> output_dim, ReLU between layers
    (reference models.py:6-47; xavier-normal weights, every bias 0.1)."""

    def __init__(self, input_dim: int = 2048, output_dim: int = 32, hidden_dims: list = None, precision: str = "fp32"):
        super().__init__()
        self.precision = precision
        self.linear = nn.Sequential()
        dims = [input_dim] + list(hidden_dims)
        for i in range(len(hidden_dims)):
            self.linear.add_module(f"linear_{i}", nn.Linear(dims[i], dims[i + 1]))
            self.linear.add_module(f"relu_{i}", nn.ReLU())
        self.linear.add_module("output", nn.Linear(dims[-1], output_dim))
        self.initialize_weights()

    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0.1)

    def _params(self):
        out = []
        for m in self.linear:
            if isinstance(m, nn.Linear):
                out += [m.weight, m.bias]
        return out

    def forward_table(self, table, mean, std, starts, W: int, events=None, out_bf16: bool = False):
        """Features [B, W, output_dim] of the windows starting at rows ``starts`` of the fp32 ``table``, standardised with
        (mean, std): the gather runs inside the first layer's kernel (bf16 mode; see ops.gather_linear_supported).
        out_bf16: the features leave in bf16 (see TableWindows)."""
        y = _MLPFunction.apply(TableWindows(table, mean, std, starts, W, events, out_bf16), self.precision, *self._params())
        return y.reshape(starts.numel(), W, y.shape[-1])

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("b200med FeatureExtractor runs on CUDA tensors only (no CPU fallback)")
        lead = x.shape[:-1]
        y = _MLPFunction.apply(x.reshape(-1, x.shape[-1]), self.precision, *self._params())
        return y.reshape(*lead, y.shape[-1])


class CNN(nn.Module):
    """Window CNN head (reference models.py:49-131): 2 (W=10) or 3 (W=30) conv/pool/dropout/BN blocks,
    then Linear 256/32/16/C with ReLU + BN.  Like the reference, only W in {10, 30} defines layers."""

    def __init__(self, in_features: int = 58, window_size: int = 30, n_classes: int = 1):
        super().__init__()
        self.name = "SimpleCNN"
        self.window_size, self.in_features, self.n_classes = window_size, in_features, n_classes
        widths = {10: [64, 128], 30: [64, 128, 256]}.get(window_size)
        if widths is None:
            # the reference defines no conv stack for other window sizes and fails with this error (models.py:66-97)
            raise AttributeError("'CNN' object has no attribute 'convolutional_layers'")
        mods, cin, length = [], in_features, window_size
        for cout in widths:
            mods += [nn.Conv1d(cin, cout, kernel_size=3, stride=1), nn.MaxPool1d(2, 2), nn.Dropout(p=0.2),
                     nn.BatchNorm1d(cout)]
            cin, length = cout, (length - 2) // 2
        self.convolutional_layers = nn.Sequential(*mods, nn.Flatten())
        self.linear_layers = nn.Sequential(
            nn.Linear(cin * length, 256), nn.ReLU(), nn.BatchNorm1d(256),
            nn.Linear(256, 32), nn.ReLU(), nn.BatchNorm1d(32),
            nn.Linear(32, 16), nn.ReLU(), nn.BatchNorm1d(16),
            nn.Linear(16, n_classes))
        self.register_buffer("_drop_seed", torch.zeros(1, dtype=torch.int32), persistent=False)  # not in state_dict
        self.initialize_weights()

    def forward(self, x):
        """x [B, F, W] -> [B, n_classes] on the fused head kernels (heads.py): convolutions as GEMMs over overlapping
        time-major rows, pool + dropout, deterministic BatchNorm, Linear / ReLU / BatchNorm tail."""
        if not x.is_cuda:
            raise RuntimeError("b200med CNN head runs on CUDA tensors only (no CPU fallback)")
        from ..heads import conv_stack, mlp_tail
        if self.training:
            self._drop_seed.add_(1)
        feats = conv_stack(x, self.convolutional_layers, self.training, self._drop_seed)
        return mlp_tail(feats, self.linear_layers, relu_in=False, training=self.training)

    def initialize_weights(self):
        # reference quirk kept for weight parity: only the LAST module's bias becomes 0.1 (models.py:130-131)
        last = None
        for m in self.modules():
            if isinstance(m, nn.Conv1d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
            last = m
        if last.bias is not None:
            nn.init.constant_(last.bias, 0.1)


class LSTM(nn.Module):
    """Window LSTM head (reference models.py:135-220): [B,F,W] -> transpose -> nn.LSTM(F, H, layers,
    dropout .2) -> ReLU -> last step -> Linear 256/64/C with ReLU + BN.

    ``precision`` (set by ``define_model_objects`` from ``exp_kwargs['precision']``): "fp32" = exact-math fp32 recurrence
    kernels (1e-5 parity mode), "bf16" = persistent tcgen05 recurrence (2e-2 throughput mode).  The ``nn.LSTM`` member only
    holds the weights; it is never called."""

    def __init__(self, in_features: int = 58, window_size: int = 30, num_layers: int = 3, hidden_size: int = 128,
                 n_classes: int = 1):
        super().__init__()
        self.name = "SimpleLSTM"
        self.window_size, self.in_features = window_size, in_features
        self.layer_dim, self.hidden_size, self.n_classes = num_layers, hidden_size, n_classes
        self.lstm = nn.LSTM(input_size=in_features, hidden_size=hidden_size, num_layers=num_layers, batch_first=True,
                            dropout=0.2)
        self.precision = "fp32"
        self.register_buffer("_drop_seed", torch.zeros(1, dtype=torch.int32), persistent=False)  # not in state_dict
        self.linear_layers = nn.Sequential(
            nn.Flatten(), nn.Linear(hidden_size, 256), nn.ReLU(), nn.BatchNorm1d(256),
            nn.Linear(256, 64), nn.ReLU(), nn.BatchNorm1d(64), nn.Linear(64, n_classes))
        self.initialize_weights()

    def prepack(self, B: int, W: int):
        """Train step hook (engine.WindowTrainStep): start the weight-only preparation of the next forward on the side stream."""
        if self.precision == "bf16" and self.training:
            from .. import lstm_stack
            if lstm_stack.prepack(self.lstm, B, W, self.in_features, need_grad=True) and self.lstm.dropout > 0:
                # the dropout seed of this forward advances on the side stream too (joined before the first recurrence kernel)
                with torch.cuda.stream(lstm_stack._side_stream(self._drop_seed.device)):
                    self._drop_seed.add_(1)
                self._seed_advanced = True

    def accepts_parts(self) -> bool:
        """True when forward(feats, parts=WindowParts) is served: bf16 mode on the persistent recurrence (hidden size 128)."""
        return self.precision == "bf16" and self.hidden_size == 128

    def forward(self, l, parts=None):
        """l: the reference's head input [B, F, W]; or, with ``parts`` (lstm_stack.WindowParts: the window's kinematics still in
        the frame table), the FeatureExtractor output [B, W, Ca] -- the concatenation of define_inputs then happens inside the
        kernel that builds the recurrence's first operand."""
        if not l.is_cuda:
            raise RuntimeError("b200med LSTM head runs on CUDA tensors only (no CPU fallback)")
        from ..heads import mlp_tail
        from ..lstm_stack import lstm_last_hidden
        if self.training and self.lstm.dropout > 0:
            if getattr(self, "_seed_advanced", False):
                self._seed_advanced = False          # prepack() has advanced it already
            else:
                self._drop_seed.add_(1)
        # only h_{W-1} of the top layer is needed: the reference takes F.relu(out)[:, -1, :] (models.py:205-206)
        h = lstm_last_hidden(l, self.lstm, self.training, self._drop_seed, precision=self.precision, parts=parts)
        return mlp_tail(h, self.linear_layers, relu_in=True, training=self.training, precision=self.precision)

    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)
