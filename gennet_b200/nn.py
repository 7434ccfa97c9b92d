"""Host-side mirror of the Keras object protocol the reference scripts use.

Same class / function names, argument meaning and return shapes as Keras 2.2.4 so that
``bbhMahoGANy.py``-style code (``Sequential().add(Dense(...))``, functional ``Model(inputs, outputs)``,
``compile`` / ``train_on_batch`` / ``predict`` / ``save_weights``) runs unchanged, but every layer's
forward, backward and optimizer update is a hand-written sm_100a kernel reached through the C ABI
(``include/gennet_b200.h``).  There is no autograd and no CPU path: backward is explicit.

Layout: activations are channels-last float32 CUDA tensors (B, L, C); parameters of a compiled model
live in one flat arena (weights / gradients / Adam moments) so the optimizer step is one launch and a
data-parallel gradient all-reduce is one NCCL call.
"""
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream

# ----------------------------------------------------------------------------- session state
_NAME_COUNTS = {}
_STATE = {'seed': 1234, 'noise_counter': 0, 'init_rng': np.random.RandomState(1234), 'dp': None,
          'dtype': 'float32', 'wver': 0, 'bn_zero_debias': True, 'f32_chain': True}


def clear_session():
    _NAME_COUNTS.clear()


def set_seed(seed):
    """Seed for weight initialisation (host MT19937, as Keras) and the device Philox streams."""
    _STATE['seed'] = int(seed)
    _STATE['noise_counter'] = 0
    _STATE['init_rng'] = np.random.RandomState(int(seed))


def set_compute_dtype(name):
    """'float32': every GEMM on the float32 SIMT kernels (default).
    'bf16x3': float32 activations / weights / gradients as in 'float32', but every Conv1D whose channel counts tile
    runs on the tcgen05 tensor cores with its operands split into three bf16 planes (six plane products per K step
    accumulated in fp32 tensor memory: float32-class accuracy, the mode the parity tests and the benchmark headline
    use).  'bf16x2': two planes, three products (~2^-16 relative).
    'f16x2': as 'bf16x3' with the operands as scaled fp16 pairs (two fp16 planes of the tensor times a power of two taken
    from its largest element: 22-23 bits relative to the tensor's scale, three plane products per K step = half the
    tensor-core work of 'bf16x3').
    'bfloat16': activations and conv operands in bf16 with fp32 accumulation on the tcgen05 tensor cores, fp32
    master weights / gradients / optimizer (throughput mode, stated tolerance in tests/test_gpu_models.py)."""
    assert name in ('float32', 'bfloat16', 'bf16x3', 'bf16x2', 'f16x2')
    _STATE['dtype'] = name


def _split_planes():
    """Number of 16-bit planes of the split tensor-core mode (0 = not in that mode)."""
    return {'bf16x3': 3, 'bf16x2': 2, 'f16x2': 2}.get(_STATE['dtype'], 0)


def _f16s():
    """True when the planes are scaled fp16 pairs (tensor attribute _gn_amax = device scalar max |x|)."""
    return _STATE['dtype'] == 'f16x2'


def set_bn_zero_debias(flag):
    """True (default): BatchNormalization moving statistics follow Keras 2.2.4 on TensorFlow 1.12
    (assign_moving_average(zero_debias=True)); False: plain exponential moving average (tf.keras behaviour)."""
    _STATE['bn_zero_debias'] = bool(flag)


def compute_dtype():
    return _STATE['dtype']


def _uid(prefix):
    _NAME_COUNTS[prefix] = _NAME_COUNTS.get(prefix, 0) + 1
    return '%s_%d' % (prefix, _NAME_COUNTS[prefix])


def device():
    _lib.require_device()
    return torch.device('cuda', torch.cuda.current_device())


def _empty(shape):
    return torch.empty(shape, dtype=torch.float32, device=device())


BF16 = torch.bfloat16


def _empty_bf16(shape):
    return torch.empty(shape, dtype=BF16, device=device())


def _as_f32(x):
    """bf16 activations entering a layer that only has a float32 kernel are widened on the device."""
    if x is None or x.dtype == torch.float32:
        return x
    y = _empty(x.shape)
    call('gn_cast_bf16_to_f32', ptr(x.contiguous(), BF16), ptr(y), x.numel(), stream())
    return y


def _bias_sink(layer, ctx, features):
    """(pointer, channels) of the bias-gradient buffer of the tensor-core Conv1D feeding `layer`, when `layer`'s
    data-gradient kernel can produce it as the per-channel sum of its output (fused activation mask, the
    producer is trainable and ran on the tensor-core path); else (None, 0)."""
    src = layer.bias_src
    if src is None or layer.in_act is None or id(src) not in ctx.trainable_ids or \
            getattr(src, '_mode', None) not in ('tc', 'tc3'):
        return None, 0
    C = src.filters
    if C % 8 != 0 or features % C != 0:
        return None, 0
    return ptr(src.params[1].grad), C


def _as_bf16(x):
    if x.dtype == BF16:
        return x
    y = _empty_bf16(x.shape)
    call('gn_cast_f32_to_bf16', ptr(x.contiguous()), ptr(y, BF16), x.numel(), stream())
    return y


F16 = torch.float16


def _pdt():
    """dtype of the operand planes in the current split mode."""
    return F16 if _f16s() else BF16


def _empty_planes(shape):
    return torch.empty(shape, dtype=_pdt(), device=device())


def _split(x, nc):
    """float32 tensor -> (nc,) + shape 16-bit planes: bf16 planes whose sum is x to float32 accuracy, or (mode 'f16x2')
    the scaled fp16 pair with the device scalar max |x| attached as planes._gn_amax (taken from x._gn_amax when the kernel
    that produced x already accumulated it)."""
    planes = _empty_planes((nc,) + tuple(x.shape))
    if _f16s():
        amax = getattr(x, '_gn_amax', None)
        have = amax is not None
        if not have:
            amax = _empty((1,))
        call('gn_split_f32_f16x2', ptr(x), ptr(planes, F16), ptr(amax), 1 if have else 0, x.numel(), stream())
        planes._gn_amax = amax
    else:
        call('gn_split_f32_bf16', ptr(x), ptr(planes, BF16), x.numel(), nc, stream())
    return planes


def _split_grad(dy, nc, db):
    """Operand planes of a gradient tensor dy (..., C); in mode 'f16x2' the same pass also leaves the bias gradient
    db (C) = column sums of dy when `db` is given and C allows it.  Returns (planes, db_written)."""
    C = dy.shape[-1]
    c8 = C // 8
    if db is not None and _f16s() and C % 8 == 0 and (256 % c8 == 0 or c8 % 256 == 0):
        planes = _empty_planes((nc,) + tuple(dy.shape))
        amax = getattr(dy, '_gn_amax', None)
        have = amax is not None
        if not have:
            amax = _empty((1,))
        call('gn_split_colsum_f32_f16x2', ptr(dy), ptr(planes, F16), ptr(amax), 1 if have else 0, dy.numel() // C, C, ptr(db),
             stream())
        planes._gn_amax = amax
        return planes, True
    return _split(dy, nc), False


def _w_split(w, k, cin, cout, nc):
    """Conv weights f32 (k,cin,cout) -> (wk planes (nc,k,cin,cout), wt planes (nc,k,cout,cin))."""
    wk = _empty_planes((nc, k, cin, cout))
    wt = _empty_planes((nc, k, cout, cin))
    if _f16s():
        amax = _empty((1,))
        call('gn_conv_w_split_f16x2', ptr(w), ptr(wk, F16), ptr(wt, F16), ptr(amax), k, cin, cout, stream())
        wk._gn_amax = wt._gn_amax = amax
    else:
        call('gn_conv_w_split_bf16', ptr(w), ptr(wk, BF16), ptr(wt, BF16), k, cin, cout, nc, stream())
    return wk, wt


def _conv_fwd_split(xs, wt, bias, y, ys, want_amax, geom, code, par, nc, want_stats=False):
    """Forward convolution on the split-operand tensor-core kernels; geom = (B, L, Cin, Lout, Cout, k, stride, pad).
    want_amax ('f16x2'): the epilogue also accumulates max |y| for the consumer's split.  want_stats ('f16x2'): it also
    produces the per-channel (sum y, sum y^2) the BatchNormalization behind this layer needs (y._gn_bn_sums)."""
    if _f16s() and want_stats and geom[4] <= 1024:
        sums = torch.empty(2 * geom[4], dtype=torch.float64, device=y.device)
        call('gn_conv1d_fwd_stats_f16x2', ptr(xs, F16), ptr(xs._gn_amax), ptr(wt, F16), ptr(wt._gn_amax), bias, ptr(y),
             ptr(sums, torch.float64), *geom, code, par, stream())
        y._gn_bn_sums = sums
    elif _f16s():
        am = _empty((1,)) if want_amax else None
        call('gn_conv1d_fwd_f16x2', ptr(xs, F16), ptr(xs._gn_amax), ptr(wt, F16), ptr(wt._gn_amax), bias, ptr(y), ptr(am),
             *geom, code, par, stream())
        if am is not None:
            y._gn_amax = am
    else:
        call('gn_conv1d_fwd_bf16x3', ptr(xs, BF16), ptr(wt, BF16), bias, ptr(y), ptr(ys, BF16) if ys is not None else None,
             *geom, code, par, nc, stream())


def _conv_dgrad_split(dys, wk, xin, dx, dxs, sink, want_amax, geom, code, par, nc):
    if _f16s():
        am = _empty((1,)) if want_amax else None
        call('gn_conv1d_dgrad_f16x2', ptr(dys, F16), ptr(dys._gn_amax), ptr(wk, F16), ptr(wk._gn_amax), xin, ptr(dx), sink,
             ptr(am), *geom, code, par, stream())
        if am is not None:
            dx._gn_amax = am
    else:
        call('gn_conv1d_dgrad_bf16x3', ptr(dys, BF16), ptr(wk, BF16), xin, ptr(dx), ptr(dxs, BF16) if dxs is not None else None,
             sink, *geom, code, par, nc, stream())


def _conv_wgrad_split(xs, dys, dy, dw, db, geom, nc):
    if _f16s():
        call('gn_conv1d_wgrad_f16x2', ptr(xs, F16), ptr(xs._gn_amax), ptr(dys, F16), ptr(dys._gn_amax), ptr(dy), dw, db, *geom,
             stream())
    else:
        call('gn_conv1d_wgrad_bf16x3', ptr(xs, BF16), ptr(dys, BF16), ptr(dy), dw, db, *geom, nc, stream())


def _chain_fwd(sfx, x, y, dt, mean, scale, gamma, beta, use_var, eps, code, par, kind, rate, r, seed, off, rows, C,
               planes_bound=None):
    """gn_chain_fwd_*; in mode 'f16x2' the float32 form also accumulates max |y| (y._gn_amax) for the consumer's split,
    or -- planes_bound = an a-priori bound on |y| -- writes the consumer's operand planes itself (y._gn_planes)."""
    if dt == torch.float32 and _f16s() and planes_bound is not None:
        planes = _empty_planes((2,) + tuple(y.shape))
        planes._gn_amax = _empty((1,))
        call('gn_chain_fwd_planes_f32', ptr(x), ptr(y), mean, scale, gamma, beta, use_var, eps, code, par, kind, rate, r,
             seed, off, rows, C, ptr(planes, F16), ptr(planes._gn_amax), float(planes_bound), stream())
        y._gn_planes = planes
    elif dt == torch.float32 and _f16s():
        y._gn_amax = _empty((1,))
        call('gn_chain_fwd_amax_f32', ptr(x), ptr(y), mean, scale, gamma, beta, use_var, eps, code, par, kind, rate, r, seed,
             off, rows, C, ptr(y._gn_amax), stream())
    else:
        call('gn_chain_fwd' + sfx, ptr(x, dt), ptr(y, dt), mean, scale, gamma, beta, use_var, eps, code, par, kind, rate, r,
             seed, off, rows, C, stream())


def _chain_bwd(sfx, x, dy, dx, dt, mean, invstd, gamma, beta, sums, n_total, code, par, kind, rate, r, seed, off, dgamma,
               dbeta, rows, C):
    if dt == torch.float32 and _f16s():
        dx._gn_amax = _empty((1,))
        call('gn_chain_bwd_amax_f32', ptr(x), ptr(dy), ptr(dx), mean, invstd, gamma, beta, sums, n_total, code, par, kind,
             rate, r, seed, off, dgamma, dbeta, rows, C, ptr(dx._gn_amax), stream())
    else:
        call('gn_chain_bwd' + sfx, ptr(x, dt), ptr(dy, dt), ptr(dx, dt), mean, invstd, gamma, beta, sums, n_total, code, par,
             kind, rate, r, seed, off, dgamma, dbeta, rows, C, stream())


def _act_bwd(dy, y, code, param):
    """dy * act'(y) in the dtype the tensors are in."""
    if dy.dtype == BF16 and y.dtype == BF16:
        g = _empty_bf16(dy.shape)
        call('gn_act_bwd_bf16', ptr(dy, BF16), ptr(y, BF16), ptr(g, BF16), dy.numel(), code, param, stream())
        return g
    dy, y = _as_f32(dy), _as_f32(y)
    g = _empty(dy.shape)
    call('gn_act_bwd_f32', ptr(dy), ptr(y), ptr(g), dy.numel(), code, param, stream())
    return g


def _same_pad(L, k, s):
    out = -(-L // s)
    total = max((out - 1) * s + k - L, 0)
    return total // 2, out


_ACTS = {None: _lib.ACT_NONE, 'linear': _lib.ACT_NONE, 'relu': _lib.ACT_RELU, 'tanh': _lib.ACT_TANH,
         'sigmoid': _lib.ACT_SIGMOID, 'elu': _lib.ACT_ELU}


class Param:
    """A weight tensor + its gradient; both may be re-homed into a flat arena at compile()."""

    def __init__(self, name, value, trainable=True):
        self.name = name
        self.trainable = trainable
        self.data = torch.as_tensor(np.ascontiguousarray(value, dtype=np.float32)).to(device())
        self.grad = None
        self.arena = None
        self.offset = 0

    @property
    def shape(self):
        return tuple(self.data.shape)

    def numel(self):
        return self.data.numel()


class Ctx:
    """Per-call execution context."""

    def __init__(self, training, trainable_ids=None, noise=None):
        self.training = training
        self.trainable_ids = trainable_ids or set()
        self.noise = noise if noise is not None else {}
        dp = _STATE['dp']
        self.world = dp.world if dp is not None else 1
        self.dp = dp
        self.reg_loss = None            # device double: regularisation losses of this call (keras `model.losses`)

    def add_reg(self, x, g, l1, l2, with_loss=True):
        """keras.regularizers.l1/l2 terms of tensor x: loss accumulated into self.reg_loss, gradient added into g."""
        if with_loss and self.reg_loss is None:
            self.reg_loss = torch.zeros(1, dtype=torch.float64, device=x.device)
        call('gn_reg_terms_f32', ptr(x), ptr(g) if g is not None else None, x.numel(), float(l1), float(l2),
             ptr(self.reg_loss, torch.float64) if with_loss else None, stream())


class Regularizer:
    """keras.regularizers.L1L2 (weight_version/subtract_model.py:217): ``regularizers.l1(0.001)``, ``regularizers.l2(0.01)``."""

    def __init__(self, l1=0.0, l2=0.0):
        self.l1, self.l2 = float(l1), float(l2)

    def get_config(self):
        return {'l1': self.l1, 'l2': self.l2}


class regularizers:
    """Namespace mirroring ``from keras import regularizers``."""
    L1L2 = Regularizer

    @staticmethod
    def l1(l=0.01):
        return Regularizer(l1=l)

    @staticmethod
    def l2(l=0.01):
        return Regularizer(l2=l)

    @staticmethod
    def l1_l2(l1=0.01, l2=0.01):
        return Regularizer(l1=l1, l2=l2)


# ----------------------------------------------------------------------------- symbolic graph
class KTensor:
    def __init__(self, shape, node):
        self.shape = tuple(shape)          # without the batch axis
        self.node = node

    @property
    def _keras_shape(self):
        return (None,) + self.shape


class Node:
    def __init__(self, layer, inputs):
        self.layer = layer
        self.inputs = inputs               # list of Nodes


class Layer:
    prefix = 'layer'

    def __init__(self, name=None, input_shape=None, trainable=True, **kwargs):
        self.name = name
        self.trainable = trainable
        self._input_shape_arg = tuple(input_shape) if input_shape is not None else None
        self.params = []
        self.built = False
        self.input_shape = None
        self.output_shape = None

    # Keras-style properties
    @property
    def weights(self):
        return self.params

    @property
    def trainable_weights(self):
        return [p for p in self.params if p.trainable and self.trainable]

    def get_weights(self):
        return [p.data.detach().cpu().numpy().copy() for p in self.params]

    def set_weights(self, ws):
        assert len(ws) == len(self.params), '%s expects %d arrays' % (self.name, len(self.params))
        for p, w in zip(self.params, ws):
            w = np.asarray(w, dtype=np.float32)
            assert tuple(w.shape) == p.shape, 'shape mismatch for %s: %s vs %s' % (p.name, w.shape, p.shape)
            p.data.copy_(torch.from_numpy(np.ascontiguousarray(w)))
        _STATE['wver'] += 1

    def count_params(self):
        return sum(p.numel() for p in self.params)

    def _ensure_built(self, in_shape):
        if not self.built:
            if self.name is None:
                self.name = _uid(self.prefix)
            self.input_shape = tuple(in_shape)
            self.output_shape = tuple(self.build(tuple(in_shape)))
            self.built = True
        return self.output_shape

    def build(self, in_shape):
        return in_shape

    def __call__(self, x):
        out_shape = self._ensure_built(x.shape)
        return KTensor(out_shape, Node(self, [x.node]))

    def all_layers(self):
        return [self]

    def forward(self, x, ctx):
        raise NotImplementedError

    def backward(self, dy, ctx, need_dx=True):
        raise NotImplementedError


class InputLayer(Layer):
    prefix = 'input'

    def __init__(self, shape, name=None):
        super().__init__(name=name)
        self._ensure_built(tuple(shape))


def Input(shape=None, name=None, **kwargs):
    layer = InputLayer(tuple(shape), name=name)
    return KTensor(tuple(shape), Node(layer, []))


def _glorot(shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return _STATE['init_rng'].uniform(-lim, lim, size=shape).astype(np.float32)


class Dense(Layer):
    """Keras Dense: kernel (in, out), bias (out); y = act(x @ kernel + bias)."""
    prefix = 'dense'

    def __init__(self, units, activation=None, kernel_initializer='glorot_uniform', **kw):
        super().__init__(**kw)
        self.units = int(units)
        self.activation = activation
        self.in_act = None           # (code, param) of the fused activation that produced our input (set by _fuse)
        self.bias_src = None         # the Conv1D that produced our input when in_act is fused (set by _fuse)

    def build(self, in_shape):
        # Keras Dense acts on the last axis; leading axes (no_weight_code/subtract_model.py:262: Dense on (1, noise_dim))
        # are folded into the rows of the GEMM
        fin = in_shape[-1]
        self._lead = tuple(in_shape[:-1])
        self.params = [Param(self.name + '/kernel:0', _glorot((fin, self.units), fin, self.units)),
                       Param(self.name + '/bias:0', np.zeros(self.units, np.float32))]
        return self._lead + (self.units,)

    def _tc_planes(self):
        """Number of bf16 planes when this layer runs on the tensor-core kernels (0 = it does not): any mode but
        'float32', N a multiple of 64, a GEMM large enough to be worth two operand conversions."""
        nc = _split_planes() or (1 if _STATE['dtype'] == 'bfloat16' else 0)
        K, N = self.input_shape[-1], self.units
        if nc == 0 or N % 64 != 0 or K < 32 or K * N < (1 << 20):
            return 0
        return nc

    def _kp(self):
        K, N = self.input_shape[-1], self.units
        if K <= 64 and N % 128 == 0:
            return 64
        return -(-K // 128) * 128

    def _tc_weights(self, nc):
        key = (_STATE['wver'], nc)
        if getattr(self, '_wsplit', None) is None or self._wsplit[0] != key:
            K, N, Kp = self.input_shape[-1], self.units, self._kp()
            wk = _empty_planes((nc, Kp, N)) if nc > 1 else _empty_bf16((nc, Kp, N))
            wt = _empty_planes((nc, N, Kp)) if nc > 1 else _empty_bf16((nc, N, Kp))
            if nc > 1 and _f16s():
                amax = _empty((1,))
                call('gn_dense_w_split_f16x2', ptr(self.params[0].data), ptr(wk, F16), ptr(wt, F16), ptr(amax), K, Kp, N,
                     stream())
                wk._gn_amax = wt._gn_amax = amax
            else:
                call('gn_dense_w_split_bf16', ptr(self.params[0].data), ptr(wk, BF16), ptr(wt, BF16), K, Kp, N, nc, stream())
            self._wsplit = (key, wk, wt)
        return self._wsplit[1], self._wsplit[2]

    def _forward_tc(self, x, nc):
        B, K = x.shape
        N, Kp = self.units, self._kp()
        x = _as_f32(x).contiguous()
        wk, wt = self._tc_weights(nc)
        y = _empty((B, N))
        if nc > 1 and _f16s():
            xs = _empty_planes((nc, B, Kp))
            xs._gn_amax = _empty((1,))
            call('gn_split_pad_f32_f16x2', ptr(x), ptr(xs, F16), ptr(xs._gn_amax), B, K, Kp, stream())
            call('gn_dense_fwd_f16x2', ptr(xs, F16), ptr(xs._gn_amax), ptr(wt, F16), ptr(wt._gn_amax), ptr(self.params[1].data),
                 ptr(y), B, Kp, N, _ACTS[self.activation], 0.0, stream())
        else:
            xs = _empty_bf16((nc, B, Kp))
            call('gn_split_pad_f32_bf16', ptr(x), ptr(xs, BF16), B, K, Kp, nc, stream())
            call('gn_dense_fwd_bf16x3', ptr(xs, BF16), ptr(wt, BF16), ptr(self.params[1].data), ptr(y), None, B, Kp, N,
                 _ACTS[self.activation], 0.0, nc, stream())
        self._xs = xs
        return x, y

    def _backward_tc(self, dy, ctx, need_dx, nc):
        x, xs = self._x, self._xs
        B, K = x.shape
        N, Kp = self.units, self._kp()
        dy = _as_f32(dy).contiguous()
        f16 = nc > 1 and _f16s()
        trn = id(self) in ctx.trainable_ids
        dys, wrote = _split_grad(dy, nc, self.params[1].grad if (f16 and trn) else None)
        if trn:
            if f16:
                call('gn_dense_wgrad_f16x2', ptr(xs, F16), ptr(xs._gn_amax), ptr(dys, F16), ptr(dys._gn_amax), ptr(dy),
                     ptr(self.params[0].grad), None if wrote else ptr(self.params[1].grad), B, K, N, Kp, stream())
            else:
                call('gn_dense_wgrad_bf16x3', ptr(xs, BF16), ptr(dys, BF16), ptr(dy), ptr(self.params[0].grad),
                     ptr(self.params[1].grad), B, K, N, Kp, nc, stream())
        dx = None
        if need_dx:
            dx = _empty((B, K))
            if K == Kp:
                wk, wt = self._tc_weights(nc)
                code, par = self.in_act if self.in_act is not None else (_lib.ACT_NONE, 0.0)
                sink, _ = _bias_sink(self, ctx, K)
                if f16:
                    call('gn_dense_dgrad_f16x2', ptr(dys, F16), ptr(dys._gn_amax), ptr(wk, F16), ptr(wk._gn_amax),
                         ptr(x) if self.in_act is not None else None, ptr(dx), sink if self.in_act is not None else None,
                         B, K, N, code, par, stream())
                else:
                    call('gn_dense_dgrad_bf16x3', ptr(dys, BF16), ptr(wk, BF16), ptr(x) if self.in_act is not None else None,
                         ptr(dx), sink if self.in_act is not None else None, B, K, N, code, par, nc, stream())
                dx._gn_preact = self.in_act is not None
                dx._gn_db_done = self.in_act is not None and sink is not None
            else:
                call('gn_dense_dgrad_f32', ptr(dy), ptr(self.params[0].data), ptr(dx), B, K, self.units, stream())
        self._xs = None
        return dx

    def forward(self, x, ctx):
        if self._lead:
            return self._forward2d(x.reshape(-1, x.shape[-1]), ctx).reshape((x.shape[0],) + self._lead + (self.units,))
        return self._forward2d(x, ctx)

    def backward(self, dy, ctx, need_dx=True):
        if self._lead:
            dx = self._backward2d(dy.reshape(-1, self.units), ctx, need_dx)
            return None if dx is None else dx.reshape((dy.shape[0],) + self._lead + (dx.shape[-1],))
        return self._backward2d(dy, ctx, need_dx)

    def _forward2d(self, x, ctx):
        B, K = x.shape
        self._tc = self._tc_planes()
        if self._tc:
            self._bf16 = self._small32 = False
            self._x, self._y = self._forward_tc(x, self._tc)
            return self._y
        y = _empty((B, self.units))
        # <= 4 outputs over a long feature vector: the streaming GEMV kernels (bf16 or float32 features)
        self._bf16 = x.dtype == BF16 and self.units <= 4 and K % 8 == 0
        self._small32 = x.dtype == torch.float32 and self.units <= 4 and K % 8 == 0 and K >= 1024
        if self._bf16:
            call('gn_dense_small_fwd_bf16', ptr(x, BF16), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y), B,
                 K, self.units, _ACTS[self.activation], 0.0, stream())
        elif self._small32:
            x = x.contiguous()
            call('gn_dense_small_fwd_f32', ptr(x), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y), B,
                 K, self.units, _ACTS[self.activation], 0.0, stream())
        else:
            x = _as_f32(x)
            call('gn_dense_fwd_f32', ptr(x), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y), B, K,
                 self.units, _ACTS[self.activation], 0.0, stream())
        self._x, self._y = x, y
        return y

    def _backward2d(self, dy, ctx, need_dx=True):
        x = self._x
        B, K = x.shape
        dy = _as_f32(dy)
        if _ACTS[self.activation] != _lib.ACT_NONE:
            dy = _act_bwd(dy, self._y, _ACTS[self.activation], 0.0)
        tr = id(self) in ctx.trainable_ids
        dx = None
        if self._tc:
            dx = self._backward_tc(dy, ctx, need_dx, self._tc)
            self._x = self._y = None
            return dx
        if self._bf16 or self._small32:
            sfx, dt = ('_bf16', BF16) if self._bf16 else ('_f32', torch.float32)
            if tr:
                call('gn_dense_small_wgrad' + sfx, ptr(x, dt), ptr(dy), ptr(self.params[0].grad),
                     ptr(self.params[1].grad), B, K, self.units, stream())
            if need_dx:
                dx = _empty_bf16((B, K)) if self._bf16 else _empty((B, K))
                code, par = self.in_act if self.in_act is not None else (_lib.ACT_NONE, 0.0)
                sink, C = _bias_sink(self, ctx, K)
                call('gn_dense_small_dgrad' + sfx, ptr(dy), ptr(self.params[0].data), ptr(x, dt), ptr(dx, dt), sink, C, B, K,
                     self.units, code, par, stream())
                dx._gn_preact = self.in_act is not None
                dx._gn_db_done = sink is not None
        else:
            if tr:
                call('gn_dense_wgrad_f32', ptr(x), ptr(dy), ptr(self.params[0].grad), ptr(self.params[1].grad), B, K,
                     self.units, stream())
            if need_dx:
                dx = _empty((B, K))
                call('gn_dense_dgrad_f32', ptr(dy), ptr(self.params[0].data), ptr(dx), B, K, self.units, stream())
        self._x = self._y = None
        return dx


class Conv1D(Layer):
    """Keras Conv1D (channels_last): kernel (k, Cin, Cout), bias (Cout), 'valid' | 'same' (TF rule).

    Execution paths (all CUDA): float32 SIMT implicit GEMM (exact parity); under
    ``set_compute_dtype('bfloat16')`` the tcgen05/TMA kernels when Cin and Cout are multiples of 64, and the
    bandwidth-bound first-layer kernel when Cin <= 2.  A directly following Activation/LeakyReLU/ReLU layer is
    folded into the epilogue (``post_act``); its backward is folded into the consumer's data-gradient epilogue
    when the consumer can do it (``in_act``), else applied here."""
    prefix = 'conv1d'

    def __init__(self, filters, kernel_size, strides=1, padding='valid', activation=None,
                 kernel_initializer='glorot_uniform', **kw):
        super().__init__(**kw)
        self.filters = int(filters)
        self.k = int(kernel_size[0] if isinstance(kernel_size, (tuple, list)) else kernel_size)
        self.s = int(strides[0] if isinstance(strides, (tuple, list)) else strides)
        self.padding = padding
        self.activation = activation
        self.fused_up = 1        # set to 2 by Model._fuse when an UpSampling1D(2) directly precedes
        self.post_act = None     # (code, param) of a following activation layer folded into the epilogue
        self.in_act = None       # (code, param) of the fused activation that produced our input
        self.bias_src = None     # the Conv1D that produced our input when in_act is fused (set by _fuse)
        self.plane_consumer = None   # the Conv1D that alone consumes our (activated) output (set by _fuse)
        self.bn_consumer = None      # the BatchNormalization(axis=-1) that alone consumes our output (set by _fuse)
        self._wcache = None
        self._wsplit = None

    def build(self, in_shape):
        L, cin = in_shape
        k = self.k
        self.params = [Param(self.name + '/kernel:0', _glorot((k, cin, self.filters), k * cin, k * self.filters)),
                       Param(self.name + '/bias:0', np.zeros(self.filters, np.float32))]
        if self.padding == 'same':
            self.pad, self.Lout = _same_pad(L, k, self.s)
        else:
            self.pad, self.Lout = 0, (L - k) // self.s + 1
        return (self.Lout, self.filters)

    def _act(self):
        if self.post_act is not None:
            return self.post_act
        return (_ACTS[self.activation], 0.0)

    def _path(self):
        L, cin = self.input_shape
        co = self.filters
        tiles = cin % 64 == 0 and co % 64 == 0 and (cin % 128 == 0 or (cin == 64 and co % 128 == 0))
        # bandwidth-bound edge layers (first convolution Cin <= 2, last convolution Cout = 1): streaming kernels
        edge_in = self.fused_up == 1 and cin <= 2 and (co in (8, 16, 32, 64) or (co % 128 == 0 and co <= 1024)) and self.k <= 5
        edge_out = self.fused_up == 1 and co == 1 and self.s == 1 and self.k <= 5 and cin % 8 == 0 and \
            self._act()[0] == _lib.ACT_NONE
        if _STATE['dtype'] != 'bfloat16':
            if _split_planes() and tiles and self.k <= 8 and self.s <= 2:
                return 'tc3'
            # float32 activations: the streaming first-layer kernels take any filter count and up to 16 taps
            # (Conv1D(50, 16) / Conv1D(25, 5) on a single input channel: 2_model_version, train_on_wvf_version)
            edge_in = edge_in or (self.fused_up == 1 and cin <= 2 and self.k <= 16 and co <= 1024 and
                                  (self.k * cin + 1) * co <= 12000)
            return 'smallcin32' if edge_in else ('cout1_32' if edge_out else 'f32')
        if self.k > 8 or self.s > 2:
            return 'f32'
        if tiles:
            return 'tc'          # a fused UpSampling1D(2) is materialised in bf16 first (cheap next to the GEMM)
        if self.fused_up != 1:
            return 'f32'
        if cin <= 2 and (co in (8, 16, 32, 64) or (co % 128 == 0 and co <= 1024)) and self.k <= 5:
            return 'smallcin'
        if co == 1 and self.s == 1 and self.k <= 5 and cin % 8 == 0 and self._act()[0] == _lib.ACT_NONE:
            return 'cout1'
        return 'f32'

    def _bf16_weights(self):
        if self._wcache is None or self._wcache[0] != _STATE['wver']:
            L, cin = self.input_shape
            wk = _empty_bf16((self.k, cin, self.filters))
            wt = _empty_bf16((self.k, self.filters, cin))
            call('gn_conv_w_to_bf16', ptr(self.params[0].data), ptr(wk, BF16), ptr(wt, BF16), self.k, cin, self.filters,
                 stream())
            self._wcache = (_STATE['wver'], wk, wt)
        return self._wcache[1], self._wcache[2]

    def _split_weights(self, nc):
        """(wk planes (nc,k,Cin,Cout), wt planes (nc,k,Cout,Cin)) of the split tensor-core mode, per weight version."""
        if self._wsplit is None or self._wsplit[0] != (_STATE['wver'], nc):
            L, cin = self.input_shape
            wk, wt = _w_split(self.params[0].data, self.k, cin, self.filters, nc)
            self._wsplit = ((_STATE['wver'], nc), wk, wt)
        return self._wsplit[1], self._wsplit[2]

    def _forward_tc3(self, x, ctx):
        """Split tensor-core mode: x float32 (with the planes the producer's epilogue already wrote, if any) ->
        y float32, plus y's planes when the only consumer is another split-mode convolution."""
        nc = _split_planes()
        B = x.shape[0]
        L, cin = self.input_shape
        code, par = self._act()
        xs = getattr(x, '_gn_planes', None)
        x = _as_f32(x).contiguous()
        if xs is None or xs.shape[0] != nc or xs.dtype != _pdt() or tuple(xs.shape[1:]) != tuple(x.shape):
            xs = _split(x, nc)
        if self.fused_up != 1:
            Lp = L // self.fused_up
            xu = _empty_planes((nc, B, L, cin))
            # a 16-bit copy: the planes of the repeated tensor are the repeated planes (and share the scale)
            call('gn_upsample1d_fwd_bf16', ptr(xs, None), ptr(xu, None), nc * B, Lp, cin, self.fused_up, stream())
            if _f16s():
                xu._gn_amax = xs._gn_amax
            xs = xu
        wk, wt = self._split_weights(nc)
        y = _empty((B, self.Lout, self.filters))
        cons = self.plane_consumer
        feeds_tc3 = cons is not None and cons.fused_up == 1 and cons._path() == 'tc3'
        ys = None
        if feeds_tc3 and not _f16s():
            ys = _empty_bf16((nc, B, self.Lout, self.filters))
        _conv_fwd_split(xs, wt, ptr(self.params[1].data), y, ys, feeds_tc3,
                        (B, L, cin, self.Lout, self.filters, self.k, self.s, self.pad), code, par, nc,
                        want_stats=ctx.training and self.bn_consumer is not None)
        if ys is not None:
            y._gn_planes = ys
        self._xs = xs
        return x, y

    def _backward_tc3(self, dy, ctx, need_dx, db_done):
        nc = _split_planes()
        x, xs = self._x, self._xs
        B = x.shape[0]
        L, cin = self.input_shape
        tr = id(self) in ctx.trainable_ids
        dys = getattr(dy, '_gn_planes', None)
        dy = _as_f32(dy).contiguous()
        if dys is None or dys.shape[0] != nc or dys.dtype != _pdt():
            # the split pass over dy also yields the bias gradient (column sums) where it can
            dys, wrote = _split_grad(dy, nc, self.params[1].grad if (tr and not db_done) else None)
            db_done = db_done or wrote
        geom = (B, L, cin, self.Lout, self.filters, self.k, self.s, self.pad)
        if tr:
            db = None if db_done else ptr(self.params[1].grad)
            _conv_wgrad_split(xs, dys, dy, ptr(self.params[0].grad), db, geom, nc)
        dx = None
        if need_dx:
            wk, wt = self._split_weights(nc)
            dx = _empty((B, L, cin))
            icode, ipar = self.in_act if self.in_act is not None else (_lib.ACT_NONE, 0.0)
            sink, _ = _bias_sink(self, ctx, cin)
            # the producer takes the planes of dx as they are when the activation mask was applied here
            emit = self.fused_up == 1 and self.in_act is not None and getattr(self.bias_src, '_mode', None) == 'tc3'
            dxs = _empty_bf16((nc, B, L, cin)) if (emit and not _f16s()) else None
            xin = x if (self.in_act is not None and self.fused_up == 1) else None
            _conv_dgrad_split(dys, wk, ptr(xin) if xin is not None else None, dx, dxs, sink if xin is not None else None,
                              emit, geom, icode if xin is not None else _lib.ACT_NONE, ipar, nc)
            if self.fused_up != 1:
                dxp = _empty((B, L // self.fused_up, cin))
                call('gn_upsample1d_bwd_f32', ptr(dx), ptr(dxp), B, L // self.fused_up, cin, self.fused_up, stream())
                dx = dxp
            else:
                dx._gn_preact = xin is not None
                dx._gn_db_done = xin is not None and sink is not None
                if dxs is not None:
                    dx._gn_planes = dxs
        self._x = self._y = self._xs = None
        return dx

    def forward(self, x, ctx):
        B = x.shape[0]
        L, cin = self.input_shape
        code, par = self._act()
        self._mode = self._path()
        if self._mode == 'tc3':
            x, y = self._forward_tc3(x, ctx)
        elif self._mode == 'tc':
            x = _as_bf16(x)
            if self.fused_up != 1:
                xu = _empty_bf16((B, L, cin))
                call('gn_upsample1d_fwd_bf16', ptr(x.contiguous(), BF16), ptr(xu, BF16), B, L // self.fused_up, cin,
                     self.fused_up, stream())
                x = xu
            wk, wt = self._bf16_weights()
            y = _empty_bf16((B, self.Lout, self.filters))
            call('gn_conv1d_fwd_bf16', ptr(x, BF16), ptr(wt, BF16), ptr(self.params[1].data), ptr(y, BF16), B, L, cin,
                 self.Lout, self.filters, self.k, self.s, self.pad, code, par, stream())
        elif self._mode == 'cout1':
            x = _as_bf16(x).contiguous()
            y = _empty((B, self.Lout, 1))
            call('gn_conv1d_cout1_fwd_bf16', ptr(x, BF16), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y), B,
                 L, cin, self.Lout, self.k, self.pad, stream())
        elif self._mode == 'cout1_32':
            x = _as_f32(x).contiguous()
            y = _empty((B, self.Lout, 1))
            call('gn_conv1d_cout1_fwd_f32', ptr(x), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y), B,
                 L, cin, self.Lout, self.k, self.pad, stream())
        elif self._mode == 'smallcin32':
            x = _as_f32(x).contiguous()
            y = _empty((B, self.Lout, self.filters))
            call('gn_conv1d_smallcin_fwd_f32', ptr(x), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y),
                 B, L, cin, self.Lout, self.filters, self.k, self.s, self.pad, code, par, stream())
        elif self._mode == 'smallcin':
            x = _as_f32(x)
            y = _empty_bf16((B, self.Lout, self.filters))
            call('gn_conv1d_smallcin_fwd_bf16', ptr(x), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y, BF16),
                 B, L, cin, self.Lout, self.filters, self.k, self.s, self.pad, code, par, stream())
        else:
            x = _as_f32(x)
            y = _empty((B, self.Lout, self.filters))
            call('gn_conv1d_fwd_f32', ptr(x), ptr(self.params[0].data), ptr(self.params[1].data), ptr(y), B, L, cin,
                 self.Lout, self.filters, self.k, self.s, self.pad, self.fused_up, code, par, stream())
        self._x, self._y = x, y
        return y

    def backward(self, dy, ctx, need_dx=True):
        x = self._x
        B = x.shape[0]
        L, cin = self.input_shape
        code, par = self._act()
        db_done = getattr(dy, '_gn_db_done', False)
        if code != _lib.ACT_NONE and not getattr(dy, '_gn_preact', False):
            dy = _act_bwd(dy.contiguous(), self._y, code, par)
        tr = id(self) in ctx.trainable_ids
        dx = None
        if self._mode == 'tc3':
            return self._backward_tc3(dy, ctx, need_dx, db_done)
        if self._mode == 'tc':
            dy = _as_bf16(dy.contiguous())
            if tr:
                # the bias gradient may already have been taken by the consumer's data-gradient epilogue
                db = None if db_done else ptr(self.params[1].grad)
                call('gn_conv1d_wgrad_bf16', ptr(x, BF16), ptr(dy, BF16), ptr(self.params[0].grad), db, B, L, cin,
                     self.Lout, self.filters, self.k, self.s, self.pad, stream())
            if need_dx:
                wk, wt = self._bf16_weights()
                dx = _empty_bf16(x.shape)
                icode, ipar = self.in_act if self.in_act is not None else (_lib.ACT_NONE, 0.0)
                sink, _ = _bias_sink(self, ctx, cin)
                call('gn_conv1d_dgrad_bf16', ptr(dy, BF16), ptr(wk, BF16), ptr(x, BF16), ptr(dx, BF16), sink, B, L, cin,
                     self.Lout, self.filters, self.k, self.s, self.pad, icode, ipar, stream())
                if self.fused_up != 1:
                    dxp = _empty_bf16((B, L // self.fused_up, cin))
                    call('gn_upsample1d_bwd_bf16', ptr(dx, BF16), ptr(dxp, BF16), B, L // self.fused_up, cin,
                         self.fused_up, stream())
                    dx = dxp
                dx._gn_preact = self.in_act is not None
                dx._gn_db_done = sink is not None
        elif self._mode in ('cout1', 'cout1_32'):
            sfx, dt = ('_bf16', BF16) if self._mode == 'cout1' else ('_f32', torch.float32)
            dy = _as_f32(dy).contiguous()
            if tr:
                call('gn_conv1d_cout1_wgrad' + sfx, ptr(x, dt), ptr(dy), ptr(self.params[0].grad),
                     ptr(self.params[1].grad), B, L, cin, self.Lout, self.k, self.pad, stream())
            if need_dx:
                dx = _empty_bf16(x.shape) if self._mode == 'cout1' else _empty(x.shape)
                call('gn_conv1d_cout1_dgrad' + sfx, ptr(dy), ptr(self.params[0].data), ptr(dx, dt), B, L, cin, self.Lout,
                     self.k, self.pad, stream())
        elif self._mode == 'smallcin32':
            dy = _as_f32(dy.contiguous())
            if tr:
                call('gn_conv1d_smallcin_wgrad_f32', ptr(x), ptr(dy), ptr(self.params[0].grad),
                     ptr(self.params[1].grad), B, L, cin, self.Lout, self.filters, self.k, self.s, self.pad, stream())
            if need_dx:
                dx = _empty(x.shape)
                call('gn_conv1d_smallcin_dgrad_f32', ptr(dy), ptr(self.params[0].data), ptr(dx), B, L, cin,
                     self.Lout, self.filters, self.k, self.s, self.pad, stream())
        elif self._mode == 'smallcin':
            dy = _as_bf16(dy.contiguous())
            if tr:
                call('gn_conv1d_smallcin_wgrad_bf16', ptr(x), ptr(dy, BF16), ptr(self.params[0].grad),
                     ptr(self.params[1].grad), B, L, cin, self.Lout, self.filters, self.k, self.s, self.pad, stream())
            if need_dx:
                dx = _empty(x.shape)
                call('gn_conv1d_smallcin_dgrad_bf16', ptr(dy, BF16), ptr(self.params[0].data), ptr(dx), B, L, cin,
                     self.Lout, self.filters, self.k, self.s, self.pad, stream())
        else:
            dy = _as_f32(dy.contiguous())
            if tr:
                call('gn_conv1d_wgrad_f32', ptr(x), ptr(dy), ptr(self.params[0].grad), ptr(self.params[1].grad), B, L,
                     cin, self.Lout, self.filters, self.k, self.s, self.pad, self.fused_up, stream())
            if need_dx:
                dx = _empty(x.shape)
                call('gn_conv1d_dgrad_f32', ptr(dy), ptr(self.params[0].data), ptr(dx), B, L, cin, self.Lout,
                     self.filters, self.k, self.s, self.pad, self.fused_up, stream())
        self._x = self._y = None
        return dx


class Conv2D(Layer):
    """Keras Conv2D restricted to the reference's discriminator shape (bbhMahoGANy.py:439,447): input
    (L, 2, C), kernel (kh, kw) with 'same' padding and strides (s, 1).  Because the width is 2, it is
    executed as a Conv1D over L with 2C input and 2*filters output channels (only min(kw,3) of the kw
    kernel columns ever touch data); the Keras kernel layout (kh, kw, Cin, Cout) is kept externally."""
    prefix = 'conv2d'

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', kernel_initializer='glorot_uniform',
                 **kw):
        super().__init__(**kw)
        self.filters = int(filters)
        self.kh, self.kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.sh, self.sw = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.padding = padding
        self.post_act = None     # (code, param) of a following activation layer folded into the epilogue
        self._wcache = None
        self._wsplit = None

    def build(self, in_shape):
        H, W, cin = in_shape
        if not (W == 2 and self.padding == 'same' and self.sw == 1):
            raise NotImplementedError('Conv2D is implemented for the reference discriminator geometry only '
                                      '(width 2, padding same, width stride 1); got input %s' % (in_shape,))
        kh, kw = self.kh, self.kw
        self.params = [Param(self.name + '/kernel:0', _glorot((kh, kw, cin, self.filters), kh * kw * cin,
                                                              kh * kw * self.filters)),
                       Param(self.name + '/bias:0', np.zeros(self.filters, np.float32))]
        self.pad, self.Lout = _same_pad(H, kh, self.sh)
        self.pw = max(kw - 1, 0) // 2      # TF same: total = kw-1 (stride 1), left = total//2
        return (self.Lout, 2, self.filters)

    def _pack(self):
        H, W, cin = self.input_shape
        w1 = _empty((self.kh, 2 * cin, 2 * self.filters))
        b1 = _empty((2 * self.filters,))
        call('gn_conv2d_w2_pack_f32', ptr(self.params[0].data), ptr(self.params[1].data), ptr(w1), ptr(b1), self.kh,
             self.kw, cin, self.filters, self.pw, stream())
        return w1, b1

    def _act(self):
        return self.post_act if self.post_act is not None else (_lib.ACT_NONE, 0.0)

    def _path(self):
        """'tc': the packed (L, 2*Cin) -> (Lout, 2*Cout) convolution on the tcgen05 kernels (bf16 mode, channel counts
        that tile); 'smallcin': first discriminator layer (2*Cin = 2) on the streaming kernels; else float32 SIMT."""
        H, W, cin = self.input_shape
        c1, c2 = 2 * cin, 2 * self.filters
        tiles = c1 % 64 == 0 and c2 % 64 == 0 and (c1 % 128 == 0 or (c1 == 64 and c2 % 128 == 0))
        edge_in = c1 == 2 and c2 % 128 == 0 and c2 <= 1024 and self.kh <= 5
        if _STATE['dtype'] != 'bfloat16':
            if _split_planes() and tiles and self.kh <= 8 and self.sh <= 2:
                return 'tc3'
            return 'smallcin32' if edge_in else 'f32'
        if self.kh > 8 or self.sh > 2:
            return 'f32'
        if tiles:
            return 'tc'
        if c1 == 2 and c2 % 128 == 0 and c2 <= 1024 and self.kh <= 5:
            return 'smallcin'
        return 'f32'

    def _packed_bf16(self):
        """(w1 f32, b1 f32, wk bf16, wt bf16) of the packed convolution, rebuilt when any weight changed."""
        if self._wcache is None or self._wcache[0] != _STATE['wver']:
            H, W, cin = self.input_shape
            w1, b1 = self._pack()
            wk = _empty_bf16(tuple(w1.shape))
            wt = _empty_bf16((self.kh, 2 * self.filters, 2 * cin))
            call('gn_conv_w_to_bf16', ptr(w1), ptr(wk, BF16), ptr(wt, BF16), self.kh, 2 * cin, 2 * self.filters, stream())
            self._wcache = (_STATE['wver'], w1, b1, wk, wt)
        return self._wcache[1:]

    def _packed_split(self, nc):
        """(w1 f32, b1 f32, wk planes, wt planes) of the packed convolution in the split tensor-core mode."""
        if self._wsplit is None or self._wsplit[0] != (_STATE['wver'], nc):
            H, W, cin = self.input_shape
            w1, b1 = self._pack()
            wk, wt = _w_split(w1, self.kh, 2 * cin, 2 * self.filters, nc)
            self._wsplit = ((_STATE['wver'], nc), w1, b1, wk, wt)
        return self._wsplit[1:]

    def forward(self, x, ctx):
        B = x.shape[0]
        H, W, cin = self.input_shape
        c1, c2 = 2 * cin, 2 * self.filters
        code, par = self._act()
        self._mode = self._path()
        if self._mode == 'tc3':
            nc = _split_planes()
            x = _as_f32(x).contiguous()
            w1, b1, wk, wt = self._packed_split(nc)
            self._xs = _split(x, nc)
            y = _empty((B, self.Lout, 2, self.filters))
            _conv_fwd_split(self._xs, wt, ptr(b1), y, None, False, (B, H, c1, self.Lout, c2, self.kh, self.sh, self.pad),
                            code, par, nc)
        elif self._mode == 'tc':
            x = _as_bf16(x).contiguous()
            w1, b1, wk, wt = self._packed_bf16()
            y = _empty_bf16((B, self.Lout, 2, self.filters))
            call('gn_conv1d_fwd_bf16', ptr(x, BF16), ptr(wt, BF16), ptr(b1), ptr(y, BF16), B, H, c1, self.Lout, c2,
                 self.kh, self.sh, self.pad, code, par, stream())
        elif self._mode in ('smallcin', 'smallcin32'):
            x = _as_f32(x).contiguous()
            w1, b1 = self._pack()
            if self._mode == 'smallcin':
                y = _empty_bf16((B, self.Lout, 2, self.filters))
                call('gn_conv1d_smallcin_fwd_bf16', ptr(x), ptr(w1), ptr(b1), ptr(y, BF16), B, H, c1, self.Lout, c2, self.kh,
                     self.sh, self.pad, code, par, stream())
            else:
                y = _empty((B, self.Lout, 2, self.filters))
                call('gn_conv1d_smallcin_fwd_f32', ptr(x), ptr(w1), ptr(b1), ptr(y), B, H, c1, self.Lout, c2, self.kh,
                     self.sh, self.pad, code, par, stream())
        else:
            x = _as_f32(x).contiguous()
            w1, b1 = self._pack()
            y = _empty((B, self.Lout, 2, self.filters))
            call('gn_conv1d_fwd_f32', ptr(x), ptr(w1), ptr(b1), ptr(y), B, H, c1, self.Lout, c2, self.kh, self.sh,
                 self.pad, 1, code, par, stream())
        self._x, self._y, self._w1 = x, y, w1
        return y

    def backward(self, dy, ctx, need_dx=True):
        x, w1 = self._x, self._w1
        B = x.shape[0]
        H, W, cin = self.input_shape
        c1, c2 = 2 * cin, 2 * self.filters
        code, par = self._act()
        if code != _lib.ACT_NONE and not getattr(dy, '_gn_preact', False):
            dy = _act_bwd(dy.contiguous(), self._y, code, par)
        tr = id(self) in ctx.trainable_ids
        dw1 = db1 = None
        if tr:
            dw1 = _empty(w1.shape)
            db1 = _empty((c2,))
        dx = None
        if self._mode == 'tc3':
            nc = _split_planes()
            dy = _as_f32(dy).contiguous()
            d2 = dy.reshape(B, self.Lout, c2)
            d2._gn_amax = getattr(dy, '_gn_amax', None)
            dys, wrote = _split_grad(d2, nc, db1 if tr else None)
            geom = (B, H, c1, self.Lout, c2, self.kh, self.sh, self.pad)
            if tr:
                _conv_wgrad_split(self._xs, dys, dy, ptr(dw1), None if wrote else ptr(db1), geom, nc)
            if need_dx:
                wk = self._packed_split(nc)[2]
                dx = _empty(x.shape)
                _conv_dgrad_split(dys, wk, None, dx, None, None, False, geom, _lib.ACT_NONE, 0.0, nc)
            self._xs = None
        elif self._mode == 'tc':
            dy = _as_bf16(dy.contiguous())
            if tr:
                call('gn_conv1d_wgrad_bf16', ptr(x, BF16), ptr(dy, BF16), ptr(dw1), ptr(db1), B, H, c1, self.Lout, c2,
                     self.kh, self.sh, self.pad, stream())
            if need_dx:
                wk = self._packed_bf16()[2]
                dx = _empty_bf16(x.shape)
                call('gn_conv1d_dgrad_bf16', ptr(dy, BF16), ptr(wk, BF16), None, ptr(dx, BF16), None, B, H, c1, self.Lout,
                     c2, self.kh, self.sh, self.pad, _lib.ACT_NONE, 0.0, stream())
        elif self._mode in ('smallcin', 'smallcin32'):
            sfx, dt = ('_bf16', BF16) if self._mode == 'smallcin' else ('_f32', torch.float32)
            dy = _as_bf16(dy.contiguous()) if self._mode == 'smallcin' else _as_f32(dy.contiguous())
            if tr:
                call('gn_conv1d_smallcin_wgrad' + sfx, ptr(x), ptr(dy, dt), ptr(dw1), ptr(db1), B, H, c1, self.Lout, c2,
                     self.kh, self.sh, self.pad, stream())
            if need_dx:
                dx = _empty(x.shape)
                call('gn_conv1d_smallcin_dgrad' + sfx, ptr(dy, dt), ptr(w1), ptr(dx), B, H, c1, self.Lout, c2, self.kh,
                     self.sh, self.pad, stream())
        else:
            dy = _as_f32(dy.contiguous())
            if tr:
                call('gn_conv1d_wgrad_f32', ptr(x), ptr(dy), ptr(dw1), ptr(db1), B, H, c1, self.Lout, c2, self.kh,
                     self.sh, self.pad, 1, stream())
            if need_dx:
                dx = _empty(x.shape)
                call('gn_conv1d_dgrad_f32', ptr(dy), ptr(w1), ptr(dx), B, H, c1, self.Lout, c2, self.kh, self.sh,
                     self.pad, 1, stream())
        if tr:
            call('gn_conv2d_w2_unpack_f32', ptr(dw1), ptr(db1), ptr(self.params[0].grad), ptr(self.params[1].grad),
                 self.kh, self.kw, cin, self.filters, self.pw, stream())
        self._x = self._y = self._w1 = None
        return dx


class Conv2DTranspose(Layer):
    """Keras Conv2DTranspose (channels_last), kernel (kh, kw, Cout, Cin), for the geometry the reference uses
    (2_model_version/*/no_mode_collapse_network.py:79-90): kh = 1, strides (1, 1), padding 'valid'.  Every image
    row is a full 1-D convolution along W, so the layer runs on the Conv1D kernels (see gn_flip_transpose_f32)."""
    prefix = 'conv2d_transpose'

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', activation=None,
                 kernel_initializer='glorot_uniform', dilation_rate=(1, 1), kernel_regularizer=None,
                 activity_regularizer=None, **kw):
        super().__init__(**kw)
        self.filters = int(filters)
        self.kh, self.kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.sh, self.sw = (strides, strides) if isinstance(strides, int) else tuple(strides)
        self.padding = padding
        self.activation = activation
        self.dilation_rate = (dilation_rate, dilation_rate) if isinstance(dilation_rate, int) else tuple(dilation_rate)
        self.kernel_regularizer, self.activity_regularizer = kernel_regularizer, activity_regularizer

    def build(self, in_shape):
        H, W, cin = in_shape
        if self.dilation_rate != (1, 1):
            # Keras 2.2.4, keras/backend/tensorflow_backend.py conv2d_transpose: a dilated transposed convolution goes to
            # tf.nn.atrous_conv2d_transpose after `assert dilation_rate[0] == dilation_rate[1]`; the generator of
            # 2_model_version/no_weight_code/subtract_model.py:278-297 asks for (1, 9), (1, 7), (1, 2), (1, 3) and fails there
            assert self.dilation_rate[0] == self.dilation_rate[1], \
                'Conv2DTranspose: dilation_rate %s is rejected by Keras 2.2.4 (rates must be equal)' % (self.dilation_rate,)
            raise NotImplementedError('dilated Conv2DTranspose is not used by any network of the reference that builds')
        if not (self.kh == 1 and self.sh == 1 and self.sw == 1 and self.padding == 'valid'):
            raise NotImplementedError('Conv2DTranspose is implemented for the reference generator geometry only '
                                      '(kernel (1, kw), strides (1, 1), padding valid)')
        kw = self.kw
        # Keras fan computation for a transposed kernel (kh, kw, out, in): fan_in = kh*kw*in ... via _compute_fans on
        # the stored shape: receptive field * shape[-2] and * shape[-1]
        self.params = [Param(self.name + '/kernel:0', _glorot((1, kw, self.filters, cin), kw * self.filters, kw * cin)),
                       Param(self.name + '/bias:0', np.zeros(self.filters, np.float32))]
        self.Wout = W + kw - 1
        return (H, self.Wout, self.filters)

    def _w1(self):
        H, W, cin = self.input_shape
        w1 = _empty((self.kw, cin, self.filters))
        call('gn_flip_transpose_f32', ptr(self.params[0].data), ptr(w1), self.kw, cin, self.filters, stream())
        return w1

    def forward(self, x, ctx):
        x = _as_f32(x).contiguous()
        B = x.shape[0]
        H, W, cin = self.input_shape
        w1 = self._w1()
        y = _empty((B, H, self.Wout, self.filters))
        code = _ACTS[self.activation]
        call('gn_conv1d_fwd_f32', ptr(x), ptr(w1), ptr(self.params[1].data), ptr(y), B * H, W, cin, self.Wout,
             self.filters, self.kw, 1, self.kw - 1, 1, code, 0.0, stream())
        if ctx.training:
            # regularisation losses enter the reported loss (keras `model.losses`); a kernel penalty is the same on every
            # rank of a data-parallel run, an activity penalty sums over the ranks' batches
            if self.activity_regularizer is not None:
                ctx.add_reg(y, None, self.activity_regularizer.l1, self.activity_regularizer.l2)
            if self.kernel_regularizer is not None:
                r = self.kernel_regularizer
                ctx.add_reg(self.params[0].data, None, r.l1 / ctx.world, r.l2 / ctx.world)
        self._x, self._y, self._w = x, y, w1
        return y

    def backward(self, dy, ctx, need_dx=True):
        x, w1 = self._x, self._w
        B = x.shape[0]
        H, W, cin = self.input_shape
        dy = _as_f32(dy).contiguous()
        code = _ACTS[self.activation]
        if self.activity_regularizer is not None:
            dy = dy.clone()          # d(l1 sum|y| + l2 sum y^2)/dy joins the incoming gradient of the layer OUTPUT
            ctx.add_reg(self._y, dy, self.activity_regularizer.l1, self.activity_regularizer.l2, with_loss=False)
        if code != _lib.ACT_NONE:
            dy = _act_bwd(dy, self._y, code, 0.0)
        if id(self) in ctx.trainable_ids:
            dw1 = _empty(w1.shape)
            call('gn_conv1d_wgrad_f32', ptr(x), ptr(dy), ptr(dw1), ptr(self.params[1].grad), B * H, W, cin, self.Wout,
                 self.filters, self.kw, 1, self.kw - 1, 1, stream())
            call('gn_flip_transpose_f32', ptr(dw1), ptr(self.params[0].grad), self.kw, self.filters, cin, stream())
            if self.kernel_regularizer is not None:
                r = self.kernel_regularizer
                ctx.add_reg(self.params[0].data, self.params[0].grad, r.l1 / ctx.world, r.l2 / ctx.world, with_loss=False)
        dx = None
        if need_dx:
            dx = _empty(x.shape)
            call('gn_conv1d_dgrad_f32', ptr(dy), ptr(w1), ptr(dx), B * H, W, cin, self.Wout, self.filters, self.kw, 1,
                 self.kw - 1, 1, stream())
        self._x = self._y = self._w = None
        return dx


class BatchNormalization(Layer):
    """Keras 2.2.4 BatchNormalization(axis=-1): gamma, beta, moving_mean, moving_variance."""
    prefix = 'batch_normalization'

    def __init__(self, momentum=0.99, epsilon=1e-3, axis=-1, **kw):
        super().__init__(**kw)
        assert axis in (-1, 1), 'BatchNormalization: axis -1 (channels) or 1 (no_weight_code/subtract_model.py:264-357)'
        self.momentum, self.epsilon = float(momentum), float(epsilon)
        self.axis = axis

    def _mid(self):
        """True when the normalised axis is axis 1 of an input with more axes behind it: it is then moved innermost
        (gn_transpose_f32), normalised by the channels-last kernels and moved back."""
        return self.axis == 1 and len(self.input_shape) > 1

    def build(self, in_shape):
        c = in_shape[0] if (self.axis == 1 and len(in_shape) > 1) else in_shape[-1]
        n = self.name
        self.params = [Param(n + '/gamma:0', np.ones(c, np.float32)), Param(n + '/beta:0', np.zeros(c, np.float32)),
                       Param(n + '/moving_mean:0', np.zeros(c, np.float32), trainable=False),
                       Param(n + '/moving_variance:0', np.ones(c, np.float32), trainable=False)]
        return in_shape

    chain = (None, None)      # (activation layer, dropout layer) directly following, set by Model._fuse
    planes_consumer = None    # the Conv1D that alone consumes the chain's output, set by Model._fuse

    def _finalize(self, ssq, n_total, C, stats, ctx):
        """1/sqrt(var + eps) of the batch and the moving-statistics update of a training-mode call.  Keras 2.2.4 on
        TF 1.12 updates them through assign_moving_average(zero_debias=True) (K.moving_average_update): a
        zero-initialised shadow accumulator per statistic and moving = accumulator / (1 - momentum^step); the shadow
        state is not a layer weight (it is not saved, as in Keras).  set_bn_zero_debias(False) selects the plain
        exponential average.  A layer frozen when the running model was compiled does not update (Keras skips the
        updates of non-trainable layers)."""
        mm, mv = self.params[2].data, self.params[3].data
        if id(self) not in ctx.trainable_ids:
            call('gn_bn_finalize_f32', None, ptr(ssq, torch.float64), n_total, C, self.epsilon, self.momentum, ptr(stats),
                 None, None, 1, None, 1.0, stream())
            return
        biased, debias = None, 1.0
        if _STATE['bn_zero_debias']:
            if getattr(self, '_biased', None) is None:
                self._biased = torch.zeros(2 * C, dtype=torch.float32, device=mm.device)
                self._local_step = 0
            self._local_step += 1
            biased = self._biased
            debias = 1.0 / (1.0 - self.momentum ** self._local_step)
        call('gn_bn_finalize_f32', None, ptr(ssq, torch.float64), n_total, C, self.epsilon, self.momentum, ptr(stats),
             ptr(mm), ptr(mv), 1, ptr(biased) if biased is not None else None, debias, stream())

    def _chain_codes(self, ctx):
        act, noise = self.chain
        code, par = (act.code, act.param) if act is not None else (_lib.ACT_NONE, 0.0)
        kind, rate = (noise.kind, noise.rate) if (noise is not None and ctx.training) else (-1, 0.0)
        return code, par, kind, rate

    def _forward_chain(self, x, ctx):
        """Chain kernels (bf16 or float32 activations): statistics in one pass, then y = drop(act(bn(x))) in one pass
        (gn_chain_*); the activation and dropout layers of the chain become pass-throughs for this call."""
        C = x.shape[-1]
        rows = x.numel() // C
        dt = x.dtype
        sfx = '_bf16' if dt == BF16 else '_f32'
        g, b, mm, mv = [p.data for p in self.params]
        act, noise = self.chain
        for l in (act, noise):
            if l is not None:
                l._chain_skip = True
        code, par, kind, rate = self._chain_codes(ctx)
        y = torch.empty(x.shape, dtype=dt, device=x.device)
        # bounded activation feeding a split-operand convolution: the apply pass writes that convolution's operand planes
        bound = None
        cons = self.planes_consumer
        if _f16s() and dt == torch.float32 and cons is not None and cons._path() == 'tc3' and \
                kind in (-1, _lib.NOISE_DROPOUT) and code in (_lib.ACT_TANH, _lib.ACT_SIGMOID, _lib.ACT_RELU_MAX):
            bound = (par if code == _lib.ACT_RELU_MAX else 1.0) / ((1.0 - rate) if kind >= 0 else 1.0)
            bound = bound if bound > 0 else None
        if not ctx.training:
            _chain_fwd(sfx, x, y, dt, ptr(mm), ptr(mv), ptr(g), ptr(b), 1, self.epsilon, code, par, -1, 0.0, None, 0, 0,
                       rows, C, planes_bound=bound)
            return y
        stats = _empty((2 * C,))
        n_total = float(rows * ctx.world)
        # (sum x, sum x^2): from the epilogue of the convolution that produced x when it gathered them, else one pass
        sums = getattr(x, '_gn_bn_sums', None)
        if sums is None or sums.numel() != 2 * C:
            sums = torch.empty(2 * C, dtype=torch.float64, device=x.device)
            call('gn_bn_stats_bf16' if dt == BF16 else 'gn_bn_sums_f32', ptr(x, dt), rows, C, ptr(sums, torch.float64),
                 stream())
        if ctx.world > 1:
            ctx.dp.all_reduce(sums)
        # centred second moment from the raw one (double): sum (x-mean)^2 = sum x^2 - (sum x)^2 / n
        ssq = (sums[C:] - sums[:C] * sums[:C] / n_total).clamp_(min=0.0).contiguous()
        call('gn_bn_finalize_f32', ptr(sums, torch.float64), None, n_total, C, self.epsilon, self.momentum,
             ptr(stats), None, None, 0, None, 1.0, stream())
        self._finalize(ssq, n_total, C, stats, ctx)
        r, seed, off = None, 0, 0
        if kind >= 0:
            fed = ctx.noise.get(noise.name)
            if fed is not None:
                r = torch.as_tensor(np.ascontiguousarray(fed, dtype=np.float32)).to(x.device).reshape(x.shape).contiguous()
            else:
                n = x.numel()
                off = _STATE['noise_counter'] + ((ctx.dp.rank << 48) if ctx.dp is not None else 0)
                _STATE['noise_counter'] += (n + 3) // 4 * 4
                seed = _STATE['seed']
        _chain_fwd(sfx, x, y, dt, ptr(stats[:C]), ptr(stats[C:]), ptr(g), ptr(b), 0, self.epsilon, code, par, kind, rate,
                   ptr(r) if r is not None else None, seed, off, rows, C, planes_bound=bound)
        self._x, self._stats, self._n = x, stats, n_total
        self._chain_state = (code, par, kind, rate, r, seed, off)
        return y

    def _backward_chain(self, dy, ctx):
        x, stats = self._x, self._stats
        C = x.shape[-1]
        rows = x.numel() // C
        dt = x.dtype
        sfx = '_bf16' if dt == BF16 else '_f32'
        code, par, kind, rate, r, seed, off = self._chain_state
        dy = (_as_bf16(dy.contiguous()) if dt == BF16 else _as_f32(dy).contiguous()).reshape(x.shape)
        g, b = self.params[0].data, self.params[1].data
        rp = ptr(r) if r is not None else None
        sums = torch.empty(2 * C, dtype=torch.float64, device=x.device)
        call('gn_chain_bwd_sums' + sfx, ptr(x, dt), ptr(dy, dt), ptr(stats[:C]), ptr(stats[C:]), ptr(g), ptr(b), code,
             par, kind, rate, rp, seed, off, rows, C, ptr(sums, torch.float64), stream())
        if ctx.world > 1:
            ctx.dp.all_reduce(sums)
        dx = torch.empty(x.shape, dtype=dt, device=x.device)
        tr = id(self) in ctx.trainable_ids
        _chain_bwd(sfx, x, dy, dx, dt, ptr(stats[:C]), ptr(stats[C:]), ptr(g), ptr(b), ptr(sums, torch.float64), self._n,
                   code, par, kind, rate, rp, seed, off, ptr(self.params[0].grad) if tr else None,
                   ptr(self.params[1].grad) if tr else None, rows, C)
        if tr and ctx.world > 1:
            self.params[0].grad.mul_(1.0 / ctx.world)
            self.params[1].grad.mul_(1.0 / ctx.world)
        self._x = self._stats = self._chain_state = None
        return dx

    def forward(self, x, ctx):
        if not self._mid():
            return self._forward_last(x, ctx)
        x = _as_f32(x).contiguous()
        B, A = x.shape[0], x.shape[1]
        inner = x.numel() // (B * A)
        xt = _empty((B, inner, A))
        call('gn_transpose_f32', ptr(x), ptr(xt), B, A, inner, stream())
        yt = self._forward_last(xt, ctx).contiguous()
        y = _empty(x.shape)
        call('gn_transpose_f32', ptr(_as_f32(yt)), ptr(y), B, inner, A, stream())
        return y

    def backward(self, dy, ctx, need_dx=True):
        if not self._mid():
            return self._backward_last(dy, ctx, need_dx)
        dy = _as_f32(dy).contiguous()
        B, A = dy.shape[0], dy.shape[1]
        inner = dy.numel() // (B * A)
        dyt = _empty((B, inner, A))
        call('gn_transpose_f32', ptr(dy), ptr(dyt), B, A, inner, stream())
        dxt = _as_f32(self._backward_last(dyt, ctx, need_dx)).contiguous()
        dx = _empty(dy.shape)
        call('gn_transpose_f32', ptr(dxt), ptr(dx), B, inner, A, stream())
        return dx

    def _forward_last(self, x, ctx):
        for l in self.chain:
            if l is not None:
                l._chain_skip = False
        self._fused_call = x.shape[-1] % 8 == 0 and (x.dtype == BF16 or _STATE['f32_chain']) and \
            not any(isinstance(l, GaussianNoise) for l in self.chain if l is not None)
        if self._fused_call:
            return self._forward_chain(x.contiguous(), ctx)
        x = _as_f32(x)
        C = x.shape[-1]
        rows = x.numel() // C
        g, b, mm, mv = [p.data for p in self.params]
        y = _empty(x.shape)
        if not ctx.training:
            call('gn_bn_apply_f32', ptr(x), ptr(mm), ptr(mv), ptr(g), ptr(b), ptr(y), rows, C, self.epsilon, 1,
                 stream())
            return y
        sums = torch.empty(2 * C, dtype=torch.float64, device=x.device)
        stats = _empty((2 * C,))
        n_total = float(rows * ctx.world)
        call('gn_bn_stats_f32', ptr(x), rows, C, ptr(sums, torch.float64), None, stream())
        if ctx.world > 1:
            ctx.dp.all_reduce(sums)
        call('gn_bn_finalize_f32', ptr(sums, torch.float64), None, n_total, C, self.epsilon, self.momentum,
             ptr(stats), None, None, 0, None, 1.0, stream())
        call('gn_bn_stats_f32', ptr(x), rows, C, ptr(sums, torch.float64), ptr(stats), stream())
        if ctx.world > 1:
            ctx.dp.all_reduce(sums)
        self._finalize(sums[C:], n_total, C, stats, ctx)
        call('gn_bn_apply_f32', ptr(x), ptr(stats[:C]), ptr(stats[C:]), ptr(g), ptr(b), ptr(y), rows, C,
             self.epsilon, 0, stream())
        self._x, self._stats, self._n = x, stats, n_total
        return y

    def _backward_last(self, dy, ctx, need_dx=True):
        if getattr(self, '_fused_call', False):
            return self._backward_chain(dy, ctx)
        dy = _as_f32(dy)
        x, stats = self._x, self._stats
        C = x.shape[-1]
        rows = x.numel() // C
        sums = torch.empty(2 * C, dtype=torch.float64, device=x.device)
        call('gn_bn_bwd_sums_f32', ptr(x), ptr(dy), ptr(stats), rows, C, ptr(sums, torch.float64), stream())
        if ctx.world > 1:
            ctx.dp.all_reduce(sums)
        dx = _empty(x.shape)
        tr = id(self) in ctx.trainable_ids
        # dgamma/dbeta written here are GLOBAL sums when data-parallel: scale so the later
        # gradient all-reduce(sum) does not count them world times
        call('gn_bn_bwd_apply_f32', ptr(x), ptr(dy), ptr(stats), ptr(self.params[0].data), ptr(sums, torch.float64),
             self._n, ptr(dx), ptr(self.params[0].grad) if tr else None, ptr(self.params[1].grad) if tr else None,
             rows, C, stream())
        if tr and ctx.world > 1:
            self.params[0].grad.mul_(1.0 / ctx.world)
            self.params[1].grad.mul_(1.0 / ctx.world)
        self._x = self._stats = None
        return dx


class _ActLayer(Layer):
    code, param = _lib.ACT_NONE, 0.0
    fused = False        # True when the preceding Conv1D applies this activation in its epilogue

    _chain_skip = False   # True for the current call when the preceding BatchNormalization ran the bf16 chain

    def forward(self, x, ctx):
        if self.code == _lib.ACT_NONE or self.fused or self._chain_skip:
            return x
        x = _as_f32(x)
        y = _empty(x.shape)
        call('gn_act_fwd_f32', ptr(x), ptr(y), x.numel(), self.code, self.param, stream())
        self._y = y
        return y

    def backward(self, dy, ctx, need_dx=True):
        if self.code == _lib.ACT_NONE or self.fused or self._chain_skip or not need_dx:
            return dy
        dx = _act_bwd(dy, self._y, self.code, self.param)
        self._y = None
        return dx


class Activation(_ActLayer):
    prefix = 'activation'

    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.activation = activation
        self.code = _ACTS[activation]


class LeakyReLU(_ActLayer):
    prefix = 'leaky_re_lu'

    def __init__(self, alpha=0.3, **kw):
        super().__init__(**kw)
        self.code, self.param = _lib.ACT_LEAKY, float(alpha)


class ReLU(_ActLayer):
    prefix = 're_lu'

    def __init__(self, max_value=None, **kw):
        super().__init__(**kw)
        if max_value is None:
            self.code = _lib.ACT_RELU
        else:
            self.code, self.param = _lib.ACT_RELU_MAX, float(max_value)


class _NoiseLayer(Layer):
    """Dropout family; active only in training mode.  The random tensor comes from ``ctx.noise[name]``
    when a caller feeds it (parity tests), else from the device Philox stream."""
    kind = _lib.NOISE_DROPOUT

    def __init__(self, rate, **kw):
        super().__init__(**kw)
        self.rate = float(rate)

    _chain_skip = False   # True for the current call when the preceding BatchNormalization ran the bf16 chain
    pre_act = None        # (code, param) of a ReLU / LeakyReLU fused into the convolution that feeds us (set by _fuse)

    def forward(self, x, ctx):
        self._bf16_state = None
        if not ctx.training or self._chain_skip:
            self._r = None
            return x
        if (x.dtype == BF16 or _STATE['f32_chain']) and x.shape[-1] % 8 == 0 and self.kind != _lib.NOISE_GNOISE:
            # chain kernel: y = x * factor with the mask recomputed from the Philox counter in backward (never stored)
            x = x.contiguous()
            dt = x.dtype
            sfx = '_bf16' if dt == BF16 else '_f32'
            C = x.shape[-1]
            rows = x.numel() // C
            fed = ctx.noise.get(self.name)
            r, seed, off = None, 0, 0
            if fed is not None:
                r = torch.as_tensor(np.ascontiguousarray(fed, dtype=np.float32)).to(x.device).reshape(x.shape).contiguous()
            else:
                off = _STATE['noise_counter'] + ((ctx.dp.rank << 48) if ctx.dp is not None else 0)
                _STATE['noise_counter'] += (x.numel() + 3) // 4 * 4
                seed = _STATE['seed']
            y = torch.empty(x.shape, dtype=dt, device=x.device)
            _chain_fwd(sfx, x, y, dt, None, None, None, None, 0, 0.0, _lib.ACT_NONE, 0.0, self.kind, self.rate,
                       ptr(r) if r is not None else None, seed, off, rows, C)
            self._bf16_state = (r, seed, off, dt)
            # the sign activation fused into the producing convolution's epilogue gets its derivative in OUR backward
            # pass (x is its output: same sign as the pre-activation), which saves the separate mask pass
            self._xin = x if (self.pre_act is not None and dt == torch.float32) else None
            self._r = None
            return y
        x = _as_f32(x)
        fed = ctx.noise.get(self.name)
        if fed is not None:
            r = torch.as_tensor(np.ascontiguousarray(fed, dtype=np.float32)).to(x.device).reshape(x.shape).contiguous()
        else:
            r = _empty(x.shape)
            n = x.numel()
            off = _STATE['noise_counter']
            _STATE['noise_counter'] += (n + 3) // 4 * 4
            rank_off = (ctx.dp.rank << 48) if ctx.dp is not None else 0
            call('gn_noise_draw_f32', ptr(r), n, self.kind, self.rate, _STATE['seed'], off + rank_off, stream())
        y = _empty(x.shape)
        call('gn_noise_fwd_f32', ptr(x), ptr(r), ptr(y), x.numel(), self.kind, self.rate, stream())
        self._r = r
        return y

    def backward(self, dy, ctx, need_dx=True):
        if getattr(self, '_bf16_state', None) is not None and need_dx:
            r, seed, off, dt = self._bf16_state
            sfx = '_bf16' if dt == BF16 else '_f32'
            dy = _as_bf16(dy.contiguous()) if dt == BF16 else _as_f32(dy).contiguous()
            C = dy.shape[-1]
            rows = dy.numel() // C
            dx = torch.empty(dy.shape, dtype=dt, device=dy.device)
            xin = getattr(self, '_xin', None)
            code, par = self.pre_act if xin is not None else (_lib.ACT_NONE, 0.0)
            _chain_bwd(sfx, xin if xin is not None else dy, dy, dx, dt, None, None, None, None, None, 1.0, code, par,
                       self.kind, self.rate, ptr(r) if r is not None else None, seed, off, None, None, rows, C)
            if xin is not None:
                dx._gn_preact = True
            self._bf16_state = self._xin = None
            return dx
        if self._r is None or not need_dx:
            return dy
        dy = _as_f32(dy)
        dx = _empty(dy.shape)
        call('gn_noise_bwd_f32', ptr(dy), ptr(self._r), ptr(dx), dy.numel(), self.kind, self.rate, stream())
        self._r = None
        return dx


class Dropout(_NoiseLayer):
    prefix = 'dropout'
    kind = _lib.NOISE_DROPOUT


class GaussianDropout(_NoiseLayer):
    prefix = 'gaussian_dropout'
    kind = _lib.NOISE_GDROPOUT


class GaussianNoise(_NoiseLayer):
    prefix = 'gaussian_noise'
    kind = _lib.NOISE_GNOISE


class Reshape(Layer):
    prefix = 'reshape'

    def __init__(self, target_shape, **kw):
        super().__init__(**kw)
        self.target = tuple(target_shape)

    def build(self, in_shape):
        n = int(np.prod(in_shape))
        t = list(self.target)
        if -1 in t:
            t[t.index(-1)] = n // int(-np.prod(t))
        assert int(np.prod(t)) == n, 'cannot reshape %s to %s' % (in_shape, self.target)
        return tuple(t)

    def forward(self, x, ctx):
        return x.reshape((x.shape[0],) + self.output_shape)

    def backward(self, dy, ctx, need_dx=True):
        dx = dy.reshape((dy.shape[0],) + self.input_shape)
        if getattr(dy, '_gn_preact', False):
            dx._gn_preact = True
        if getattr(dy, '_gn_db_done', False):
            dx._gn_db_done = True
        return dx


class Flatten(Reshape):
    prefix = 'flatten'

    def __init__(self, **kw):
        Layer.__init__(self, **kw)

    def build(self, in_shape):
        return (int(np.prod(in_shape)),)


class UpSampling1D(Layer):
    prefix = 'up_sampling1d'

    def __init__(self, size=2, **kw):
        super().__init__(**kw)
        self.size = int(size)
        self.fused = False       # True when the following Conv1D reads through the upsampling

    def build(self, in_shape):
        return (in_shape[0] * self.size, in_shape[1])

    def forward(self, x, ctx):
        if self.fused:
            return x
        x = _as_f32(x)
        B, L, C = x.shape
        y = _empty((B, L * self.size, C))
        call('gn_upsample1d_fwd_f32', ptr(x), ptr(y), B, L, C, self.size, stream())
        return y

    def backward(self, dy, ctx, need_dx=True):
        if self.fused:
            return dy
        dy = _as_f32(dy)
        B = dy.shape[0]
        L, C = self.input_shape
        dx = _empty((B, L, C))
        call('gn_upsample1d_bwd_f32', ptr(dy), ptr(dx), B, L, C, self.size, stream())
        return dx


class MaxPooling1D(Layer):
    prefix = 'max_pooling1d'

    def __init__(self, pool_size=2, **kw):
        super().__init__(**kw)
        self.pool = int(pool_size)

    def build(self, in_shape):
        return (in_shape[0] // self.pool, in_shape[1])

    def forward(self, x, ctx):
        x = _as_f32(x)
        B, L, C = x.shape
        y = _empty((B, L // self.pool, C))
        call('gn_maxpool1d_fwd_f32', ptr(x), ptr(y), B, L, C, self.pool, stream())
        self._x, self._y = x, y
        return y

    def backward(self, dy, ctx, need_dx=True):
        x = self._x
        dy = _as_f32(dy)
        B, L, C = x.shape
        dx = _empty(x.shape)
        call('gn_maxpool1d_bwd_f32', ptr(x), ptr(self._y), ptr(dy), ptr(dx), B, L, C, self.pool, stream())
        self._x = self._y = None
        return dx


class GlobalAveragePooling1D(Layer):
    """Keras GlobalAveragePooling1D / 2D (channels_last): mean over every axis between batch and channels
    (2_model_version/no_weight_code/subtract_model.py:330,371)."""
    prefix = 'global_average_pooling1d'

    def build(self, in_shape):
        return (in_shape[-1],)

    def forward(self, x, ctx):
        x = _as_f32(x).contiguous()
        B, C = x.shape[0], x.shape[-1]
        self._L = x.numel() // (B * C)
        y = _empty((B, C))
        call('gn_gap_fwd_f32', ptr(x), ptr(y), B, self._L, C, stream())
        self._shape = tuple(x.shape)
        return y

    def backward(self, dy, ctx, need_dx=True):
        dy = _as_f32(dy).contiguous()
        dx = _empty(self._shape)
        call('gn_gap_bwd_f32', ptr(dy), ptr(dx), self._shape[0], self._L, self._shape[-1], stream())
        return dx


class GlobalAveragePooling2D(GlobalAveragePooling1D):
    prefix = 'global_average_pooling2d'


class StackResidual(Layer):
    """bbhMahoGANy.py:164-188 ``MyLayer``: stack([x, const - x], axis=2) -> (B, n_pix, 2, 1)."""
    prefix = 'my_layer'

    def __init__(self, const, **kw):
        super().__init__(**kw)
        self._const_host = np.asarray(const, dtype=np.float32)

    def build(self, in_shape):
        L = in_shape[0]
        self.const = torch.from_numpy(np.ascontiguousarray(self._const_host.reshape(L))).to(device())
        return (L, 2, 1)

    def forward(self, x, ctx):
        x = _as_f32(x)
        B = x.shape[0]
        L = self.input_shape[0]
        y = _empty((B, L, 2, 1))
        call('gn_stack_residual_fwd_f32', ptr(x), ptr(self.const), ptr(y), B, L, stream())
        return y

    def backward(self, dy, ctx, need_dx=True):
        dy = _as_f32(dy)
        B = dy.shape[0]
        L = self.input_shape[0]
        dx = _empty((B,) + self.input_shape)
        call('gn_stack_residual_bwd_f32', ptr(dy), ptr(dx), B, L, stream())
        return dx


class ResidualMoments(Layer):
    """tests/burstMahoGANy.py:100-125 ``MyLayer``: [mean(const-x), mean((const-x)^2)] over the whole
    batch; the output has shape (2,) (no batch axis), as in the reference."""
    prefix = 'my_layer'
    batch_global = True

    def __init__(self, const, **kw):
        super().__init__(**kw)
        self._const_host = np.asarray(const, dtype=np.float32)

    def build(self, in_shape):
        L = int(np.prod(in_shape))
        self.const = torch.from_numpy(np.ascontiguousarray(self._const_host.reshape(L))).to(device())
        return (2,)

    def forward(self, x, ctx):
        x = _as_f32(x)
        B = x.shape[0]
        L = x.numel() // B
        sums = torch.empty(2, dtype=torch.float64, device=x.device)
        call('gn_residual_moments_fwd_f32', ptr(x), ptr(self.const), ptr(sums, torch.float64), B, L, stream())
        if ctx.world > 1:
            ctx.dp.all_reduce(sums)
        self._n = float(B * L * ctx.world)
        self._x = x
        return (sums / self._n).to(torch.float32)

    def backward(self, dy, ctx, need_dx=True):
        x = self._x
        B = x.shape[0]
        L = x.numel() // B
        dx = _empty(x.shape)
        call('gn_residual_moments_bwd_f32', ptr(x), ptr(self.const), ptr(dy.contiguous()), ptr(dx), B, L, self._n,
             stream())
        self._x = None
        return dx


# ----------------------------------------------------------------------------- optimizers
class Optimizer:
    def __init__(self):
        self.iterations = 0
        self.slots = {}

    def get_config(self):
        return {}


class Adam(Optimizer):
    """Keras 2.2.4 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps), eps = 1e-7."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, **kw):
        super().__init__()
        self.lr, self.beta_1, self.beta_2 = float(lr), float(beta_1), float(beta_2)
        self.epsilon = 1e-7 if epsilon is None else float(epsilon)
        self.decay = float(decay)

    def apply(self, segments, grad_scale):
        lr = self.lr
        if self.decay > 0:
            lr = lr * (1.0 / (1.0 + self.decay * self.iterations))
        t = self.iterations + 1
        lr_t = lr * (math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t))
        for key, p, g in segments:
            if key not in self.slots:
                self.slots[key] = (torch.zeros_like(p), torch.zeros_like(p))
            m, v = self.slots[key]
            call('gn_adam_step_f32', ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr_t, self.beta_1, self.beta_2,
                 self.epsilon, grad_scale, stream())
        self.iterations += 1
        _STATE['wver'] += 1


class SGD(Optimizer):
    def __init__(self, lr=0.01, momentum=0.0, decay=0.0, nesterov=False, **kw):
        super().__init__()
        assert momentum == 0.0 and not nesterov, 'the reference uses plain SGD (nn.py:79)'
        self.lr, self.decay = float(lr), float(decay)

    def apply(self, segments, grad_scale):
        lr = self.lr
        if self.decay > 0:
            lr = lr * (1.0 / (1.0 + self.decay * self.iterations))
        for key, p, g in segments:
            call('gn_sgd_step_f32', ptr(p), ptr(g), p.numel(), lr, grad_scale, stream())
        self.iterations += 1
        _STATE['wver'] += 1


# ----------------------------------------------------------------------------- losses
class _LossSpec:
    def __init__(self, loss):
        self.param = 0.0
        if callable(loss) and getattr(loss, 'gn_kind', None) is not None:
            self.kind, self.param, self.name = loss.gn_kind, float(loss.gn_param), 'chisquare_Loss'
        elif loss in ('binary_crossentropy',):
            self.kind, self.name = _lib.LOSS_BCE, loss
        elif loss in ('mean_squared_error', 'mse'):
            self.kind, self.name = _lib.LOSS_MSE, 'mean_squared_error'
        else:
            raise ValueError('unsupported loss %r (the reference uses binary_crossentropy, mean_squared_error '
                             'and chisquare_Loss)' % (loss,))


def chisquare_Loss(n_sig=1.0):
    """bbhMahoGANy.py:146-162: K.sum(K.square(yTrue - yPred)/n_sig**2, axis=-1)."""
    def loss(y_true, y_pred):
        raise RuntimeError('chisquare_Loss is evaluated on the device')
    loss.gn_kind, loss.gn_param = _lib.LOSS_CHISQ, n_sig
    return loss


# ----------------------------------------------------------------------------- models
class Model(Layer):
    """Functional model: ``Model(inputs, outputs)``.  Also the base of Sequential."""
    prefix = 'model'

    def __init__(self, inputs=None, outputs=None, name=None):
        super().__init__(name=name)
        self._compiled = None
        self.optimizer = None
        if inputs is not None:
            self._init_graph(inputs, outputs)

    def _init_graph(self, inputs, outputs):
        self._multi_out = isinstance(outputs, (list, tuple))
        ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        outs = list(outputs) if self._multi_out else [outputs]
        assert len(ins) == 1, 'single-input models only (all reference models are)'
        self._in_node = ins[0].node
        self._out_nodes = [o.node for o in outs]
        order, seen = [], set()

        def visit(n):
            if id(n) in seen:
                return
            seen.add(id(n))
            for i in n.inputs:
                visit(i)
            order.append(n)
        for o in self._out_nodes:
            visit(o)
        assert id(self._in_node) in seen, 'outputs do not depend on the input'
        self._order = [n for n in order if n is not self._in_node and not isinstance(n.layer, InputLayer)]
        self.layers = self._keras_layer_order()
        if self.name is None:
            self.name = _uid(self.prefix)
        self.input_shape = ins[0].shape
        self.output_shape = [o.shape for o in outs] if self._multi_out else outs[0].shape
        self.built = True
        self._fuse()

    def _keras_layer_order(self):
        """``model.layers`` as Keras 2.2.4 orders it (keras/engine/network.py ``_map_graph_network``): by depth from
        the outputs, deepest first, ties broken by the pre-order index of the traversal that starts at the first
        output.  ``layer_names`` of a saved file, ``get_weights`` / ``set_weights`` and the optimizer state follow
        this order, so a two-branch model (bbhMahoGANy.py:357-404, saved at :1173, reloaded at :1135) interleaves its
        branches exactly as a Keras-written ``signal_pe.h5`` does.  A functional model lists its InputLayer first; a
        Sequential model does not list it (``Sequential.layers`` of 2.2.4)."""
        layer_index, finished, post = {}, set(), []

        def build_map(n):
            if id(n) in finished:
                return
            if id(n.layer) not in layer_index:
                layer_index[id(n.layer)] = len(layer_index)
            for i in n.inputs:
                build_map(i)
            finished.add(id(n))
            post.append(n)
        for o in self._out_nodes:
            build_map(o)
        node_depth, layer_depth, by_id = {}, {}, {}
        for n in reversed(post):
            d = max(node_depth.get(id(n), 0), layer_depth.get(id(n.layer), 0))
            node_depth[id(n)] = layer_depth[id(n.layer)] = d
            by_id[id(n.layer)] = n.layer
            for i in n.inputs:
                node_depth[id(i)] = max(d + 1, node_depth.get(id(i), 0))
        top = max(layer_depth.values())
        for lid, l in by_id.items():
            if isinstance(l, InputLayer):
                layer_depth[lid] = top
        ordered = sorted(by_id, key=lambda lid: (-layer_depth[lid], layer_index[lid]))
        layers = [by_id[lid] for lid in ordered]
        if isinstance(self, Sequential):
            layers = [l for l in layers if not isinstance(l, InputLayer)]
        return layers

    def _fuse(self):
        """UpSampling1D(2) feeding exactly one Conv1D is folded into the convolution's loader."""
        users = {}
        for n in self._order:
            for i in n.inputs:
                users.setdefault(id(i), []).append(n)
        for n in self._order:
            if isinstance(n.layer, UpSampling1D) and n.layer.size == 2:
                u = users.get(id(n), [])
                if len(u) == 1 and type(u[0].layer) is Conv1D and n not in self._out_nodes:
                    n.layer.fused = True
                    u[0].layer.fused_up = 2
        # Conv1D -> (Activation | LeakyReLU | ReLU) with a single user: the activation runs in the conv epilogue
        for n in self._order:
            if type(n.layer) in (Conv1D, Conv2D) and getattr(n.layer, 'activation', None) is None \
                    and n not in self._out_nodes:
                u = users.get(id(n), [])
                if len(u) == 1 and isinstance(u[0].layer, _ActLayer) and u[0].layer.code != _lib.ACT_NONE:
                    n.layer.post_act = (u[0].layer.code, u[0].layer.param)
                    u[0].layer.fused = True
        # Conv1D -> BatchNormalization(axis=-1) with a single user: the convolution epilogue gathers the statistics
        for n in self._order:
            if type(n.layer) is Conv1D and n not in self._out_nodes:
                u = users.get(id(n), [])
                if len(u) == 1 and type(u[0].layer) is BatchNormalization and u[0].layer.axis != 1:
                    n.layer.bn_consumer = u[0].layer
        # Conv (ReLU | LeakyReLU in the epilogue) -> Dropout: the dropout's backward pass applies the activation mask too
        for n in self._order:
            if type(n.layer) in (Conv1D, Conv2D) and n.layer.post_act is not None and \
                    n.layer.post_act[0] in (_lib.ACT_RELU, _lib.ACT_LEAKY):
                a = users.get(id(n), [])[0]
                u = users.get(id(a), [])
                if len(u) == 1 and isinstance(u[0].layer, (Dropout, GaussianDropout)) and a not in self._out_nodes:
                    u[0].layer.pre_act = n.layer.post_act
        # BatchNormalization -> [activation] -> [dropout]: candidates for the bf16 chain kernels (decided per call by the
        # dtype of the tensor reaching the BatchNormalization; float32 tensors keep the exact per-layer kernels)
        for n in self._order:
            if type(n.layer) is BatchNormalization and n.layer.axis != 1:
                act = noise = None
                cur = n
                u = users.get(id(cur), [])
                if len(u) == 1 and isinstance(u[0].layer, _ActLayer) and not u[0].layer.fused and cur not in self._out_nodes:
                    act, cur = u[0].layer, u[0]
                    u = users.get(id(cur), [])
                if len(u) == 1 and isinstance(u[0].layer, (Dropout, GaussianDropout)) and cur not in self._out_nodes:
                    noise, cur = u[0].layer, u[0]
                    u = users.get(id(cur), [])
                n.layer.chain = (act, noise)
                # the Conv1D that alone consumes the chain's output (through a fused UpSampling1D): candidate for
                # operand planes written by the chain's apply pass
                if len(u) == 1 and isinstance(u[0].layer, UpSampling1D) and u[0].layer.fused and cur not in self._out_nodes:
                    cur = u[0]
                    u = users.get(id(cur), [])
                n.layer.planes_consumer = u[0].layer if (len(u) == 1 and type(u[0].layer) is Conv1D and
                                                         cur not in self._out_nodes) else None
        # consumer of a fused conv+activation (through views / the now-identity activation layer): it may apply
        # the activation derivative in its own data-gradient epilogue, using its input as the mask source
        for n in self._order:
            if type(n.layer) in (Conv1D, Dense):
                src = n.inputs[0]
                ok = True
                reshaped = False
                while ok and src is not self._in_node and (
                        (isinstance(src.layer, _ActLayer) and src.layer.fused) or isinstance(src.layer, Reshape)):
                    ok = len(users.get(id(src), [])) == 1 and src not in self._out_nodes
                    reshaped = reshaped or isinstance(src.layer, Reshape)
                    src = src.inputs[0]
                if ok and src is not self._in_node and type(src.layer) is Conv1D and src.layer.post_act is not None \
                        and len(users.get(id(src), [])) == 1:
                    n.layer.in_act = src.layer.post_act
                    n.layer.bias_src = src.layer     # its bias gradient = column sums of our data gradient
                    if type(n.layer) is Conv1D and not reshaped:
                        src.layer.plane_consumer = n.layer

    # a model can be used as a layer
    def __call__(self, x):
        assert self.built, 'model has no input shape yet'
        return KTensor(self.output_shape, Node(self, [x.node]))

    def _ensure_built(self, in_shape):
        assert self.built
        return self.output_shape

    def all_layers(self):
        out = []
        for l in self.layers:
            for s in l.all_layers():
                if s not in out:
                    out.append(s)
        return out

    @property
    def params(self):
        return [p for l in self.all_layers() for p in l.params]

    @params.setter
    def params(self, v):
        pass

    def get_weights(self):
        return [w for l in self.all_layers() for w in l.get_weights()]

    def set_weights(self, ws):
        i = 0
        for l in self.all_layers():
            n = len(l.params)
            l.set_weights(ws[i:i + n])
            i += n
        assert i == len(ws), 'expected %d arrays, got %d' % (i, len(ws))

    def count_params(self):
        return sum(l.count_params() for l in self.all_layers())

    def summary(self, print_fn=print):
        print_fn('_' * 65)
        print_fn('%-30s%-22s%-13s' % ('Layer (type)', 'Output Shape', 'Param #'))
        print_fn('=' * 65)
        for l in self.layers:
            shp = l.output_shape
            print_fn('%-30s%-22s%-13d' % ('%s (%s)' % (l.name, type(l).__name__), str((None,) + tuple(shp)) if not
                                          isinstance(shp, list) else str(shp), l.count_params()))
        tot = self.count_params()
        tr = sum(p.numel() for l in self.all_layers() if l.trainable for p in l.params if p.trainable)
        print_fn('=' * 65)
        print_fn('Total params: {:,}'.format(tot))
        print_fn('Trainable params: {:,}'.format(tr))
        print_fn('Non-trainable params: {:,}'.format(tot - tr))
        print_fn('_' * 65)

    # ---- execution ------------------------------------------------------------
    def forward(self, x, ctx):
        vals = {id(self._in_node): x}
        for n in self._order:
            vals[id(n)] = n.layer.forward(vals[id(n.inputs[0])], ctx)
        outs = [vals[id(o)] for o in self._out_nodes]
        return outs if self._multi_out else outs[0]

    def backward(self, dy, ctx, need_dx=True):
        dys = dy if self._multi_out else [dy]
        grads = {}
        for o, g in zip(self._out_nodes, dys):
            grads[id(o)] = g
        first_users = self._nodes_needing_dx(ctx, need_dx)
        for n in reversed(self._order):
            g = grads.pop(id(n), None)
            if g is None:
                continue
            src = n.inputs[0]
            want = id(n) in first_users
            dx = n.layer.backward(g, ctx, want)
            if not want or dx is None:
                continue
            if id(src) in grads:
                acc = grads[id(src)]
                if acc.dtype != torch.float32:
                    acc = grads[id(src)] = _as_f32(acc)
                call('gn_axpy_f32', ptr(acc.reshape(-1)), ptr(_as_f32(dx).reshape(-1).contiguous()), 1.0,
                     dx.numel(), stream())
                acc._gn_amax = acc._gn_planes = acc._gn_bn_sums = None      # side products of the kernel that wrote acc are stale now
            else:
                grads[id(src)] = dx
        return grads.get(id(self._in_node))

    def _nodes_needing_dx(self, ctx, need_dx):
        """A node must produce dx iff something upstream of it has trainable parameters (or the
        caller wants the model's input gradient)."""
        need = set()
        upstream_trainable = {id(self._in_node): need_dx}
        for n in self._order:
            src = n.inputs[0]
            up = upstream_trainable.get(id(src), False)
            if up:
                need.add(id(n))
            own = any(id(l) in ctx.trainable_ids and l.params for l in n.layer.all_layers())
            upstream_trainable[id(n)] = up or own
        return need

    # ---- Keras protocol ---------------------------------------------------------
    def compile(self, loss=None, optimizer=None, metrics=None, **kw):
        self.loss = loss
        self.optimizer = optimizer
        self.metrics = metrics or []
        spec = _LossSpec(loss)
        train_layers = [l for l in self.all_layers() if l.trainable and l.params]
        tparams = [p for l in train_layers for p in l.params if p.trainable]
        _home_params(tparams)
        self._compiled = {'loss': spec, 'trainable_ids': set(id(l) for l in train_layers),
                          'segments': _segments(tparams)}
        self.metrics_names = ['loss'] + (['acc'] if self.metrics else [])

    def predict(self, x, batch_size=1024, verbose=0):
        ctx = Ctx(False)
        xs = _to_device(x)
        outs = None
        for i in range(0, xs.shape[0], batch_size):
            o = self.forward(xs[i:i + batch_size].contiguous(), ctx)
            o = o if isinstance(o, list) else [o]
            if outs is None:
                outs = [[] for _ in o]
            for k, t in enumerate(o):
                outs[k].append(t)
        # batch-global outputs (burst MyLayer) have no batch axis to concatenate along
        res = [ts[0] if ts[0].dim() == 1 else torch.cat(ts, 0) for ts in outs]
        res = [r.detach().cpu().numpy() for r in res]
        return res if self._multi_out else res[0]

    def train_on_batch(self, x, y, sample_weight=None, class_weight=None, _noise=None, _return_device=False):
        """One optimizer step; returns [loss, acc] (single output) or [total, loss_1.., acc_1..]."""
        assert self._compiled is not None, 'compile() the model first'
        c = self._compiled
        ctx = Ctx(True, c['trainable_ids'], _noise)
        xs = _to_device(x)
        B = xs.shape[0]
        out = self.forward(xs, ctx)
        outs = out if isinstance(out, list) else [out]
        ys = y if self._multi_out else [y]
        assert len(ys) == len(outs), 'expected %d target arrays' % len(outs)
        res = torch.zeros(2 * len(outs), dtype=torch.float32, device=xs.device)
        dys, invs = [], []
        for k, (o, t) in enumerate(zip(outs, ys)):
            vec = o.dim() == 1           # batch-global output (burst MyLayer)
            D = o.shape[-1]
            # Keras: mean over the last axis, then over every remaining axis (batch and, e.g., time)
            rows = B if vec else o.numel() // D
            tt = _to_device(t).reshape(rows, -1)
            assert tt.shape[1] == D, 'target shape %s does not match output %s' % (tuple(tt.shape), tuple(o.shape))
            inv_rows = 1.0 / (rows * ctx.world)
            d = torch.zeros_like(o) if vec else _empty(o.shape)
            metric_kind = 0 if (D == 1 or c['loss'].kind == _lib.LOSS_BCE) else 1
            call('gn_loss_fwd_bwd_f32', ptr(o.contiguous()), ptr(tt.contiguous()), ptr(res[2 * k:2 * k + 2]), ptr(d),
                 rows, D, c['loss'].kind, c['loss'].param, inv_rows, 1 if vec else 0, metric_kind, stream())
            if vec and ctx.world > 1:
                ctx.dp.all_reduce(d)
            dys.append(d)
            invs += [inv_rows, inv_rows]
        self.backward(dys if self._multi_out else dys[0], ctx, need_dx=False)
        segs = c['segments']
        if ctx.world > 1:
            for _, p, g in segs:
                ctx.dp.all_reduce(g)
            ctx.dp.all_reduce(res)
        self.optimizer.apply(segs, 1.0)
        if len(set(invs)) == 1:
            res = res * invs[0]
        else:
            res = res * torch.tensor(invs, dtype=torch.float32, device=res.device)
        reg = ctx.reg_loss              # keras: total loss = output losses + sum(model.losses)
        if reg is not None and ctx.world > 1:
            ctx.dp.all_reduce(reg)
        if _return_device:
            if reg is not None:
                res[0] += reg[0].to(torch.float32)
            return res
        r = res.detach().cpu().numpy().astype(np.float64)      # the D2H read of the step's loss / metric
        regv = float(reg.item()) if reg is not None else 0.0
        losses = [float(r[2 * k]) for k in range(len(outs))]
        accs = [float(r[2 * k + 1]) for k in range(len(outs))]
        if len(outs) == 1:
            return [losses[0] + regv, accs[0]] if self.metrics else losses[0] + regv
        return [float(sum(losses)) + regv] + losses + (accs if self.metrics else [])

    def fit(self, x, y, batch_size=32, epochs=1, verbose=0, shuffle=True, **kw):
        x = np.asarray(x)
        y = np.asarray(y)
        n = x.shape[0]
        hist = []
        rs = np.random.RandomState(_STATE['seed'])
        for e in range(epochs):
            idx = rs.permutation(n) if shuffle else np.arange(n)
            for i in range(0, n, batch_size):
                j = idx[i:i + batch_size]
                hist.append(self.train_on_batch(x[j], y[j]))
        return hist

    def get_gradients(self):
        """Gradients of the last train_on_batch, Keras weight order (parity-test hook)."""
        return [p.grad.detach().cpu().numpy().copy() for l in self.all_layers() if id(l) in
                self._compiled['trainable_ids'] for p in l.params if p.trainable]

    # persistence: see gennet_b200.io (Keras HDF5 layout)
    def save_weights(self, path, overwrite=True):
        from . import io
        io.save_weights(self, path, overwrite)

    def load_weights(self, path):
        from . import io
        io.load_weights(self, path)

    def save(self, path, overwrite=True):
        from . import io
        io.save_model(self, path, overwrite)


class Sequential(Model):
    prefix = 'sequential'

    def __init__(self, layers=None, name=None):
        Layer.__init__(self, name=name)
        self._compiled = None
        self.optimizer = None
        self._pending = []
        self._tensor = None
        self._input = None
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        if self._tensor is None:
            shape = None
            if isinstance(layer, Model):
                shape = layer.input_shape if layer.built else None
            elif layer._input_shape_arg is not None:
                shape = layer._input_shape_arg
            if shape is None:
                raise ValueError('the first layer of a Sequential model needs input_shape=...')
            self._input = Input(shape=shape)
            self._tensor = self._input
        self._tensor = layer(self._tensor)
        self._init_graph(self._input, self._tensor)


def set_trainable(model, trainable):
    """bbhMahoGANy.py:797-809 / nn.py:95-98 (set_trainability)."""
    model.trainable = trainable
    for l in model.all_layers():
        l.trainable = trainable


set_trainability = set_trainable


def load_model(path, custom_objects=None):
    from . import io
    return io.load_model(path, custom_objects)


# ----------------------------------------------------------------------------- parameter arenas
def _home_params(params):
    """Move parameters that are not yet in an arena into one new flat arena (values preserved)."""
    fresh = [p for p in params if p.arena is None]
    if not fresh:
        return
    n = sum(p.numel() for p in fresh)
    # keep every tensor 16-byte aligned inside the arena
    offs, tot = [], 0
    for p in fresh:
        offs.append(tot)
        tot += (p.numel() + 3) // 4 * 4
    data = torch.zeros(tot, dtype=torch.float32, device=device())
    grad = torch.zeros(tot, dtype=torch.float32, device=device())
    arena = {'data': data, 'grad': grad}
    for p, o in zip(fresh, offs):
        view = data[o:o + p.numel()].view(p.shape)
        view.copy_(p.data)
        p.data = view
        p.grad = grad[o:o + p.numel()].view(p.shape)
        p.arena, p.offset = arena, o


def _segments(params):
    """Maximal runs of parameters that are adjacent in the same arena -> [(key, data, grad)]."""
    segs = []
    cur = None
    for p in params:
        end = p.offset + (p.numel() + 3) // 4 * 4
        if cur is not None and cur[0] is p.arena and cur[2] == p.offset:
            cur[2] = end
        else:
            if cur is not None:
                segs.append(cur)
            cur = [p.arena, p.offset, end]
    if cur is not None:
        segs.append(cur)
    return [((id(a), lo, hi), a['data'][lo:hi], a['grad'][lo:hi]) for a, lo, hi in segs]


_PINNED = {}


def _to_device(x):
    """NumPy / list / tensor -> contiguous float32 CUDA tensor (pinned staging for host arrays)."""
    if isinstance(x, torch.Tensor):
        t = x
        if not t.is_cuda:
            t = t.to(device(), non_blocking=True)
        return t.to(torch.float32).contiguous()
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    key = a.shape
    slot = _PINNED.get(key)
    if slot is None:
        if len(_PINNED) > 64:
            _PINNED.clear()
        # two pinned staging buffers per shape, each guarded by the event of the last copy out of it: the host never
        # waits for the compute stream, only (rarely) for the H2D copy issued two calls ago from the same buffer
        slot = {'bufs': [torch.empty(a.shape, dtype=torch.float32).pin_memory() for _ in range(2)],
                'events': [None, None], 'next': 0}
        _PINNED[key] = slot
    i = slot['next']
    slot['next'] = 1 - i
    if slot['events'][i] is not None:
        slot['events'][i].synchronize()
    buf = slot['bufs'][i]
    buf.copy_(torch.from_numpy(a))
    t = buf.to(device(), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    slot['events'][i] = ev
    return t
