"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the network half of the hot path.

torch-CPU (float64 by default) restatement of the Keras 2.2.4 / TensorFlow 1.12
semantics the reference's builders and ``train_on_batch`` loops rely on
(requirements.txt:16,47).  Keras/TF are third-party and absent from the
checkout and from this image, so these semantics are restated from the
published Keras 2.2.4 sources ([A1]-[A12] in SURVEY.md section 8c):
**parity unpinned** for this file -- no reference test or fixture holds expected
network outputs.  Autograd supplies the gradients; nothing here is used by the
product.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import it.

Builders cite the reference lines whose layer list they reproduce.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-7  # K.epsilon()


def same_pad(L, k, s):
    """TF 'SAME': out=ceil(L/s), total=max((out-1)s+k-L,0), left=total//2 [A2]."""
    out = -(-L // s)
    total = max((out - 1) * s + k - L, 0)
    return total // 2, total - total // 2


class Layer:
    kind = 'layer'

    def __init__(self):
        self.trainable = True
        self.weights = []        # trainable tensors, Keras order
        self.state = []          # non-trainable tensors (BN moving stats)
        self.name = None

    def build(self, in_shape, gen, dtype):
        return in_shape

    def forward(self, x, training, noise):
        raise NotImplementedError

    def all_layers(self):
        return [self]


def _kink(noise, layer):
    k = noise.get('__kinks__') if isinstance(noise, dict) else None
    return None if k is None else k.get(layer.name)


def glorot_uniform(shape, fan_in, fan_out, gen, dtype):
    lim = math.sqrt(6.0 / (fan_in + fan_out))  # [A12]
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype).requires_grad_(True)


class Dense(Layer):
    kind = 'dense'

    def __init__(self, units, activation=None):
        super().__init__()
        self.units, self.activation = units, activation

    def build(self, in_shape, gen, dtype):
        fin = in_shape[-1]
        self.weights = [glorot_uniform((fin, self.units), fin, self.units, gen, dtype),
                        torch.zeros(self.units, dtype=dtype, requires_grad=True)]
        return in_shape[:-1] + (self.units,)

    def forward(self, x, training, noise):
        return apply_activation(x @ self.weights[0] + self.weights[1], self.activation, _kink(noise, self))


class Conv1D(Layer):
    kind = 'conv1d'

    def __init__(self, filters, kernel_size, strides=1, padding='valid', activation=None):
        super().__init__()
        self.filters, self.k, self.s, self.padding, self.activation = filters, kernel_size, strides, padding, activation

    def build(self, in_shape, gen, dtype):
        L, cin = in_shape
        k = self.k
        self.weights = [glorot_uniform((k, cin, self.filters), k * cin, k * self.filters, gen, dtype),
                        torch.zeros(self.filters, dtype=dtype, requires_grad=True)]
        if self.padding == 'same':
            Lo = -(-L // self.s)
        else:
            Lo = (L - k) // self.s + 1
        return (Lo, self.filters)

    def forward(self, x, training, noise):        # x (B, L, Cin) NLC [A1]
        xt = x.permute(0, 2, 1)
        if self.padding == 'same':
            xt = F.pad(xt, same_pad(x.shape[1], self.k, self.s))
        y = F.conv1d(xt, self.weights[0].permute(2, 1, 0), self.weights[1], stride=self.s)
        return apply_activation(y.permute(0, 2, 1), self.activation, _kink(noise, self))


class Conv2D(Layer):
    kind = 'conv2d'

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid'):
        super().__init__()
        self.filters, self.k, self.s, self.padding = filters, tuple(kernel_size), tuple(strides), padding

    def build(self, in_shape, gen, dtype):
        H, W, cin = in_shape
        kh, kw = self.k
        self.weights = [glorot_uniform((kh, kw, cin, self.filters), kh * kw * cin, kh * kw * self.filters, gen, dtype),
                        torch.zeros(self.filters, dtype=dtype, requires_grad=True)]
        if self.padding == 'same':
            return (-(-H // self.s[0]), -(-W // self.s[1]), self.filters)
        return ((H - kh) // self.s[0] + 1, (W - kw) // self.s[1] + 1, self.filters)

    def forward(self, x, training, noise):        # x (B,H,W,C) NHWC
        xt = x.permute(0, 3, 1, 2)
        if self.padding == 'same':
            ph = same_pad(x.shape[1], self.k[0], self.s[0])
            pw = same_pad(x.shape[2], self.k[1], self.s[1])
            xt = F.pad(xt, (pw[0], pw[1], ph[0], ph[1]))
        y = F.conv2d(xt, self.weights[0].permute(3, 2, 0, 1), self.weights[1], stride=self.s)
        return y.permute(0, 2, 3, 1)


class Conv2DTranspose(Layer):
    """Keras Conv2DTranspose, channels_last, kernel (kh, kw, Cout, Cin), 'valid': out = (in-1)*s + k
    (2_model_version/weight_version/no_mode_collapse_network.py:79-90).  TF's conv2d_transpose is the gradient of
    conv2d w.r.t. its input, i.e. torch's conv_transpose2d with weight (Cin, Cout, kh, kw)."""
    kind = 'conv2d_transpose'

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', activation=None, kernel_regularizer=None,
                 activity_regularizer=None):
        super().__init__()
        assert padding == 'valid'
        self.filters, self.k, self.s, self.activation = filters, tuple(kernel_size), tuple(strides), activation
        # (l1, l2) pairs: keras.regularizers.L1L2 -> K.sum(l1 * |x|) + K.sum(l2 * x^2), added to the model's total loss
        # (keras/engine/base_layer.py __call__ / add_weight, Keras 2.2.4: no division by the batch size)
        self.kernel_regularizer, self.activity_regularizer = kernel_regularizer, activity_regularizer

    def build(self, in_shape, gen, dtype):
        H, W, cin = in_shape
        kh, kw = self.k
        self.weights = [glorot_uniform((kh, kw, self.filters, cin), kh * kw * self.filters, kh * kw * cin, gen, dtype),
                        torch.zeros(self.filters, dtype=dtype, requires_grad=True)]
        return ((H - 1) * self.s[0] + kh, (W - 1) * self.s[1] + kw, self.filters)

    def forward(self, x, training, noise):        # x (B,H,W,C) NHWC
        y = F.conv_transpose2d(x.permute(0, 3, 1, 2).contiguous(), self.weights[0].permute(3, 2, 0, 1).contiguous(), self.weights[1],
                               stride=self.s)
        y = apply_activation(y.permute(0, 2, 3, 1), self.activation, _kink(noise, self))
        if training and isinstance(noise, dict):
            for reg, t in ((self.kernel_regularizer, self.weights[0]), (self.activity_regularizer, y)):
                if reg is not None:
                    noise.setdefault('__losses__', []).append(reg[0] * t.abs().sum() + reg[1] * (t * t).sum())
        return y


ZERO_DEBIAS = True      # Keras 2.2.4 + TF 1.12 moving averages (see BatchNormalization.forward); False = plain EMA


class BatchNormalization(Layer):
    """[A5] eps 1e-3, axis -1, batch mean + biased var in training, moving stats in
    inference; moving_var is fed the n/(n-(1+eps)) 'sample variance'."""
    kind = 'bn'

    def __init__(self, momentum=0.99, epsilon=1e-3, axis=-1):
        super().__init__()
        self.momentum, self.epsilon, self.axis = momentum, epsilon, axis
        self.biased = None          # shadow accumulators of the zero-debiased moving average (not layer weights)
        self.local_step = 0

    def build(self, in_shape, gen, dtype):
        c = in_shape[0] if (self.axis == 1 and len(in_shape) > 1) else in_shape[-1]
        self.mid = self.axis == 1 and len(in_shape) > 1
        self.weights = [torch.ones(c, dtype=dtype, requires_grad=True), torch.zeros(c, dtype=dtype, requires_grad=True)]
        self.state = [torch.zeros(c, dtype=dtype), torch.ones(c, dtype=dtype)]
        return in_shape

    def forward(self, x, training, noise):
        if getattr(self, 'mid', False):
            # axis = 1 (no_weight_code/subtract_model.py:264-357): the statistics run over every other axis; move the
            # normalised axis last, normalise, move it back
            perm = [0] + list(range(2, x.dim())) + [1]
            inv = [0, x.dim() - 1] + list(range(1, x.dim() - 1))
            return self._forward_last(x.permute(perm), training, noise).permute(inv)
        return self._forward_last(x, training, noise)

    def _forward_last(self, x, training, noise):
        g, b = self.weights
        if training:
            axes = tuple(range(x.dim() - 1))
            mean = x.mean(dim=axes)
            var = ((x - mean) ** 2).mean(dim=axes)
            n = x.numel() / x.shape[-1]
            with torch.no_grad():
                m = self.momentum
                sv = var.detach() * (n / (n - (1.0 + self.epsilon)))
                upd = noise.get('__updates__') if isinstance(noise, dict) else None
                if upd is not None and id(self) not in upd:
                    pass            # Keras: a layer that was non-trainable for this compiled model contributes no updates
                elif ZERO_DEBIAS:
                    # K.moving_average_update (keras/backend/tensorflow_backend.py, 2.2.4) ->
                    # tf.python.training.moving_averages.assign_moving_average(x, value, momentum, zero_debias=True)
                    # (TF 1.12 _zero_debias): biased -= (biased - value)*(1-m); local_step += 1;
                    # x -= x - biased / (1 - m**local_step), i.e. x is overwritten with the debiased average
                    if self.biased is None:
                        self.biased = [torch.zeros_like(self.state[0]), torch.zeros_like(self.state[1])]
                        self.local_step = 0
                    self.biased[0] = self.biased[0] * m + mean.detach() * (1 - m)
                    self.biased[1] = self.biased[1] * m + sv * (1 - m)
                    self.local_step += 1
                    self.state[0] = self.biased[0] / (1 - m ** self.local_step)
                    self.state[1] = self.biased[1] / (1 - m ** self.local_step)
                else:
                    self.state[0] = self.state[0] * m + mean.detach() * (1 - m)
                    self.state[1] = self.state[1] * m + sv * (1 - m)
        else:
            mean, var = self.state
        return (x - mean) / torch.sqrt(var + self.epsilon) * g + b


KINK_TOL = 1e-5      # of the tensor's scale: the forward tolerance of a float32-class implementation


def _fed_mask(x, own, ykink, cond):
    """Piecewise-linear activations have kinks where float32 and float64 can land on different sides.
    Like RNG draws, the side taken by the implementation under test may be fed in (its OUTPUT `ykink`):
    the oracle then differentiates the same linear piece -- but only where its own pre-activation is
    within rounding distance (KINK_TOL of the tensor's scale) of the kink; any other disagreement is an
    error of the implementation and raises."""
    fed = cond(torch.as_tensor(np.asarray(ykink)).reshape(x.shape))
    diff = fed != own
    if diff.any():
        lim = KINK_TOL * max(float(x.detach().abs().max()), 1e-30)
        worst = float(x.detach()[diff].abs().max())
        return fed, worst, lim
    return fed, 0.0, 1.0


def relu_like(x, lo_slope=0.0, max_value=None, ykink=None):
    """max(x,0) (or leaky) optionally clipped at max_value, with optionally fed kink decisions."""
    pos = x > 0 if lo_slope == 0.0 else x >= 0
    if ykink is not None:
        pos, worst, lim = _fed_mask(x, pos, ykink, (lambda y: y > 0) if lo_slope == 0.0 else (lambda y: y >= 0))
        assert worst <= lim, 'activation sign disagrees with the oracle away from the kink (|x|=%g > %g)' % (worst, lim)
    y = torch.where(pos, x, x * lo_slope)
    if max_value is not None:
        below = x < max_value
        if ykink is not None:
            fed = torch.as_tensor(np.asarray(ykink)).reshape(x.shape) < max_value
            diff = fed != below
            if diff.any():
                worst = float((x.detach()[diff] - max_value).abs().max())
                assert worst <= KINK_TOL * max(float(x.detach().abs().max()), 1e-30), 'clip side disagrees'
            below = fed
        y = torch.where(below, y, torch.full_like(y, max_value))
    return y


def apply_activation(x, act, ykink=None):
    if act in (None, 'linear'):
        return x
    if act == 'relu':
        return relu_like(x, ykink=ykink)
    if act == 'tanh':
        return torch.tanh(x)
    if act == 'sigmoid':
        return torch.sigmoid(x)
    if act == 'elu':            # keras.activations.elu, alpha = 1.0
        return torch.where(x > 0, x, torch.expm1(x))
    raise ValueError(act)


class Activation(Layer):
    kind = 'act'

    def __init__(self, act):
        super().__init__()
        self.act = act

    def forward(self, x, training, noise):
        return apply_activation(x, self.act, _kink(noise, self))


class LeakyReLU(Layer):
    kind = 'act'

    def __init__(self, alpha=0.3):
        super().__init__()
        self.alpha = alpha

    def forward(self, x, training, noise):
        return relu_like(x, lo_slope=self.alpha, ykink=_kink(noise, self))  # [A7]


class ReLU(Layer):
    kind = 'act'

    def __init__(self, max_value=None):
        super().__init__()
        self.max_value = max_value

    def forward(self, x, training, noise):
        return relu_like(x, max_value=self.max_value, ykink=_kink(noise, self))


class _NoiseLayer(Layer):
    """Dropout family [A6]; the random tensor is taken from ``noise[name]`` when
    given, else drawn and recorded in ``noise[name]`` for the caller to reuse."""
    kind = 'noise'

    def draw(self, x, gen):
        raise NotImplementedError

    def forward(self, x, training, noise):
        if not training:
            return x
        r = noise.get(self.name)
        if r is None:
            r = self.draw(x, noise.get('__gen__'))
            noise[self.name] = r
        return self.apply(x, r.to(x.dtype))


class Dropout(_NoiseLayer):
    def __init__(self, rate):
        super().__init__()
        self.rate = rate

    def draw(self, x, gen):   # keep mask (1 = keep)
        return (torch.rand(x.shape, generator=gen, dtype=torch.float64) >= self.rate).to(torch.float64)

    def apply(self, x, r):
        return x * r / (1.0 - self.rate)


class GaussianDropout(_NoiseLayer):
    def __init__(self, rate):
        super().__init__()
        self.rate = rate

    def draw(self, x, gen):   # standard normals
        return torch.randn(x.shape, generator=gen, dtype=torch.float64)

    def apply(self, x, r):
        return x * (1.0 + r * math.sqrt(self.rate / (1.0 - self.rate)))


class GaussianNoise(_NoiseLayer):
    def __init__(self, stddev):
        super().__init__()
        self.stddev = stddev

    def draw(self, x, gen):
        return torch.randn(x.shape, generator=gen, dtype=torch.float64)

    def apply(self, x, r):
        return x + r * self.stddev


class Reshape(Layer):
    kind = 'shape'

    def __init__(self, target):
        super().__init__()
        self.target = tuple(target)

    def build(self, in_shape, gen, dtype):
        n = int(np.prod(in_shape))
        t = list(self.target)
        if -1 in t:
            t[t.index(-1)] = n // int(-np.prod(t))
        self.out = tuple(t)
        return self.out

    def forward(self, x, training, noise):
        return x.reshape((x.shape[0],) + self.out)   # row-major [A3]


class Flatten(Layer):
    kind = 'shape'

    def build(self, in_shape, gen, dtype):
        return (int(np.prod(in_shape)),)

    def forward(self, x, training, noise):
        return x.reshape(x.shape[0], -1)


class UpSampling1D(Layer):
    kind = 'shape'

    def __init__(self, size=2):
        super().__init__()
        self.size = size

    def build(self, in_shape, gen, dtype):
        return (in_shape[0] * self.size, in_shape[1])

    def forward(self, x, training, noise):
        return x.repeat_interleave(self.size, dim=1)   # [A4]


class MaxPooling1D(Layer):
    kind = 'shape'

    def __init__(self, pool_size=2):
        super().__init__()
        self.p = pool_size

    def build(self, in_shape, gen, dtype):
        return (in_shape[0] // self.p, in_shape[1])

    def forward(self, x, training, noise):
        ykink = _kink(noise, self)
        if ykink is None:
            return F.max_pool1d(x.permute(0, 2, 1), self.p).permute(0, 2, 1)
        # Like a ReLU kink, a near-tie inside a pooling window is a discontinuity of the gradient: the window element
        # the implementation under test selected (its INDEX within the window, fed in) is differentiated here too --
        # but only where it is within KINK_TOL (of the tensor's scale) of the true maximum; anything else raises.
        B, L, C = x.shape
        n = L // self.p
        win = x[:, :n * self.p].reshape(B, n, self.p, C)
        idx = torch.as_tensor(np.asarray(ykink['pool_idx'])).to(torch.int64).reshape(B, n, 1, C)
        chosen = torch.gather(win, 2, idx).squeeze(2)
        gap = float((win.detach().max(dim=2).values - chosen.detach()).max())
        lim = KINK_TOL * max(float(x.detach().abs().max()), 1e-30)
        assert gap <= lim, 'pooling selection disagrees with the oracle away from a tie (gap %g > %g)' % (gap, lim)
        return chosen


class GlobalAveragePooling(Layer):
    """Keras GlobalAveragePooling1D / 2D, channels_last: mean over every axis between batch and channels."""
    kind = 'shape'

    def build(self, in_shape, gen, dtype):
        return (in_shape[-1],)

    def forward(self, x, training, noise):
        return x.mean(dim=tuple(range(1, x.dim() - 1)))


class StackResidual(Layer):
    """bbhMahoGANy.py:164-188 MyLayer: stack([x, const-x], axis=2) -> (B,L,2,1)."""
    kind = 'mylayer'

    def __init__(self, const):
        super().__init__()
        self.const = torch.as_tensor(np.asarray(const))

    def build(self, in_shape, gen, dtype):
        self.const = self.const.to(dtype).reshape(in_shape)
        return (in_shape[0], 2, 1)

    def forward(self, x, training, noise):
        return torch.stack([x, self.const - x], dim=2)


class ResidualMoments(Layer):
    """tests/burstMahoGANy.py:100-125 MyLayer: stack([mean(d), mean(d^2)]), d=const-x,
    batch-global scalars, output shape (2,)."""
    kind = 'mylayer'

    def __init__(self, const):
        super().__init__()
        self.const = torch.as_tensor(np.asarray(const))

    def build(self, in_shape, gen, dtype):
        self.const = self.const.to(dtype).reshape(in_shape)
        return (2,)

    def forward(self, x, training, noise):
        d = self.const - x
        return torch.stack([d.mean(), (d * d).mean()])


class Sequential(Layer):
    kind = 'model'

    def __init__(self, layers=None):
        super().__init__()
        self.layers = []
        self.in_shape = None
        self.out_shape = None
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        self.layers.append(layer)

    def all_layers(self):
        out = []
        for l in self.layers:
            out += l.all_layers()
        return out

    def build(self, in_shape, gen=None, dtype=torch.float64):
        if self.out_shape is not None:      # already built (shared sub-model)
            return self.out_shape
        self.in_shape = tuple(in_shape)
        s = tuple(in_shape)
        for l in self.layers:
            s = tuple(l.build(s, gen, dtype))
        self.out_shape = s
        _name_layers(self)
        return s

    def forward(self, x, training, noise):
        for l in self.layers:
            x = l.forward(x, training, noise)
        return x

    # ---- Keras protocol -------------------------------------------------
    def compile(self, loss, optimizer, metrics=None):
        _compile(self, loss, optimizer)

    def predict(self, x):
        with torch.no_grad():
            y = self.forward(_t(x, self), False, {})
        return _np(y)

    def train_on_batch(self, x, y, noise=None):
        return _train_on_batch(self, x, y, noise)

    def get_weights(self):
        out = []
        for l in self.all_layers():
            out += [w.detach().numpy().copy() for w in l.weights] + [s.numpy().copy() for s in l.state]
        return out

    def set_weights(self, ws):
        i = 0
        for l in self.all_layers():
            for w in l.weights:
                w.data = torch.as_tensor(np.asarray(ws[i])).to(w.dtype).reshape(w.shape).clone()
                i += 1
            for k in range(len(l.state)):
                l.state[k] = torch.as_tensor(np.asarray(ws[i])).to(l.state[k].dtype).clone()
                i += 1
        assert i == len(ws)


class BranchModel(Sequential):
    """Functional ``Model(inputs, [out_a, out_b])`` with branches sharing the input
    (bbhMahoGANy.py:357-404)."""

    def __init__(self, branches):
        Layer.__init__(self)
        self.branches = branches
        # Keras 2.2.4 `model.layers` (keras/engine/network.py _map_graph_network): by depth from the outputs, deepest
        # first; ties go to the layer met first by the traversal that starts at the first output, i.e. to the
        # earlier branch.  get_weights / set_weights / the gradient list follow this order.
        keyed = [(-(len(b) - 1 - i), bi, l) for bi, b in enumerate(branches) for i, l in enumerate(b)]
        self.layers = [l for _, _, l in sorted(keyed, key=lambda t: (t[0], t[1]))]
        self.in_shape = self.out_shape = None

    def build(self, in_shape, gen=None, dtype=torch.float64):
        self.in_shape = tuple(in_shape)
        outs = []
        for b in self.branches:
            s = tuple(in_shape)
            for l in b:
                s = tuple(l.build(s, gen, dtype))
            outs.append(s)
        self.out_shape = outs
        _name_layers(self)
        return outs

    def forward(self, x, training, noise):
        outs = []
        for b in self.branches:
            h = x
            for l in b:
                h = l.forward(h, training, noise)
            outs.append(h)
        return outs

    def predict(self, x):
        with torch.no_grad():
            ys = self.forward(_t(x, self), False, {})
        return [_np(y) for y in ys]


_PREFIX = {Dense: 'dense', Conv1D: 'conv1d', Conv2D: 'conv2d', Conv2DTranspose: 'conv2d_transpose', BatchNormalization: 'batch_normalization',
           Activation: 'activation', LeakyReLU: 'leaky_re_lu', ReLU: 're_lu', Dropout: 'dropout',
           GaussianDropout: 'gaussian_dropout', GaussianNoise: 'gaussian_noise', Reshape: 'reshape',
           Flatten: 'flatten', UpSampling1D: 'up_sampling1d', MaxPooling1D: 'max_pooling1d',
           StackResidual: 'my_layer', ResidualMoments: 'my_layer', GlobalAveragePooling: 'global_average_pooling1d'}


_NAME_COUNTS = {}


def clear_session():
    """Reset Keras' per-session layer-name counters."""
    _NAME_COUNTS.clear()


def _name_layers(model):
    cnt = _NAME_COUNTS          # names are unique per session, as in Keras
    for l in model.all_layers():
        if l.name is None:
            p = _PREFIX.get(type(l), 'layer')
            cnt[p] = cnt.get(p, 0) + 1
            l.name = '%s_%d' % (p, cnt[p])


def _dtype(model):
    for l in model.all_layers():
        if l.weights:
            return l.weights[0].dtype
    return torch.float64


def _t(x, model):
    return torch.as_tensor(np.asarray(x)).to(_dtype(model))


def _np(y):
    return y.detach().numpy()


# ---- optimizers [A8] ---------------------------------------------------------

class Adam:
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=EPS, decay=0.0):
        self.lr, self.b1, self.b2, self.eps, self.decay = lr, beta_1, beta_2, epsilon, decay
        self.iterations = 0
        self.m, self.v = {}, {}

    def step(self, params, grads):
        lr = self.lr
        if self.decay > 0:
            lr = lr * (1.0 / (1.0 + self.decay * self.iterations))
        t = self.iterations + 1
        lr_t = lr * (math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t))
        with torch.no_grad():
            for p, g in zip(params, grads):
                k = id(p)
                m = self.m.get(k, torch.zeros_like(p))
                v = self.v.get(k, torch.zeros_like(p))
                m = self.b1 * m + (1 - self.b1) * g
                v = self.b2 * v + (1 - self.b2) * g * g
                p -= lr_t * m / (torch.sqrt(v) + self.eps)
                self.m[k], self.v[k] = m, v
        self.iterations += 1


class SGD:
    def __init__(self, lr=0.01, decay=0.0):
        self.lr, self.decay, self.iterations = lr, decay, 0

    def step(self, params, grads):
        lr = self.lr
        if self.decay > 0:
            lr = lr * (1.0 / (1.0 + self.decay * self.iterations))
        with torch.no_grad():
            for p, g in zip(params, grads):
                p -= lr * g
        self.iterations += 1


# ---- losses / metrics [A9][A10] ------------------------------------------------

def binary_crossentropy(y_true, y_pred):
    p = torch.clamp(y_pred, EPS, 1 - EPS)
    x = torch.log(p / (1 - p))
    l = torch.clamp(x, min=0) - x * y_true + torch.log1p(torch.exp(-torch.abs(x)))
    return l.mean(dim=-1)


def mean_squared_error(y_true, y_pred):
    return ((y_pred - y_true) ** 2).mean(dim=-1)


def chisquare_loss(n_sig):
    """bbhMahoGANy.py:146-162."""
    return lambda y_true, y_pred: (((y_true - y_pred) ** 2) / (n_sig ** 2)).sum(dim=-1)


LOSSES = {'binary_crossentropy': binary_crossentropy, 'mean_squared_error': mean_squared_error, 'mse': mean_squared_error}


def accuracy(y_true, y_pred, loss_name):
    if y_pred.shape[-1] == 1 or loss_name == 'binary_crossentropy':
        return (y_true == torch.round(y_pred)).to(y_pred.dtype).mean()
    yt = y_true if y_true.dim() > 1 else y_true[None]
    yp = y_pred if y_pred.dim() > 1 else y_pred[None].expand_as(yt)
    return (yt.argmax(-1) == yp.argmax(-1)).to(y_pred.dtype).mean()


def _compile(model, loss, optimizer):
    model.loss_name = loss if isinstance(loss, str) else 'custom'
    model.loss_fn = LOSSES[loss] if isinstance(loss, str) else loss
    model.optimizer = optimizer
    # trainable set is frozen at compile time [A11]
    model.collected = [w for l in model.all_layers() if l.trainable for w in l.weights]
    model.update_ids = set(id(l) for l in model.all_layers() if l.trainable)


def _labels(y, like):
    y = torch.as_tensor(np.asarray(y, dtype=np.float64)).to(like.dtype)
    if y.dim() == 1 and like.dim() == 2:
        y = y[:, None]
    return y


def _train_on_batch(model, x, y, noise=None):
    noise = dict({} if noise is None else noise)
    noise['__updates__'] = model.update_ids
    noise['__losses__'] = []          # regularisation terms of THIS forward pass
    out = model.forward(_t(x, model), True, noise)
    outs = out if isinstance(out, list) else [out]
    ys = y if isinstance(out, list) else [y]
    losses, accs = [], []
    for o, yy in zip(outs, ys):
        yt = _labels(yy, o)
        losses.append(model.loss_fn(yt, o).mean())
        accs.append(accuracy(yt, o.detach(), model.loss_name))
    total = sum(losses) + sum(noise.get('__losses__', []))        # + keras `model.losses` (regularisers)
    grads = torch.autograd.grad(total, model.collected, allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for p, g in zip(model.collected, grads)]
    model.last_grads = [g.detach().numpy().copy() for g in grads]
    model.optimizer.step(model.collected, grads)
    if len(outs) == 1:
        return [float(total.detach()), float(accs[0])]
    return [float(total.detach())] + [float(l.detach()) for l in losses] + [float(a) for a in accs]


def set_trainable(model, trainable):
    """bbhMahoGANy.py:797-809."""
    model.trainable = trainable
    for l in model.all_layers():
        l.trainable = trainable


# =============================================================================
# Builders
# =============================================================================

def bbh_generator_model(n_pix=1024):
    """bbhMahoGANy.py:212-295."""
    act, mom, dr = 'tanh', 0.99, 0.2
    L = [Dense(256 * int(n_pix / 2)), BatchNormalization(mom), Activation(act), Dropout(dr), Reshape((int(n_pix / 2), 256))]
    for i, (f, s, up) in enumerate([(64, 2, True), (128, 1, True), (256, 1, False), (512, 1, False), (1024, 1, False)]):
        if up:
            L.append(UpSampling1D(2))
        L += [Conv1D(f, 5, strides=s, padding='same'), BatchNormalization(mom), Activation(act), Dropout(dr)]
    L += [Conv1D(1, 5, padding='same'), Activation('linear')]
    m = Sequential(L)
    m.input_shape = (100,)
    return m


def bbh_signal_pe_model(n_pix=1024):
    """bbhMahoGANy.py:356-404 (comb_pe_model=False branch)."""
    mc = [Conv1D(64, 5, strides=2, padding='same'), Activation('relu')]
    for f in (128, 256, 512):
        mc += [Conv1D(f, 5, strides=2), Activation('relu')]
    mc += [Flatten(), Dense(1), Activation('relu')]
    q = [Conv1D(64, 5, strides=1, padding='same'), Activation('relu')]
    for f, s in ((128, 1), (256, 1), (512, 2), (1024, 2)):
        q += [Conv1D(f, 5, strides=s), Activation('relu')]
    q += [Flatten(), Dense(1), ReLU(max_value=1.0)]
    # Keras creates layers in call order: mc branch first, then q branch
    m = BranchModel([mc, q])
    m.input_shape = (n_pix, 1)
    return m


def bbh_signal_discriminator_model(n_pix=1024):
    """bbhMahoGANy.py:408-498."""
    L = [Conv2D(256, (5, 5), strides=(2, 1), padding='same'), LeakyReLU(0.2), Dropout(0.4),
         Conv2D(512, (5, 5), strides=(2, 1), padding='same'), LeakyReLU(0.2), Dropout(0.4),
         Flatten(), Dense(1), Activation('sigmoid')]
    m = Sequential(L)
    m.input_shape = (n_pix, 2, 1)
    return m


def burst_generator_model(n_pix=512):
    """tests/burstMahoGANy.py:127-251."""
    L = [Dense(256 * int(n_pix / 2)), Activation('relu'), Reshape((int(n_pix / 2), 256)), UpSampling1D(2)]
    for f in (64, 64, 256, 512):
        L += [Conv1D(f, 5, strides=1, padding='same'), Activation('relu'), GaussianDropout(0.3)]
    L += [Conv1D(1, 5, padding='same'), Activation('tanh')]
    m = Sequential(L)
    m.input_shape = (100,)
    return m


def burst_signal_pe_model(n_pix=512):
    """tests/burstMahoGANy.py:263-293."""
    m = Sequential([Conv1D(64, 5, strides=2, padding='same'), Activation('relu'), Conv1D(128, 5, strides=2),
                    Activation('relu'), Flatten(), Dense(1024), Activation('relu'), Dense(2), Activation('linear')])
    m.input_shape = (n_pix, 1)
    return m


def burst_signal_discriminator_model(n_pix=512):
    """tests/burstMahoGANy.py:295-402."""
    m = Sequential([Conv1D(64, 5, strides=1, padding='same'), Activation('tanh'), MaxPooling1D(2),
                    Conv1D(128, 5, strides=1), Activation('tanh'), MaxPooling1D(2), Flatten(),
                    Dense(1024), Activation('tanh'), Dense(1), Activation('sigmoid')])
    m.input_shape = (n_pix, 1)
    return m


def wvf_get_generative(noise_dim=10, dense_dim=300, out_dim=8192):
    """train_on_wvf_version/nn.py:72-81."""
    m = Sequential([Dense(dense_dim), Activation('relu'), Dense(150), Activation('relu'), Dense(out_dim, activation='tanh')])
    m.input_shape = (noise_dim,)
    return m


def wvf_get_discriminative(in_dim=8192, drate=.25, n_channels=25, conv_sz=5):
    """train_on_wvf_version/nn.py:83-93."""
    m = Sequential([Reshape((-1, 1)), Conv1D(n_channels, conv_sz, activation='relu'), Dropout(drate), Flatten(),
                    Dense(n_channels), Dense(2, activation='sigmoid')])
    m.input_shape = (in_dim,)
    return m


def two_model_get_generative(noise_dim=1, out_dim=50):
    """2_model_version/weight_version/no_mode_collapse_network.py:62-106 (the transposed-convolution network)."""
    L = [Reshape((-1, 1, 1)), BatchNormalization()]
    for f, k in ((128, 4), (64, 8), (32, 16), (16, 32)):
        L += [Conv2DTranspose(f, (1, k), strides=(1, 1), padding='valid', activation='relu'), BatchNormalization()]
    L += [Flatten(), BatchNormalization(), Dense(out_dim, activation='relu'), BatchNormalization(), Dense(out_dim)]
    m = Sequential(L)
    m.input_shape = (1, noise_dim)
    return m


def two_model_get_discriminative(in_dim=50, n_channels=50, conv_sz=16, leak=0.2):
    """The discriminator stored in 2_model_version/weight_version/d_model.hdf5 (model_config attribute)."""
    m = Sequential([Reshape((-1, 1)), Conv1D(n_channels, conv_sz), LeakyReLU(leak), Flatten(), Dense(50, activation='tanh'),
                    Dense(2, activation='sigmoid')])
    m.input_shape = (in_dim,)
    return m


def subtract_get_generative(noise_dim=10, out_dim=50):
    """2_model_version/weight_version/subtract_model.py:199-251: the transposed-convolution generator with ELU
    activations and, on its first Conv2DTranspose, activity_regularizer=l1(0.001) and kernel_regularizer=l2(0.01)."""
    L = [Reshape((-1, 1, 1)), BatchNormalization(),
         Conv2DTranspose(128, (1, 4), activation='elu', kernel_regularizer=(0.0, 0.01), activity_regularizer=(0.001, 0.0)),
         BatchNormalization()]
    for f, k in ((64, 8), (32, 16), (16, 32)):
        L += [Conv2DTranspose(f, (1, k), activation='elu'), BatchNormalization()]
    L += [Flatten(), BatchNormalization(), Dense(out_dim, activation='elu'), BatchNormalization(), Dense(out_dim)]
    m = Sequential(L)
    m.input_shape = (1, noise_dim)
    return m


def subtract_get_discriminative(in_dim=50, n_channels=50, drate=0.3):
    """2_model_version/weight_version/subtract_model.py:253-291."""
    m = Sequential([Reshape((-1, 1)), Conv1D(50, 16), LeakyReLU(0.2), Dropout(drate), Flatten(), Dense(n_channels),
                    Dropout(drate), Dense(2, activation='sigmoid')])
    m.input_shape = (in_dim,)
    return m


def nw_get_discriminative(in_dim=50, gauss_noise=1.6):
    """2_model_version/no_weight_code/subtract_model.py:322-390: three Conv1D(.., 8, tanh) -> LeakyReLU(0.2) ->
    GaussianNoise(1.6) -> BatchNormalization(axis=1) blocks, GlobalAveragePooling1D, Dense(2, sigmoid)."""
    L = [Reshape((-1, 1))]
    for f in (128, 256, 512):
        L += [Conv1D(f, 8, activation='tanh'), LeakyReLU(0.2), GaussianNoise(gauss_noise), BatchNormalization(axis=1)]
    L += [GlobalAveragePooling(), Dense(2, activation='sigmoid')]
    m = Sequential(L)
    m.input_shape = (in_dim,)
    return m


def build(model, seed=0, dtype=torch.float64):
    gen = torch.Generator().manual_seed(seed)
    model.build(model.input_shape, gen, dtype)
    return model
