"""File formats of the reference pipeline: Keras HDF5 weights / models and Python-2 cPickle
template banks, so that files written by the reference load here and files written here load there.

* ``save_weights / load_weights / save_model / load_model`` -- the layout Keras 2.1-2.2 writes
  (``keras/engine/saving.py``): root attrs ``layer_names, backend, keras_version``; one group per layer with
  attr ``weight_names`` and one dataset per weight (``dense_1/dense_1/kernel:0``); full models add the
  ``model_config`` / ``training_config`` JSON attrs, keep the weights under ``model_weights`` and the optimizer
  state under ``optimizer_weights``.  Call sites: bbhMahoGANy.py:1133-1142 (load), :1173 (``signal_pe.save``),
  :1373-1375 (``save_weights``).  HDF5 itself is handled by ``gennet_b200.hdf5`` (no h5py in this stack).
* ``load_pickle / dump_pickle`` -- ``cPickle`` protocol-2 ``.sav/.pkl`` files of
  ``[ndarray (size,1,fs) float64, ndarray]`` and of ``list[bbhparams]`` (gw_template_maker.py:840-849,
  bbhMahoGANy.py:968-998,1027-1029).  ``bbhparams`` is an old-style class pickled as ``__main__.bbhparams``;
  it is mapped onto ``gennet_b200.synth.bbhparams`` on load and written back under the same global name with
  the classic-instance opcodes, and NumPy globals are written under their Python-2-era module path
  (``numpy.core.multiarray``), so Python 2 + NumPy 1.15 can read the result.
* ``load_npy / save_npy`` -- plain ``.npy`` (an addition: the reference itself never uses it).
"""
import io as _io
import json
import os
import pickle

import numpy as np

from . import hdf5

KERAS_VERSION = '2.2.4'
BACKEND = 'tensorflow'


# ======================================================================================================= pickles
class _Py2Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == 'bbhparams':
            from . import synth
            return synth.bbhparams
        if module.startswith('numpy.core'):          # NumPy >= 2 keeps these as deprecated aliases of numpy._core
            import importlib
            try:
                mod = importlib.import_module(module.replace('numpy.core', 'numpy._core', 1))
                return getattr(mod, name)
            except (ImportError, AttributeError):
                pass
        return super().find_class(module, name)


def load_pickle(path_or_file):
    """Load a ``.sav/.pkl`` file written by Python 2 ``cPickle`` (or by :func:`dump_pickle`)."""
    if hasattr(path_or_file, 'read'):
        return _Py2Unpickler(path_or_file, encoding='latin1').load()
    with open(path_or_file, 'rb') as f:
        return _Py2Unpickler(f, encoding='latin1').load()


class _Py2Pickler(pickle._Pickler):
    """Protocol-2 pickler whose output Python 2.7 + NumPy 1.15 can load."""

    def __init__(self, f):
        super().__init__(f, protocol=2, fix_imports=True)

    def save_global(self, obj, name=None):
        module = getattr(obj, '__module__', None) or ''
        if module.startswith('numpy._core'):
            nm = name or getattr(obj, '__qualname__', None) or obj.__name__
            self.write(pickle.GLOBAL + module.replace('numpy._core', 'numpy.core', 1).encode('ascii') + b'\n' +
                       nm.encode('ascii') + b'\n')
            self.memoize(obj)
            return
        super().save_global(obj, name)

    def save(self, obj, save_persistent_id=True):
        from . import synth
        if type(obj) is synth.bbhparams:
            # classic (old-style) instance: MARK, class global, OBJ, then the attribute dict and BUILD --
            # what cPickle wrote for `__main__.bbhparams` (gw_template_maker.py:69-85, bbhMahoGANy.py:129-144)
            x = self.memo.get(id(obj))
            if x is not None:
                self.write(self.get(x[0]))
                return
            self.write(pickle.MARK + pickle.GLOBAL + b'__main__\nbbhparams\n' + pickle.OBJ)
            self.memoize(obj)
            state = dict(obj.__dict__)
            self.save({k: _py2_scalar(v) for k, v in state.items()})
            self.write(pickle.BUILD)
            return
        super().save(obj, save_persistent_id)


def _py2_scalar(v):
    if isinstance(v, np.generic):
        return v.item()
    return v


def dump_pickle(obj, path_or_file):
    """Write ``obj`` the way ``cPickle.dump(obj, f, protocol=cPickle.HIGHEST_PROTOCOL)`` does under Python 2
    (gw_template_maker.py:840-849)."""
    if hasattr(path_or_file, 'write'):
        _Py2Pickler(path_or_file).dump(obj)
        return
    with open(path_or_file, 'wb') as f:
        _Py2Pickler(f).dump(obj)


def load_template_bank(ts_path, par_path=None):
    """``[ts (size,1,fs) float64, yval]`` and the parameter list of one template block
    (bbhMahoGANy.py:968-998): returns (ts array, yval, params list or None)."""
    ts = load_pickle(ts_path)
    par = load_pickle(par_path) if par_path is not None else None
    return np.asarray(ts[0]), np.asarray(ts[1]), par


def save_template_bank(ts, yval, params, ts_path, par_path=None):
    dump_pickle([np.asarray(ts, dtype=np.float64), np.asarray(yval)], ts_path)
    if par_path is not None:
        dump_pickle(list(params), par_path)


def save_npy(path, arr):
    np.save(path, np.asarray(arr))


def load_npy(path):
    return np.load(path, allow_pickle=False)


# ================================================================================================== Keras configs
def _initializer(name):
    if name == 'glorot_uniform':
        return {'class_name': 'VarianceScaling', 'config': {'scale': 1.0, 'mode': 'fan_avg', 'distribution': 'uniform',
                                                            'seed': None}}
    return {'class_name': str(name), 'config': {}}


_ZEROS = {'class_name': 'Zeros', 'config': {}}
_ONES = {'class_name': 'Ones', 'config': {}}


def _kernel_cfg(layer):
    return {'use_bias': True, 'kernel_initializer': _initializer('glorot_uniform'), 'bias_initializer': _ZEROS,
            'kernel_regularizer': None, 'bias_regularizer': None, 'activity_regularizer': None,
            'kernel_constraint': None, 'bias_constraint': None}


def layer_config(layer):
    """(Keras class name, config dict) of one layer, in the form ``keras.layers.deserialize`` accepts."""
    from . import nn
    cfg = {'name': layer.name, 'trainable': bool(layer.trainable)}
    if layer._input_shape_arg is not None:
        cfg['batch_input_shape'] = [None] + list(layer._input_shape_arg)
        cfg['dtype'] = 'float32'
    t = type(layer)
    if t is nn.InputLayer:
        return 'InputLayer', {'batch_input_shape': [None] + list(layer.output_shape), 'dtype': 'float32',
                              'sparse': False, 'name': layer.name}
    if t is nn.Dense:
        cfg.update(units=layer.units, activation=layer.activation or 'linear', **_kernel_cfg(layer))
        return 'Dense', cfg
    if t is nn.Conv1D:
        cfg.update(filters=layer.filters, kernel_size=[layer.k], strides=[layer.s], padding=layer.padding,
                   dilation_rate=[1], activation=layer.activation or 'linear', **_kernel_cfg(layer))
        return 'Conv1D', cfg
    if t is nn.Conv2D:
        cfg.update(filters=layer.filters, kernel_size=[layer.kh, layer.kw], strides=[layer.sh, layer.sw],
                   padding=layer.padding, data_format='channels_last', dilation_rate=[1, 1], activation='linear',
                   **_kernel_cfg(layer))
        return 'Conv2D', cfg
    if t is nn.Conv2DTranspose:
        cfg.update(filters=layer.filters, kernel_size=[layer.kh, layer.kw], strides=[layer.sh, layer.sw],
                   padding=layer.padding, data_format='channels_last', activation=layer.activation or 'linear',
                   **_kernel_cfg(layer))
        for key in ('kernel_regularizer', 'activity_regularizer'):
            r = getattr(layer, key, None)
            if r is not None:
                cfg[key] = {'class_name': 'L1L2', 'config': r.get_config()}
        return 'Conv2DTranspose', cfg
    if t in (nn.GlobalAveragePooling1D, nn.GlobalAveragePooling2D):
        if t is nn.GlobalAveragePooling2D:
            cfg.update(data_format='channels_last')
        return t.__name__, cfg
    if t is nn.BatchNormalization:
        cfg.update(axis=layer.axis, momentum=layer.momentum, epsilon=layer.epsilon, center=True, scale=True,
                   beta_initializer=_ZEROS, gamma_initializer=_ONES, moving_mean_initializer=_ZEROS,
                   moving_variance_initializer=_ONES, beta_regularizer=None, gamma_regularizer=None,
                   beta_constraint=None, gamma_constraint=None)
        return 'BatchNormalization', cfg
    if t is nn.Activation:
        cfg.update(activation=layer.activation)
        return 'Activation', cfg
    if t is nn.LeakyReLU:
        cfg.update(alpha=layer.param)
        return 'LeakyReLU', cfg
    if t is nn.ReLU:
        cfg.update(max_value=(layer.param if layer.code == nn._lib.ACT_RELU_MAX else None))
        return 'ReLU', cfg
    if t is nn.Dropout:
        cfg.update(rate=layer.rate, noise_shape=None, seed=None)
        return 'Dropout', cfg
    if t is nn.GaussianDropout:
        cfg.update(rate=layer.rate)
        return 'GaussianDropout', cfg
    if t is nn.GaussianNoise:
        cfg.update(stddev=layer.rate)
        return 'GaussianNoise', cfg
    if t is nn.Flatten:
        cfg.update(data_format='channels_last')
        return 'Flatten', cfg
    if t is nn.Reshape:
        cfg.update(target_shape=list(layer.target))
        return 'Reshape', cfg
    if t is nn.UpSampling1D:
        cfg.update(size=layer.size)
        return 'UpSampling1D', cfg
    if t is nn.MaxPooling1D:
        cfg.update(pool_size=[layer.pool], strides=[layer.pool], padding='valid')
        return 'MaxPooling1D', cfg
    if t in (nn.StackResidual, nn.ResidualMoments):
        # the reference's custom `MyLayer` (bbhMahoGANy.py:164-188): Keras needs custom_objects to load it
        cfg.update(const=np.asarray(layer._const_host, dtype=np.float64).ravel().tolist(),
                   gn_kind=t.__name__)
        return 'MyLayer', cfg
    if isinstance(layer, nn.Model):
        return model_config(layer)['class_name'], model_config(layer)['config']
    raise ValueError('no Keras config for layer type %s' % t.__name__)


def model_config(model):
    """``{'class_name': 'Sequential'|'Model', 'config': ...}`` as ``model.to_json()`` of Keras 2.2.4 gives it."""
    from . import nn
    if isinstance(model, nn.Sequential):
        layers = []
        for l in model.layers:
            cn, cfg = layer_config(l)
            layers.append({'class_name': cn, 'config': cfg})
        if layers and 'batch_input_shape' not in layers[0]['config']:
            layers[0]['config']['batch_input_shape'] = [None] + list(model.input_shape)
            layers[0]['config'].setdefault('dtype', 'float32')
        return {'class_name': 'Sequential', 'config': {'name': model.name, 'layers': layers}}
    # functional graph: nodes in topological order, every layer called exactly once in the reference models
    in_layer = model._in_node.layer
    names = {id(model._in_node): in_layer.name}
    out = [{'name': in_layer.name, 'class_name': 'InputLayer', 'config': layer_config(in_layer)[1], 'inbound_nodes': []}]
    for n in model._order:
        cn, cfg = layer_config(n.layer)
        names[id(n)] = n.layer.name
        inbound = [[[names[id(i)], 0, 0, {}] for i in n.inputs]]
        out.append({'name': n.layer.name, 'class_name': cn, 'config': cfg, 'inbound_nodes': inbound})
    return {'class_name': 'Model',
            'config': {'name': model.name, 'layers': out, 'input_layers': [[in_layer.name, 0, 0]],
                       'output_layers': [[n.layer.name, 0, 0] for n in model._out_nodes]}}


def _layer_from_config(class_name, cfg, custom_objects):
    from . import nn
    cfg = dict(cfg)
    kw = {'name': cfg.get('name'), 'trainable': cfg.get('trainable', True)}
    if cfg.get('batch_input_shape') is not None and class_name != 'InputLayer':
        kw['input_shape'] = tuple(cfg['batch_input_shape'][1:])

    def act(a):
        return None if a in (None, 'linear') else a
    first = lambda v: v[0] if isinstance(v, (list, tuple)) else v
    if class_name == 'Dense':
        return nn.Dense(cfg['units'], activation=act(cfg.get('activation')), **kw)
    if class_name in ('Conv1D', 'Convolution1D'):
        return nn.Conv1D(cfg['filters'], first(cfg['kernel_size']), strides=first(cfg.get('strides', 1)),
                         padding=cfg.get('padding', 'valid'), activation=act(cfg.get('activation')), **kw)
    if class_name in ('Conv2D', 'Convolution2D'):
        return nn.Conv2D(cfg['filters'], tuple(cfg['kernel_size']), strides=tuple(cfg.get('strides', (1, 1))),
                         padding=cfg.get('padding', 'valid'), **kw)
    if class_name in ('Conv2DTranspose', 'Deconvolution2D'):
        def reg(c):
            return None if not c else nn.Regularizer(l1=c['config'].get('l1', 0.0), l2=c['config'].get('l2', 0.0))
        return nn.Conv2DTranspose(cfg['filters'], tuple(cfg['kernel_size']), strides=tuple(cfg.get('strides', (1, 1))),
                                  padding=cfg.get('padding', 'valid'), activation=act(cfg.get('activation')),
                                  kernel_regularizer=reg(cfg.get('kernel_regularizer')),
                                  activity_regularizer=reg(cfg.get('activity_regularizer')), **kw)
    if class_name == 'GlobalAveragePooling1D':
        return nn.GlobalAveragePooling1D(**kw)
    if class_name == 'GlobalAveragePooling2D':
        return nn.GlobalAveragePooling2D(**kw)
    if class_name == 'BatchNormalization':
        ax = cfg.get('axis', -1)
        ax = ax[0] if isinstance(ax, (list, tuple)) else ax
        return nn.BatchNormalization(momentum=cfg.get('momentum', 0.99), epsilon=cfg.get('epsilon', 1e-3), axis=ax, **kw)
    if class_name == 'Activation':
        return nn.Activation(cfg['activation'], **kw)
    if class_name == 'LeakyReLU':
        return nn.LeakyReLU(alpha=float(cfg.get('alpha', 0.3)), **kw)
    if class_name == 'ReLU':
        return nn.ReLU(max_value=cfg.get('max_value'), **kw)
    if class_name == 'Dropout':
        return nn.Dropout(cfg['rate'], **kw)
    if class_name == 'GaussianDropout':
        return nn.GaussianDropout(cfg['rate'], **kw)
    if class_name == 'GaussianNoise':
        return nn.GaussianNoise(cfg['stddev'], **kw)
    if class_name == 'Flatten':
        return nn.Flatten(**kw)
    if class_name == 'Reshape':
        return nn.Reshape(tuple(cfg['target_shape']), **kw)
    if class_name == 'UpSampling1D':
        return nn.UpSampling1D(first(cfg.get('size', 2)), **kw)
    if class_name == 'MaxPooling1D':
        return nn.MaxPooling1D(first(cfg.get('pool_size', 2)), **kw)
    if class_name == 'MyLayer':
        if 'const' in cfg:
            cls = getattr(nn, cfg.get('gn_kind', 'StackResidual'))
            return cls(np.asarray(cfg['const'], dtype=np.float32), **kw)
        if custom_objects and 'MyLayer' in custom_objects:
            return custom_objects['MyLayer'](**{k: v for k, v in cfg.items() if k not in ('trainable',)})
        raise ValueError("MyLayer needs custom_objects={'MyLayer': ...} (its constant is not stored by the reference)")
    if class_name in ('Sequential', 'Model'):
        return model_from_config({'class_name': class_name, 'config': cfg}, custom_objects)
    if custom_objects and class_name in custom_objects:
        return custom_objects[class_name](**cfg)
    raise ValueError('layer class %r is not on the hot path of this package' % class_name)


def model_from_config(mc, custom_objects=None):
    """Rebuild a model from the ``model_config`` JSON of Keras 2.1 (Sequential config = list) or 2.2 (dict)."""
    from . import nn
    cn, cfg = mc['class_name'], mc['config']
    if cn == 'Sequential':
        layer_cfgs = cfg if isinstance(cfg, list) else cfg['layers']
        name = None if isinstance(cfg, list) else cfg.get('name')
        model = nn.Sequential(name=name)
        for lc in layer_cfgs:
            if lc['class_name'] == 'InputLayer':
                continue
            model.add(_layer_from_config(lc['class_name'], lc['config'], custom_objects))
        return model
    if cn == 'Model':
        tensors = {}
        inputs = []
        for lc in cfg['layers']:
            if lc['class_name'] == 'InputLayer':
                shape = tuple(lc['config']['batch_input_shape'][1:])
                t = nn.Input(shape=shape, name=lc['config'].get('name'))
                tensors[lc['name']] = t
                inputs.append(t)
                continue
            layer = _layer_from_config(lc['class_name'], lc['config'], custom_objects)
            inbound = lc['inbound_nodes'][0]
            srcs = [tensors[i[0]] for i in inbound]
            assert len(srcs) == 1, 'single-input layers only'
            tensors[lc['name']] = layer(srcs[0])
        ins = [tensors[i[0]] for i in cfg['input_layers']]
        outs = [tensors[o[0]] for o in cfg['output_layers']]
        return nn.Model(inputs=ins[0] if len(ins) == 1 else ins, outputs=outs if len(outs) > 1 else outs[0],
                        name=cfg.get('name'))
    raise ValueError('unsupported model class %r' % cn)


# ================================================================================================ weights <-> HDF5
def _flat_layers(model):
    """Keras stores nested models flattened: ``model.layers`` one level deep, a nested model's weights under the
    nested model's own name."""
    return list(model.layers)


def _layer_weights(layer):
    """[(weight name, ndarray)] in Keras order: trainable weights first, then non-trainable ones."""
    from . import nn
    if isinstance(layer, nn.Model):
        ps = [p for l in layer.all_layers() if l is not layer for p in l.params]
    else:
        ps = list(layer.params)
    if isinstance(layer, nn.Conv2D) and hasattr(layer, 'keras_weights'):
        return layer.keras_weights()
    ordered = [p for p in ps if p.trainable] + [p for p in ps if not p.trainable]
    return [(p.name, p.data.detach().cpu().numpy()) for p in ordered]


def _write_weight_group(g, model):
    layers = _flat_layers(model)
    g.attrs['layer_names'] = np.array([l.name.encode('utf8') for l in layers])
    g.attrs['backend'] = BACKEND.encode('utf8')
    g.attrs['keras_version'] = KERAS_VERSION.encode('utf8')
    for l in layers:
        lg = g.create_group(l.name)
        ws = _layer_weights(l)
        names = [n.encode('utf8') for n, _ in ws]
        lg.attrs['weight_names'] = np.array(names) if names else np.zeros((0,), dtype='S1')
        for n, v in ws:
            lg.create_dataset(n, np.asarray(v, dtype=np.float32))


def save_weights(model, path, overwrite=True):
    """``model.save_weights(path, overwrite)`` (bbhMahoGANy.py:1373-1375)."""
    if os.path.exists(path) and not overwrite:
        raise IOError('%s exists and overwrite is False' % path)
    f = hdf5.File(path, 'w')
    _write_weight_group(f._w, model)
    f.close()


def _set_layer_weights(layer, arrays, names):
    from . import nn
    if isinstance(layer, nn.Model):
        ps = [p for l in layer.all_layers() if l is not layer for p in l.params]
    else:
        ps = list(layer.params)
    ordered = [p for p in ps if p.trainable] + [p for p in ps if not p.trainable]
    if len(arrays) != len(ordered):
        raise ValueError('layer %s: file has %d weight arrays, the model expects %d' % (layer.name, len(arrays), len(ordered)))
    import torch
    for p, a, n in zip(ordered, arrays, names):
        a = np.asarray(a, dtype=np.float32)
        if tuple(a.shape) != p.shape:
            raise ValueError('layer %s weight %s: file shape %s, model shape %s' % (layer.name, n, a.shape, p.shape))
        p.data.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    nn._STATE['wver'] += 1


def _read_weight_group(g, model):
    """Keras ``load_weights_from_hdf5_group``: topological matching of the layers that HAVE weights."""
    def dec(x):
        return x.decode('utf8') if isinstance(x, (bytes, np.bytes_)) else str(x)
    layer_names = [dec(n) for n in np.asarray(g.attrs['layer_names']).ravel()]
    filtered = []
    for ln in layer_names:
        wn = [dec(n) for n in np.asarray(g[ln].attrs.get('weight_names', [])).ravel()]
        if wn:
            filtered.append((ln, wn))
    targets = [l for l in _flat_layers(model) if _layer_weights(l)]
    if len(filtered) != len(targets):
        raise ValueError('file contains %d layers with weights, the model has %d' % (len(filtered), len(targets)))
    for (ln, wn), layer in zip(filtered, targets):
        arrays = [np.asarray(g[ln][n][...]) for n in wn]
        _set_layer_weights(layer, arrays, wn)


def load_weights(model, path):
    """``model.load_weights(path)`` (bbhMahoGANy.py:1136-1138): accepts a weights file or a full-model file."""
    f = hdf5.File(path, 'r')
    g = f['model_weights'] if 'model_weights' in f else f
    _read_weight_group(g, model)


# ================================================================================================== whole models
def _optimizer_config(opt):
    from . import nn
    if isinstance(opt, nn.Adam):
        return {'class_name': 'Adam', 'config': {'lr': opt.lr, 'beta_1': opt.beta_1, 'beta_2': opt.beta_2,
                                                 'decay': opt.decay, 'epsilon': opt.epsilon, 'amsgrad': False}}
    if isinstance(opt, nn.SGD):
        return {'class_name': 'SGD', 'config': {'lr': opt.lr, 'momentum': 0.0, 'decay': opt.decay, 'nesterov': False}}
    raise ValueError('unsupported optimizer %r' % (opt,))


def _optimizer_from_config(oc):
    from . import nn
    c = oc['config']
    if oc['class_name'] == 'Adam':
        return nn.Adam(lr=c['lr'], beta_1=c['beta_1'], beta_2=c['beta_2'], epsilon=c.get('epsilon'), decay=c.get('decay', 0.0))
    if oc['class_name'] == 'SGD':
        return nn.SGD(lr=c['lr'], momentum=c.get('momentum', 0.0), decay=c.get('decay', 0.0), nesterov=c.get('nesterov', False))
    raise ValueError('unsupported optimizer class %r' % oc['class_name'])


def _trainable_params(model):
    c = model._compiled
    return [p for l in model.all_layers() if id(l) in c['trainable_ids'] for p in l.params if p.trainable]


def _optimizer_state(model):
    """Keras order: iterations, then (Adam) all first moments, all second moments [, all vhats] in weight order."""
    from . import nn
    opt = model.optimizer
    out = [('%s/iterations:0' % type(opt).__name__, np.asarray(opt.iterations, dtype=np.int64))]
    if isinstance(opt, nn.Adam) and opt.slots:
        params = _trainable_params(model)
        ms, vs = [], []
        for p in params:
            m, v = _param_slots(model, opt, p)
            ms.append(m)
            vs.append(v)
        k = 0
        for arrs in (ms, vs):
            for a in arrs:
                out.append(('training/Adam/Variable%s:0' % ('' if k == 0 else '_%d' % k), a))
                k += 1
        for p in params:      # vhats: Keras keeps (1,)-shaped zeros when amsgrad is off
            out.append(('training/Adam/Variable_%d:0' % k, np.zeros((1,), np.float32)))
            k += 1
    return out


def _param_slots(model, opt, p):
    """(m, v) of one parameter out of the optimizer's flat per-segment slot tensors."""
    for key, seg_p, _g in model._compiled['segments']:
        if key in opt.slots:
            base = seg_p.data_ptr()
            off = (p.data.data_ptr() - base) // 4
            if 0 <= off and off + p.numel() <= seg_p.numel():
                m, v = opt.slots[key]
                sl = slice(off, off + p.numel())
                return (m[sl].reshape(p.shape).detach().cpu().numpy(), v[sl].reshape(p.shape).detach().cpu().numpy())
    return np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)


def save_model(model, path, overwrite=True):
    """``model.save(path, overwrite)`` (bbhMahoGANy.py:1173): architecture + weights + training config + optimizer."""
    if os.path.exists(path) and not overwrite:
        raise IOError('%s exists and overwrite is False' % path)
    f = hdf5.File(path, 'w')
    root = f._w
    root.attrs['keras_version'] = KERAS_VERSION.encode('utf8')
    root.attrs['backend'] = BACKEND.encode('utf8')
    root.attrs['model_config'] = json.dumps(model_config(model)).encode('utf8')
    _write_weight_group(root.create_group('model_weights'), model)
    if model.optimizer is not None and model._compiled is not None:
        c = model._compiled
        tc = {'optimizer_config': _optimizer_config(model.optimizer), 'loss': c['loss'].name,
              'metrics': list(getattr(model, 'metrics', None) or []), 'sample_weight_mode': None, 'loss_weights': None}
        root.attrs['training_config'] = json.dumps(tc).encode('utf8')
        state = _optimizer_state(model)
        og = root.create_group('optimizer_weights')
        og.attrs['weight_names'] = np.array([n.encode('utf8') for n, _ in state])
        for n, v in state:
            og.create_dataset(n, v)
    f.close()


def load_model(path, custom_objects=None, compile=True):
    """``keras.models.load_model(path)`` (bbhMahoGANy.py:1135,1142)."""
    from . import nn
    f = hdf5.File(path, 'r')
    mc = f.attrs.get('model_config')
    if mc is None:
        raise ValueError('%s holds weights only (no model_config): build the model and call load_weights' % path)
    if isinstance(mc, bytes):
        mc = mc.decode('utf8')
    model = model_from_config(json.loads(mc), custom_objects)
    _read_weight_group(f['model_weights'], model)
    tc = f.attrs.get('training_config')
    if compile and tc is not None:
        tc = json.loads(tc.decode('utf8') if isinstance(tc, bytes) else tc)
        loss = tc['loss']
        if isinstance(loss, str) and loss == 'chisquare_Loss':
            loss = (custom_objects or {}).get('chisquare_Loss', nn.chisquare_Loss())
        model.compile(loss=loss, optimizer=_optimizer_from_config(tc['optimizer_config']), metrics=tc.get('metrics') or None)
        if 'optimizer_weights' in f:
            _restore_optimizer(model, f['optimizer_weights'])
    return model


def _restore_optimizer(model, og):
    from . import nn
    import torch

    def dec(x):
        return x.decode('utf8') if isinstance(x, (bytes, np.bytes_)) else str(x)
    names = [dec(n) for n in np.asarray(og.attrs.get('weight_names', [])).ravel()]
    vals = [np.asarray(og[n][...]) for n in names]
    opt = model.optimizer
    if not vals:
        return
    opt.iterations = int(np.asarray(vals[0]).ravel()[0])
    if isinstance(opt, nn.Adam):
        params = _trainable_params(model)
        n = len(params)
        if len(vals) < 1 + 2 * n:
            return
        ms, vs = vals[1:1 + n], vals[1 + n:1 + 2 * n]
        for key, seg_p, _g in model._compiled['segments']:
            m = torch.zeros_like(seg_p)
            v = torch.zeros_like(seg_p)
            base = seg_p.data_ptr()
            for p, pm, pv in zip(params, ms, vs):
                off = (p.data.data_ptr() - base) // 4
                if 0 <= off and off + p.numel() <= seg_p.numel() and tuple(pm.shape) == p.shape:
                    m[off:off + p.numel()] = torch.from_numpy(np.ascontiguousarray(pm, dtype=np.float32)).reshape(-1).to(m.device)
                    v[off:off + p.numel()] = torch.from_numpy(np.ascontiguousarray(pv, dtype=np.float32)).reshape(-1).to(v.device)
            opt.slots[key] = (m, v)
