"""File formats of the reference pipeline: Keras HDF5 weights and models through the
pure-Python HDF5 layer, and Python-2 cPickle template banks.  CPU only (host logic; the C ABI is answered by
tests/fake_backend.py where a model is needed)."""
import hashlib
import io
import json
import os
import pickle
import pickletools

import numpy as np
import pytest

from tests import fake_backend

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'keras_hdf5_digest.json')
REF_DIR = '/root/reference/2_model_version/weight_version'


@pytest.fixture
def fake(monkeypatch):
    return fake_backend.install(monkeypatch)


def _digest(path):
    from gennet_b200 import hdf5
    f = hdf5.File(path)
    out = {'attrs': sorted(f.attrs.keys()), 'datasets': {}}
    for p, d in f.visit_datasets():
        a = np.ascontiguousarray(d[...])
        out['datasets'][p] = [list(a.shape), str(a.dtype), hashlib.sha256(a.tobytes()).hexdigest()[:16]]
    g = f['model_weights'] if 'model_weights' in f else f
    out['layer_names'] = [n.decode() for n in np.asarray(g.attrs['layer_names']).ravel()]
    out['keras_version'] = f.attrs['keras_version']
    return out


def test_hdf5_roundtrip_all_supported_types(tmp_path):
    from gennet_b200 import hdf5
    rs = np.random.RandomState(0)
    p = str(tmp_path / 'a.h5')
    f = hdf5.File(p, 'w')
    f.wattrs['title'] = 'gennet'
    f.wattrs['raw'] = b'tensorflow'
    f.wattrs['names'] = np.array([b'dense_1', b'conv1d_12', b'x'])
    f.wattrs['vec'] = np.arange(5, dtype=np.int64)
    f.wattrs['long_json'] = json.dumps({'k': list(range(3000))})       # > 4 KiB: its own global heap collection
    arrays = {'g1/w:0': rs.normal(size=(7, 3, 5)).astype(np.float32), 'g1/sub/b:0': rs.normal(size=(11,)),
              'scalar': np.asarray(2532, dtype=np.int64), 'empty': np.zeros((0, 4), np.float32),
              'half': rs.normal(size=(4, 4)).astype(np.float16), 'u8': np.arange(20, dtype=np.uint8).reshape(4, 5)}
    for k, v in arrays.items():
        f.create_dataset(k, v)
    g = f.create_group('g1')
    g.attrs['weight_names'] = np.array([b'w:0', b'sub/b:0'])
    many = f.create_group('many')                                        # more links than one symbol-table node holds
    for i in range(150):
        many.create_dataset('d%03d' % i, np.full((2,), i, np.float32))
    f.close()
    r = hdf5.File(p)
    assert r.attrs['title'] == 'gennet' and r.attrs['raw'] == 'tensorflow'
    assert list(r.attrs['names']) == [b'dense_1', b'conv1d_12', b'x']
    assert np.array_equal(r.attrs['vec'], np.arange(5))
    assert json.loads(r.attrs['long_json'])['k'][-1] == 2999
    for k, v in arrays.items():
        got = r[k][...]
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
    assert list(r['g1'].attrs['weight_names']) == [b'w:0', b'sub/b:0']
    assert len(r['many'].keys()) == 150 and float(r['many/d149'][...][0]) == 149.0
    with pytest.raises(KeyError):
        r['nope']


def test_hdf5_rejects_foreign_bytes(tmp_path):
    from gennet_b200 import hdf5
    p = tmp_path / 'x.h5'
    p.write_bytes(b'not an hdf5 file at all')
    with pytest.raises(ValueError):
        hdf5.File(str(p))


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason='reference checkout not mounted')
def test_reader_matches_golden_digest_of_reference_files():
    """The four Keras files shipped by the reference (written by h5py/libhdf5) parse to the committed digest
    (tests/golden/make_hdf5_digest.py): names, shapes, dtypes, data hashes, attributes."""
    gold = json.load(open(GOLDEN))
    for name, want in gold.items():
        got = _digest(os.path.join(REF_DIR, name))
        assert got == want, name


def test_golden_digest_is_self_consistent():
    gold = json.load(open(GOLDEN))
    d = gold['d_model.hdf5']
    assert d['layer_names'] == ['input_2', 'reshape_2', 'conv1d_1', 'leaky_re_lu_1', 'flatten_2', 'dense_3', 'dense_4']
    assert d['datasets']['model_weights/conv1d_1/conv1d_1/kernel:0'][0] == [16, 1, 50]      # (k, Cin, Cout)
    assert d['datasets']['model_weights/dense_3/dense_3/kernel:0'][0] == [1750, 50]          # (in, out)
    assert d['datasets']['optimizer_weights/Adam_1/iterations:0'][1] == 'int64'


def gold_names(fname):
    return json.load(open(GOLDEN))[fname]['layer_names']


PE_LAYER_NAMES = ['input_1', 'conv1d_5', 'activation_6', 'conv1d_1', 'conv1d_6', 'activation_1', 'activation_7', 'conv1d_2',
                  'conv1d_7', 'activation_2', 'activation_8', 'conv1d_3', 'conv1d_8', 'activation_3', 'activation_9',
                  'conv1d_4', 'conv1d_9', 'activation_4', 'activation_10', 'flatten_1', 'flatten_2', 'dense_1', 'dense_2',
                  'activation_5', 're_lu_1']


def test_save_and_load_weights_roundtrip(fake, tmp_path):
    from gennet_b200 import nn, hdf5
    from tests import parity_cases as pc
    prod, orc, x, y = pc.pe_case(128, 4)
    p = str(tmp_path / 'signal_pe_weights.h5')
    prod.save_weights(p, True)
    f = hdf5.File(p)
    names = [n.decode() for n in f.attrs['layer_names']]
    assert names == [l.name for l in prod.layers]
    # Keras 2.2.4 orders a functional model's layers by depth from the outputs (network.py _map_graph_network), ties to
    # the branch of the first output: the two towers of signal_pe_model (bbhMahoGANy.py:357-404) interleave, and
    # the deeper q tower's first convolution comes before the mc tower's although it was created later
    assert names == PE_LAYER_NAMES
    assert f.attrs['backend'] == 'tensorflow' and f.attrs['keras_version'] == '2.2.4'
    conv = [l for l in prod.layers if isinstance(l, nn.Conv1D) and l.params[0].shape == (5, 64, 128)][0]
    assert [n.decode() for n in f[conv.name].attrs['weight_names']] == [conv.name + '/kernel:0', conv.name + '/bias:0']
    k = f[conv.name][conv.name + '/kernel:0'][...]
    assert k.shape == (5, 64, 128) and k.dtype == np.float32                                  # Keras (k, Cin, Cout)
    w0 = prod.get_weights()
    before = prod.predict(x)
    for l in prod.layers:                      # scramble, then restore from the file
        if l.params:
            l.set_weights([np.zeros(p_.shape, np.float32) for p_ in l.params])
    prod.load_weights(p)
    for a, b in zip(w0, prod.get_weights()):
        assert np.array_equal(a, b)
    after = prod.predict(x)
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    with pytest.raises(IOError):
        prod.save_weights(p, overwrite=False)


def test_save_and_load_full_model(fake, tmp_path):
    from gennet_b200 import nn
    from tests import parity_cases as pc
    prod, orc, x, y = pc.pe_case(128, 4)
    prod.train_on_batch(x, y)
    prod.train_on_batch(x, y)
    p = str(tmp_path / 'signal_pe.h5')
    prod.save(p, True)
    nn.clear_session()
    m2 = nn.load_model(p)
    assert [type(l).__name__ for l in m2.layers] == [type(l).__name__ for l in prod.layers]
    for a, b in zip(prod.get_weights(), m2.get_weights()):
        assert np.array_equal(a, b)
    for a, b in zip(prod.predict(x), m2.predict(x)):
        assert np.array_equal(a, b)
    # the optimizer continues where it stopped: same third step on both
    assert m2.optimizer.iterations == 2 and abs(m2.optimizer.lr - 9e-5) < 1e-12 and m2.optimizer.beta_1 == 0.5
    r1 = prod.train_on_batch(x, y)
    r2 = m2.train_on_batch(x, y)
    assert np.allclose(r1, r2, rtol=1e-6, atol=1e-7)
    for a, b in zip(prod.get_weights(), m2.get_weights()):
        assert np.allclose(a, b, rtol=1e-6, atol=1e-8)


def test_sequential_gan_models_roundtrip(fake, tmp_path):
    from gennet_b200 import nn
    from tests import parity_cases as pc
    (g, d, dg), _, z, sX, sy = pc.gan_case(64, 4)
    for m, tag in ((g, 'generator'), (d, 'discriminator')):
        p = str(tmp_path / (tag + '.h5'))
        m.save_weights(p, True)
        w0 = m.get_weights()
        for l in m.layers:
            if l.params:
                l.set_weights([np.ones(p_.shape, np.float32) for p_ in l.params])
        m.load_weights(p)
        for a, b in zip(w0, m.get_weights()):
            assert np.array_equal(a, b)
    # BatchNormalization: Keras order gamma, beta, moving_mean, moving_variance
    from gennet_b200 import hdf5
    f = hdf5.File(str(tmp_path / 'generator.h5'))
    bn = [l for l in g.layers if isinstance(l, nn.BatchNormalization)][0]
    assert [n.decode().split('/')[-1] for n in f[bn.name].attrs['weight_names']] == \
        ['gamma:0', 'beta:0', 'moving_mean:0', 'moving_variance:0']


# ------------------------------------------------------------------------------------------- Python-2 pickles
# what `cPickle.dump([ts, yval], f, 2)` / `cPickle.dump([bbhparams(...)], f, 2)` write under Python 2.7 + NumPy 1.15
PY2_TS = (b'\x80\x02]q\x01(cnumpy.core.multiarray\n_reconstruct\nq\x02cnumpy\nndarray\nq\x03K\x00\x85U\x01b\x87Rq\x04'
          b'(K\x01K\x02K\x01K\x03\x87cnumpy\ndtype\nq\x05U\x02f8K\x00K\x01\x87Rq\x06(K\x03U\x01<NNNJ\xff\xff\xff\xffJ\xff'
          b'\xff\xff\xffK\x00tb\x89U0\x00\x00\x00\x00\x00\x00\xf0?\x00\x00\x00\x00\x00\x00\x00@\x00\x00\x00\x00\x00\x00'
          b'\x08@\x00\x00\x00\x00\x00\x00\x10@\x00\x00\x00\x00\x00\x00\x14@\x00\x00\x00\x00\x00\x00\x18@tbh\x02h\x03K\x00'
          b'\x85U\x01b\x87Rq\x07(K\x01K\x02\x85h\x06\x89U\x10\x00\x00\x00\x00\x00\x00\xf0?\x00\x00\x00\x00\x00\x00\x00\x00'
          b'tbe.')
PY2_PARS = (b'\x80\x02]q\x01(c__main__\nbbhparams\nq\x02)\x81q\x03}q\x04(U\x02mcq\x05G@>\x00\x00\x00\x00\x00\x00U\x02m1'
            b'q\x06G@B\x00\x00\x00\x00\x00\x00U\x02m2q\x07G@=\x00\x00\x00\x00\x00\x00U\x03idxq\x08M\x00\x02U\x03snrq\tNube.')
# classic-instance form (old-style class): (c__main__ bbhparams o } ... b
PY2_PARS_OLD = (b'\x80\x02]q\x01(c__main__\nbbhparams\nq\x02oq\x03}q\x04(U\x02mcq\x05G@>\x00\x00\x00\x00\x00\x00U\x02m1'
                b'q\x06G@B\x00\x00\x00\x00\x00\x00U\x02m2q\x07G@=\x00\x00\x00\x00\x00\x00U\x03idxq\x08M\x00\x02U\x03snrq\tNuba.')


def test_load_python2_template_pickles():
    from gennet_b200 import io as gio, synth
    ts = gio.load_pickle(io.BytesIO(PY2_TS))
    assert isinstance(ts, list) and ts[0].shape == (2, 1, 3) and ts[0].dtype == np.float64
    assert np.array_equal(ts[0].ravel(), [1, 2, 3, 4, 5, 6]) and np.array_equal(ts[1], [1.0, 0.0])
    par = gio.load_pickle(io.BytesIO(PY2_PARS_OLD))
    assert len(par) == 1 and isinstance(par[0], synth.bbhparams)
    assert (par[0].mc, par[0].m1, par[0].m2, par[0].idx, par[0].snr) == (30.0, 36.0, 29.0, 512, None)


def test_dump_pickle_is_python2_loadable(tmp_path):
    from gennet_b200 import io as gio, synth
    ts = [np.arange(12, dtype=np.float64).reshape(2, 1, 6), np.array([1, 0])]
    pars = [synth.bbhparams(np.float64(30.0), 65.0, 0.24, 36.0, 29.0, 2.2, -1.2, 0.1, 0.2, 0.3, np.int64(512), None, None, 40.0)]
    p1, p2 = str(tmp_path / 'gw150914_ts_0_2Samp.sav'), str(tmp_path / 'gw150914_params_0_2Samp.sav')
    gio.save_template_bank(ts[0], ts[1], pars, p1, p2)
    raw1, raw2 = open(p1, 'rb').read(), open(p2, 'rb').read()
    assert raw1[:2] == b'\x80\x02' and raw2[:2] == b'\x80\x02'
    ops1 = [(op.name, arg) for op, arg, _ in pickletools.genops(raw1)]
    glob = [a for n, a in ops1 if n == 'GLOBAL']
    # Python 2 / NumPy 1.15 resolvable globals only; no Python-3-only opcodes
    assert 'numpy.core.multiarray _reconstruct' in glob and all('_core' not in g for g in glob)
    assert not any(n in ('BINBYTES', 'SHORT_BINBYTES', 'FRAME', 'SHORT_BINUNICODE', 'STACK_GLOBAL', 'NEWOBJ_EX') for n, _ in ops1)
    ops2 = [(op.name, arg) for op, arg, _ in pickletools.genops(raw2)]
    assert ('GLOBAL', '__main__ bbhparams') in ops2 and any(n == 'OBJ' for n, _ in ops2)
    assert not any(n in ('NEWOBJ', 'REDUCE') for n, _ in ops2)          # classic instances: no __new__, no numpy scalars
    a, b, par = gio.load_template_bank(p1, p2)
    assert np.array_equal(a, ts[0]) and a.dtype == np.float64 and np.array_equal(b, ts[1])
    assert par[0].mc == 30.0 and par[0].idx == 512 and par[0].fmin == 40.0 and par[0].snr is None
    assert type(par[0].mc) is float and type(par[0].idx) is int


def test_npy_roundtrip(tmp_path):
    from gennet_b200 import io as gio
    a = np.random.RandomState(1).normal(size=(3, 1, 8))
    p = str(tmp_path / 't.npy')
    gio.save_npy(p, a)
    assert np.array_equal(gio.load_npy(p), a)


def test_writer_encodings_equal_libhdf5_bytes():
    """Message bodies as libhdf5 1.8/1.10 (through h5py) wrote them in the reference's own files
    (d_model.hdf5: dumped with gennet_b200.hdf5._Reader.messages) -- the writer must emit the same bytes."""
    from gennet_b200.hdf5 import _Writer as W
    assert W.dtype_msg(np.float32).hex() == '11201f000400000000002000170800177f000000'       # dense_3/bias:0
    assert W.dtype_msg(np.int64).hex() == '100800000800000000004000'                         # Adam_1/iterations:0
    assert W.dtype_msg('S13').hex() == '130100000d000000'                                    # attr layer_names
    assert W.vlen_str_msg(utf8=False).hex() == '1901000010000000100000000100000000000800'    # attr keras_version
    assert W.space_msg((50,)).hex() == '010101000000000032000000000000003200000000000000'
    assert W.space_msg(()).hex() == '0100000000000000'


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason='reference checkout not mounted')
@pytest.mark.parametrize('fname', ['g_model.hdf5', 'd_model.hdf5'])
def test_load_reference_keras_models(fake, fname):
    """keras.models.load_model on the two full models the reference ships (Keras 2.1.x files): the architecture is
    rebuilt from model_config, weights land by Keras' topological order, the optimizer comes back from
    training_config / optimizer_weights, and predict() agrees with the float64 oracle carrying the same weights."""
    import torch
    from gennet_b200 import nn, hdf5
    from oracle import keras_oracle as ko
    nn.clear_session()
    ko.clear_session()
    m = nn.load_model(os.path.join(REF_DIR, fname))
    f = hdf5.File(os.path.join(REF_DIR, fname))
    rs = np.random.RandomState(0)
    if fname == 'g_model.hdf5':
        assert [type(l).__name__ for l in m.layers].count('Conv2DTranspose') == 4 and m.output_shape == (50,)
        assert type(m.optimizer).__name__ == 'SGD' and abs(m.optimizer.lr - 0.004) < 1e-8
        orc = ko.build(ko.two_model_get_generative(1, 50))
        x = rs.uniform(-5, 5, (6, 1, 1)).astype(np.float32)
        k = np.asarray(f['model_weights/conv2d_transpose_2/conv2d_transpose_2/kernel:0'][...])
        tl = [l for l in m.layers if type(l).__name__ == 'Conv2DTranspose'][1]
        assert np.array_equal(tl.get_weights()[0], k) and k.shape == (1, 8, 64, 128)          # (kh, kw, Cout, Cin)
    else:
        assert type(m.optimizer).__name__ == 'Adam' and m.optimizer.iterations == 2532 and m.optimizer.beta_1 == 0.5
        orc = ko.build(ko.two_model_get_discriminative(50))
        x = rs.normal(size=(6, 50)).astype(np.float32)
        # the file was saved after set_trainability(D, False) (make_gan): every layer comes back frozen, as in Keras,
        # so there is nothing for the stored Adam moments to attach to until the model is unfrozen and recompiled
        assert not any(l.trainable for l in m.layers if l.params)
        from gennet_b200 import io as gio
        nn.set_trainability(m, True)
        m.compile(loss='binary_crossentropy', optimizer=m.optimizer)
        gio._restore_optimizer(m, f['optimizer_weights'])
        mom = np.asarray(f['optimizer_weights/training/Adam/Variable:0'][...])       # first moment of the conv kernel
        vel = np.asarray(f['optimizer_weights/training/Adam/Variable_6:0'][...])     # its second moment (6 weights)
        assert [l.name for l in m.layers] == [n for n in gold_names(fname)]       # the file's own layer_names, InputLayer first
        conv = [l for l in m.layers if type(l).__name__ == 'Conv1D'][0]
        got_m, got_v = gio._param_slots(m, m.optimizer, conv.params[0])
        assert np.array_equal(got_m, mom) and np.array_equal(got_v, vel) and m.optimizer.iterations == 2532
    orc.set_weights([w.astype(np.float64) for w in m.get_weights()])
    got, ref = m.predict(x), orc.predict(x)
    assert got.shape == ref.shape and np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-3)
