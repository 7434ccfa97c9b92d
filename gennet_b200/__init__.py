"""gennet_b200: B200-native (sm_100a) implementation of hagabbar/GenNet's data-parallel hot path.

    gennet_b200.synth   -- BBH_version/gw_template_maker.py synthesis API (noise, whitening, injection)
    gennet_b200.nn      -- the Keras object protocol the reference scripts use (layers, Model, Adam ...)
    gennet_b200.bbh     -- builders / steps of BBH_version/bbhMahoGANy.py
    gennet_b200.burst   -- builders of tests/burstMahoGANy.py
    gennet_b200.wvf     -- builders / loop of train_on_wvf_version/nn.py
    gennet_b200.parallel-- data-parallel wiring (torch.distributed, NCCL)

Everything numeric runs in hand-written CUDA behind the C ABI in include/gennet_b200.h; importing the
package without the built library raises (no CPU fallback).
"""
from . import _lib

_lib.load()   # fail loudly if libgennet_b200.so has not been built

__version__ = '0.1.0'
