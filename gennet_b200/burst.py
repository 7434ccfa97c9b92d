"""Model builders of tests/burstMahoGANy.py (sine-Gaussian burst GAN, BASELINE config 1), same names."""
import numpy as np

from . import nn
from .nn import (Activation, Conv1D, Dense, Flatten, GaussianDropout, MaxPooling1D, Reshape, Sequential,
                 UpSampling1D, Adam, set_trainable)
from .synth import make_burst_waveforms  # noqa: F401  (burstMahoGANy.py:76-98)

# tests/burstMahoGANy.py:31-48
n_colors = 1
n_pix = 512
n_sig = 0.25
batch_size = 64
lr = 2e-4
chi_loss = False


class MyLayer(nn.ResidualMoments):
    """tests/burstMahoGANy.py:100-125."""


def generator_model():
    """tests/burstMahoGANy.py:127-251."""
    model = Sequential()
    act, drate = 'relu', 0.3
    model.add(Dense(256 * 1 * int(n_pix / 2), input_shape=(100,)))
    model.add(Activation(act))
    model.add(Reshape((int(n_pix / 2), 256)))
    model.add(UpSampling1D(size=2))
    for filters in (64, 64, 256, 512):
        model.add(Conv1D(filters, 5, strides=1, padding='same'))
        model.add(Activation(act))
        model.add(GaussianDropout(drate))
    model.add(Conv1D(n_colors, 5, padding='same'))
    model.add(Activation('tanh'))
    return model


def data_subtraction_model(noise_signal, npix):
    """tests/burstMahoGANy.py:253-261."""
    model = Sequential()
    model.add(MyLayer(noise_signal, input_shape=(npix, 1)))
    return model


def signal_pe_model():
    """tests/burstMahoGANy.py:263-293."""
    model = Sequential()
    act = 'relu'
    model.add(Conv1D(64, 5, strides=2, input_shape=(n_pix, 1), padding='same'))
    model.add(Activation(act))
    model.add(Conv1D(128, 5, strides=2))
    model.add(Activation(act))
    model.add(Flatten())
    model.add(Dense(1024))
    model.add(Activation(act))
    model.add(Dense(2))
    model.add(Activation('linear'))
    return model


def signal_discriminator_model():
    """tests/burstMahoGANy.py:295-402."""
    act = 'tanh'
    model = Sequential()
    model.add(Conv1D(64, 5, input_shape=(n_pix, 1), strides=1, padding='same'))
    model.add(Activation(act))
    model.add(MaxPooling1D(pool_size=2))
    model.add(Conv1D(128, 5, strides=1))
    model.add(Activation(act))
    model.add(MaxPooling1D(pool_size=2))
    model.add(Flatten())
    model.add(Dense(1024))
    model.add(Activation(act))
    model.add(Dense(1))
    model.add(Activation('sigmoid'))
    return model


def generator_after_subtracting_noise(generator, data_subtraction):
    """tests/burstMahoGANy.py:404-413."""
    model = Sequential()
    model.add(generator)
    model.add(data_subtraction)
    return model


def generator_containing_signal_discriminator(generator, signal_discriminator):
    """tests/burstMahoGANy.py:415-423."""
    model = Sequential()
    model.add(generator)
    model.add(signal_discriminator)
    return model


def build_gan(noise_signal):
    """Model set-up of main(), tests/burstMahoGANy.py:641-673."""
    signal_discriminator = signal_discriminator_model()
    data_subtraction = data_subtraction_model(noise_signal, n_pix)
    generator = generator_model()
    data_subtraction_on_generator = generator_after_subtracting_noise(generator, data_subtraction)
    data_subtraction_on_generator.compile(loss='mean_squared_error', optimizer=Adam(lr=lr, beta_1=0.5),
                                          metrics=['accuracy'])
    signal_discriminator_on_generator = generator_containing_signal_discriminator(generator, signal_discriminator)
    set_trainable(signal_discriminator, False)
    signal_discriminator_on_generator.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5),
                                              metrics=['accuracy'])
    set_trainable(signal_discriminator, True)
    signal_discriminator.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5), metrics=['accuracy'])
    return generator, signal_discriminator, signal_discriminator_on_generator, data_subtraction_on_generator
