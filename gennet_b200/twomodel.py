"""Builders of 2_model_version (BASELINE config 5), same names as
2_model_version/weight_version/no_mode_collapse_network.py: transposed-convolution generator (:62-106),
Conv1D discriminator (:115-149), ``set_trainability`` / ``make_gan`` (:160-172), the sampling helpers (:181-221).
``g_model.hdf5`` / ``d_model.hdf5`` shipped next to that script hold exactly these two networks and load through
``gennet_b200.nn.load_model``."""
import numpy as np

from .nn import (Adam, BatchNormalization, Conv1D, Conv2DTranspose, Dense, Flatten, Input, LeakyReLU, Model, Reshape,
                 SGD, set_trainability)


def get_generative(G_in, dense_dim=128, drate=0.1, out_dim=50, lr=1e-3):
    """no_mode_collapse_network.py:62-106: widths 1 -> 4 -> 11 -> 26 -> 57 along the transposed-conv axis."""
    x = Reshape((-1, 1, 1))(G_in)
    x = BatchNormalization()(x)
    for f, k in ((128, 4), (64, 8), (32, 16), (16, 32)):
        x = Conv2DTranspose(f, (1, k), strides=(1, 1), padding='valid', activation='relu')(x)
        x = BatchNormalization()(x)
    x = Flatten()(x)
    x = BatchNormalization()(x)
    x = Dense(out_dim, activation='relu')(x)
    x = BatchNormalization()(x)
    G_out = Dense(out_dim, activation='linear')(x)
    G = Model(G_in, G_out)
    G.compile(loss='binary_crossentropy', optimizer=SGD(lr=lr))
    return G, G_out


def get_discriminative(D_in, lr=1e-3, drate=.25, n_channels=50, conv_sz=5, leak=.2):
    """no_mode_collapse_network.py:115-149 (Dense(n_channels) is linear in the script; the shipped d_model.hdf5 was
    saved from a variant with tanh there -- load_model follows the file)."""
    x = Reshape((-1, 1))(D_in)
    x = Conv1D(50, 16)(x)
    x = LeakyReLU(alpha=0.2)(x)
    x = Flatten()(x)
    x = Dense(n_channels)(x)
    D_out = Dense(2, activation='sigmoid')(x)
    D = Model(D_in, D_out)
    D.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5))
    return D, D_out


def make_gan(GAN_in, G, D):
    """no_mode_collapse_network.py:165-172."""
    set_trainability(D, False)
    x = G(GAN_in)
    GAN_out = D(x)
    GAN = Model(GAN_in, GAN_out)
    GAN.compile(loss='binary_crossentropy', optimizer=G.optimizer)
    return GAN, GAN_out


def sample_noise(G, noise_dim=10, n_samples=10000, rng=np.random):
    """no_mode_collapse_network.py:201-205."""
    X = rng.uniform(-5, 5, size=[n_samples, 1, noise_dim])
    y = np.zeros((n_samples, 2))
    y[:, 1] = 1
    return X, y


def sample_data_and_gen(G, XT, noise_dim=10, noise_samples=100, rng=np.random):
    """no_mode_collapse_network.py:181-192 with the training set XT (n, out_dim) passed in."""
    XT = np.asarray(XT)
    n_samples = XT.shape[0]
    XN_noise = rng.uniform(-5, 5, size=[noise_samples, 1, noise_dim])
    XN = G.predict(XN_noise)
    X = np.vstack((XT, XN))
    y = np.zeros((n_samples + len(XN_noise), 2))
    y[:n_samples, 1] = 1
    y[n_samples:, 0] = 1
    return X, y
