"""Builders and loop helpers of train_on_wvf_version/nn.py (BASELINE config 4), same names."""
import numpy as np

from .nn import (Activation, Conv1D, Dense, Dropout, Flatten, Input, Model, Reshape, Adam, SGD, set_trainability)


def sample_data(n_samples=10000, x_vals=np.arange(0, 5, .1), max_offset=100, mul_range=[1, 2], rng=np.random):
    """nn.py:58-70 (host; a 50-point toy generator, not on the device path)."""
    vectors = []
    for i in range(n_samples):
        offset = rng.random_sample() * max_offset
        mul = mul_range[0] + rng.random_sample() * (mul_range[1] - mul_range[0])
        vectors.append(np.sin(offset + x_vals * mul) / 2 + .5)
    return np.array(vectors)


def get_generative(G_in, dense_dim=300, out_dim=8192, lr=0.425e-1):
    """nn.py:72-81."""
    x = Dense(dense_dim)(G_in)
    x = Activation('relu')(x)
    x = Dense(150)(x)
    x = Activation('relu')(x)
    G_out = Dense(out_dim, activation='tanh')(x)
    G = Model(G_in, G_out)
    G.compile(loss='binary_crossentropy', optimizer=SGD(lr=lr))
    return G, G_out


def get_discriminative(D_in, lr=1e-6, drate=.25, n_channels=25, conv_sz=5, leak=.2):
    """nn.py:83-93."""
    x = Reshape((-1, 1))(D_in)
    x = Conv1D(n_channels, conv_sz, activation='relu')(x)
    x = Dropout(drate)(x)
    x = Flatten()(x)
    x = Dense(n_channels)(x)
    D_out = Dense(2, activation='sigmoid')(x)
    D = Model(D_in, D_out)
    D.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr))
    return D, D_out


def make_gan(GAN_in, G, D):
    """nn.py:100-106."""
    set_trainability(D, False)
    x = G(GAN_in)
    GAN_out = D(x)
    GAN = Model(GAN_in, GAN_out)
    GAN.compile(loss='binary_crossentropy', optimizer=G.optimizer)
    return GAN, GAN_out


def sample_data_and_gen(G, x_train, noise_dim=10, n_samples=10000, rng=np.random):
    """nn.py:108-122."""
    XT = x_train[:n_samples]
    XN_noise = rng.uniform(0, 1, size=[n_samples, noise_dim])
    XN = G.predict(XN_noise)
    X = np.concatenate((XT, XN))
    y = np.zeros((2 * n_samples, 2))
    y[:n_samples, 1] = 1
    y[n_samples:, 0] = 1
    return X, y


def pretrain(G, D, x_train, noise_dim=10, n_samples=10000, batch_size=32):
    """nn.py:124-128."""
    X, y = sample_data_and_gen(G, x_train, n_samples=n_samples, noise_dim=noise_dim)
    set_trainability(D, True)
    D.fit(X, y, epochs=1, batch_size=batch_size)


def sample_noise(G, noise_dim=10, n_samples=10000, rng=np.random):
    """nn.py:130-134."""
    X = rng.uniform(0, 1, size=[n_samples, noise_dim])
    y = np.zeros((n_samples, 2))
    y[:, 1] = 1
    return X, y


def train(GAN, G, D, x_train, epochs=1000, n_samples=10000, noise_dim=10, batch_size=32, verbose=False, v_freq=1):
    """nn.py:136-152."""
    d_loss, g_loss = [], []
    for epoch in range(epochs):
        X, y = sample_data_and_gen(G, x_train, n_samples=n_samples, noise_dim=noise_dim)
        set_trainability(D, True)
        d_loss.append(D.train_on_batch(X, y))
        X, y = sample_noise(G, n_samples=n_samples, noise_dim=noise_dim)
        set_trainability(D, False)
        g_loss.append(GAN.train_on_batch(X, y))
        if verbose and (epoch + 1) % v_freq == 0:
            print("Epoch #{}: Generative Loss: {}, Discriminative Loss: {}".format(epoch + 1, g_loss[-1], d_loss[-1]))
    return d_loss, g_loss
