"""2_model_version/weight_version/noise_gan.py: the discriminator is pre-trained to tell Gaussian noise N(0, nstd) from
generator output, then saved (``best_d_weights.hdf5`` / ``d_model.hdf5``, the files the subtract stage loads)."""
import numpy as np

from ..nn import (Adam, BatchNormalization, Conv1D, Conv2DTranspose, Dense, Flatten, Input, LeakyReLU, Model, Reshape,
                  set_trainability)

# noise_gan.py:29-38
n_total = 500
n_samples = int(n_total * 0.5)
noise_samples = int(n_total * 0.5)
noise_dim = 1
batch_size = 16
epochs = 2500
g_lr = 40e-4
d_lr = 40e-4
nstd = 1


def sample_data(n_samples=10000, x_vals=np.arange(0, 5, .1), max_offset=2 * np.pi, mul_range=[1, 2], rng=np.random):
    """noise_gan.py:41-51."""
    vectors = []
    for i in range(n_samples):
        offset = rng.random_sample() * max_offset
        mul = (2 * np.pi) / 5
        vectors.append(np.sin(offset + x_vals * mul))
    return np.array(vectors)


def get_generative(G_in, dense_dim=128, drate=0.6, out_dim=50, lr=1e-3):
    """noise_gan.py:63-114: the transposed-convolution generator (ReLU), Adam(lr, beta_1=0.5)."""
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
    G.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5))
    return G, G_out


def get_discriminative(D_in, lr=1e-3, drate=.6, n_channels=50, conv_sz=5, leak=.2):
    """noise_gan.py:123-158 (the network stored in the shipped d_model.hdf5: Dense(n_channels, tanh))."""
    x = Reshape((-1, 1))(D_in)
    x = Conv1D(50, 16)(x)
    x = LeakyReLU(alpha=0.2)(x)
    x = Flatten()(x)
    x = Dense(n_channels, activation='tanh')(x)
    D_out = Dense(2, activation='sigmoid')(x)
    D = Model(D_in, D_out)
    D.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5))
    return D, D_out


def make_gan(GAN_in, G, D):
    """noise_gan.py:174-180."""
    set_trainability(D, False)
    x = G(GAN_in)
    GAN_out = D(x)
    GAN = Model(GAN_in, GAN_out)
    GAN.compile(loss='binary_crossentropy', optimizer=G.optimizer)
    return GAN, GAN_out


def sample_data_and_gen(G, noise_dim=10, n_samples=10000, noise_samples=100, rng=np.random):
    """noise_gan.py:194-207: "real" rows are pure noise N(0, nstd) of the data's shape, fake rows G(z / max z)."""
    XT = sample_data(n_samples=n_samples, rng=rng)
    XT = rng.normal(0, nstd, size=[XT.shape[0], 1, XT.shape[1]])
    XN_noise = rng.normal(0, 1, size=[noise_samples, 1, noise_dim])
    XN = G.predict(XN_noise / np.max(XN_noise))
    XT = np.resize(XT, (XT.shape[0], XT.shape[2]))
    X = np.vstack((XT, XN))
    y = np.zeros((n_samples + len(XN_noise), 2))
    y[:n_samples, 1] = 1
    y[n_samples:, 0] = 1
    return X, y


def pretrain(G, D, noise_dim=10, n_samples=10000, noise_samples=10000, batch_size=32, rng=np.random):
    """noise_gan.py:209-214."""
    X, y = sample_data_and_gen(G, n_samples=n_samples, noise_samples=noise_samples, noise_dim=noise_dim, rng=rng)
    set_trainability(D, True)
    return D.fit(X, y, epochs=1, batch_size=batch_size)


def sample_noise(G, noise_dim=10, n_samples=10000, rng=np.random):
    """noise_gan.py:221-225."""
    X = rng.normal(0, 1, size=[n_samples, 1, noise_dim])
    y = np.zeros((n_samples, 2))
    y[:, 1] = 1
    return X, y


def train(GAN, G, D, epochs=500, n_samples=10000, noise_samples=noise_samples, noise_dim=10, batch_size=32, verbose=False,
          v_freq=1, rng=np.random, save_to=None):
    """noise_gan.py:227-264: D step on (noise, generated), G step through the frozen D; then D is saved
    (``best_d_weights.hdf5`` and ``d_model.hdf5`` under ``save_to`` when given)."""
    d_loss, g_loss = [], []
    for epoch in range(epochs):
        X, y = sample_data_and_gen(G, n_samples=n_samples, noise_samples=noise_samples, noise_dim=noise_dim, rng=rng)
        set_trainability(D, True)
        d_loss.append(D.train_on_batch(X, y))
        X, y = sample_noise(G, n_samples=noise_samples, noise_dim=noise_dim, rng=rng)
        set_trainability(D, False)
        g_loss.append(GAN.train_on_batch(X, y))
        if verbose and (epoch + 1) % v_freq == 0:
            print("Epoch #{}: Generative Loss: {}, Discriminative Loss: {}".format(epoch + 1, g_loss[-1], d_loss[-1]))
    if save_to is not None:
        import os
        D.save_weights(os.path.join(save_to, 'best_d_weights.hdf5'))
        D.save(os.path.join(save_to, 'd_model.hdf5'))
    return d_loss, g_loss
