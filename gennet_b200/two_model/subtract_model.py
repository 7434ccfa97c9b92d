"""2_model_version/weight_version/subtract_model.py: the subtract stage.  The generator learns the signal hidden in one
noisy series x_t: the discriminator (pre-trained on noise, noise_gan.py) is shown pure noise as "real" and the residual
x_t - G(z) as "fake", so fooling it means generating the signal."""
import numpy as np

from ..nn import (Adam, BatchNormalization, Conv1D, Conv2DTranspose, Dense, Dropout, Flatten, Input, LeakyReLU, Model,
                  Reshape, regularizers, set_trainability)


class hyperparams:
    """subtract_model.py:26-54."""
    n_total = 500
    n_samples = int(n_total * 0.5)
    noise_dim = 1
    noise_samples = int(n_total * 0.5)
    batch_size = 16
    epochs = 2500
    g_lr = 1e-4
    d_lr = 1e-4
    loss = 'binary_crossentropy'
    snr = 5
    outdim = 50


def sample_data(n_samples=10000, x_vals=np.arange(0, 5, .1), max_offset=2 * np.pi, mul_range=[1, 2], snr=None,
                rng=np.random):
    """subtract_model.py:55-64."""
    snr = hyperparams.snr if snr is None else snr
    vectors = []
    for i in range(n_samples):
        offset = rng.random_sample() * max_offset
        mul = (2 * np.pi) / 5
        vectors.append(np.sin(offset + x_vals * mul) * snr)
    return np.array(vectors)


def make_gan(GAN_in, G, D):
    """subtract_model.py:72-78."""
    set_trainability(D, False)
    x = G(GAN_in)
    GAN_out = D(x)
    GAN = Model(GAN_in, GAN_out)
    GAN.compile(loss=hyperparams.loss, optimizer=G.optimizer)
    return GAN, GAN_out


def sample_data_and_gen(G, xt_train, encoder, noise_dim=10, n_samples=10000, noise_samples=100, rng=np.random):
    """subtract_model.py:80-98: noise rows N(0, 5), then the residuals x_t - G(z).  The labels are what the script
    builds: ``np.ones`` with the class columns set to 1 again, i.e. all ones."""
    XT = rng.normal(0, 5, size=[n_samples, hyperparams.outdim])
    XN_noise = rng.normal(0, 1, size=[noise_samples, 1, noise_dim])
    XN = G.predict(XN_noise)
    for s in range(noise_samples):
        XN[s] = np.subtract(xt_train[0], XN[s])
    X = np.vstack((XT, XN))
    y = np.ones((n_samples + len(XN_noise), 2))
    y[:n_samples, 1] = 1
    y[n_samples:, 0] = 1
    return X, y


def pretrain(G, D, xt_train, encoder, noise_dim=10, n_samples=10000, noise_samples=10000, batch_size=32, rng=np.random):
    """subtract_model.py:100-105."""
    X, y = sample_data_and_gen(G, xt_train, encoder, n_samples=n_samples, noise_samples=noise_samples, noise_dim=noise_dim,
                               rng=rng)
    set_trainability(D, True)
    return D.fit(X, y, epochs=1, batch_size=batch_size)


def sample_noise(G, xt_train, encoder, noise_dim=10, n_samples=10000, rng=np.random):
    """subtract_model.py:108-116."""
    X = rng.normal(0, 1, size=[n_samples, 1, noise_dim])
    y = np.ones((n_samples, 2))
    y[:, 1] = 1
    return X, y


def train(GAN, G, D, xt_train, encoder, epochs=500, n_samples=10000, noise_samples=None, noise_dim=10, batch_size=32,
          verbose=False, v_freq=1, rng=np.random):
    """subtract_model.py:118-172."""
    noise_samples = hyperparams.noise_samples if noise_samples is None else noise_samples
    d_loss, g_loss = [], []
    for epoch in range(epochs):
        X, y = sample_data_and_gen(G, xt_train, encoder, n_samples=n_samples, noise_samples=noise_samples,
                                   noise_dim=noise_dim, rng=rng)
        set_trainability(D, True)
        d_loss.append(D.train_on_batch(X, y))
        X, y = sample_noise(G, xt_train, encoder, n_samples=noise_samples, noise_dim=noise_dim, rng=rng)
        set_trainability(D, False)
        g_loss.append(GAN.train_on_batch(X, y))
        if verbose and (epoch + 1) % v_freq == 0:
            print("Epoch #{}: Generative Loss: {}, Discriminative Loss: {}".format(epoch + 1, g_loss[-1], d_loss[-1]))
    return d_loss, g_loss


def test_data_and_gen(G, xt_train, encoder, noise_dim=10, n_samples=10000, noise_samples=100, rng=np.random):
    """subtract_model.py:174-197: returns (noise rows + generated rows, residuals x_t - G(z))."""
    XT = rng.normal(0, 5, size=[n_samples, hyperparams.outdim])
    XN_noise = rng.normal(0, 1, size=[noise_samples, 1, noise_dim])
    XN = G.predict(XN_noise)
    residuals = np.array([xt_train - XN[s] for s in range(noise_samples)])
    X = np.vstack((XT, XN))
    return X, residuals


test_data_and_gen.__test__ = False      # a reference function name, not a pytest test


def get_generative(G_in, dense_dim=128, drate=0.1, out_dim=50, lr=1e-3):
    """subtract_model.py:199-251: ELU transposed-convolution generator; the first Conv2DTranspose carries
    activity_regularizer=l1(0.001) and kernel_regularizer=l2(0.01); Adam(lr, beta_1=0.5)."""
    act = 'elu'
    x = Reshape((-1, 1, 1))(G_in)
    x = BatchNormalization()(x)
    x = Conv2DTranspose(128, (1, 4), activity_regularizer=regularizers.l1(0.001), kernel_regularizer=regularizers.l2(0.01),
                        strides=(1, 1), padding='valid', activation=act)(x)
    x = BatchNormalization()(x)
    for f, k in ((64, 8), (32, 16), (16, 32)):
        x = Conv2DTranspose(f, (1, k), strides=(1, 1), padding='valid', activation=act)(x)
        x = BatchNormalization()(x)
    x = Flatten()(x)
    x = BatchNormalization()(x)
    x = Dense(out_dim, activation=act)(x)
    x = BatchNormalization()(x)
    G_out = Dense(out_dim, activation='linear')(x)
    G = Model(G_in, G_out)
    G.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5))
    return G, G_out


def get_discriminative(D_in, lr=1e-3, drate=.3, n_channels=50, conv_sz=5, leak=.2):
    """subtract_model.py:253-291."""
    x = Reshape((-1, 1))(D_in)
    x = Conv1D(50, 16)(x)
    x = LeakyReLU(alpha=0.2)(x)
    x = Dropout(drate)(x)
    x = Flatten()(x)
    x = Dense(n_channels)(x)
    x = Dropout(drate)(x)
    D_out = Dense(2, activation='sigmoid')(x)
    D = Model(D_in, D_out)
    D.compile(loss='binary_crossentropy', optimizer=Adam(lr=lr, beta_1=0.5))
    return D, D_out


def main(g_weights=None, d_weights=None, epochs=None, rng=np.random, verbose=False):
    """subtract_model.py:292-348 without the plots: one noisy series x_t = signal + N(0, 5), generator and discriminator
    built and (optionally) loaded from ``best_g_weights.hdf5`` / ``best_d_weights.hdf5``, D pre-trained on the residuals,
    then the adversarial loop.  Returns a dict with the models, the data and the loss histories."""
    ht_train = sample_data(1, rng=rng)
    xt_train = ht_train + rng.normal(0, 5, size=[1, ht_train.shape[1]])
    G_in = Input(shape=(1, hyperparams.noise_dim))
    G, G_out = get_generative(G_in, lr=hyperparams.g_lr)
    if g_weights is not None:
        G.load_weights(g_weights)
    D_in = Input(shape=(hyperparams.outdim,))
    D, D_out = get_discriminative(D_in, lr=hyperparams.d_lr)
    if d_weights is not None:
        D.load_weights(d_weights)
    GAN_in = Input((1, hyperparams.noise_dim))
    GAN, GAN_out = make_gan(GAN_in, G, D)
    encoder = []
    pretrain(G, D, xt_train, encoder, n_samples=hyperparams.n_samples, noise_samples=hyperparams.noise_samples,
             noise_dim=hyperparams.noise_dim, batch_size=hyperparams.batch_size, rng=rng)
    d_loss, g_loss = train(GAN, G, D, xt_train, encoder, epochs=hyperparams.epochs if epochs is None else epochs,
                           n_samples=hyperparams.n_samples, noise_samples=hyperparams.noise_samples,
                           noise_dim=hyperparams.noise_dim, batch_size=hyperparams.batch_size, verbose=verbose, rng=rng)
    N_VIEWED_SAMPLES = 25
    data_and_gen, residuals = test_data_and_gen(G, xt_train, encoder, noise_samples=N_VIEWED_SAMPLES,
                                                n_samples=N_VIEWED_SAMPLES, noise_dim=hyperparams.noise_dim, rng=rng)
    return {'G': G, 'D': D, 'GAN': GAN, 'ht_train': ht_train, 'xt_train': xt_train, 'd_loss': d_loss, 'g_loss': g_loss,
            'data_and_gen': data_and_gen, 'residuals': residuals}
