"""2_model_version/no_weight_code/subtract_model.py: the newer subtract stage (no pre-trained weights; soft labels;
MSE discriminator with GaussianNoise + BatchNormalization(axis=1) blocks and global average pooling).

``get_generative`` of that script asks Conv2DTranspose for dilation rates (1, 9), (1, 7), (1, 2), (1, 3).  Keras 2.2.4
(requirements.txt:16) routes a dilated transposed convolution to ``tf.nn.atrous_conv2d_transpose`` after
``assert dilation_rate[0] == dilation_rate[1]`` (keras/backend/tensorflow_backend.py, conv2d_transpose), so the script's
generator fails while the graph is built; the builder here fails the same way and says why."""
import numpy as np

from ..nn import (Adam, BatchNormalization, Conv1D, Conv2DTranspose, Dense, GaussianNoise, GlobalAveragePooling1D,
                  GlobalAveragePooling2D, Input, LeakyReLU, Model, Reshape, set_trainability)


class hyperparams:
    """no_weight_code/subtract_model.py:44-58."""
    n_total = 100
    n_samples = int(n_total * 0.5)
    noise_dim = 100
    noise_samples = int(n_total * 0.5)
    epochs = 1000
    batch_size = 4
    g_lr = 1e-4
    d_lr = 1e-4
    loss = 'binary_crossentropy'
    snr = 5
    outdim = 50
    outdir = 'output/'


def sample_data(n_samples=10000, x_vals=np.arange(0, 5, .1), max_offset=2 * np.pi, mul_range=[1, 2], snr=None,
                rng=np.random):
    """:60-69."""
    snr = hyperparams.snr if snr is None else snr
    vectors = []
    for i in range(n_samples):
        offset = rng.random_sample() * max_offset
        mul = (2 * np.pi) / 5
        vectors.append(np.sin(offset + x_vals * mul) * snr)
    return np.array(vectors)


def make_gan(GAN_in, G, D):
    """:77-84."""
    set_trainability(D, False)
    x = G(GAN_in)
    GAN_out = D(x)
    GAN = Model(GAN_in, GAN_out)
    GAN.compile(loss=hyperparams.loss, optimizer=G.optimizer, metrics=['accuracy'])
    return GAN, GAN_out


def sample_data_and_gen(G, xt_train, encoder, epoch, noise_dim=10, n_samples=10000, noise_samples=100, rng=np.random):
    """:86-121: half a batch of noise N(0, snr), half a batch of residuals G(z) - x_t, soft labels (one uniform draw
    per label group, in the script's order)."""
    h = int(hyperparams.batch_size / 2)
    XT = rng.normal(0, hyperparams.snr, size=[h, hyperparams.outdim])
    XN_noise = rng.normal(0, 1, size=[h, 1, noise_dim])
    XN = G.predict(XN_noise)
    for s in range(h):
        XN[s] = np.subtract(XN[s], xt_train[0])
    X = np.vstack((XT, XN))
    y = np.zeros((hyperparams.batch_size, 2))
    y[:h, 0] = rng.uniform(0.7, 1)
    y[h:, 1] = rng.uniform(0.7, 1)
    y[:h, 1] = rng.uniform(0, 0.3)
    y[h:, 0] = rng.uniform(0, 0.3)
    return X, y


def pretrain(G, D, xt_train, encoder, noise_dim=10, n_samples=10000, noise_samples=10000, batch_size=32, rng=np.random):
    """:123-129."""
    X, y = sample_data_and_gen(G, xt_train, encoder, 1, n_samples=n_samples, noise_samples=noise_samples,
                               noise_dim=noise_dim, rng=rng)
    set_trainability(D, True)
    return D.fit(X, y, epochs=1, batch_size=batch_size)


def sample_noise(G, xt_train, encoder, noise_dim=10, n_samples=10000, rng=np.random):
    """:132-141."""
    X = rng.normal(0, 1, size=[hyperparams.batch_size, 1, noise_dim])
    y = np.zeros((hyperparams.batch_size, 2))
    y[:, 0] = 1
    y[:, 1] = 0
    return X, y


def train(GAN, G, D, xt_train, encoder, epochs=500, n_samples=10000, noise_samples=None, noise_dim=10, batch_size=32,
          verbose=False, v_freq=1, rng=np.random):
    """:143-201: as the script, D and GAN each take TWO optimizer steps per epoch (``train_on_batch(X, y)[0]`` and
    ``train_on_batch(X, y)[1]`` are separate calls)."""
    d_loss, d_acc, g_loss, g_acc = [], [], [], []
    for epoch in range(epochs):
        X, y = sample_data_and_gen(G, xt_train, encoder, epoch, n_samples=n_samples, noise_samples=noise_samples,
                                   noise_dim=noise_dim, rng=rng)
        set_trainability(D, True)
        d_loss.append(D.train_on_batch(X, y)[0])
        d_acc.append(D.train_on_batch(X, y)[1])
        X, y = sample_noise(G, xt_train, encoder, n_samples=noise_samples, noise_dim=noise_dim, rng=rng)
        set_trainability(D, False)
        g_loss.append(GAN.train_on_batch(X, y)[0])
        g_acc.append(GAN.train_on_batch(X, y)[1])
        if verbose and (epoch + 1) % v_freq == 0:
            print("Epoch #{}: Generative Loss: {}, Acc: {} Discriminative Loss: {}, Acc: {}".format(
                epoch + 1, g_loss[-1], g_acc[-1], d_loss[-1], d_acc[-1]))
    return d_loss, g_loss, d_acc, g_acc


def test_data_and_gen(G, xt_train, encoder, noise_dim=10, n_samples=10000, noise_samples=100, rng=np.random):
    """:203-224."""
    XT = rng.normal(0, hyperparams.snr, size=[n_samples, hyperparams.outdim])
    XN_noise = rng.normal(0, 1, size=[noise_samples, 1, noise_dim])
    XN = G.predict(XN_noise)
    residuals = np.array([xt_train - XN[s] for s in range(noise_samples)])
    X = np.vstack((XT, XN))
    return X, residuals


test_data_and_gen.__test__ = False


def get_generative(G_in, dense_dim=128, drate=0.5, out_dim=50, lr=1e-3):
    """:226-320 as written: Dense(1024, tanh) -> BN(axis=1) -> Reshape -> BN(axis=1) -> dilated, strided
    Conv2DTranspose blocks -> GlobalAveragePooling2D -> Dense(out_dim).  Raises at the first Conv2DTranspose, as it does
    under the reference's pinned Keras (see the module docstring)."""
    act, padding = 'tanh', 'same'
    x = Dense(1024, activation=act)(G_in)
    x = BatchNormalization(axis=1)(x)
    x = Reshape((-1, 1, 1))(x)
    x = BatchNormalization(axis=1)(x)
    for f, k, s, d, noise in ((512, 9, 2, 9, True), (256, 7, 1, 7, False), (128, 3, 1, 2, False), (64, 2, 1, 3, False)):
        x = Conv2DTranspose(f, (1, k), strides=(1, s), dilation_rate=(1, d), padding=padding, activation=act)(x)
        x = LeakyReLU(alpha=0.2)(x)
        if noise:
            x = GaussianNoise(1)(x)
        x = BatchNormalization(axis=1)(x)
    x = GlobalAveragePooling2D()(x)
    G_out = Dense(out_dim, activation='linear')(x)
    G = Model(G_in, G_out)
    G.compile(loss=hyperparams.loss, optimizer=Adam(lr=lr, beta_1=0.5, decay=1e-4), metrics=['accuracy'])
    return G, G_out


def get_discriminative(D_in, lr=1e-3, drate=.3, n_channels=50, conv_sz=5, leak=.2):
    """:322-390: Conv1D(128|256|512, 8, tanh) -> LeakyReLU(0.2) -> GaussianNoise(1.6) -> BatchNormalization(axis=1),
    GlobalAveragePooling1D, Dense(2, sigmoid); mean_squared_error, Adam(lr, beta_1=0.5, decay=1e-4), accuracy."""
    padding, act, strides, gauss_noise = 'valid', 'tanh', 1, 1.6
    x = Reshape((-1, 1))(D_in)
    for f in (128, 256, 512):
        x = Conv1D(f, 8, padding=padding, strides=strides, activation=act)(x)
        x = LeakyReLU(alpha=0.2)(x)
        x = GaussianNoise(gauss_noise)(x)
        x = BatchNormalization(axis=1)(x)
    x = GlobalAveragePooling1D()(x)
    D_out = Dense(2, activation='sigmoid')(x)
    D = Model(D_in, D_out)
    D.compile(loss='mean_squared_error', optimizer=Adam(lr=lr, beta_1=0.5, decay=1e-4), metrics=['accuracy'])
    return D, D_out
