"""2_model_version of the reference (BASELINE config 5), one module per reference script:

    no_mode_collapse_network   2_model_version/*/no_mode_collapse_network.py   (= gennet_b200.twomodel)
    noise_gan                  2_model_version/weight_version/noise_gan.py     discriminator pre-training on pure noise
    subtract_model             2_model_version/weight_version/subtract_model.py   the subtract stage (x_t - G(z) into D)
    subtract_model_nw          2_model_version/no_weight_code/subtract_model.py   the newer variant (MSE discriminator)

Same function names, arguments, label layouts and RNG call order as the scripts; plotting and hard-coded output paths
are left out, model persistence goes through gennet_b200.io (Keras HDF5)."""
from .. import twomodel as no_mode_collapse_network     # noqa: F401
from . import noise_gan, subtract_model, subtract_model_nw     # noqa: F401
