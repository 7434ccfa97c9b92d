"""Digest (names, shapes, dtypes, data hashes, attributes) of the four Keras HDF5 files the reference ships
(2_model_version/weight_version/*.hdf5, written by h5py/libhdf5), as read by gennet_b200.hdf5.
Run in the build container (the reference is mounted at /root/reference):  python tests/golden/make_hdf5_digest.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from tests.test_io import _digest, REF_DIR, GOLDEN      # noqa: E402

out = {}
for name in sorted(os.listdir(REF_DIR)):
    if name.endswith('.hdf5'):
        out[name] = _digest(os.path.join(REF_DIR, name))
with open(GOLDEN, 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
print('wrote', GOLDEN, {k: len(v['datasets']) for k, v in out.items()})
