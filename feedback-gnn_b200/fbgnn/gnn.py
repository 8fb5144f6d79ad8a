"""Weight (de)serialisation for the feedback GNN -- ``save_weights`` / ``load_weights`` of the
reference (``sionna/fec/ldpc/gnn.py:755-791``) without TensorFlow.

The reference pickles ``system.get_weights()``.  The four files it ships under ``weights/``
were written from a list of ``tf.Tensor`` -- the pickle stream calls
``tensorflow.python.framework.ops.convert_to_tensor(ndarray)`` and refers to numpy's old
``numpy.core.multiarray`` module path.  ``read_weights`` maps both to plain
numpy so the shipped files load as a list of 12 float32 arrays; ``save_weights`` writes a
pickle of plain ndarrays, which the reference's ``load_weights`` (``pickle.load`` +
``set_weights``) reads unchanged.
"""
import io
import os
import pickle

import numpy as np

WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


class _WeightsUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("tensorflow"):
            if name == "convert_to_tensor":
                return lambda x, *a, **k: np.asarray(x)
            raise pickle.UnpicklingError(f"unsupported TensorFlow object {module}.{name} in weights file")
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core", 1)
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                return super().find_class(module.replace("numpy._core", "numpy.core", 1), name)
        return super().find_class(module, name)


def read_weights(model_path):
    """Return the list of float32 arrays stored in a reference weights file."""
    with open(model_path, "rb") as f:
        data = f.read()
    weights = _WeightsUnpickler(io.BytesIO(data)).load()
    return [np.ascontiguousarray(np.asarray(w), dtype=np.float32) for w in weights]


def save_weights(system, model_path):
    """Save ``system.get_weights()`` to ``model_path`` (gnn.py:757-772)."""
    weights = [np.asarray(w) for w in system.get_weights()]
    with open(model_path, "wb") as f:
        pickle.dump(weights, f)


def load_weights(system, model_path):
    """Load the weights stored at ``model_path`` into ``system`` (gnn.py:774-791)."""
    system.set_weights(read_weights(model_path))
