"""Drop-in for the reference's pnn/PredictionNeuralNetwork.py (inference side).

Same constructor arguments, attributes and `initialization(sess, path_to_restore)` call; the graph is
not built with TensorFlow but loaded into libpnn_cuda from a PNNW flat binary.
"""
import os
import tempfile

from .. import engine as _engine
from .. import weights as _weights

# reference pnn/PredictionNeuralNetwork.py:8
NB_ITERS_TRAINING = 800000

_ENGINES = {}


def get_engine(device=0, mean_training=_engine.MEAN_TRAINING_LUMINANCE):
    """One engine per (device, mean) -- plays the role of the `tf.Session` the reference passes around."""
    key = (device, float(mean_training))
    if key not in _ENGINES:
        _ENGINES[key] = _engine.Engine(mean_training=mean_training, device=device)
    return _ENGINES[key]


class PredictionNeuralNetwork(object):
    """Prediction neural network (reference pnn/PredictionNeuralNetwork.py:16-200)."""

    def __init__(self, batch_size, width_target, is_fully_connected, tuple_coeffs=None, dict_reading=None,
                 device=0, mean_training=_engine.MEAN_TRAINING_LUMINANCE):
        # reference pnn/PredictionNeuralNetwork.py:73-74
        if dict_reading is not None and tuple_coeffs is None:
            raise ValueError('`dict_reading` is not None while `tuple_coeffs` is None')
        if tuple_coeffs is not None:
            raise NotImplementedError('libpnn_cuda is an inference engine: the optimisation graph '
                                      '(reference pnn/components.py:263-368) is out of scope')
        self.batch_size = batch_size
        self.width_target = width_target
        self.is_fully_connected = is_fully_connected
        if not is_fully_connected:
            # reference pnn/PredictionNeuralNetwork.py:126-133 (KeyError for an unsupported width, as there)
            self.strides_branch = _weights.STRIDES_BRANCH[width_target]
        elif width_target not in (4, 8, 16, 32, 64):
            raise ValueError('`width_target` does not belong to {4, 8, 16, 32, 64}')
        self.engine = get_engine(device, mean_training)
        self.path_to_flat_binary = None

    def initialization(self, sess, path_to_restore, seed=0):
        """Restores a model (reference pnn/PredictionNeuralNetwork.py:185-200).

        `sess` is accepted and ignored.  `path_to_restore` ends with ".ckpt" (or ".pnnw"); the flat
        binary is its sibling `<stem>.pnnw`, exported on the fly when only a TensorFlow V2 bundle
        (`.index` + `.data-00000-of-00001`) is found.  An empty string initialises the variables like
        `tf.global_variables_initializer()` does, with the reference's initialisers (seeded).
        """
        if path_to_restore:
            if path_to_restore.endswith('.pnnw'):
                path = path_to_restore
            else:
                stem = path_to_restore[:-5] if path_to_restore.endswith('.ckpt') else path_to_restore
                path = stem + '.pnnw'
                if not os.path.exists(path):
                    if not os.path.exists(path_to_restore + '.index'):
                        raise IOError('neither "%s" nor the checkpoint bundle "%s.index" exists'
                                      % (path, path_to_restore))
                    target = path if os.access(os.path.dirname(path) or '.', os.W_OK) else \
                        os.path.join(tempfile.mkdtemp(prefix='pnnw_'), os.path.basename(path))
                    _weights.export_checkpoint(path_to_restore, self.width_target, self.is_fully_connected, target)
                    path = target
        else:
            path = os.path.join(tempfile.mkdtemp(prefix='pnnw_'), 'init_%d_%d.pnnw'
                                % (self.width_target, int(self.is_fully_connected)))
            _weights.save_flat(path, self.width_target, self.is_fully_connected,
                               _weights.init_weights(self.width_target, self.is_fully_connected, seed))
        self.engine.load_net(path)
        self.path_to_flat_binary = path
