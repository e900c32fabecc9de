"""Drop-in for the reference's pnn/batching.py."""
import numpy


def predict_by_batch_via_pnn(tuple_batches_float32, sess, predictor, batch_size):
    """Computes a prediction of each target patch via PNN (reference pnn/batching.py:7-88).

    Same arguments, return value and exceptions as the reference; `sess` is ignored and the whole set
    goes through libpnn_cuda in one call (the reference's loop of `sess.run` over batches of
    `batch_size` is an artefact of its static graph).
    """
    nb_predictions = tuple_batches_float32[0].shape[0]
    # reference tools/tools.py:403-434 (divide_ints_check_divisible), called at pnn/batching.py:56-58
    if not isinstance(nb_predictions, int):
        raise TypeError('`numerator` is not an instance of `int`.')
    if not isinstance(batch_size, int):
        raise TypeError('`denominator` is not an instance of `int`.')
    if nb_predictions % batch_size != 0:
        raise ValueError('`numerator` is not divisible by `denominator`.')
    if predictor.is_fully_connected:
        # reference pnn/batching.py:64-69
        width_float = numpy.sqrt(float(tuple_batches_float32[0].shape[1])/5.).item()
        if not width_float.is_integer():
            raise ValueError('`numpy.sqrt(float(tuple_batches_float32[0].shape[1])/5.)` is not a whole number.')
        width_target = int(width_float)
        return predictor.engine.predict_batch(width_target, True, tuple_batches_float32[0])
    width_target = tuple_batches_float32[0].shape[1]
    return predictor.engine.predict_batch(width_target, False, tuple_batches_float32[0], tuple_batches_float32[1])
