"""B200-native engine for the prediction neural networks (PNN) of
thierrydumas/context_adaptive_neural_network_based_prediction.

The product is `csrc/libpnn_cuda.so` (hand-written sm_100a kernels behind the C ABI of
`include/pnn_cuda.h`); this package is its Python host side:

  engine.Engine                    one handle = the nets loaded on one GPU
  weights                          flat-binary weight format, seeded initialisers, checkpoint export
  pnn.PredictionNeuralNetwork      drop-in for the reference's pnn/PredictionNeuralNetwork.py
  pnn.batching                     drop-in for the reference's pnn/batching.py
  offline                          image-sharded batched evaluation (one NCCL gather of the statistics)
"""
from .engine import Engine, PnnError, MEAN_TRAINING_LUMINANCE  # noqa: F401
