"""CPU oracle for the PNN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, on the CPU, the reference's algorithm for the prediction
neural network (PNN) forward pass (reference repository
thierrydumas/context_adaptive_neural_network_based_prediction; every function
cites the reference file:line it follows).  It exists only to CHECK the CUDA
path.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package
(`context_adaptive_neural_network_based_prediction_b200/`) never does and fails
loudly when its CUDA library is missing.

Parity pinning status (see DESIGN.md "Oracle"):
  * context gather / mask / mean subtraction: PINNED -- checked against the
    known-answer strings of the reference's own C++ tests
    (hevc/hm_common/c++/source_test/tests.cpp:248-644) and against the
    reference's `extraction_context.cpp` compiled unmodified into
    `oracle/_ref/libextract_ref.so` (recipe: oracle/Makefile).
  * network arithmetic (FC / conv / merger / transposed conv): PARITY UNPINNED
    against TensorFlow -- the arithmetic lives in TensorFlow 1.x (Python
    1.4.2/1.5.1, C++ 1.9.0; not vendored in the reference, not installable
    offline) and the reference's tests hold no numeric golden for it.  The
    restatement follows TensorFlow's published op semantics (SAME padding,
    conv2d_transpose as the gradient of conv2d) and is cross-checked by an
    independent loop implementation and by prediction PSNR on the two shipped
    pretrained checkpoints (CONV-4, CONV-8).
"""
