"""Stand-in for the un-vendored third-party package `torch_geometric` (TEST INFRASTRUCTURE).

The reference (neuraloperator/graph-pde) imports torch_geometric / torch_scatter, which are not
installed here and cannot be (no network, no wheels).  This package restates the minimal part of the
PyG ~1.3 API that graph-neural-operator/nn_conv.py:3-4,242,261-271 touches, so that the reference's
own files can be executed UNMODIFIED in this container to produce golden vectors
(oracle/gen_golden.py).  Parity against PyG itself is therefore *unpinned* at this boundary; the two
judgement calls are (i) mean over an empty neighbourhood = 0 (count clamped to >= 1) and
(ii) summation order unspecified.  Nothing in the product imports this package.
"""
