"""covomix_b200: B200-native (sm_100a) implementation of the CoVoMix inference hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI library
``libcovomix_b200.so``), ``_native`` (ctypes binding), ``packing`` (checkpoint -> packed
device weights), ``flow`` / ``vocoder`` (host-side mirrors of the reference's
``ConditionalFlowMatcherWrapper.sample`` and ``Generator.__call__``), ``dropin`` (swap-in
for the reference's entry scripts), ``synthetic`` (seeded random checkpoints/inputs).
"""
__version__ = "0.1.0"
