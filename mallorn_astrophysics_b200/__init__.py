"""B200-native per-object light-curve featurisation (MALLORN 222-feature table).

Host side is Python (mirroring the reference's ``src/features`` API); all compute is
hand-written sm_100a CUDA behind the C-ABI declared in ``include/mallorn_b200.h``.
There is no CPU fallback: importing the compute entry points without the built
``libmallorn_b200.so`` or without a CUDA device raises.
"""
__version__ = "0.1.0"
