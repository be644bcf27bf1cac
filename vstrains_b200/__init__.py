"""vstrains_b200 -- B200-native paired-end link inference for VStrains (one hot path).

Public surface: :mod:`vstrains_b200.pe_inference` (host-side mirror of the reference script,
backed by the C ABI in ``include/vspe.h``) and :mod:`vstrains_b200.synth` (test/bench inputs)."""
__all__ = ["pe_inference", "synth"]
