"""B200-native GenCast denoiser + DPM-Solver++ 2S sampler hot path.

Host side mirrors the reference's Python interface (gencast.GenCast,
denoiser.Denoiser, dpm_solver_plus_plus_2s.Sampler, rollout.chunked_prediction);
the compute is hand-written sm_100a CUDA behind the C ABI declared in
include/gencast_b200.h.  There is no CPU fallback: importing the compute layer
without the built shared library raises.
"""

__version__ = "0.1.0"
