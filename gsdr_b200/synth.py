"""Re-export of harness.synth (kept so that `from gsdr_b200 import synth` keeps working in tests and tools)."""
from harness.synth import *  # noqa: F401,F403
from harness.synth import hash_uniform, lowpass_taps, random_taps, tone_plus_noise  # noqa: F401
