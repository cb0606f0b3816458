from .synth import synth_reads, synth_reads_fixed  # noqa: F401
from .weights import load_weights, default_weights_path, STATE_KEYS  # noqa: F401
