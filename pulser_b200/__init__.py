"""pulser_b200: B200-native time-evolution emulator for Pulser sequences."""
from ._compat import HAVE_PULSER  # noqa: F401  (installs the import hooks)

__version__ = "0.1.0"
