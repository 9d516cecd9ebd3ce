from . import signal, stats  # noqa: F401
