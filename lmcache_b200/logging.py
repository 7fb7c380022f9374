"""Module loggers (mirrors lmcache/logging.py:4-14: one formatted stream handler, DEBUG level)."""
import logging

_FORMAT = "%(levelname)s lmcache_b200: %(message)s [%(asctime)s]"
_configured = False


def init_logger(name: str) -> logging.Logger:
    global _configured
    if not _configured:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(_FORMAT))
        root = logging.getLogger("lmcache_b200")
        root.addHandler(handler)
        root.setLevel(logging.WARNING)
        root.propagate = False
        _configured = True
    return logging.getLogger(name if name.startswith("lmcache_b200") else f"lmcache_b200.{name}")
