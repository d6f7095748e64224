"""setup_logging(name) -> logging.Logger, as imported by src/main.py:16 and src/ply/ply.py (the reference keeps it in
src/utils/setup_logging/setup_loggin.py:14-43): INFO level, one stderr handler per logger, the same line format
("2024-01-15 12:34:56 - ply.ply - INFO - message")."""
from __future__ import annotations

import logging

_FORMAT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"
_DATEFMT = "%Y-%m-%d %H:%M:%S"


def setup_logging(name: str) -> logging.Logger:
    log = logging.getLogger(name)
    log.setLevel(logging.INFO)
    if log.hasHandlers():  # also true when an ancestor (the root logger) already has one: nothing is added twice
        return log
    h = logging.StreamHandler()
    h.setLevel(logging.INFO)
    h.setFormatter(logging.Formatter(fmt=_FORMAT, datefmt=_DATEFMT))
    log.addHandler(h)
    return log


__all__ = ["setup_logging"]
