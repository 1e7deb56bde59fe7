"""Logging helper (reference: munk/munk/io.py:11-20).

Log lines look like the reference's - time, file name padded to 15, level padded to 10, a colon,
the message - because the CLI's output is what users of the reference grep.  As there, importing
the module attaches one stream handler with that layout to the root logger if it has none yet.
"""
import logging

_COLUMNS = ("%(asctime)s", "%(filename)-15s", "%(levelname)-10s:", "%(message)s")


def _attach_root_handler():
    root = logging.getLogger()
    if root.handlers:
        return
    handler = logging.StreamHandler()
    handler.setFormatter(logging.Formatter(" ".join(_COLUMNS)))
    root.addHandler(handler)


_attach_root_handler()


def get_logger(verbosity=logging.INFO):
    """The package logger at level `verbosity` (the level is reset by every call, as in the reference)."""
    package_log = logging.getLogger(__name__)
    package_log.setLevel(verbosity)
    return package_log
