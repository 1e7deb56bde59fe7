"""Logging helper with the reference's format (reference: munk/munk/io.py:11-20)."""
import logging

LOG_FORMAT = '%(asctime)s %(filename)-15s %(levelname)-10s: %(message)s'
logging.basicConfig(format=LOG_FORMAT)


def get_logger(verbosity=logging.INFO):
    """One package-wide logger, level reset on every call (as munk.io.get_logger does)."""
    log = logging.getLogger(__name__)
    log.setLevel(verbosity)
    return log
