"""Logging with the reference's level names incl. PERF (devito/logger.py:29)."""
import logging
import sys

__all__ = ['logger', 'debug', 'info', 'perf', 'warning', 'error', 'set_log_level']

PERF = 19
logging.addLevelName(PERF, 'PERF')
logger = logging.getLogger('devito_b200')
if not logger.handlers:
    _h = logging.StreamHandler(sys.stderr)
    _h.setFormatter(logging.Formatter('%(message)s'))
    logger.addHandler(_h)
_levels = {'DEBUG': logging.DEBUG, 'PERF': PERF, 'INFO': logging.INFO,
           'WARNING': logging.WARNING, 'ERROR': logging.ERROR, 'CRITICAL': logging.CRITICAL}


def set_log_level(level):
    logger.setLevel(_levels.get(str(level).upper(), logging.INFO))


def debug(msg, *a): logger.debug(msg, *a)
def info(msg, *a): logger.info(msg, *a)
def perf(msg, *a): logger.log(PERF, msg, *a)
def warning(msg, *a): logger.warning(msg, *a)
def error(msg, *a): logger.error(msg, *a)


from .parameters import configuration  # noqa: E402
set_log_level(configuration['log-level'])
