"""``mx.log`` — logger factory with the glog-like one-letter level prefix (parity: python/mxnet/log.py: ``get_logger(name, filename, filemode,
level)``; format ``L MMDD HH:MM:SS pid file:line] message``)."""
from __future__ import annotations

import logging
import sys

__all__ = ["get_logger", "getLogger", "CRITICAL", "ERROR", "WARNING", "INFO", "DEBUG", "NOTSET"]

CRITICAL, ERROR, WARNING, INFO, DEBUG, NOTSET = logging.CRITICAL, logging.ERROR, logging.WARNING, logging.INFO, logging.DEBUG, logging.NOTSET


class _Formatter(logging.Formatter):
    def __init__(self, colored=True):
        self._colored = colored
        super().__init__(datefmt="%m%d %H:%M:%S")

    def format(self, record):
        letter = {logging.CRITICAL: "C", logging.ERROR: "E", logging.WARNING: "W", logging.INFO: "I", logging.DEBUG: "D"}.get(record.levelno, "U")
        color = {logging.CRITICAL: "\x1b[31m", logging.ERROR: "\x1b[31m", logging.WARNING: "\x1b[33m", logging.DEBUG: "\x1b[32m"}.get(record.levelno, "\x1b[34m")
        head = "%s%s %s %d %s:%d]" % (letter, self.formatTime(record, self.datefmt), "", record.process, record.filename, record.lineno)
        if self._colored:
            head = color + head + "\x1b[0m"
        self._style._fmt = head + " %(message)s"
        return super().format(record)


def get_logger(name=None, filename=None, filemode=None, level=WARNING):
    logger = logging.getLogger(name)
    if name is not None and not getattr(logger, "_init_done", None):
        logger._init_done = True
        if filename:
            hdlr = logging.FileHandler(filename, filemode or "a")
            hdlr.setFormatter(_Formatter(colored=False))
        else:
            hdlr = logging.StreamHandler()
            hdlr.setFormatter(_Formatter(colored=sys.stderr.isatty()))
        logger.addHandler(hdlr)
        logger.setLevel(level)
    return logger


getLogger = get_logger
