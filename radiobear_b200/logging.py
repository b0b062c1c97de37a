"""Append-only text log (reference: logging.py:8-40)."""
import os


class LogIt:
    def __init__(self, log):
        if isinstance(log, str):
            self.logfile = log
            d = os.path.dirname(log)
            if d and not os.path.isdir(d):
                os.makedirs(d)
            self.fp = open(log, 'a')
        else:
            self.logfile = None
            self.fp = log

    def add(self, msg, printOut=True):
        if self.fp is not None:
            self.fp.write(msg + '\n')
            if printOut:
                print(msg)

    def close(self):
        if self.fp is not None:
            self.fp.close()
            self.fp = None


def setup(log):
    return log if isinstance(log, LogIt) else LogIt(log)
