"""Host utilities with the semantics of mdsea/helpers.py (logging, n-sphere volume, dt, progress)."""
import logging
import math
import os
import statistics
import time
from logging.handlers import RotatingFileHandler

from scipy.special import gamma

from mdsea_b200.constants.fileman import DIR_SIMFILES


def setup_logging(format_file="%(asctime)s - [%(levelname)s:%(name)s] %(message)s",
                  format_stout="[%(levelname)s:%(name)s] %(message)s",
                  level=logging.INFO,
                  filename=f"{DIR_SIMFILES}/_mdsea_logdump.log",
                  maxbytes=500000) -> None:
    """Stream + rotating-file handlers on the package logger (mdsea/helpers.py:18-51)."""
    root = logging.getLogger("mdsea_b200")
    root.setLevel(level)
    stream = logging.StreamHandler()
    stream.setFormatter(logging.Formatter(format_stout))
    stream.setLevel(level)
    root.addHandler(stream)
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    fileh = RotatingFileHandler(filename, maxBytes=maxbytes, backupCount=1)
    fileh.setFormatter(logging.Formatter(format_file))
    fileh.setLevel(level)
    root.addHandler(fileh)


def nsphere_volume(n, r):
    """pi^(n/2) r^n / Gamma(n/2 + 1)  (mdsea/helpers.py:54-56)."""
    return math.pi ** (n / 2) * (r ** n) / gamma(n / 2 + 1)


def get_dt(radius: float, mean_speed: float, drpf: float = 0.01) -> float:
    """dt such that a particle at the mean speed moves drpf radii per step (mdsea/helpers.py:89-100)."""
    if mean_speed == 0:
        mean_speed = 1
    return drpf * radius / mean_speed


class ProgressBar:
    """Progress / ETA log lines in the reference's format (mdsea/helpers.py:113-179)."""

    def __init__(self, name: str, stop: int, loggername: str = __name__, step: int = 10) -> None:
        self.name, self.stop, self.step = name, stop, step
        self.log = logging.getLogger(loggername)
        self.prefix = f"ProgressBar ({name}):"
        self.percentage = 0
        self._cycle_times = []
        self.t_lastcycle = self.tstart = time.time()
        self.tfinish = 0.0

    def set_start(self, t=None):
        self.tstart = t or time.time()
        self.log.debug("%s %s", self.prefix, {"start-time": self.tstart})

    def set_finish(self, t=0.0):
        self.tfinish = t or time.time()
        self.log.debug("%s %s", self.prefix, {"end-time": self.tfinish})

    def get_duration(self, ndigits: int = 1) -> float:
        if not self.tfinish:
            self.log.warning("You forgot to finish the '%s' ProgressBar! Setting the finished time now.", self.name)
            self.set_finish()
        return float(round(self.tfinish - self.tstart, ndigits))

    def log_duration(self, ndigits: int = 1) -> None:
        self.log.info("%s %s", self.prefix, {"status": "finished", "lifetime": f"{self.get_duration(ndigits)}s"})

    def log_progress(self, step) -> None:
        percent = round(100 * (step + 1) / self.stop)
        rounded = int(self.step * round(float(percent) / self.step))
        if self.percentage >= rounded or rounded % self.step:
            return
        cycles_left = (100 - percent) / self.step
        now = time.time()
        if step <= 1:
            eta = round(cycles_left * (now - self.t_lastcycle))
        else:
            self._cycle_times.append(now - self.t_lastcycle)
            eta = round(cycles_left * statistics.mean(self._cycle_times))
        self.log.info("%s %s", self.prefix, {"step": f"{step + 1}/{self.stop}", "percentage": percent, "ETA": f"{eta}s"})
        self.percentage = rounded
        self.t_lastcycle = time.time()
