"""Process CPU time for the timing loops of example scripts (``pysparse.tools.sptime``)."""
import time

__all__ = ["cputime"]


def cputime():
    "Seconds of CPU time (user + system) this process has used so far."
    return time.process_time()
