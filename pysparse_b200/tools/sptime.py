"""``cputime()``: user CPU time of this process in seconds (reference: ``pysparse/tools/sptime.py``)."""
try:
    import resource

    def cputime():
        return resource.getrusage(resource.RUSAGE_SELF)[0]
except ImportError:                                   # pragma: no cover
    import time

    def cputime():
        return time.process_time()
