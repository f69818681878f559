"""Stand-in for the un-vendored PyTimer package (setup-phase stopwatch).

Derived from usage only: Timer(name).start()/.stop() and Timer.report().
Used ONLY by tests/golden/make_golden.py in the build container.
"""
import time


class Timer:
    _totals = {}

    def __init__(self, name):
        self._name = name
        self._t0 = None

    def start(self):
        self._t0 = time.perf_counter()

    def stop(self):
        if self._t0 is not None:
            Timer._totals[self._name] = (Timer._totals.get(self._name, 0.0)
                                         + time.perf_counter() - self._t0)
            self._t0 = None

    @staticmethod
    def report():
        for k, v in Timer._totals.items():
            print("%-40s %10.4f s" % (k, v))
