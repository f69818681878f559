"""Stand-in for the un-vendored PyTab package the reference imports.

Written from the usage sites only (print indentation helper): the reference
formats a Tab() with %s / str.format and calls indent()/unindent() around
nested solves.  Used ONLY by tests/golden/make_golden.py in the build
container; it never travels into the product path.
"""


class Tab:
    _depth = 0

    def indent(self):
        Tab._depth += 1

    def unindent(self):
        Tab._depth = max(0, Tab._depth - 1)

    def __str__(self):
        return "  " * Tab._depth

    def __format__(self, spec):
        return format(str(self), spec)
