"""Minimal stand-in for the `natsort` package (absent offline) so the reference imports.

Only used by oracle/make_golden.py when it imports the unmodified reference in the
build container; never by the product path.
"""
import re
import sys


def natsorted(seq, reverse=False, **_kw):
    def key(s):
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]
    return sorted(seq, key=key, reverse=reverse)


natsort = sys.modules[__name__]  # `from natsort import natsort; natsort.natsorted(...)`
