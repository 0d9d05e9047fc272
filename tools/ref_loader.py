"""Import the unmodified reference (dobrosketchkun/dither_pie) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tools/make_golden.py and tools/extract_tables.py -- never by tests, bench.py or the package.
``pywt`` is absent here and only needed by the out-of-scope wavelet mode, so a stub module is
injected before import.
"""
import sys
import types

REF = "/root/reference"


def load():
    sys.modules.setdefault("pywt", types.ModuleType("pywt"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dithering_lib  # noqa: E402
    import video_processor  # noqa: E402
    return dithering_lib, video_processor
