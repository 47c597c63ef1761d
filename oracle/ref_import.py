"""TEST INFRASTRUCTURE ONLY — imports the *real* reference (whisperX 3.1.1) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It is used
  * by scripts/make_golden.py to generate the committed fixtures under tests/golden/, and
  * by tests marked `needs_reference` to pin oracle/ref_port.py and oracle/align_oracle.c
    against the reference itself.
Nothing in whisperx_b200/, bench.py or smoke() may import this module.

Recipe: the reference package __init__ pulls faster_whisper /
ctranslate2 / pyannote (absent here), so a bare namespace object stands in for the package and
only `whisperx.alignment` (+ its relative imports audio/utils/types) is executed.  `nltk` is not
installed, so a stub Punkt with a pluggable splitter is registered.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

# /root/reference exists in the build container only.  `python -m oracle.make_ref` (run by
# __graft_entry__.build()) copies the files this importer executes into the git-ignored oracle/_ref/,
# which travels to the GPU box with the snapshot: there the reference is imported from that copy.
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.environ.get("WXB_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref"))


def _find_root() -> str:
    for root in _CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "whisperx", "alignment.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "whisperx", "alignment.py"))


def reference_is_live_checkout() -> bool:
    """True when the importer runs the files where they lie under /root/reference (not the build-time copy)."""
    return reference_available() and os.path.abspath(REFERENCE_ROOT) != os.path.join(_HERE, "_ref")


# The splitter both sides of a parity test share.  Signature: text -> list[(start, end)].
_SPLITTER = [lambda text: [(0, len(text))]]


def set_sentence_splitter(fn) -> None:
    _SPLITTER[0] = fn


def _install_nltk_stub() -> None:
    if "nltk" in sys.modules and not getattr(sys.modules["nltk"], "_wxb_stub", False):
        return  # a real nltk is importable; leave it alone
    nltk = types.ModuleType("nltk")
    nltk._wxb_stub = True
    tok = types.ModuleType("nltk.tokenize")
    punkt = types.ModuleType("nltk.tokenize.punkt")

    class PunktParameters:  # attribute bag, like the real one
        def __init__(self):
            self.abbrev_types = set()

    class PunktSentenceTokenizer:
        def __init__(self, params=None):
            self.params = params

        def span_tokenize(self, text):
            return list(_SPLITTER[0](text))

    punkt.PunktParameters = PunktParameters
    punkt.PunktSentenceTokenizer = PunktSentenceTokenizer
    tok.punkt = punkt
    nltk.tokenize = tok
    sys.modules["nltk"] = nltk
    sys.modules["nltk.tokenize"] = tok
    sys.modules["nltk.tokenize.punkt"] = punkt


_CACHED = None
_CACHED_DIARIZE = None


def load_reference_diarize():
    """The reference's `whisperx.diarize` module (unmodified).  pyannote is not installed, and the
    module only needs its `Pipeline` name at import time, so an empty stand-in is registered."""
    global _CACHED_DIARIZE
    if _CACHED_DIARIZE is not None:
        return _CACHED_DIARIZE
    load_reference_alignment()                 # registers the bare `whisperx` package object
    if "pyannote.audio" not in sys.modules:
        pa = types.ModuleType("pyannote")
        paa = types.ModuleType("pyannote.audio")
        paa.Pipeline = type("Pipeline", (), {})
        pa.audio = paa
        sys.modules.setdefault("pyannote", pa)
        sys.modules["pyannote.audio"] = paa
    _CACHED_DIARIZE = importlib.import_module("whisperx.diarize")
    return _CACHED_DIARIZE



def load_reference_utils():
    """The reference's `whisperx.utils` module (writers, `format_timestamp`), unmodified."""
    load_reference_alignment()
    return importlib.import_module("whisperx.utils")


def load_reference_alignment():
    """Return the reference's `whisperx.alignment` module (unmodified code)."""
    global _CACHED
    if _CACHED is not None:
        return _CACHED
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    try:
        import nltk  # noqa: F401
    except Exception:
        _install_nltk_stub()
    if "whisperx" not in sys.modules or not hasattr(sys.modules["whisperx"], "__path__"):
        pkg = types.ModuleType("whisperx")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "whisperx")]
        sys.modules["whisperx"] = pkg
    _CACHED = importlib.import_module("whisperx.alignment")
    return _CACHED
