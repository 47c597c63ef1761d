"""Host side of align(): transcript preparation before the kernels and result-dict assembly after.

Nothing here touches emissions or the trellis — that arithmetic lives in the CUDA library.  This
module turns text into the token / word-group arrays the C ABI takes (include/wxb_align.h) and
turns the per-character / per-word arrays it returns into the reference's result records,
without pandas (reference: whisperx/alignment.py:128-177 text preparation, :252-347 tables).

What the reference does with DataFrames is restated on plain lists:
  * the per-sentence character window is inclusive of the span end          (alignment.py:288)
  * a word is the run of characters sharing a word index *inside* that window (:297-320)
  * sentence start = min start over all characters, end = max end over non-space ones (:292-294)
  * NaN sentence bounds are interpolated, then sentences with equal (start, end) are merged in
    key order and sentences whose key is still NaN are dropped              (:336-347, utils.py:433-437)
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np

LANGUAGES_WITHOUT_SPACES = ("ja", "zh")                  # alignment.py:22
PUNKT_ABBREVIATIONS = ("dr", "vs", "mr", "mrs", "prof")  # alignment.py:20

_NAN = float("nan")


# --------------------------------------------------------------------------- sentence spans
def _fallback_spans(text: str):
    """Sentence spans when nltk is not installed: a sentence ends at a run of . ! ? that is
    followed by whitespace and more text, unless the word before a single '.' is one of the
    reference's Punkt abbreviations (alignment.py:20) or a one-letter initial.  Spans exclude the
    separating whitespace, as Punkt's span_tokenize does."""
    spans = []
    start = 0
    n = len(text)
    while start < n and text[start].isspace():
        start += 1
    for m in re.finditer(r"[.!?]+(?=\s+\S)", text):
        if m.end() <= start:
            continue
        if m.group() == ".":
            word = re.search(r"(\w+)$", text[start:m.start()])
            if word and (word.group(1).lower() in PUNKT_ABBREVIATIONS or len(word.group(1)) == 1 and word.group(1).isalpha()):
                continue
        spans.append((start, m.end()))
        nxt = m.end()
        while nxt < n and text[nxt].isspace():
            nxt += 1
        start = nxt
    if start < n or not spans:
        end = len(text.rstrip()) if text.strip() else n
        spans.append((min(start, end), end))
    return spans


def _punkt_spans():
    try:
        from nltk.tokenize.punkt import PunktParameters, PunktSentenceTokenizer
    except Exception:
        return None
    params = PunktParameters()
    params.abbrev_types = set(PUNKT_ABBREVIATIONS)
    splitter = PunktSentenceTokenizer(params)
    return lambda text: list(splitter.span_tokenize(text))


_SPLITTER: List[Optional[Callable[[str], Iterable]]] = [None]


def set_sentence_splitter(fn: Optional[Callable[[str], Iterable]]) -> None:
    """Override the sentence splitter (text -> iterable of (start, end)); None restores the
    default (nltk Punkt with the reference's abbreviations when nltk is importable)."""
    _SPLITTER[0] = fn


def sentence_spans(text: str) -> list:
    fn = _SPLITTER[0]
    if fn is None:
        fn = _punkt_spans() or _fallback_spans
        _SPLITTER[0] = fn
    return [tuple(s) for s in fn(text)]


# --------------------------------------------------------------------------- stage 1
WILDCARD = "*"     # clean_char entry of a character the model does not know (wildcard extension only)


def wildcard_symbol(dictionary: dict) -> str:
    return WILDCARD if WILDCARD not in dictionary else "\0" + WILDCARD


def prepare_transcript(transcript, dictionary: dict, lang: str, print_progress: bool = False,
                       combined_progress: bool = False, wildcard: bool = False) -> None:
    """align() stage 1 (alignment.py:128-177): adds clean_char / clean_cdx / clean_wdx /
    sentence_spans to every input segment dict (the reference mutates its input the same way).
    `wildcard` (extension, off by default): characters the model does not know are kept as wildcard
    entries instead of being dropped, so that e.g. digits receive timestamps too."""
    wild = wildcard_symbol(dictionary)
    spaced = lang not in LANGUAGES_WITHOUT_SPACES
    total = len(transcript)
    for sdx, seg in enumerate(transcript):
        if print_progress:
            base = ((sdx + 1) / total) * 100
            print(f"Progress: {(50 + base / 2) if combined_progress else base:.2f}%...")
        text = seg["text"]
        first = len(text) - len(text.lstrip())
        last = len(text.rstrip()) - 1                     # index of the last non-space character
        chars, cdx = [], []
        for i in range(first, last + 1):
            c = text[i].lower()
            if spaced:
                c = c.replace(" ", "|")
            if c in dictionary:
                chars.append(c)
                cdx.append(i)
            elif wildcard:
                chars.append(wild)
                cdx.append(i)
        units = text.split(" ") if spaced else text
        seg["clean_char"] = chars
        seg["clean_cdx"] = cdx
        seg["clean_wdx"] = [k for k, w in enumerate(units) if (len(w) > 0 if wildcard else any(c in dictionary for c in w))]
        seg["sentence_spans"] = sentence_spans(text)


def blank_id_of(dictionary: dict) -> int:
    """alignment.py:237-240: the id of '[pad]' / '<pad>' (last one in dict order wins), else 0."""
    blank = 0
    for ch, code in dictionary.items():
        if ch == "[pad]" or ch == "<pad>":
            blank = code
    return blank


# --------------------------------------------------------------------------- word groups
@dataclass
class SegmentLayout:
    """Sentence windows and word groups of one segment, in the order the reference visits them."""
    word_index: List[int]                      # per character of the text (alignment.py:256-280)
    windows: List[tuple]                       # per sentence span: (lo, hi) inclusive char range
    sentence_words: List[List[tuple]]          # per sentence: [(word_text, group_id), ...]
    group_off: List[int] = field(default_factory=lambda: [0])   # [G+1] offsets into group_idx
    group_idx: List[int] = field(default_factory=list)          # clean-char index or -1 per non-space char

    @property
    def n_groups(self) -> int:
        return len(self.group_off) - 1


def layout_segment(text: str, clean_cdx: Sequence[int], spans: Sequence[tuple], lang: str) -> SegmentLayout:
    """Word index per character (:277-280) and, per sentence window (:286-288), the word groups
    whose start / end / score the device reduces (:297-320).  A group lists, for every non-space
    character of the word inside the window, its aligned-character index or -1 (a NaN row)."""
    L = len(text)
    spaced = lang not in LANGUAGES_WITHOUT_SPACES
    widx = [0] * L
    w = 0
    for i in range(L):
        widx[i] = w
        if not spaced or i == L - 1 or text[i + 1] == " ":
            w += 1
    clean_pos = {c: k for k, c in enumerate(clean_cdx)}
    lay = SegmentLayout(word_index=widx, windows=[], sentence_words=[])
    for (s0, s1) in spans:
        lo, hi = max(int(s0), 0), min(int(s1), L - 1)
        lay.windows.append((lo, hi))
        words = []
        i = lo
        while i <= hi:
            j = i
            while j + 1 <= hi and widx[j + 1] == widx[i]:
                j += 1
            wtext = text[i:j + 1].strip()
            if wtext:
                for c in range(i, j + 1):
                    if text[c] != " ":
                        lay.group_idx.append(clean_pos.get(c, -1))
                lay.group_off.append(len(lay.group_idx))
                words.append((wtext, lay.n_groups - 1))
            i = j + 1
        lay.sentence_words.append(words)
    return lay


# --------------------------------------------------------------------------- NaN interpolation
def interpolate_nans(values: Sequence[float], method: str = "nearest") -> List[float]:
    """utils.interpolate_nans (utils.py:433-437) on a list: Series.interpolate(method) when more
    than one value is known, then forward fill, then backward fill.
    'nearest' is scipy's interp1d(kind='nearest') over the row index (an equidistant row takes the
    earlier neighbour), 'linear' is positional linear interpolation; both leave leading NaNs to
    the backward fill, and 'linear' carries the last known value forward."""
    v = [float(x) for x in values]
    known = [i for i, x in enumerate(v) if not math.isnan(x)]
    if len(known) > 1:
        if method == "nearest":
            k = 0
            for i in range(known[0] + 1, known[-1]):
                while known[k + 1] < i:
                    k += 1
                if math.isnan(v[i]):
                    a, b = known[k], known[k + 1]
                    v[i] = v[a] if (i - a) <= (b - i) else v[b]
        elif method == "linear":
            xs = np.asarray(known, dtype=np.float64)
            ys = np.asarray([v[i] for i in known], dtype=np.float64)
            for i in range(known[0] + 1, len(v)):
                if math.isnan(v[i]):
                    v[i] = float(np.interp(float(i), xs, ys))
        else:
            import pandas as pd               # uncommon methods: defer to the library the reference uses
            v = pd.Series(v).interpolate(method=method).tolist()
    # ffill then bfill
    last = _NAN
    for i in range(len(v)):
        if math.isnan(v[i]):
            v[i] = last
        else:
            last = v[i]
    nxt = _NAN
    for i in range(len(v) - 1, -1, -1):
        if math.isnan(v[i]):
            v[i] = nxt
        else:
            nxt = v[i]
    return v


# --------------------------------------------------------------------------- assembly
def _fmin(vals):
    m = _NAN
    for x in vals:
        if x == x and not (m <= x):
            m = x
    return m


def _fmax(vals):
    m = _NAN
    for x in vals:
        if x == x and not (m >= x):
            m = x
    return m


def assemble_segment(text: str, clean_cdx: Sequence[int], lay: SegmentLayout, spans: Sequence[tuple], lang: str,
                     char_start: Sequence[float], char_end: Sequence[float], char_score: Sequence[float],
                     word_start: Sequence[float], word_end: Sequence[float], word_score: Sequence[float],
                     interpolate_method: str = "nearest", return_char_alignments: bool = False) -> list:
    """Records of one aligned segment (alignment.py:282-347).

    char_* are this segment's per-aligned-character values (seconds, rounded; indexable by the
    clean-char index), word_* this segment's per-group values in `lay` order (NaN = no aligned
    character in the word).  Returns the list of sentence records after NaN interpolation and
    merging of equal (start, end) keys."""
    clean_pos = {c: k for k, c in enumerate(clean_cdx)}
    subs = []
    for k, (lo, hi) in enumerate(lay.windows):
        s0, s1 = spans[k]
        idx = [clean_pos.get(c) for c in range(lo, hi + 1)]
        sent_start = _fmin(char_start[p] for p in idx if p is not None)
        sent_end = _fmax(char_end[p] for c, p in zip(range(lo, hi + 1), idx) if p is not None and text[c] != " ")
        words = []
        for wtext, g in lay.sentence_words[k]:
            rec = {"word": wtext}
            ws, we, wsc = word_start[g], word_end[g], word_score[g]
            if ws == ws:
                rec["start"] = ws
            if we == we:
                rec["end"] = we
            if wsc == wsc:
                rec["score"] = wsc
            words.append(rec)
        sub = {"text": text[s0:s1], "start": sent_start, "end": sent_end, "words": words}
        if return_char_alignments:
            chars = []
            for c, p in zip(range(lo, hi + 1), idx):
                rec = {"char": text[c]}
                if p is not None:
                    # the reference fills NaN with -1 and then drops every value equal to -1 (:329-334)
                    for key, val in (("start", char_start[p]), ("end", char_end[p]), ("score", char_score[p])):
                        if val == val and val != -1:
                            rec[key] = val
                chars.append(rec)
            sub["chars"] = chars
        subs.append(sub)
    if not subs:
        return []
    starts = interpolate_nans([s["start"] for s in subs], interpolate_method)
    ends = interpolate_nans([s["end"] for s in subs], interpolate_method)
    # groupby(["start", "end"]): keys ascending, NaN keys dropped, members in original order
    groups: dict = {}
    for s, a, b in zip(subs, starts, ends):
        if a != a or b != b:
            continue
        groups.setdefault((a, b), []).append(s)
    joiner = "" if lang in LANGUAGES_WITHOUT_SPACES else " "
    out = []
    for key in sorted(groups):
        members = groups[key]
        rec = {"start": key[0], "end": key[1], "text": joiner.join(m["text"] for m in members),
               "words": [w for m in members for w in m["words"]]}
        if return_char_alignments:
            rec["chars"] = [c for m in members for c in m["chars"]]
        out.append(rec)
    return out
