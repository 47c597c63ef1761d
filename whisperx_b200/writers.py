"""Result writers consuming the word timings `align()` produces — host-side mirror of the reference's
writer interface (`whisperx/utils.py:170-431`).

Same names, options and output bytes as the reference (`get_writer(fmt, output_dir)(result, audio_path,
{"max_line_width", "max_line_count", "highlight_words"})`, called at `whisperx/transcribe.py:156,227`), but
organised differently: the subtitle layout is one pure function (`layout_cues`) that turns the word stream
into cues of (words, owning segment), and each format is a thin renderer over it.  Pinned by
`tests/golden/writer_cases.json` (generated from the reference's own classes) and by a live differential
test against the reference when it is present.

Layout rules restated from `utils.py:224-282`:
  * `max_line_width` None means 1000; without both limits every input segment is one cue (":231").
  * a word continues the current line when the line is not empty, the word still fits, there was no
    pause of more than 3 s since the last timed word (only when both limits are given) and the word
    does not open a new segment in keep-segments mode (":247-255").
  * otherwise the word is stripped and starts a new cue — when the cue already holds words and
    (`max_line_count` is set and (long pause or the cue is full)) or a new segment begins — or a new
    line inside the cue ("\\n" prefix) when the line was not empty (":257-275").
  * a cue is stamped with the start / end / speaker of the segment of its *first* word (":285").
"""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass, field
from typing import Callable, Iterator, List, Optional, TextIO, Tuple

from .alignment import LANGUAGES_WITHOUT_SPACES

__all__ = ["format_timestamp", "layout_cues", "Cue", "ResultWriter", "WriteTXT", "SubtitlesWriter", "WriteVTT",
           "WriteSRT", "WriteTSV", "WriteAudacity", "WriteJSON", "get_writer"]

PAUSE_SECONDS = 3.0          # utils.py:245
UNLIMITED_WIDTH = 1000       # utils.py:230


def format_timestamp(seconds: float, always_include_hours: bool = False, decimal_marker: str = ".") -> str:
    """`[HH:]MM:SS<marker>mmm` of a non-negative time (utils.py:170-188): milliseconds are rounded with
    Python's round() first, then split."""
    assert seconds >= 0, "non-negative timestamp expected"
    total_ms = round(seconds * 1000.0)
    rest, ms = divmod(total_ms, 1000)
    rest, sec = divmod(rest, 60)
    hours, minutes = divmod(rest, 60)
    head = f"{hours:02d}:" if (always_include_hours or hours > 0) else ""
    return f"{head}{minutes:02d}:{sec:02d}{decimal_marker}{ms:03d}"


@dataclass
class Cue:
    """One subtitle: its words (copies; the `word` text carries the line breaks) and the segment that owns it."""
    words: List[dict] = field(default_factory=list)
    start: float = 0.0
    end: float = 0.0
    speaker: Optional[str] = None


def layout_cues(segments: List[dict], max_line_width: Optional[int], max_line_count: Optional[int]) -> Iterator[Cue]:
    """Break the word stream of `segments` into cues and lines (utils.py:235-282)."""
    width = UNLIMITED_WIDTH if max_line_width is None else max_line_width
    keep_segments = max_line_count is None or max_line_width is None
    if not segments:
        return
    cue = Cue()
    used = 0                  # characters on the current line
    lines = 1                 # lines in the current cue
    last_start = segments[0]["start"]
    for seg in segments:
        for pos, source in enumerate(seg["words"]):
            word = dict(source)
            timed = "start" in word
            paused = (not keep_segments) and timed and (word["start"] - last_start > PAUSE_SECONDS)
            fits = used + len(word["word"]) <= width
            opens_segment = keep_segments and pos == 0 and bool(cue.words)
            if used > 0 and fits and not paused and not opens_segment:
                used += len(word["word"])
            else:
                word["word"] = word["word"].strip()
                cue_full = max_line_count is not None and (paused or lines >= max_line_count)
                if (cue.words and cue_full) or opens_segment:
                    yield cue
                    cue = Cue()
                    lines = 1
                elif used > 0:
                    lines += 1
                    word["word"] = "\n" + word["word"]
                used = len(word["word"].strip())
            if not cue.words:
                cue.start, cue.end, cue.speaker = seg["start"], seg["end"], seg.get("speaker")
            cue.words.append(word)
            if timed:
                last_start = word["start"]
    if cue.words:
        yield cue


_UNDERLINE = re.compile(r"^(\s*)(.*)$")


class ResultWriter:
    """Base: `writer(result, audio_path, options)` writes `<output_dir>/<audio basename>.<extension>` (utils.py:191-208)."""
    extension: str

    def __init__(self, output_dir: str):
        self.output_dir = output_dir

    def __call__(self, result: dict, audio_path: str, options: dict):
        stem = os.path.splitext(os.path.basename(audio_path))[0]
        with open(os.path.join(self.output_dir, f"{stem}.{self.extension}"), "w", encoding="utf-8") as fh:
            self.write_result(result, file=fh, options=options)

    def write_result(self, result: dict, file: TextIO, options: dict):
        raise NotImplementedError


class WriteTXT(ResultWriter):
    extension = "txt"

    def write_result(self, result: dict, file: TextIO, options: dict):
        for seg in result["segments"]:
            file.write(seg["text"].strip() + "\n")
        file.flush()


class SubtitlesWriter(ResultWriter):
    always_include_hours: bool
    decimal_marker: str

    def format_timestamp(self, seconds: float) -> str:
        return format_timestamp(seconds, self.always_include_hours, self.decimal_marker)

    def iterate_result(self, result: dict, options: dict) -> Iterator[Tuple[str, str, str]]:
        """(start, end, text) triples (utils.py:224-338)."""
        width, count, highlight = options["max_line_width"], options["max_line_count"], options["highlight_words"]
        segments = result["segments"]
        if len(segments) == 0:
            return
        if "words" not in segments[0]:
            # un-aligned result: one cue per segment, "-->" defused (utils.py:331-338)
            for seg in segments:
                text = seg["text"].strip().replace("-->", "->")
                if "speaker" in seg:
                    text = f"[{seg['speaker']}]: {text}"
                yield self.format_timestamp(seg["start"]), self.format_timestamp(seg["end"]), text
            return
        joiner = "" if result["language"] in LANGUAGES_WITHOUT_SPACES else " "
        for cue in layout_cues(segments, width, count):
            begin, end = self.format_timestamp(cue.start), self.format_timestamp(cue.end)
            label = f"[{cue.speaker}]: " if cue.speaker is not None else ""
            texts = [w["word"] for w in cue.words]
            whole = label + joiner.join(texts)
            if not (highlight and any("start" in w for w in cue.words)):
                yield begin, end, whole
                continue
            # one cue per timed word with that word underlined, plus the plain text over the gaps (utils.py:302-327)
            cursor = begin
            for k, w in enumerate(cue.words):
                if "start" not in w:
                    continue
                w0, w1 = self.format_timestamp(w["start"]), self.format_timestamp(w["end"])
                if cursor != w0:
                    yield cursor, w0, whole
                marked = [_UNDERLINE.sub(r"\1<u>\2</u>", t) if j == k else t for j, t in enumerate(texts)]
                yield w0, w1, label + " ".join(marked)
                cursor = w1


class WriteVTT(SubtitlesWriter):
    extension = "vtt"
    always_include_hours = False
    decimal_marker = "."

    def write_result(self, result: dict, file: TextIO, options: dict):
        file.write("WEBVTT\n\n")
        for begin, end, text in self.iterate_result(result, options):
            file.write(f"{begin} --> {end}\n{text}\n\n")
        file.flush()


class WriteSRT(SubtitlesWriter):
    extension = "srt"
    always_include_hours = True
    decimal_marker = ","

    def write_result(self, result: dict, file: TextIO, options: dict):
        number = 0
        for begin, end, text in self.iterate_result(result, options):
            number += 1
            file.write(f"{number}\n{begin} --> {end}\n{text}\n\n")
        file.flush()


class WriteTSV(ResultWriter):
    """start / end in integer milliseconds, tab-separated, with a header line (utils.py:355-376)."""
    extension = "tsv"

    def write_result(self, result: dict, file: TextIO, options: dict):
        file.write("start\tend\ttext\n")
        for seg in result["segments"]:
            text = seg["text"].strip().replace("\t", " ")
            file.write(f"{round(1000 * seg['start'])}\t{round(1000 * seg['end'])}\t{text}\n")
        file.flush()


class WriteAudacity(ResultWriter):
    """Audacity label track: seconds, no header, speaker in [[...]] before the text (utils.py:378-397)."""
    extension = "aud"

    def write_result(self, result: dict, file: TextIO, options: dict):
        for seg in result["segments"]:
            who = f"[[{seg['speaker']}]]" if "speaker" in seg else ""
            text = seg["text"].strip().replace("\t", " ")
            file.write(f"{seg['start']}\t{seg['end']}\t{who}{text}\n")
        file.flush()


class WriteJSON(ResultWriter):
    extension = "json"

    def write_result(self, result: dict, file: TextIO, options: dict):
        json.dump(result, file, ensure_ascii=False)


_STANDARD = {"txt": WriteTXT, "vtt": WriteVTT, "srt": WriteSRT, "tsv": WriteTSV, "json": WriteJSON}
_OPTIONAL = {"aud": WriteAudacity}          # not part of "all" (utils.py:415-417)


def get_writer(output_format: str, output_dir: str) -> Callable[[dict, str, dict], None]:
    """Writer for one format, or for every standard format with "all" (utils.py:406-431)."""
    if output_format == "all":
        every = [cls(output_dir) for cls in _STANDARD.values()]

        def write_all(result: dict, file, options: dict):
            for w in every:
                w(result, file, options)

        return write_all
    if output_format in _OPTIONAL:
        return _OPTIONAL[output_format](output_dir)
    return _STANDARD[output_format](output_dir)
