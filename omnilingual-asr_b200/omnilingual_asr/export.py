"""Wire and export formats on the far side of the path: what the reference's web layer does with the
segments the pipeline returns, restated server-side so that the existing UI / tools work on the CTC backend.

  result_dict      the JSON shape of workflows/wav2elan_web/app.py:_run_transcription (:111-154)
  build_srt        public/static/app.js:buildSRT      (:1741-1771)
  build_textgrid   public/static/app.js:buildTextGrid (:1582-1688)
  build_eaf        public/static/app.js:buildEAF      (:1381-1576)

Behaviour follows the JavaScript, quirks included (they are what downstream files already look like):
millisecond fields are Math.round((t % 1) * 1000), so 1.9996 s prints as "00:00:01,1000"; an EAF translation tier
is only emitted for translations that differ from the text; TextGrid gaps shorter than 1 ms are not filled.
"""
from __future__ import annotations

import math
from datetime import datetime, timezone
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence


def _get(seg: Any, name: str, default: Any = None) -> Any:
    return seg.get(name, default) if isinstance(seg, Mapping) else getattr(seg, name, default)


def _js_round(x: float) -> int:
    """Math.round: half up (towards +infinity)."""
    return int(math.floor(x + 0.5))


def result_dict(segments: Iterable[Any], summary: Optional[str] = None,
                detected_languages: Optional[Sequence[Mapping[str, Any]]] = None) -> Dict[str, Any]:
    """{"segments": [...], "summary"?, "detected_languages"?}; optional fields appear only when truthy
    (app.py:119-154)."""
    out: List[Dict[str, Any]] = []
    for seg in segments:
        d: Dict[str, Any] = {
            "start": _get(seg, "start"), "end": _get(seg, "end"), "speaker": _get(seg, "speaker"),
            "text": _get(seg, "text"),
            "words": [{"word": _get(w, "word"), "start": _get(w, "start"), "end": _get(w, "end")}
                      for w in (_get(seg, "words") or [])],
        }
        for key in ("language", "language_code", "languages", "emotion", "translation"):
            v = _get(seg, key)
            if v:
                d[key] = v
        out.append(d)
    result: Dict[str, Any] = {"segments": out}
    if summary:
        result["summary"] = summary
    if detected_languages:
        result["detected_languages"] = list(detected_languages)
    return result


def _srt_time(seconds: float) -> str:
    h = int(seconds // 3600)
    m = int((seconds % 3600) // 60)
    s = int(seconds % 60)
    ms = _js_round((seconds % 1) * 1000)
    return f"{h:02d}:{m:02d}:{s:02d},{ms:03d}"


def build_srt(data: Mapping[str, Any]) -> str:
    """Numbered cues `HH:MM:SS,mmm --> HH:MM:SS,mmm`; a `[speaker] ` prefix when there is more than one speaker."""
    segs = data["segments"]
    many = len({_get(s, "speaker") for s in segs}) > 1
    lines: List[str] = []
    for i, seg in enumerate(segs, 1):
        lines.append(str(i))
        lines.append(f"{_srt_time(_get(seg, 'start'))} --> {_srt_time(_get(seg, 'end'))}")
        lines.append(f"[{_get(seg, 'speaker')}] {_get(seg, 'text')}" if many else _get(seg, "text"))
        lines.append("")
    return "\n".join(lines)


def build_textgrid(data: Mapping[str, Any]) -> str:
    """Praat long TextGrid: one IntervalTier per speaker, words as intervals when a segment has them, gaps (> 1 ms)
    filled with empty intervals up to the last segment end."""
    segs = data["segments"]
    max_time = 0.0
    for seg in segs:
        max_time = max(max_time, _get(seg, "end"))
    speakers: List[str] = []
    for seg in segs:
        if _get(seg, "speaker") not in speakers:
            speakers.append(_get(seg, "speaker"))
    tiers: Dict[str, List[Dict[str, Any]]] = {sp: [] for sp in speakers}
    for seg in segs:
        words = _get(seg, "words") or []
        if words:
            for w in words:
                tiers[_get(seg, "speaker")].append({"xmin": _get(w, "start"), "xmax": _get(w, "end"), "text": _get(w, "word")})
        else:
            tiers[_get(seg, "speaker")].append({"xmin": _get(seg, "start"), "xmax": _get(seg, "end"), "text": _get(seg, "text")})
    for sp in speakers:
        filled: List[Dict[str, Any]] = []
        last_end = 0.0
        for iv in sorted(tiers[sp], key=lambda v: v["xmin"]):
            if iv["xmin"] > last_end + 0.001:
                filled.append({"xmin": last_end, "xmax": iv["xmin"], "text": ""})
            filled.append(iv)
            last_end = iv["xmax"]
        if last_end < max_time - 0.001:
            filled.append({"xmin": last_end, "xmax": max_time, "text": ""})
        tiers[sp] = filled

    def t(v: float) -> str:
        return f"{v:.6f}"

    def esc(s: str) -> str:
        return s.replace('"', '""')

    tg = ('File type = "ooTextFile"\nObject class = "TextGrid"\n\nxmin = 0 \n'
          f"xmax = {t(max_time)}\n\ntiers? <exists> \nsize = {len(speakers)}\nitem []:\n")
    for ti, sp in enumerate(speakers, 1):
        ivs = tiers[sp]
        tg += (f"    item [{ti}]:\n        class = \"IntervalTier\" \n        name = \"{esc(sp)}\"\n"
               f"        xmin = 0 \n        xmax = {t(max_time)}\n        intervals: size = {len(ivs)}\n")
        for ii, iv in enumerate(ivs, 1):
            tg += (f"        intervals [{ii}]:\n            xmin = {t(iv['xmin'])} \n            xmax = {t(iv['xmax'])}\n"
                   f"            text = \"{esc(iv['text'])}\"\n")
    return tg


def _xml(s: str) -> str:
    return s.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;").replace('"', "&quot;")


def build_eaf(data: Mapping[str, Any], date: Optional[str] = None) -> str:
    """ELAN 3.0 document: two time slots per segment (ms), a `transcription` tier per speaker, and `<speaker>_language`
    / `_emotion` / `_translation` tiers when any segment carries those fields.  `date` defaults to now (UTC, ISO)."""
    segs = data["segments"]

    def valid_translation(seg: Any) -> bool:
        tr = _get(seg, "translation")
        return bool(tr) and tr != "null" and tr.strip() != ""

    has_language = any(_get(s, "language") for s in segs)
    has_emotion = any(_get(s, "emotion") for s in segs)
    has_translation = any(valid_translation(s) for s in segs)
    speakers: List[str] = []
    slots: List[str] = []
    ann: Dict[str, List[tuple]] = {"transcription": [], "language": [], "emotion": [], "translation": []}
    ts_id = ann_id = 1
    for seg in segs:
        sp = _get(seg, "speaker")
        if sp not in speakers:
            speakers.append(sp)
        ts1, ts2 = f"ts{ts_id}", f"ts{ts_id + 1}"
        ts_id += 2
        slots.append(f'        <TIME_SLOT TIME_SLOT_ID="{ts1}" TIME_VALUE="{_js_round(_get(seg, "start") * 1000)}"/>')
        slots.append(f'        <TIME_SLOT TIME_SLOT_ID="{ts2}" TIME_VALUE="{_js_round(_get(seg, "end") * 1000)}"/>')

        def add(kind: str, tier: str, value: str) -> None:
            nonlocal ann_id
            ann[kind].append((tier, ts1, ts2, value, f"a{ann_id}"))
            ann_id += 1

        add("transcription", sp, _get(seg, "text"))
        if _get(seg, "language"):
            add("language", f"{sp}_language", _get(seg, "language_code") or _get(seg, "language"))
        if _get(seg, "emotion"):
            add("emotion", f"{sp}_emotion", _get(seg, "emotion"))
        if valid_translation(seg) and _get(seg, "translation") != _get(seg, "text"):
            add("translation", f"{sp}_translation", _get(seg, "translation"))

    def tier_xml(kind: str, tier_id: str) -> str:
        body = "\n".join(
            "            <ANNOTATION>\n"
            f'                <ALIGNABLE_ANNOTATION ANNOTATION_ID="{aid}" TIME_SLOT_REF1="{a}" TIME_SLOT_REF2="{b}">\n'
            f"                    <ANNOTATION_VALUE>{_xml(value)}</ANNOTATION_VALUE>\n"
            "                </ALIGNABLE_ANNOTATION>\n"
            "            </ANNOTATION>"
            for (tier, a, b, value, aid) in ann[kind] if tier == tier_id)
        if not body and kind != "transcription":
            return ""
        return f'        <TIER LINGUISTIC_TYPE_REF="{kind}" TIER_ID="{tier_id}">\n{body}\n        </TIER>'

    transcript_tiers = "\n".join(tier_xml("transcription", sp) for sp in speakers)
    additional = ""
    if has_language:
        additional += "\n".join(x for x in (tier_xml("language", f"{sp}_language") for sp in speakers) if x) + "\n"
    if has_emotion:
        additional += "\n".join(x for x in (tier_xml("emotion", f"{sp}_emotion") for sp in speakers) if x) + "\n"
    if has_translation:
        additional += "\n".join(x for x in (tier_xml("translation", f"{sp}_translation") for sp in speakers) if x)
    types = '    <LINGUISTIC_TYPE LINGUISTIC_TYPE_ID="transcription" TIME_ALIGNABLE="true"/>'
    for flag, kind in ((has_language, "language"), (has_emotion, "emotion"), (has_translation, "translation")):
        if flag:
            types += f'\n    <LINGUISTIC_TYPE LINGUISTIC_TYPE_ID="{kind}" TIME_ALIGNABLE="true"/>'
    if date is None:
        date = datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.") + f"{datetime.now(timezone.utc).microsecond // 1000:03d}Z"
    slots_xml = "\n".join(slots)
    return (
        '<?xml version="1.0" encoding="UTF-8"?>\n'
        f'<ANNOTATION_DOCUMENT AUTHOR="OmniTranscribe" DATE="{date}" FORMAT="3.0" VERSION="3.0" '
        'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" '
        'xsi:noNamespaceSchemaLocation="http://www.mpi.nl/tools/elan/EAFv3.0.xsd">\n'
        '    <HEADER MEDIA_FILE="" TIME_UNITS="milliseconds">\n'
        f'        <MEDIA_DESCRIPTOR MEDIA_URL="{_xml(data.get("audio_url", ""))}" MIME_TYPE="audio/x-wav"/>\n'
        "    </HEADER>\n    <TIME_ORDER>\n"
        f"{slots_xml}\n    </TIME_ORDER>\n{transcript_tiers}\n{additional}\n{types}\n</ANNOTATION_DOCUMENT>")


def build_eaf_with_words(data: Mapping[str, Any], *, media_url: str = "", relative_media_url: str = "",
                         date: Optional[str] = None) -> str:
    """ELAN 3.0 document in the layout of the reference's bundled `gettysburg.eaf` (written by the former LOCAL pipeline,
    wav2elan: /root/reference/gettysburg.eaf:1-135): a `transcription` tier per speaker plus a time-aligned `word` tier
    `<speaker>_words`, both carrying PARTICIPANT; annotation and time-slot ids run segment by segment, each segment
    followed by its words (a1 ts1/ts2, then a2.. for its words, then the next segment); times in integer
    milliseconds.  Needs `word_timestamps=True` results for the word tier; segments without words get none.
    tests/test_export.py rebuilds the reference file byte for byte from its own content (tests/golden/gettysburg_eaf.json).
    """
    segs = data["segments"]
    speakers: List[str] = []
    slots: List[str] = []
    seg_ann: Dict[str, List[str]] = {}
    word_ann: Dict[str, List[str]] = {}
    ts_id = ann_id = 1

    def ms(seg: Any, name: str) -> int:
        v = _get(seg, name + "_ms")
        return int(v) if v is not None else _js_round(_get(seg, name) * 1000)

    def annotation(value: str, t0: int, t1: int) -> str:
        nonlocal ts_id, ann_id
        a, b = f"ts{ts_id}", f"ts{ts_id + 1}"
        slots.append(f'    <TIME_SLOT TIME_SLOT_ID="{a}" TIME_VALUE="{t0}"/>')
        slots.append(f'    <TIME_SLOT TIME_SLOT_ID="{b}" TIME_VALUE="{t1}"/>')
        xml = ("    <ANNOTATION>\n"
               f'      <ALIGNABLE_ANNOTATION ANNOTATION_ID="a{ann_id}" TIME_SLOT_REF1="{a}" TIME_SLOT_REF2="{b}">\n'
               f"        <ANNOTATION_VALUE>{_xml(value)}</ANNOTATION_VALUE>\n"
               "      </ALIGNABLE_ANNOTATION>\n"
               "    </ANNOTATION>")
        ts_id += 2
        ann_id += 1
        return xml

    for seg in segs:
        sp = _get(seg, "speaker")
        if sp not in speakers:
            speakers.append(sp)
            seg_ann[sp] = []
            word_ann[sp] = []
        seg_ann[sp].append(annotation(_get(seg, "text"), ms(seg, "start"), ms(seg, "end")))
        for w in (_get(seg, "words") or []):
            word_ann[sp].append(annotation(_get(w, "word"), ms(w, "start"), ms(w, "end")))
    if date is None:
        date = datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")
    out = ['<?xml version="1.0" encoding="utf-8"?>',
           f'<ANNOTATION_DOCUMENT xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" AUTHOR="" DATE="{date}" '
           'FORMAT="3.0" VERSION="3.0" xsi:noNamespaceSchemaLocation="http://www.mpi.nl/tools/elan/EAFv3.0.xsd">',
           '  <HEADER MEDIA_FILE="" TIME_UNITS="milliseconds">',
           f'    <MEDIA_DESCRIPTOR MEDIA_URL="{_xml(media_url)}" MIME_TYPE="audio/wav" '
           f'RELATIVE_MEDIA_URL="{_xml(relative_media_url)}"/>',
           "  </HEADER>",
           '  <LOCALE LANG_ID="und"/>',
           '  <LINGUISTIC_TYPE LINGUISTIC_TYPE_ID="transcription" TIME_ALIGNABLE="true" GRAPHIC_REFERENCES="false"/>',
           '  <LINGUISTIC_TYPE LINGUISTIC_TYPE_ID="word" TIME_ALIGNABLE="true" GRAPHIC_REFERENCES="false"/>',
           "  <TIME_ORDER>"] + slots + ["  </TIME_ORDER>"]
    for sp in speakers:
        out.append(f'  <TIER TIER_ID="{_xml(sp)}" LINGUISTIC_TYPE_REF="transcription" PARTICIPANT="{_xml(sp)}">')
        out += seg_ann[sp]
        out.append("  </TIER>")
        if word_ann[sp]:
            out.append(f'  <TIER TIER_ID="{_xml(sp)}_words" LINGUISTIC_TYPE_REF="word" PARTICIPANT="{_xml(sp)}">')
            out += word_ann[sp]
            out.append("  </TIER>")
    out.append("</ANNOTATION_DOCUMENT>")
    return "\n".join(out) + "\n"


__all__ = ["result_dict", "build_srt", "build_textgrid", "build_eaf", "build_eaf_with_words"]
