"""Extracts the CONTENT of the reference's bundled ELAN file (/root/reference/gettysburg.eaf, written by the former local
pipeline) into tests/golden/gettysburg_eaf.json: header fields, segments with their words (integer milliseconds) and
the sha256 of the file itself.  tests/test_export.py rebuilds the document from this content with
omnilingual_asr.export.build_eaf_with_words and compares the hash - the reference file does not travel, its content does.
    python tests/golden/make_eaf_fixture.py        (build container only: reads /root/reference)"""
import hashlib
import json
import xml.etree.ElementTree as ET
from pathlib import Path

REF = Path("/root/reference/gettysburg.eaf")
raw = REF.read_bytes()
root = ET.fromstring(raw)
slot = {t.attrib["TIME_SLOT_ID"]: int(t.attrib["TIME_VALUE"]) for t in root.find("TIME_ORDER")}
md = root.find("HEADER/MEDIA_DESCRIPTOR").attrib
tiers = {t.attrib["TIER_ID"]: t for t in root.findall("TIER")}
segments = []
for tid, tier in tiers.items():
    if tier.attrib["LINGUISTIC_TYPE_REF"] != "transcription":
        continue
    words = []
    wt = tiers.get(tid + "_words")
    if wt is not None:
        for a in wt.findall("ANNOTATION/ALIGNABLE_ANNOTATION"):
            words.append({"word": a.find("ANNOTATION_VALUE").text, "start_ms": slot[a.attrib["TIME_SLOT_REF1"]],
                          "end_ms": slot[a.attrib["TIME_SLOT_REF2"]], "id": int(a.attrib["ANNOTATION_ID"][1:])})
    for a in tier.findall("ANNOTATION/ALIGNABLE_ANNOTATION"):
        segments.append({"speaker": tier.attrib["PARTICIPANT"], "text": a.find("ANNOTATION_VALUE").text,
                         "start_ms": slot[a.attrib["TIME_SLOT_REF1"]], "end_ms": slot[a.attrib["TIME_SLOT_REF2"]],
                         "id": int(a.attrib["ANNOTATION_ID"][1:]), "words": []})
    segments.sort(key=lambda s: s["id"])
    for i, s in enumerate(segments):      # a word belongs to the segment whose annotation id precedes it
        hi = segments[i + 1]["id"] if i + 1 < len(segments) else 1 << 30
        s["words"] = [{k: w[k] for k in ("word", "start_ms", "end_ms")} for w in words if s["id"] < w["id"] < hi]
for s in segments:
    del s["id"]
out = {"source": "gettysburg.eaf of Nathan-Roll1/omnilingual-asr", "sha256": hashlib.sha256(raw).hexdigest(),
       "date": root.attrib["DATE"], "media_url": md["MEDIA_URL"], "relative_media_url": md["RELATIVE_MEDIA_URL"],
       "n_time_slots": len(slot), "segments": segments}
Path(__file__).with_name("gettysburg_eaf.json").write_text(json.dumps(out, indent=1) + "\n")
print(len(segments), "segments,", sum(len(s["words"]) for s in segments), "words,", len(slot), "time slots")
