"""Export / wire formats: expected strings derived by hand from the reference's JavaScript
(public/static/app.js:buildSRT :1741, buildTextGrid :1582, buildEAF :1381) and app.py:119-154."""
import xml.etree.ElementTree as ET

from omnilingual_asr.diarization.pipeline import DiarizedTranscriptSegment, WordTimestamp
from omnilingual_asr.export import build_eaf, build_srt, build_textgrid, result_dict

SEGS = [
    DiarizedTranscriptSegment(start=0.352, end=3.6605, speaker="Speaker 1", text='four "score" & seven', words=[
        WordTimestamp("four", 0.352, 0.612), WordTimestamp("score", 0.772, 1.072)]),
    DiarizedTranscriptSegment(start=3661.9996, end=3663.25, speaker="Speaker 2", text="now <we>", language="English",
                              language_code="en", translation="now <we>"),
]


def test_result_dict_matches_the_web_app_shape():
    d = result_dict(SEGS, summary="s", detected_languages=None)
    assert list(d) == ["segments", "summary"]
    assert d["segments"][0] == {"start": 0.352, "end": 3.6605, "speaker": "Speaker 1", "text": 'four "score" & seven',
                                "words": [{"word": "four", "start": 0.352, "end": 0.612},
                                          {"word": "score", "start": 0.772, "end": 1.072}]}
    assert d["segments"][1]["language"] == "English" and d["segments"][1]["words"] == []
    assert "emotion" not in d["segments"][1]


def test_srt_cues_speaker_prefix_and_the_rounding_quirk():
    d = result_dict(SEGS)
    # 3.6605 % 1 is 0.66049999... in IEEE doubles (JavaScript and Python alike): 660, not 661
    want = ("1\n00:00:00,352 --> 00:00:03,660\n[Speaker 1] four \"score\" & seven\n\n"
            "2\n01:01:01,1000 --> 01:01:03,250\n[Speaker 2] now <we>\n")      # (t % 1) * 1000 rounds to 1000: as the JS
    assert build_srt(d) == want
    one = result_dict(SEGS[:1])
    assert build_srt(one).splitlines()[2] == 'four "score" & seven'            # single speaker: no prefix


def test_textgrid_tiers_words_and_gap_filling():
    tg = build_textgrid(result_dict(SEGS))
    lines = tg.splitlines()
    assert lines[0] == 'File type = "ooTextFile"' and lines[3] == "xmin = 0 " and lines[4] == "xmax = 3663.250000"
    assert "size = 2" in lines
    # speaker 1: gap [0, 0.352), word, gap, word, gap to the end
    i = lines.index('        name = "Speaker 1"')
    assert lines[i + 3] == "        intervals: size = 5"
    assert lines[i + 5:i + 8] == ["            xmin = 0.000000 ", "            xmax = 0.352000", '            text = ""']
    assert '            text = "four"' in lines and '            text = "score"' in lines
    # speaker 2 has no words: the whole segment is one interval, quotes doubled in names/text
    j = lines.index('        name = "Speaker 2"')
    assert lines[j + 3] == "        intervals: size = 2"
    assert lines[-1] == '            text = "now <we>"'


def test_eaf_is_wellformed_and_follows_the_tier_rules():
    eaf = build_eaf(dict(result_dict(SEGS), audio_url='a&b".wav'), date="2026-01-01T00:00:00.000Z")
    root = ET.fromstring(eaf)
    assert root.tag == "ANNOTATION_DOCUMENT" and root.attrib["AUTHOR"] == "OmniTranscribe"
    slots = root.find("TIME_ORDER").findall("TIME_SLOT")
    assert [s.attrib["TIME_VALUE"] for s in slots] == ["352", "3661", "3662000", "3663250"]   # Math.round(t * 1000)
    tiers = {t.attrib["TIER_ID"]: t for t in root.findall("TIER")}
    assert set(tiers) == {"Speaker 1", "Speaker 2", "Speaker 2_language"}    # translation == text: no translation tier
    assert tiers["Speaker 2_language"].find("ANNOTATION/ALIGNABLE_ANNOTATION/ANNOTATION_VALUE").text == "en"
    assert tiers["Speaker 1"].find("ANNOTATION/ALIGNABLE_ANNOTATION/ANNOTATION_VALUE").text == 'four "score" & seven'
    assert 'MEDIA_URL="a&amp;b&quot;.wav"' in eaf
    types = [t.attrib["LINGUISTIC_TYPE_ID"] for t in root.findall("LINGUISTIC_TYPE")]
    assert types == ["transcription", "language", "translation"]           # hasTranslation is true, the tier is empty


def test_eaf_with_words_rebuilds_the_reference_golden_byte_for_byte():
    """The one reference-held golden of the export layer: /root/reference/gettysburg.eaf (written by the former local
    pipeline; ELAN 3.0 with a transcription tier and a time-aligned word tier per speaker).  Its CONTENT is committed as
    tests/golden/gettysburg_eaf.json (tests/golden/make_eaf_fixture.py); build_eaf_with_words must reproduce the FILE -
    tier / linguistic-type / time-slot layout, id numbering, attribute order, whitespace - checked by sha256."""
    import hashlib
    import json
    from pathlib import Path
    from omnilingual_asr.export import build_eaf_with_words
    g = json.loads((Path(__file__).parent / "golden" / "gettysburg_eaf.json").read_text())
    doc = build_eaf_with_words({"segments": g["segments"]}, media_url=g["media_url"],
                               relative_media_url=g["relative_media_url"], date=g["date"])
    assert hashlib.sha256(doc.encode("utf-8")).hexdigest() == g["sha256"]
    assert doc.count("<TIME_SLOT ") == g["n_time_slots"] == 2 * (len(g["segments"]) + sum(len(s["words"]) for s in g["segments"]))


def test_eaf_with_words_from_pipeline_segments():
    """The same writer on what the pipeline returns (seconds, WordTimestamp objects): integer milliseconds, one word tier
    per speaker, segments without words contribute none."""
    import xml.etree.ElementTree as ET
    from omnilingual_asr.export import build_eaf_with_words
    from omnilingual_asr.models.inference.ctc_pipeline import CTCTranscriptSegment, WordTimestamp
    segs = [CTCTranscriptSegment(0.352, 1.0, "Speaker 1", "a <b>", words=[WordTimestamp("a", 0.4, 0.5), WordTimestamp("<b>", 0.6, 0.95)]),
            CTCTranscriptSegment(2.0, 3.5, "Speaker 1", "c", words=None)]
    root = ET.fromstring(build_eaf_with_words({"segments": segs}, date="2026-01-01T00:00:00Z"))
    tiers = {t.attrib["TIER_ID"]: t for t in root.findall("TIER")}
    assert set(tiers) == {"Speaker 1", "Speaker 1_words"}
    slots = {t.attrib["TIME_SLOT_ID"]: int(t.attrib["TIME_VALUE"]) for t in root.find("TIME_ORDER")}
    assert [slots[f"ts{i}"] for i in range(1, 9)] == [352, 1000, 400, 500, 600, 950, 2000, 3500]
    words = [a.find("ANNOTATION_VALUE").text for a in tiers["Speaker 1_words"].findall("ANNOTATION/ALIGNABLE_ANNOTATION")]
    assert words == ["a", "<b>"]
    ids = [a.attrib["ANNOTATION_ID"] for a in tiers["Speaker 1"].findall("ANNOTATION/ALIGNABLE_ANNOTATION")]
    assert ids == ["a1", "a4"]
