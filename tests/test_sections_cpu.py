"""Host logic of ciclope_b200.sections (deck -> sections -> hashes) on the golden decks of the reference."""
import numpy as np
import pytest

from ciclope_b200 import sections
from tests import cases

NAMES = [c["name"] for c in cases.deck_cases()]


@pytest.mark.parametrize("name", NAMES)
def test_sections_partition_the_deck(golden, name):
    deck = bytes(golden[name + "/deck"])
    sec = sections.split_deck(deck)
    spans = list(sec.values())
    assert spans[0][0] == 0 and spans[-1][1] == len(deck)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    assert deck[:sec["header"][1]].endswith(b"*NODE\n")
    assert deck[sec["elements"][0]:].startswith(b"** Elements and Element sets")
    kw = {c["name"]: c for c in cases.deck_cases()}[name]["m"].get("keywords", ["NSET", "ELSET"])
    assert ("nset" in sec) == ("NSET" in kw) and ("elset" in sec) == ("ELSET" in kw)
    if "nset" in sec:
        assert deck[sec["nset"][0]:].startswith(b"** Additional nset")
    if "elset" in sec:
        assert deck[sec["elset"][0]:].startswith(b"** Additional elset")
    tmpl = open(kw_template(name), "rb").read()
    tail = deck[sec["tail"][0]:]
    # the tail is the PROPERTY cards (if any) followed by the template copy
    assert tail.startswith(b"** User material") or tail[:20] == tmpl[:20] or b"refnode" in tmpl
    # every node line is 53 bytes in these cases unless the coordinates are wide
    if "wide" not in name:
        assert (sec["nodes"][1] - sec["nodes"][0]) % 53 == 0


def kw_template(name):
    return {c["name"]: c for c in cases.deck_cases()}[name]["m"]["templatefile"]


def test_hashes_ignore_the_timestamp_only(golden):
    deck = bytes(golden["bool_5x7x9/deck"])
    other = deck.replace(b"<timestamp>", b"2026-10-20 12:00:00.123456") if b"<timestamp>" in deck else deck
    lines = deck.split(b"\n")
    lines[1] = b"** Abaqus .INP file written by Voxel_FEM script on 2026-10-20 12:00:00"
    other = b"\n".join(lines)
    ok, per = sections.compare_decks(deck, other)
    assert ok and all(per.values())
    broken = other.replace(b",", b";", 1)
    ok, per = sections.compare_decks(deck, broken)
    assert not ok and [k for k, v in per.items() if not v] == ["nodes"]
    assert sections.section_hashes(deck)["nodes"]["bytes"] == per_len(deck)


def per_len(deck):
    sec = sections.split_deck(deck)
    return sec["nodes"][1] - sec["nodes"][0]


def test_stream_hasher_equals_section_hashes(golden):
    """tests.cases.SectionStreamHasher (how the GPU tests hash multi-GB decks on the fly) against
    sections.section_hashes of the same bytes, fed in odd-sized chunks"""
    for name in ("blobs_grey_property", "kw_nset", "kw_none", "bool_5x7x9"):
        deck = bytes(golden[name + "/deck"])
        want = sections.section_hashes(deck)
        sink = cases.SectionStreamHasher({k: v for k, v in want.items() if k != "header"})
        for i in range(0, len(deck), 997):
            sink.write(memoryview(deck)[i:i + 997])
        got = sink.result()
        assert set(got) == set(want) - {"header"}
        assert all(got[k] == want[k] for k in got)
        assert len(sink.head) == want["header"]["bytes"]
