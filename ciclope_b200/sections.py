"""Split a CalculiX voxel-FE deck (the text ``mesh2voxelfe`` writes, voxelFE.py:612-768 of the reference) into
its sections and hash them.  Host-side utility: comparing a z-slab deck written by N ranks with the
single-GPU deck (or with a CPU reference deck) section by section tells *where* two decks differ, and the
header's time-stamp line is left out of the hashes."""
import collections
import hashlib
import re

_NODE = b"*NODE\n"
_ELEM = b"** Elements and Element sets from input model\n"
_NSET = b"** Additional nset from voxel model."
_ELSET = b"** Additional elset from voxel model."
_IDS = re.compile(rb"[0-9,\n]*")


def _after_id_lines(data, pos):
    """end of the run of id lines (digits, commas, newlines) that starts at `pos`"""
    return _IDS.match(data, pos).end()


def split_deck(data):
    """-> OrderedDict section name -> (lo, hi) byte range.  Sections: ``header`` (comment lines up to and
    including ``*NODE``), ``nodes``, ``elements`` (comment line, ``*ELEMENT`` headers and element lines),
    ``nset``, ``elset`` (absent when the keyword was not given) and ``tail`` (PROPERTY cards + template)."""
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    out = collections.OrderedDict()
    i_node = data.find(_NODE)
    if i_node < 0:
        raise ValueError("not a voxel-FE deck: no *NODE line")
    i_node += len(_NODE)
    i_elem = data.find(_ELEM, i_node)
    if i_elem < 0:
        raise ValueError("not a voxel-FE deck: no element section")
    out["header"] = (0, i_node)
    out["nodes"] = (i_node, i_elem)
    i_nset = data.find(_NSET, i_elem)
    i_elset = data.find(_ELSET, i_elem)
    # the last block before the tail: ids after its last '*...' header line
    if i_elset >= 0:
        last_hdr = data.rfind(b"*ELSET, ELSET=", i_elset)
    elif i_nset >= 0:
        last_hdr = data.rfind(b"*NSET, NSET=", i_nset)
    else:
        last_hdr = data.rfind(b"*ELEMENT, TYPE=", i_elem)
    if last_hdr < 0:
        i_tail = i_elem + len(_ELEM)
    else:
        i_tail = _after_id_lines(data, data.index(b"\n", last_hdr) + 1)
    ends = [p for p in (i_nset, i_elset) if p >= 0] + [i_tail]
    out["elements"] = (i_elem, min(ends))
    if i_nset >= 0:
        out["nset"] = (i_nset, i_elset if i_elset >= 0 else i_tail)
    if i_elset >= 0:
        out["elset"] = (i_elset, i_tail)
    out["tail"] = (i_tail, len(data))
    return out


def mask_timestamp(data):
    """The deck with header line 2 (``... written by Voxel_FEM script on <datetime.now()>``) blanked."""
    data = bytes(data)
    a = data.find(b"\n") + 1
    b = data.find(b"\n", a)
    return data[:a] + b"** <timestamp>" + data[b:]


def section_hashes(data):
    """-> OrderedDict section name -> {"bytes": n, "sha256": hex}; the header is hashed without its
    time-stamp line."""
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    out = collections.OrderedDict()
    for name, (lo, hi) in split_deck(data).items():
        piece = data[lo:hi]
        if name == "header":
            piece = mask_timestamp(piece)
        out[name] = {"bytes": hi - lo, "sha256": hashlib.sha256(piece).hexdigest()}
    return out


def compare_decks(a, b):
    """-> (all sections equal, {section: equal}) for two decks (bytes)."""
    ha, hb = section_hashes(a), section_hashes(b)
    # (the header's byte count depends on the time stamp's length: its masked hash is what counts)
    per = {k: (k in hb and ha[k]["sha256"] == hb[k]["sha256"] and (k == "header" or ha[k] == hb[k])) for k in ha}
    for k in hb:
        per.setdefault(k, False)
    return all(per.values()), per
