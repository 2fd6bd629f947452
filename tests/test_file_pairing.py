"""md/tas file pairing (include/utilities/files.h:39-108) of rl_markets_b200.ingest against the reference's own header,
compiled into oracle/_ref/ref_files by oracle/Makefile (skipped where the reference is not present)."""
import os
import subprocess

import pytest

from rl_markets_b200 import ingest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILES = os.path.join(ROOT, "oracle", "_ref", "ref_files")


def _touch(path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").close()


def _ref(*args):
    r = subprocess.run([REF_FILES] + list(args), capture_output=True, text=True)
    rows = [tuple(l.split("\t")) for l in r.stdout.splitlines() if l]
    return r.returncode, rows


def _tree(tmp_path, md_name, tas_name):
    md, tas = str(tmp_path / md_name), str(tmp_path / tas_name)
    for day in ("20100104", "20100105", "20100106", "20100211"):
        _touch("%s/AAL.L/xmd_%s.csv" % (md, day))
    for day in ("20100104", "20100106", "20100211"):  # the 5th has no partner
        _touch("%s/AAL.L/xmtas%s.csv" % (tas, day))
    _touch("%s/AAL.L/xmd_notes.txt" % md)
    for day in ("20100104",):
        _touch("%s/VOD.L/ymd_%s.csv" % (md, day))
        _touch("%s/VOD.L/ymtas%s.csv" % (tas, day))
    return md, tas


def test_pairing_rule_without_the_reference(tmp_path):
    md, tas = _tree(tmp_path, "depth", "trade")
    got = ingest.file_sample(md, tas, ["AAL.L", "VOD.L"])
    assert [os.path.basename(t[1]) for t in got] == ["xmd_20100104.csv", "xmd_20100106.csv", "xmd_20100211.csv", "ymd_20100104.csv"]
    assert all(os.path.basename(t[2]).startswith(("xmtas", "ymtas")) and os.path.exists(t[2]) for t in got)
    with pytest.raises(RuntimeError, match="No such directory"):
        ingest.file_sample(md, tas, ["NONE.L"])
    _touch("%s/BAD.L/depth_20100104.csv" % md)
    os.makedirs("%s/BAD.L" % tas)
    with pytest.raises(RuntimeError, match="Unexpected file name"):
        ingest.file_sample(md, tas, ["BAD.L"])


@pytest.mark.skipif(not os.path.exists(REF_FILES), reason="oracle/_ref/ref_files not built (no /root/reference)")
@pytest.mark.parametrize("md_name,tas_name", [("depth", "trade"), ("md_data", "tas_data"), ("d", "trades_long")])
def test_file_sample_and_window_match_the_reference_header(tmp_path, md_name, tas_name):
    # directory names of equal and of different lengths, and one containing "md_" itself: the offset quirk
    md, tas = _tree(tmp_path, md_name, tas_name)
    rc, rows = _ref("sample", md, tas, "AAL.L", "VOD.L")
    assert rc == 0
    assert [tuple(r) for r in rows] == ingest.file_sample(md, tas, ["AAL.L", "VOD.L"])
    rc, rows = _ref("window", md, tas, "AAL.L", "201001", "20100211")
    try:
        mine = ingest.sample_window(md, tas, "AAL.L", ["201001", "20100211"])
    except RuntimeError as e:
        assert rc == 1 and rows and rows[-1][0] == "ERROR" and str(e) in rows[-1][1]
    else:
        assert rc == 0 and [tuple(r) for r in rows] == mine
    rc, rows = _ref("sample", md, tas, "NONE.L")
    assert rc == 1 and rows[-1][0] == "ERROR"
    with pytest.raises(RuntimeError) as ei:
        ingest.file_sample(md, tas, ["NONE.L"])
    assert str(ei.value) == rows[-1][1]
