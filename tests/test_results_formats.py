"""CPU: the track output writers against text produced by the reference's own writers
(tests/golden/formats.npz, minted by oracle/make_golden.py::gen_formats): character-identical."""
import numpy as np

from conftest import load_golden
from moyolo_b200 import results as R


def test_mot_challenge_rows_identical_to_reference_writer():
    meta, g = load_golden("formats")
    got = "".join(R.mot_challenge_lines(g["table"], meta["img_w"], meta["img_h"]))
    assert got == meta["mot"]
    assert ",-1," not in "".join(l.split(",", 2)[1] for l in got.splitlines())  # untracked rows (id < 0) are skipped
    # per-sequence filter and file writer
    t = g["table"].copy()
    t[::2, 0] = 3
    only = R.mot_challenge_lines(t, meta["img_w"], meta["img_h"], seq=3)
    assert len(only) == int(((t[:, 0] == 3) & (t[:, 2] >= 0)).sum())
    assert R.iter_sequences(t) == [0, 3]


def test_save_txt_rows_identical_to_reference_results(tmp_path):
    meta, g = load_golden("formats")
    for conf, key in ((False, "save_txt"), (True, "save_txt_conf")):
        got = R.save_txt_lines(g["table"], meta["img_w"], meta["img_h"], save_conf=conf)
        assert sorted(got) == sorted(int(f) for f in meta[key])
        for f, text in meta[key].items():
            assert "".join(got[int(f)]) == text, (conf, f)
    n = R.write_save_txt(tmp_path / "labels", g["table"], meta["img_w"], meta["img_h"], stem="seq0")
    assert n == 5 and (tmp_path / "labels" / "seq0_0.txt").read_text() == meta["save_txt"]["0"]
    k = R.write_mot_challenge(tmp_path / "seq0.txt", g["table"], meta["img_w"], meta["img_h"])
    assert k == int((g["table"][:, 2] >= 0).sum())


def test_bad_table_shape():
    import pytest
    with pytest.raises(ValueError):
        R.mot_challenge_lines(np.zeros((3, 8)), 10, 10)
