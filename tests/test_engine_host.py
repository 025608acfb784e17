"""Host-side logic of the batched engine that needs no GPU: the packed index maps (rows / positions of the
text and audio segments of every utterance) that generate() / _nar() ship to the device in one copy."""
import numpy as np

from valle_b200.engine import _seg_ranges


def _naive(starts, lens):
    rows = [s + i for s, n in zip(starts, lens) for i in range(n)]
    pos = [i for n in lens for i in range(n)]
    return rows, pos


def test_seg_ranges_matches_python_loops():
    rng = np.random.default_rng(0)
    for _ in range(50):
        B = int(rng.integers(1, 9))
        lens = rng.integers(0, 40, size=B).tolist()   # ragged, empty segments included
        starts = rng.integers(0, 1000, size=B).tolist()
        rows, pos = _seg_ranges(starts, lens)
        exp_rows, exp_pos = _naive(starts, lens)
        assert rows.tolist() == exp_rows and pos.tolist() == exp_pos
        assert rows.dtype == np.int64 and pos.dtype == np.int64


def test_seg_ranges_empty_and_packed_layout():
    rows, pos = _seg_ranges([5, 9], [0, 0])
    assert rows.size == 0 and pos.size == 0
    # the packed [text_b | audio_b] layout of generate(): text rows then audio rows of each utterance
    S, T = [3, 2], [4, 1]
    cu = np.concatenate([[0], np.cumsum(np.add(S, T))])
    trow, tpos = _seg_ranges(cu[:-1], S)
    arow, apos = _seg_ranges(cu[:-1] + np.asarray(S), T)
    assert trow.tolist() == [0, 1, 2, 7, 8] and tpos.tolist() == [0, 1, 2, 0, 1]
    assert arow.tolist() == [3, 4, 5, 6, 9] and apos.tolist() == [0, 1, 2, 3, 0]
    assert sorted(trow.tolist() + arow.tolist()) == list(range(int(cu[-1])))
