"""Streaming detokenisation (reference moondream.py:476-537, row a21 of SURVEY.md §8): the chunks must be exactly the
ones the reference's generator yields.  tests/golden/streaming.json was produced by the unmodified reference (its own
sampled tokens, a tokenizer whose pieces hit the newline / CJK / last-space rules); no GPU is involved."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


class _Pieces:
    def __init__(self, pieces):
        self.pieces = pieces

    def decode(self, ids):
        return "".join(self.pieces[int(i) % len(self.pieces)] for i in ids)


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(HERE, "golden", "streaming.json")))


def _stream(tokenizer, tokens):
    from moondream_b200.moondream import MoondreamModel

    holder = type("Holder", (), {"tokenizer": tokenizer})()
    return list(MoondreamModel._stream_text(holder, tokens))


def test_chunks_equal_the_reference_generator(gold):
    tok = _Pieces(gold["pieces"])
    assert len(gold["cases"]) >= 4
    for case in gold["cases"]:
        got = _stream(tok, case["tokens"])
        assert got == case["chunks"], (case["seed"], got[:8], case["chunks"][:8])
        assert "".join(got) == tok.decode(case["tokens"])


def test_flush_rules():
    tok = _Pieces([" ab", "cd", "\n", "漢", " ", "."])
    # nothing is printed before a space proves the word is complete; the tail is flushed at the end
    assert _stream(tok, [0, 1]) == [" ", "abcd"]
    # a newline flushes everything and restarts the cache
    assert _stream(tok, [0, 2, 1]) == [" ", "ab\n", "cd"]
    # a CJK character is printable immediately
    assert _stream(tok, [1, 3, 3]) == ["cd漢", "漢"]
    assert _stream(tok, []) == []
