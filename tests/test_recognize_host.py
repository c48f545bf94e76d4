"""Host-side control flow of shazam_b200.recognize that needs no GPU: the batching of return_matches and the a-priori
early exit (recognizer_apriori.py:246-310), against golden vectors produced by executing the reference's
own functions (tests/golden/make_golden.py -> apriori_cases.json).  The database is an in-memory stand-in answering the
cursor statement; the vote inside the exit test is the oracle's (on a GPU box the product's own vote runs:
tests/test_index_gpu.py::test_apriori_early_exit_golden)."""
import hashlib
import json
import os

import pytest

from oracle import sia_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Cursor:
    def __init__(self, table, log):
        self.table, self.rows, self.log = table, [], log

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def execute(self, query, params=None):
        q = " ".join(query.split())
        assert q.startswith("SELECT HEX(`hash`), `song_id`, `offset` FROM `fingerprints` WHERE `hash` IN (")
        assert q.count("UNHEX(%s)") == len(params)
        self.log.append(len(params))
        self.rows = list(self.table.select_multiple(list(params)))

    def __iter__(self):
        return iter(self.rows)


class _Db:
    from shazam_b200.database import GPUDatabase as _G
    SELECT_MULTIPLE, IN_MATCH = _G.SELECT_MULTIPLE, _G.IN_MATCH

    def __init__(self, table):
        self.table, self.batches = table, []

    def cursor(self, **kw):
        return _Cursor(self.table, self.batches)

    def get_song_by_id(self, sid):
        return self.table.get_song_by_id(sid)


def _table(case):
    table = O.FingerprintTable()
    for sid, s in sorted(case["songs"].items(), key=lambda kv: int(kv[0])):
        assert table.insert_song(s["song_name"], s["file_sha1"], s["total_hashes"]) == int(sid)
    by_song = {}
    for sid, h, o in case["rows"]:
        by_song.setdefault(sid, []).append((h, o))
    for sid, hs in by_song.items():
        table.insert_hashes(sid, hs)
    return table


def _strip(results):
    return [{k: (v.decode() if isinstance(v, bytes) else v) for k, v in r.items()} for r in results]


@pytest.fixture()
def cases():
    return json.load(open(os.path.join(GOLDEN, "apriori_cases.json")))


def test_apriori_early_exit_control_flow(cases, monkeypatch):
    from shazam_b200 import recognize
    exits = 0
    for case in cases:
        table = _table(case)
        db = _Db(table)
        monkeypatch.setattr(recognize, "db", db)
        monkeypatch.setattr(recognize, "align_matches",
                            lambda m, d, n, topn=recognize.TOPN, _t=table: O.align_matches(_t, m, d, n, topn))
        q = [tuple(x) for x in case["query"]]
        matches, dedup, songs_arr = recognize.return_matches(q, case["batch_size"], apriori=True)
        assert len(matches) == case["n_matches"], case["case"]
        assert hashlib.sha256(repr(sorted(matches)).encode()).hexdigest() == case["matches_sorted_sha"]
        assert {str(k): v for k, v in dedup.items()} == case["dedup"]
        assert _strip(songs_arr) == case["songs_arr"]
        # batches of `batch_size` DISTINCT hashes in first-seen order, and none is read after the exit
        n_distinct = len({h.upper() for h, _ in q})
        full = [case["batch_size"]] * (n_distinct // case["batch_size"]) + ([n_distinct % case["batch_size"]] if n_distinct % case["batch_size"] else [])
        assert db.batches == full[:len(db.batches)]
        if songs_arr:
            exits += 1
            assert len(db.batches) < len(full) or len(full) == 1
            assert songs_arr[0]["hashes_matched_in_input"] / 2 > songs_arr[1]["hashes_matched_in_input"]
        else:
            assert db.batches == full
        # without the flag every batch is read and the reference's two-element result comes back
        db.batches.clear()
        m2, d2 = recognize.return_matches(q, case["batch_size"])
        assert db.batches == full and len(m2) >= len(matches)
    assert exits >= 2
