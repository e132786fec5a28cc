"""Shared test helpers."""


def assert_occurrence_text_equal(got_lines, golden_text):
    """Occurrence CSVs are deterministic except when a read has > 20 hits at its minimum distance: the reference
    then keeps a random 20 (np.random.choice, motif_discovery.py:1467-1469).  Such cells are compared
    structurally (20 sorted distinct positions); everything else must match exactly."""
    gold = golden_text.splitlines()
    assert len(got_lines) == len(gold)
    for a, b in zip(got_lines, gold):
        if a == b:
            continue
        ca, cb = a.split(";"), b.split(";")
        assert len(ca) == len(cb) and ca[0] == cb[0] and ca[-1] == cb[-1], (a, b)
        for x, y in zip(ca[1:-1], cb[1:-1]):
            if x == y:
                continue
            px, py = [int(t) for t in x.split(",")], [int(t) for t in y.split(",")]
            assert len(px) == len(py) == 20, (a, b)
            assert px == sorted(set(px)), (a, b)
