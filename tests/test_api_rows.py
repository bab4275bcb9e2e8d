"""Output rows and consensus of the host-side mirror (reference src/sketchy.rs:358-413), no GPU: row format
"{read}\t{name}\t{shared}\t{genotype columns}", consensus row "{read}\t-\t-\t{majority per column}" over the top rows,
the error for an empty vote, and the documented tie rule (first in rank order; the reference's HashMap order is
nondeterministic there)."""
import numpy as np
import pytest

from sketchy_b200.api import PredictConfig, Sketchy, SketchyError, consensus_value, flatten_sketches, records_blob


def _host_only():
    s = Sketchy.__new__(Sketchy)   # the formatting methods touch no device state
    s.names = ["a.fa", "b.fa", "c.fa", "d.fa"]
    return s


def test_rows_and_consensus():
    s = _host_only()
    geno = {"a.fa": ["ST1", "R"], "b.fa": ["ST2", "S"], "c.fa": ["ST1", "S"], "d.fa": ["ST9", "S"]}
    idx, sums = np.array([2, 0, 1], np.uint32), np.array([9, 9, 4], np.uint64)
    assert s.format_rows(7, idx, sums, geno, consensus=False) == ["7\tc.fa\t9\tST1\tS", "7\ta.fa\t9\tST1\tR", "7\tb.fa\t4\tST2\tS"]
    assert s.format_rows(7, idx, sums, geno, consensus=True) == ["7\t-\t-\tST1\tS"]
    assert consensus_value(["x", "y", "y"]) == "y"
    assert consensus_value(["p", "q", "r"]) == "p"      # three-way tie: first in rank order
    with pytest.raises(SketchyError, match="consensus genotype could not be computed"):
        consensus_value([])
    assert PredictConfig() == PredictConfig(top=1, limit=0, stream=False, consensus=False, header=False)  # src/cli.rs defaults


def test_record_and_sketch_flattening():
    blob, off = records_blob([b"ACGT", b"", np.frombuffer(b"GG", np.uint8)])
    assert blob.tobytes() == b"ACGTGG" and off.tolist() == [0, 4, 4, 6]
    blob, off = records_blob([])
    assert off.tolist() == [0] and blob.size == 1       # a valid pointer for the ABI even when there is nothing to add
    flat, off = flatten_sketches([np.array([1, 5], np.uint64), np.zeros(0, np.uint64), np.array([7], np.uint64)])
    assert flat.tolist() == [1, 5, 7] and off.tolist() == [0, 2, 2, 3]
