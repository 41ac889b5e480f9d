"""`diceseq` drop-in CLI on the bundled demo data (demo.sh: separated and joint model) against
the reference's checked-in outputs.  Those were produced unseeded, so agreement is judged at
Monte-Carlo tolerance: isoform ratios within 0.06 absolute (the reference's own 95% intervals
are 0.1-0.25 wide), FPKM within 15%."""
import gzip
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")


def _read_dice(path):
    rows = [l.rstrip("\n").split("\t") for l in open(path)]
    return rows[0], rows[1:]


def _compare(ours, ref):
    h1, r1 = _read_dice(ours)
    h2, r2 = _read_dice(ref)
    assert h1 == h2
    assert [r[0] for r in r1] == [r[0] for r in r2]                      # transcript order = annotation order
    for a, b in zip(r1, r2):
        assert a[1] == b[1] and a[3] == b[3]
        va, vb = np.array(a[4:], float).reshape(-1, 4), np.array(b[4:], float).reshape(-1, 4)
        assert np.abs(va[:, 1] - vb[:, 1]).max() < 0.06, (a[0], va[:, 1], vb[:, 1])
        assert (va[:, 2] <= va[:, 1] + 1e-9).all() and (va[:, 1] <= va[:, 3] + 1e-9).all()
        big = vb[:, 0] > 2e4
        assert (np.abs(va[big, 0] / vb[big, 0] - 1) < 0.15).all(), (a[0], va[:, 0], vb[:, 0])
        assert abs(float(a[2]) / float(b[2]) - 1) < 0.05


def test_joint_model_like_demo(tmp_path):
    from diceseq_b200 import cli
    out = str(tmp_path / "joint")
    sams = "---".join(os.path.join(DATA, "reads_t%d.sorted.bam" % t) for t in (1, 2, 3))
    cli.main(["-a", os.path.join(DATA, "yeast_RNA_splicing.gtf"), "-s", sams, "-o", out, "-p", "1", "--add_premRNA",
              "--time_seq=1.5,2.5,5", "--mcmc", "20", "20000", "1000", "100"])
    _compare(out + ".dice", os.path.join(DATA, "joint.dice"))
    txt = gzip.open(out + ".sample.gz", "rt").read().split("\n")
    assert txt[3].startswith("@YMR116C|YMR116C.m,YMR116C.p|1.44e+00")
    assert len(txt[4].split(";")) == 3 and txt[4].split(";")[0].endswith(",0.00e+00")


def test_separated_model_like_demo(tmp_path):
    from diceseq_b200 import cli
    out = str(tmp_path / "t1")
    cli.main(["-a", os.path.join(DATA, "yeast_RNA_splicing.gtf"), "-s", os.path.join(DATA, "reads_t1.sorted.bam"),
              "-o", out, "-p", "1", "--add_premRNA"])
    _compare(out + ".dice", os.path.join(DATA, "t1.dice"))


def test_single_isoform_genes_get_rows(tmp_path):
    from diceseq_b200 import cli
    out = str(tmp_path / "single")
    cli.main(["-a", os.path.join(DATA, "yeast_RNA_splicing.gtf"), "-s", os.path.join(DATA, "reads_t1.sorted.bam"),
              "-o", out, "-p", "1"])
    _h, rows = _read_dice(out + ".dice")
    assert len(rows) == 5 and all(r[5] == "1.00e+00" for r in rows)
