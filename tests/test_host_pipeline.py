"""Host side of the `diceseq` drop-in (CPU only): annotation, BAM reader, read model, formatting.
The read model is checked bit for bit against matrices produced by the UNMODIFIED reference
read model (tests/golden/make_golden_host.py) on the bundled yeast data (tests/data)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")

from diceseq_b200.annotation import loadgene, parse_attributes  # noqa: E402
from diceseq_b200.bam import load_samfile  # noqa: E402
from diceseq_b200.read_model import gene_time_matrices  # noqa: E402
from diceseq_b200.run_utils import format_dice_lines, format_sample_lines, gene_inputs, single_isoform_result  # noqa: E402


def test_annotation():
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    assert [g.geneID for g in genes] == ["YMR116C", "YNL112W", "YOL120C", "YKL006W", "YDR381W"]
    g = genes[0]
    assert g.tranNum == 1 and g.trans[0].tranID == "YMR116C.m" and g.strand == "-"
    assert g.trans[0].exons.tolist() == [[499456, 499878], [500152, 500688]] and g.trans[0].tranL == 960
    g.add_premRNA()
    assert g.trans[1].tranID == "YMR116C.p" and g.trans[1].tranL == 1233
    assert g.get_gene_info()[-1] == "YMR116C.m,YMR116C.p"
    a = parse_attributes('ID=gene1;Name=abc;biotype=coding', "gene")
    assert a == {"ID": "gene1", "Name": "abc", "Type": "coding"}


def test_bam_reader():
    b = load_samfile(os.path.join(DATA, "reads_t3.sorted.bam"))
    assert len(b.references) == 17 and b.references[:3] == ["I", "II", "III"]
    assert b.mapped_total() == 952.0
    rs = b.fetch("XIII", 499456, 500688)
    assert len(rs) == 200 and all(r.rlen == 75 for r in rs)
    r = rs[0]
    assert (r.qname, r.pos, r.aend, r.mapq, r.qlen) == ("simulated_reads_64_2", 499463, 499538, 255, 75)
    assert r.positions[0] == 499463 and len(r.positions) == 75 and r.is_read2 and not r.is_reverse
    spliced = [x for x in rs if len(x.cigar) == 3]
    assert spliced and all(x.aend - x.pos > 75 for x in spliced)
    with pytest.raises(ValueError):
        b.fetch("nope", 0, 10)


def test_indexed_bam_fetch_equals_full_decode(tmp_path):
    """IndexedBam (BAI bins + linear index + BGZF virtual offsets) returns exactly the records of the in-memory
    decode for every gene region and for open-ended regions, and the idxstats totals from the index's
    pseudo-bins; load_samfile picks it whenever a .bai is present; unsorted files are refused."""
    import gzip
    import shutil
    import struct
    from diceseq_b200 import bam
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    for t in (1, 4):
        path = os.path.join(DATA, "reads_t%d.sorted.bam" % t)
        full, idx = bam.BamFile(path), bam.IndexedBam(path, path + ".bai")
        assert isinstance(load_samfile(path), bam.IndexedBam)
        assert full.idxstats() == idx.idxstats() and full.mapped_total() == idx.mapped_total()
        for g in genes:
            for a, b in ((g.start, g.stop), (g.start - 500, g.start + 10), (g.stop - 3, g.stop + 2000), (0, 10 ** 7)):
                ra, rb = full.fetch(g.chrom, a, b), idx.fetch(g.chrom, a, b)
                assert [(r.qname, r.pos, r.flag, r.cigar) for r in ra] == [(r.qname, r.pos, r.flag, r.cigar) for r in rb]
    plain = str(tmp_path / "noindex.bam")
    shutil.copy(os.path.join(DATA, "reads_t1.sorted.bam"), plain)
    assert isinstance(load_samfile(plain), bam.BamFile)
    # swap the first two placed records of the uncompressed stream -> not coordinate sorted any more
    data = bytearray(gzip.open(plain, "rb").read())
    off = 12 + struct.unpack_from("<i", data, 4)[0]
    for _ in range(struct.unpack_from("<i", data, off - 4)[0]):
        off += 8 + struct.unpack_from("<i", data, off)[0]
    n1 = struct.unpack_from("<i", data, off)[0]
    n2 = struct.unpack_from("<i", data, off + 4 + n1)[0]
    last = off
    while last + 4 + struct.unpack_from("<i", data, last)[0] < len(data):
        last += 4 + struct.unpack_from("<i", data, last)[0]
    tail = bytes(data[last:])
    shuffled = bytes(data[:off]) + tail + bytes(data[off:last])
    bad = str(tmp_path / "unsorted.bam")
    with gzip.open(bad, "wb") as f:
        f.write(shuffled)
    with pytest.raises(ValueError, match="not sorted"):
        bam.BamFile(bad)


def test_read_model_matches_reference_bit_for_bit():
    z = np.load(os.path.join(ROOT, "tests", "golden", "host_matrices.npz"))
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    for g in genes:
        g.add_premRNA()
        for ti in (1, 2, 3, 4):
            R, P, L, cnt = gene_time_matrices(g, [load_samfile(os.path.join(DATA, "reads_t%d.sorted.bam" % ti))])
            k = "%s_t%d" % (g.geneID, ti)
            assert np.array_equal(R, z[k + "_R"]), k
            assert np.array_equal(P, z[k + "_proU"], equal_nan=True), k
            assert np.array_equal(L, z[k + "_efflen"]) and cnt == int(z[k + "_count"]), k


def test_bias_mode_read_model_matches_reference_bit_for_bit(tmp_path):
    """end5 / end3 / both: proB and the bias-weighted effective lengths against the unmodified reference's
    TranSplice (set_sequence + set_bias("seq") + get_ready, tests/golden/make_golden_host_bias.py) on the
    bundled reads, the bundled bias file and the seeded synthetic genome."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import synth_fasta
    from diceseq_b200.bias import BiasFile, FastaFile
    z = np.load(os.path.join(ROOT, "tests", "golden", "host_matrices_bias.npz"))
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    fa = str(tmp_path / "genome.fa")
    synth_fasta.write_fasta(synth_fasta.genome_for(genes), fa)
    fasta, bias = FastaFile(fa), BiasFile(os.path.join(DATA, "yeast_4tU.bias"))
    for g in genes:
        g.add_premRNA()
        for ti in (1, 3):
            bam = load_samfile(os.path.join(DATA, "reads_t%d.sorted.bam" % ti))
            for mode in ("end5", "end3", "both"):
                R, P, L, cnt, Lb = gene_time_matrices(g, [bam], None, None, mode, fasta, bias, with_efflen_bias=True)
                k = "%s_t%d_%s" % (g.geneID, ti, mode)
                assert np.array_equal(R, z[k + "_R"]), k
                assert np.array_equal(P, z[k + "_proB"], equal_nan=True), k
                assert np.array_equal(Lb, z[k + "_efflen_bias"]), k
                assert np.array_equal(L, z[k + "_efflen_unif"]), k
    g = [x for x in genes if x.geneID == str(z["flen_case_gene"])][0]
    mu, sd = z["flen_mean_std"]
    assert (bias.flen_mean, bias.flen_std) == (mu, sd)
    R, P, L, cnt, Lb = gene_time_matrices(g, [load_samfile(os.path.join(DATA, "reads_t2.sorted.bam"))], mu, sd, "both",
                                          fasta, bias, with_efflen_bias=True)
    assert np.array_equal(P, z["flen_case_proB"], equal_nan=True) and np.array_equal(Lb, z["flen_case_efflen_bias"])


def test_gene_inputs_in_bias_mode_follows_get_psi(tmp_path):
    """run_utils.py:66-87: one bias file for all time points (or one per time point); the fragment-length
    model takes the file's mean / std; the effective length stays the uniform one."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import synth_fasta
    from diceseq_b200.bias import BiasFile, FastaFile
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    fa = str(tmp_path / "genome.fa")
    synth_fasta.write_fasta(synth_fasta.genome_for(genes), fa)
    bf = os.path.join(DATA, "yeast_4tU.bias")
    g = genes[2]
    g.add_premRNA()
    f1, f3 = (os.path.join(DATA, "reads_t%d.sorted.bam" % t) for t in (1, 3))
    R_all, L_all, P_all, count = gene_inputs(g, [[f1], [f3]], None, None, "end5", fa, [bf])
    bias = BiasFile(bf)
    for t, f in enumerate((f1, f3)):
        R, P, L, n = gene_time_matrices(g, [load_samfile(f)], bias.flen_mean, bias.flen_std, "end5", FastaFile(fa), bias)
        assert np.array_equal(R_all[t], R) and np.array_equal(P_all[t], P, equal_nan=True)
        assert np.array_equal(L_all[t], [tr.tranL for tr in g.trans]) and count[t] == n
    two = gene_inputs(g, [[f1], [f3]], None, None, "end5", fa, [bf, bf])
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(two[2], P_all))
    with pytest.raises(ValueError):
        gene_inputs(g, [[f1], [f3]], None, None, "end5", fa, [bf, bf, bf])


def test_replicates_concatenate():
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    g = genes[1]
    g.add_premRNA()
    f1, f2 = (os.path.join(DATA, "reads_t%d.sorted.bam" % t) for t in (1, 2))
    R_all, L_all, P_all, count = gene_inputs(g, [[f1, f2], [f1]])
    Ra = gene_time_matrices(g, [load_samfile(f1)])[0]
    Rb = gene_time_matrices(g, [load_samfile(f2)])[0]
    assert R_all[0].shape[0] == Ra.shape[0] + Rb.shape[0] and R_all[1].shape == Ra.shape
    assert count.tolist() == [Ra.shape[0] + Rb.shape[0], Ra.shape[0]]


def test_dice_and_sample_formatting():
    genes = loadgene(os.path.join(DATA, "yeast_RNA_splicing.gtf"))
    g = genes[0]
    g.add_premRNA()
    X = np.array([1.5, 2.5])
    psi = np.array([[0.4, 0.7], [0.6, 0.3]])
    ci = [np.array([[0.5, 0.3], [0.8, 0.6]]), np.array([[0.7, 0.5], [0.4, 0.2]])]
    txt = format_dice_lines(g, X, np.array([100, 200]), [1000.0, 2000.0], psi, ci, -3812.4)
    rows = txt.strip().split("\n")
    assert len(rows) == 2
    c = rows[0].split("\t")
    assert c[:4] == ["YMR116C.m", "YMR116C", "-3.8e+03", "960"] and len(c) == 4 + 4 * 2
    fsi = 960 * 0.4 / (960 * 0.4 + 1233 * 0.6)
    assert c[4] == "%.2e" % (100 * fsi * 1e9 / 960 / 1000.0) and c[5:8] == ["4.00e-01", "3.00e-01", "5.00e-01"]
    s = format_sample_lines(g, 1.44, np.array([[[-1.87, 0.535], [0.0, 0.0]]]), 2)
    assert s == "@YMR116C|YMR116C.m,YMR116C.p|1.44e+00\n-1.87e+00,0.00e+00;5.35e-01,0.00e+00\n"
    pm, p95, lm = single_isoform_result(3)
    assert pm.shape == (1, 3) and p95[0].shape == (3, 2) and lm == 0.0


def test_cli_refuses_unsupported_modes():
    from diceseq_b200 import cli
    gtf = os.path.join(DATA, "yeast_RNA_splicing.gtf")
    bam = os.path.join(DATA, "reads_t1.sorted.bam")
    for extra in (["--bias", "end5", "ref.fa", "x.bias"],):
        with pytest.raises(SystemExit):
            cli.main(["-a", gtf, "-s", bam, "-o", "/tmp/never"] + extra)
