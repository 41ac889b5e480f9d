"""GTF / GFF3 annotation -> genes with transcripts (host side of the `diceseq` drop-in).

Mirrors what the sampler's caller needs from `diceseq/utils/gtf_utils.py:11-93,170-249`:
genes in file order, each with its transcripts in file order, exons sorted by start,
transcript length = sum of exon lengths, `add_premRNA` (one extra single-exon transcript
spanning the gene, id `<gene>.p`), and `get_gene_info`.  Feature lines must come as
gene -> transcript -> exons like in Ensembl files (same assumption as the reference).
"""
from __future__ import annotations

import numpy as np

_ID_TAGS = {"gene": ("ID", "gene_id"), "tran": ("ID", "transcript_id", "mRNA_id")}
_NAME_TAGS = {"gene": ("Name", "gene_name"), "tran": ("Name", "transcript_name", "mRNA_name")}
_TYPE_TAGS = ("Type", "gene_type", "gene_biotype", "biotype")


def parse_attributes(text, kind):
    """ID / Name / Type of a GTF (`key "value";`) or GFF3 (`key=value;`) attribute column;
    the last matching tag wins, missing ones are "*"."""
    out = {"ID": "*", "Name": "*", "Type": "*"}
    for att in text.rstrip().split(";"):
        att = att.lstrip(" ")
        if not att:
            continue
        parts = att.split("=") if "=" in att else att.split(" ")
        if len(parts) < 2:
            print("Can't pase this attribute: %s" % att)
            continue
        key, val = parts[0], parts[1]
        if val.startswith('"'):
            val = val.split('"')[1]
        if key in _ID_TAGS[kind]:
            out["ID"] = val
        elif key in _NAME_TAGS[kind]:
            out["Name"] = val
        elif key in _TYPE_TAGS:
            out["Type"] = val
    return out


class Transcript:
    def __init__(self, chrom, strand, start, stop, tran_id, tran_name="*", biotype="*"):
        self.chrom, self.strand = chrom, strand
        self.start, self.stop = int(start), int(stop)
        self.tranID, self.tranName, self.biotype = tran_id, tran_name, biotype
        self.exons = np.zeros((0, 2), dtype=int)
        self.tranL = 0
        self.exonNum = 0

    def add_exon(self, chrom, strand, start, stop):
        if strand != self.strand or chrom != self.chrom:
            print("The exon has different chrom or strand to the transcript.")
            return
        self.exons = np.sort(np.vstack([self.exons, [[int(start), int(stop)]]]), axis=0)
        self.tranL += abs(int(stop) - int(start) + 1)
        self.exonNum += 1


class Gene:
    def __init__(self, chrom, strand, start, stop, gene_id, gene_name="*", biotype="*"):
        self.chrom, self.strand = chrom, strand
        self.start, self.stop = int(start), int(stop)
        self.geneID, self.geneName, self.biotype = gene_id, gene_name, biotype
        self.trans = []

    @property
    def tranNum(self):
        return len(self.trans)

    def add_premRNA(self):
        t = Transcript(self.chrom, self.strand, self.start, self.stop, self.geneID + ".p", self.geneName, self.biotype)
        t.add_exon(self.chrom, self.strand, self.start, self.stop)
        self.trans.append(t)

    def get_gene_info(self):
        return [self.geneID, self.geneName, self.chrom, self.strand, self.start, self.stop, self.biotype,
                ",".join(t.tranID for t in self.trans)]


def loadgene(anno_file, gene_tags=("gene",), tran_tags=("transcript", "mRNA"), exon_tags=("exon",)):
    """Genes of a GTF / GFF3 file, in file order."""
    genes, gene = [], None
    with open(anno_file) as fh:
        for line in fh:
            if line[:1] in ("#", ">"):
                continue
            col = line.split("\t")
            if len(col) < 9:
                continue
            if col[2] in gene_tags:
                if gene is not None:
                    genes.append(gene)
                a = parse_attributes(col[8], "gene")
                gene = Gene(col[0], col[6], col[3], col[4], a["ID"], a["Name"], a["Type"])
            elif col[2] in tran_tags:
                a = parse_attributes(col[8], "tran")
                if gene is None:
                    print("Gene is not ready before transcript.")
                    continue
                gene.trans.append(Transcript(col[0], col[6], col[3], col[4], a["ID"], a["Name"], a["Type"]))
            elif col[2] in exon_tags:
                if gene is None or not gene.trans:
                    print("Gene or transcript is not ready before exon.")
                    continue
                tr = gene.trans[-1]
                if col[0] != tr.chrom:
                    print("Exon from a different chrom of transcript.")
                    continue
                if col[6] != tr.strand:
                    print("Exon from a different strand of transcript.")
                    continue
                tr.add_exon(col[0], col[6], col[3], col[4])
    if gene is not None:
        genes.append(gene)
    return genes
