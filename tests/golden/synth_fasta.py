"""A synthetic genome for the bias-mode fixtures: the bundled data has no reference FASTA, so every
chromosome gets a seeded random A/C/G/T sequence (with a few N and lower-case stretches, which the k-mer
tables do not list) long enough for the annotated genes.  Used by make_golden_host_bias.py (behind the pysam
stand-in of the unmodified reference) and by tests/test_host_pipeline.py (written to a FASTA file)."""
import zlib

import numpy as np


def chrom_sequence(name, length):
    rng = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    seq = np.array(list("ACGT"))[rng.randint(0, 4, size=length)]
    for _ in range(max(1, length // 5000)):                       # sprinkle characters outside the tables
        p = rng.randint(0, length)
        seq[p:p + 3] = "N"
        p = rng.randint(0, length)
        seq[p:p + 5] = np.char.lower(seq[p:p + 5])
    return "".join(seq)


def genome_for(genes, pad=100):
    need = {}
    for g in genes:
        need[g.chrom] = max(need.get(g.chrom, 0), int(g.stop) + pad)
    return {c: chrom_sequence(c, n) for c, n in need.items()}


def write_fasta(genome, path, width=60):
    with open(path, "w") as f:
        for name, seq in genome.items():
            f.write(">%s synthetic\n" % name)
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + "\n")
