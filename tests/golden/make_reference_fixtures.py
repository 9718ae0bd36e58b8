"""Copies the reference-held DATA files the parity tests pin on into tests/golden/ (run in the authoring container, where
/root/reference exists; the GPU box only sees the copies).  Data only — no reference source code is copied.

  data/SARS-Cov2/VIC01/wuhan_coronavirus_australia.fasta.gz   MT007544.1, 29 893 bp, the reference of scripts/run_example.sh
  data/SARS-Cov2/VIC01/Coordinates.csv                        its ORF table (J/cluster/Annotation.java:138-179 reads it)
  data/SARS-Cov2/VIC01/0.annot.tsv                            an 0.annot output of a real run of the reference: the only
                                                              reference-held output sample of this path (Gene/Start/End columns,
                                                              incl. the 3'UTR end moved to seqlen - polyA by adjust3UTR)
  data/229E_CoV/WT_229_reference.fasta.gz, Coordinates.csv    config 3 (scripts/run_slurm_229E.sh)
  data/SARS-Cov2/NC/*, data/U13369/*                          two more reference + ORF-table pairs the reference ships: overlapping
                                                              ORFs, a longer reference without a leader row
"""
import os
import shutil

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [
    ("SARS-Cov2/VIC01/wuhan_coronavirus_australia.fasta.gz", "vic01/wuhan_coronavirus_australia.fasta.gz"),
    ("SARS-Cov2/VIC01/Coordinates.csv", "vic01/Coordinates.csv"),
    ("SARS-Cov2/VIC01/0.annot.tsv", "vic01/0.annot.tsv"),
    ("229E_CoV/WT_229_reference.fasta.gz", "cov229e/WT_229_reference.fasta.gz"),
    ("229E_CoV/Coordinates.csv", "cov229e/Coordinates.csv"),
    # NC_045512.2 (29 903 bp): its table has ORF7b overlapping ORF7a (27756 < 27759) and a 3'UTR row ending at 29903
    ("SARS-Cov2/NC/NC_045512.fasta.gz", "nc045512/NC_045512.fasta.gz"),
    ("SARS-Cov2/NC/Coordinates.csv", "nc045512/Coordinates.csv"),
    # U13369 human rDNA repeat (42 999 bp): 8 rows, the first one 3 656 bp long, no leader row
    ("U13369/Human_ribosomal_DNA_complete_repeating_unit.fasta.gz", "u13369/Human_ribosomal_DNA_complete_repeating_unit.fasta.gz"),
    ("U13369/Coordinates.csv", "u13369/Coordinates.csv"),
]
if __name__ == "__main__":
    for src, dst in FILES:
        os.makedirs(os.path.dirname(os.path.join(HERE, dst)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        print(dst, os.path.getsize(os.path.join(HERE, dst)), "bytes")
