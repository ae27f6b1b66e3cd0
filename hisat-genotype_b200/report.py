"""The `.report` text typing() writes, restated from the reference's print statements.

Reference: hisatgenotype_modules/hisatgenotype_typing_core.py
  header            :309-325      aligner line   :336-343
  reads/pairs line  :1592-1595    count lines    :1650-1677    abundance lines :2076-2121
The format is consumed by common.call_nuance_results (common:1984-2030) and
hisatgenotype_tools/hisatgenotype_parse_results.py, so it is reproduced byte for byte
(tests/test_gpu_report.py compares with reports written by the unmodified reference).
"""
from __future__ import annotations


def header(hisat2_version, hg_version, dbversion, cmd_call):
    """core:315-325 — every value is printed with print(), i.e. followed by one newline."""
    return ("# VERSIONS:\n# HISAT2 - %s\n# HISAT-genotype - %s\n# Database - %s\n# COMMAND:\n%s\n"
            % (hisat2_version, hg_version, dbversion, cmd_call))


def aligner_line(aligner, index_type):
    return "\n\t\t%s %s\n" % (aligner, index_type)  # core:338-343


def locus_block(num_reads, num_pairs, gene_counts, gene_prob, simulation, test_gene_names, output_allele_counts=False,
                best_alleles=False):
    """Report lines of one locus.  gene_counts: [[allele, count]] in Gene_counts dict order; gene_prob: ranked
    [[allele, prob]].  Returns (text, success flags per truth rank) — success mirrors core:2086-2103."""
    out = []
    if num_reads <= 0:  # core:1588-1589: nothing is printed for the locus
        return "", []
    out.append("\t\t\t%d reads and %d pairs are aligned\n" % (num_reads, num_pairs))
    counts = sorted(gene_counts, key=lambda x: x[1], reverse=True)  # stable: ties keep dict order (core:1651)
    for count_i, (allele, count) in enumerate(counts):
        if simulation:
            found = False
            for name in test_gene_names:
                if allele == name:
                    out.append("\t\t\t*** %d ranked %s (count: %d)\n" % (count_i + 1, name, count))
                    found = True
            if count_i < 5 and not found:
                out.append("\t\t\t\t%d %s (count: %d)\n" % (count_i + 1, allele, count))
        else:
            out.append("\t\t\t\t%d %s (count: %d)\n" % (count_i + 1, allele, count))
            if count_i >= 9 and not output_allele_counts:
                break
    out.append("\n\n")  # print("\n")
    success = [False] * len(test_gene_names) if simulation else []
    found_list = [False] * len(test_gene_names) if simulation else []
    for prob_i, (allele, prob) in enumerate(gene_prob):
        if prob < 0.01:
            break
        found = False
        if simulation:
            for name_i, name in enumerate(test_gene_names):
                if allele == name:
                    rank_i = prob_i  # the reference's tie walk-back compares a list with a float and never fires
                    out.append("\t\t\t*** %d ranked %s (abundance: %.2f%%)\n" % (rank_i + 1, name, prob * 100.0))
                    if rank_i < len(success):
                        success[rank_i] = True
                    found_list[name_i] = True
                    found = True
            if False not in found_list and prob_i >= 10:
                break
        if not found:
            out.append("\t\t\t\t%d ranked %s (abundance: %.2f%%)\n" % (prob_i + 1, allele, prob * 100.0))
            if best_alleles and prob_i < 2:
                out.append("SingleModel %s (abundance: %.2f%%)\n" % (allele, prob * 100.0))
        if not simulation and prob_i >= 9:
            break
        if prob_i >= 19:
            break
    return "".join(out), success
