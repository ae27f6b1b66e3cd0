"""TEST INFRASTRUCTURE (oracle): Python-3 restatement of the diploid pair model of the reference's legacy typer,
etc/hisatgenotype_hla_cyp.py:236-302 (joint_abundance; helpers normalize :147-150, prob_diff :155-162, HLA_prob_cmp
:167-176).  The legacy script is Python 2 (`sorted(cmp=...)`) and cannot run here; the statements below follow it line by
line, with functools.cmp_to_key in place of cmp=.  PARITY UNPINNED: no live test or golden of the reference covers this
function; the product (hgt_pair_em) is checked against THIS restatement only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
from functools import cmp_to_key


def normalize(prob):  # :147-150
    total = sum(prob.values())
    for allele, mass in prob.items():
        prob[allele] = mass / total


def prob_diff(prob1, prob2):  # :155-162
    diff = 0.0
    for allele in prob1.keys():
        if allele in prob2:
            diff += abs(prob1[allele] - prob2[allele])
        else:
            diff += prob1[allele]
    return diff


def HLA_prob_cmp(a, b):  # :167-176
    if a[1] != b[1]:
        if a[1] < b[1]:
            return 1
        else:
            return -1
    assert a[0] != b[0]
    if a[0] < b[0]:
        return -1
    else:
        return 1


def joint_abundance(HLA_cmpt, HLA_length=None, return_iters=False):  # :236-302
    allele_names = set()
    for cmpt in HLA_cmpt.keys():
        allele_names |= set(cmpt.split('-'))

    HLA_prob, HLA_prob_next = {}, {}
    for cmpt, count in HLA_cmpt.items():
        alleles = cmpt.split('-')
        for allele1 in alleles:
            for allele2 in allele_names:
                if allele1 < allele2:
                    allele_pair = "%s-%s" % (allele1, allele2)
                else:
                    allele_pair = "%s-%s" % (allele2, allele1)
                if allele_pair not in HLA_prob:
                    HLA_prob[allele_pair] = 0.0
                HLA_prob[allele_pair] += (float(count) / len(alleles))

    if len(HLA_prob) <= 0:
        return (HLA_prob, 0) if return_iters else HLA_prob

    def choose_top_alleles(HLA_prob):
        HLA_prob_list = [[allele_pair, prob] for allele_pair, prob in HLA_prob.items()]
        HLA_prob_list = sorted(HLA_prob_list, key=cmp_to_key(HLA_prob_cmp))
        HLA_prob = {}
        best_prob = HLA_prob_list[0][1]
        for i in range(len(HLA_prob_list)):
            allele_pair, prob = HLA_prob_list[i]
            if prob * 2 <= best_prob:
                break
            HLA_prob[allele_pair] = prob
        normalize(HLA_prob)
        return HLA_prob
    HLA_prob = choose_top_alleles(HLA_prob)

    def next_prob(HLA_cmpt, HLA_prob):
        HLA_prob_next = {}
        for cmpt, count in HLA_cmpt.items():
            alleles = cmpt.split('-')
            prob = 0.0
            for allele in alleles:
                for allele_pair in HLA_prob.keys():
                    if allele in allele_pair:  # NB: substring test on the pair STRING
                        prob += HLA_prob[allele_pair]
            for allele in alleles:
                for allele_pair in HLA_prob.keys():
                    if allele not in allele_pair:
                        continue
                    if allele_pair not in HLA_prob_next:
                        HLA_prob_next[allele_pair] = 0.0
                    HLA_prob_next[allele_pair] += (float(count) * HLA_prob[allele_pair] / prob)
        normalize(HLA_prob_next)
        return HLA_prob_next

    diff, iter = 1.0, 0
    while diff > 0.0001 and iter < 1000:
        HLA_prob_next = next_prob(HLA_cmpt, HLA_prob)
        diff = prob_diff(HLA_prob, HLA_prob_next)
        HLA_prob = HLA_prob_next
        HLA_prob = choose_top_alleles(HLA_prob)
        iter += 1

    HLA_prob = [[allele_pair, prob] for allele_pair, prob in HLA_prob.items()]
    HLA_prob = sorted(HLA_prob, key=cmp_to_key(HLA_prob_cmp))
    return (HLA_prob, iter) if return_iters else HLA_prob
