"""CPU oracle for HISAT-genotype's typing hot path — TEST INFRASTRUCTURE ONLY.

This file restates, in plain Python (allele sets as arbitrary-precision integers, one bit per allele),
the algorithm of the reference's stage (a) "per-read allele compatibility" and the glue around
stage (b).  It is the checker the CUDA path is compared against; the product never imports it
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do).

Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against vectors
captured from the unmodified reference (oracle/ref_rig/run_reference.py -> tests/golden/*.json.gz):
alignment lines in, Gene_cmpt / Gene_counts (three tables), num_reads / num_pairs, pileups,
alternative-haplotype tables, identify_ambigious_diffs results and EM inputs out.

All citations are relative to /root/reference/hisatgenotype_modules/ :
  core   = hisatgenotype_typing_core.py
  common = hisatgenotype_typing_common.py

Allele bit order: alleles of a locus (backbone excluded) sorted by name — the order of the names
inside a Gene_cmpt key (core:1229-1230), so a class bitset maps to its key by listing set bits.
"""
from __future__ import annotations

import re

MATCH, MISMATCH, INSERTION, DELETION = "match", "mismatch", "insertion", "deletion"
UNKNOWN = "unknown"
_CIGAR_RE = re.compile(r"(\d+)([A-Za-z=])")


def lower_bound(rows, key):
    """First index whose position is >= key (common:406-422).  rows: list of [pos, ...]."""
    lo, hi = 0, len(rows)
    while lo < hi:
        mid = (lo + hi) // 2
        if rows[mid][0] < key:
            lo = mid + 1
        else:
            hi = mid
    return lo


# ------------------------------------------------------------------------------------------------
# Per-locus tables
# ------------------------------------------------------------------------------------------------
class OracleLocus:
    """Integer view of one locus, built from the reference's own containers.

    gene_vars   {var_id: [type, pos, data]}          (Vars[gene],  common:339-368)
    var_list    [[pos, var_id], ...] sorted by pos   (Var_list[gene])
    links       {var_id: [allele, ...]}              (Links, common:388-403)
    gene_names  Gene_names[gene] (backbone first)    (core:2476-2478)
    """

    def __init__(self, base_fname, gene, ref_allele, ref_seq, gene_vars, var_list, links, gene_names,
                 gene_lengths, exons, primary_exons):
        self.base_fname = base_fname
        self.gene = gene
        self.ref_allele = ref_allele
        self.ref_seq = ref_seq
        self.gene_vars = {k: [v[0], int(v[1]), v[2]] for k, v in gene_vars.items()}
        self.var_list = [[int(p), i] for p, i in var_list]
        self.links = links
        self.gene_names = list(gene_names)
        self.gene_lengths = gene_lengths
        self.exons = [list(e) for e in exons]
        self.primary_exons = [list(e) for e in primary_exons]
        # table keys: Gene_names minus BACKBONE (core:1338-1347)
        self.table_names = [n for n in self.gene_names if n.find("BACKBONE") == -1]
        self.names = sorted(self.table_names)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.A = len(self.names)
        self.all_mask = (1 << self.A) - 1
        self.row_of = {vid: r for r, (_, vid) in enumerate(self.var_list)}
        self.V = len(self.var_list)
        self.bits = []  # allele bitset per variant row; None when the id is not in Links
        for _, vid in self.var_list:
            if vid in links:
                b = 0
                for n in links[vid]:
                    i = self.index.get(n)
                    if i is not None:
                        b |= 1 << i
                self.bits.append(b)
            else:
                self.bits.append(None)
        # running max of right ends in Var_list order (core:393-401)
        self.maxright = []
        cur = -1
        for _, vid in self.var_list:
            t, p, d = self.gene_vars[vid]
            if t == DELETION:
                p = p + int(d) - 1
            cur = max(cur, p)
            self.maxright.append(cur)
        # allele -> variant ids in Var_list order (core:476-487)
        self.allele_vars = {}
        for _, vid in self.var_list:
            if vid not in links:
                continue
            for n in links[vid]:
                if n not in self.index and n != ref_allele:
                    continue
                self.allele_vars.setdefault(n, []).append(vid)
        self.exon_vars = self._exonic(self.exons)
        self.primary_exon_vars = self._exonic(self.primary_exons)
        self.allele_reps, self.allele_rep_groups = self.rep_alleles(self.exon_vars, None)
        self.allele_rep_set = set(self.allele_reps.values())
        self.primary_reps, self.primary_rep_groups = self.rep_alleles(self.primary_exon_vars, self.allele_rep_set)
        self.primary_rep_set = set(self.primary_reps.values())
        self.exon_mask = self.mask_of(self.allele_rep_set)
        self.primary_mask = self.mask_of(self.primary_rep_set)
        self.alts_left, self.alts_right = get_alternatives(self)
        self.alts_left_list = sorted(([int(k.split("-")[-1]), k] for k in self.alts_left), key=lambda x: x[0])
        self.alts_right_list = sorted(([int(k.split("-")[0]), k] for k in self.alts_right), key=lambda x: x[0])

    def mask_of(self, names):
        m = 0
        for n in names:
            if n in self.index:
                m |= 1 << self.index[n]
        return m

    def names_of(self, bits):
        out, i = [], 0
        while bits:
            if bits & 1:
                out.append(self.names[i])
            bits >>= 1
            i += 1
        return out

    def var_right(self, vid):
        t, p, d = self.gene_vars[vid]
        return p + int(d) - 1 if t == DELETION else p

    def _exonic(self, exons):
        """ids of variants lying entirely inside an exon (core:67-78)."""
        out = set()
        for vid, (t, p, d) in self.gene_vars.items():
            r = p + int(d) - 1 if t == DELETION else p
            for el, er in exons:
                if p >= el and r <= er:
                    out.add(vid)
        return out

    def rep_alleles(self, exon_vars, in_alleles):
        """Group alleles with identical exonic variant sets; representative = first member met while
        walking Links in file order (core:86-115)."""
        avars, order = {}, []
        for vid, alleles in self.links.items():
            if vid not in exon_vars:
                continue
            for n in alleles:
                if in_alleles is not None and n not in in_alleles:
                    continue
                if n not in avars:
                    avars[n] = set()
                    order.append(n)
                avars[n].add(vid)
        groups = {}
        for n in order:
            groups.setdefault(frozenset(avars[n]), []).append(n)
        reps, rep_groups = {}, {}
        for members in groups.values():
            rep_groups[members[0]] = members
            for m in members:
                reps[m] = members[0]
        return reps, rep_groups


# ------------------------------------------------------------------------------------------------
# get_alternatives (common:1424-1657): haplotypes with identical sequence around each deletion
# ------------------------------------------------------------------------------------------------
def get_alternatives(loc):
    ref_seq, gvars, var_list = loc.ref_seq, loc.gene_vars, loc.var_list
    L = len(ref_seq)
    second = set()
    for vs in loc.allele_vars.values():
        for a, b in zip(vs[:-1], vs[1:]):
            second.add((a, b))
    by_right = []
    for _, vid in var_list:
        t, p, d = gvars[vid]
        if t == DELETION:
            p = p + int(d) - 1
        elif t == INSERTION:
            p += 1
        by_right.append([p, vid])
    by_right.sort(key=lambda x: x[0])
    out = {True: {}, False: {}}

    def extend(ht, leftward, exclude):
        """All one-base extensions of haplotype ht=[left, ids..., right] (common:1447-1527)."""
        pos = ht[0] - 1 if leftward else ht[-1] + 1
        if pos < 0 or pos >= L:
            return []
        if leftward:
            res = [([pos] + ht[1:], ref_seq[pos])]
            nxt = ht[1] if len(ht) > 2 else None
            hi = lower_bound(by_right, pos + 1)
            for j in range(hi - 1, -1, -1):
                vid = by_right[j][1]
                t, vp, d = gvars[vid]
                if t == DELETION:
                    if vp == 0:
                        continue
                    vp = vp + int(d) - 1
                if vp > pos:
                    continue
                if vp < pos:
                    break
                if vid in exclude:
                    continue
                if nxt is not None and (vid, nxt) not in second:
                    continue
                if t == "single":
                    res.append(([vp, vid] + ht[1:], d))
                elif t == DELETION:
                    res += extend([vp - int(d) + 1, vid] + ht[1:], leftward, exclude)
        else:
            res = [(ht[:-1] + [pos], ref_seq[pos])]
            prv = ht[-2] if len(ht) > 2 else None
            for j in range(lower_bound(var_list, pos), len(var_list)):
                vid = var_list[j][1]
                t, vp, d = gvars[vid]
                if vp < pos:
                    continue
                if vp > pos:
                    break
                if vid in exclude:
                    continue
                if prv is not None and (prv, vid) not in second:
                    continue
                if t == "single":
                    res.append((ht[:-1] + [vid, vp], d))
                elif t == DELETION:
                    res += extend(ht[:-1] + [vid, vp + int(d) - 1], leftward, exclude)
        return res

    def to_str(ht):
        return "-".join(str(x) for x in ht)

    def recur(orig, ht, alt, leftward, depth):
        found = False
        ext_alt = extend(alt, leftward, [orig])
        for nht, b1 in extend(ht, leftward, []):
            for nalt, b2 in ext_alt:
                if b1 != b2:
                    continue
                if (nht[0] == nalt[0]) if leftward else (nht[-1] == nalt[-1]):
                    continue
                found = True
                recur(orig, nht, nalt, leftward, depth + 1)
        if depth > 0 and not found:
            a, b = to_str(ht), to_str(alt)
            out[leftward].setdefault(a, set()).add(b)
            out[leftward].setdefault(b, set()).add(a)

    for _, vid in var_list:
        t, p, d = gvars[vid]
        if p == 0 or t != DELETION:
            continue
        n = int(d)
        if p + n >= L:
            continue
        recur(vid, [p, vid, p + n - 1], [p + n, p + n - 1], True, 0)
        recur(vid, [p, vid, p + n - 1], [p, p - 1], False, 0)
    return out[True], out[False]


# ------------------------------------------------------------------------------------------------
# Pileup (common:1059-1184) — only the parts the typing loop reads: per-position counts and nt_set
# ------------------------------------------------------------------------------------------------
def parse_cigar(s):
    return [(op, int(n)) for n, op in _CIGAR_RE.findall(s)]


def get_mpileup(sam_lines, ref_len, base_locus, allow_discordant):
    counts = [dict() for _ in range(ref_len)]
    for line in sam_lines:
        cols = line.strip().split()
        flag, pos, cigar, seq = int(cols[1]), int(cols[3]), cols[5], cols[9]
        if flag & 0x4:
            continue
        pos -= base_locus + 1
        if pos < 0:
            continue
        if not allow_discordant and not (flag & 0x2):
            continue
        rpos, gpos = 0, pos
        for op, n in parse_cigar(cigar):
            if op in "MD":
                for j in range(n):
                    nt = seq[rpos + j] if op == "M" else "D"
                    if gpos + j < ref_len:
                        counts[gpos + j][nt] = counts[gpos + j].get(nt, 0) + 1
            if op in "MND":
                gpos += n
            if op in "MIS":
                rpos += n
    nt_sets = []
    for c in counts:
        depth = sum(c.values())
        s = []
        if depth >= 20:
            for nt, k in c.items():
                if nt in "ACGT" and (k >= depth * 0.2 or k >= 7):
                    s.append(nt)
        nt_sets.append(s)
    return counts, nt_sets


# ------------------------------------------------------------------------------------------------
# Error correction of one M segment (core:119-243)
# ------------------------------------------------------------------------------------------------
def _known_single(loc, pos, base):
    """Known `single` variant at pos with this base (core:159-169, 204-214, 949-961), else UNKNOWN.
    Novel variants registered earlier resolve to nv* ids in the reference; both nv* and `unknown`
    are turned into matches downstream (core:1361-1366), so they are one state here."""
    vl = loc.var_list
    j = lower_bound(vl, pos)
    while j < len(vl) and vl[j][0] == pos:
        t, _, d = loc.gene_vars[vl[j][1]]
        if t == "single" and d == base:
            return vl[j][1]
        j += 1
    return UNKNOWN


def error_correct(loc, read_seq, read_pos, nt_sets, seg):
    ref_seq = loc.ref_seq
    out, ncorr = [], 0
    for k, ent in enumerate(seg):
        typ, left, length = ent[:3]
        if left >= len(ref_seq):
            # the reference stops correcting here and keeps the remaining entries (core:138-139)
            out.extend(seg[k:])
            break
        if typ == MATCH:
            last = 0
            for j in range(length):
                if read_pos + j >= len(read_seq) or left + j >= len(ref_seq):
                    continue
                bp = read_seq[read_pos + j]
                s = nt_sets[left + j]
                if len(s) > 0 and bp not in s:
                    bp = "N" if len(s) > 1 else s[0]
                    read_seq = read_seq[:read_pos + j] + bp + read_seq[read_pos + j + 1:]
                    ncorr += 1
                    vid = _known_single(loc, left + j, bp) if bp != "N" else UNKNOWN
                    if j > last:
                        out.append([MATCH, left + last, j - last])
                    out.append([MISMATCH, left + j, 1, vid])
                    last = j + 1
            if last < length:
                out.append([MATCH, left + last, length - last])
        else:
            bp, ref_bp = read_seq[read_pos], ref_seq[left]
            s = nt_sets[left]
            ent = list(ent)
            if len(s) > 0 and bp not in s:
                bp = "N" if len(s) > 1 else s[0]
                read_seq = read_seq[:read_pos] + bp + read_seq[read_pos + 1:]
                if bp == "N":
                    ent[3] = UNKNOWN
                elif bp == ref_bp:
                    ent = [MATCH, left, 1]
                    ncorr += 1
                else:
                    ent[3] = _known_single(loc, left, bp)
            out.append(ent)
        read_pos += length
    merged = []
    for ent in out:
        if ent[0] == MATCH and merged and merged[-1][0] == MATCH:
            merged[-1] = [MATCH, merged[-1][1], merged[-1][2] + ent[2]]
        else:
            merged.append(ent)
    return merged, read_seq, ncorr


# ------------------------------------------------------------------------------------------------
# identify_ambigious_diffs (common:1663-1955)
# ------------------------------------------------------------------------------------------------
class AmbiguityError(Exception):
    """check_amb_uniqueness (validation_check.py:313-341) would print and exit(1)."""


def identify_ambiguous_diffs(loc, cmp_list):
    gvars = loc.gene_vars
    n = len(cmp_list)
    cmp_left, cmp_right = 0, n - 1
    left = cmp_list[0][1]
    right = cmp_list[-1][1] + cmp_list[-1][2] - 1
    left_alts, right_alts = set(), set()

    def ids_and_len(part):
        ids, seqlen = [], 0
        for e in part:
            if e[0] == MATCH:
                seqlen += len(loc.ref_seq[e[1]:e[1] + e[2]])
            elif e[0] == MISMATCH:
                seqlen += 1
            if len(e) > 3 and e[3] != "" and e[3] != UNKNOWN:
                ids.append(e[3])
        return ids, seqlen

    def hv_ids_between(lo, hi):
        return [cmp_list[j][3] for j in range(lo, hi)
                if cmp_list[j][0] != MATCH and cmp_list[j][3].startswith("hv")]

    # ---- left end -------------------------------------------------------------------------------
    found = False
    for i in range(n - 1, -1, -1):
        typ, cur_left, length = cmp_list[i][:3]
        vid = cmp_list[i][3] if typ in (MISMATCH, DELETION) else ""
        if typ != MATCH and not vid.startswith("hv"):
            continue
        cur_right = cur_left + length - 1 if typ in (MATCH, DELETION) else cur_left
        cur_ids, cur_len = ids_and_len(cmp_list[:i + 1])
        joined = "-".join(cur_ids)
        al = loc.alts_left_list
        hit = False
        start = min(lower_bound(al, cur_right + 1) + 1, len(al))
        for j in range(start - 1, -1, -1):
            hpos, key = al[j]
            if hpos < cur_left:
                break
            if hpos > cur_right:
                continue
            if cur_ids and key.find(joined) == -1:
                continue
            toks = key.split("-")[:-1]
            if len(cur_ids) + 1 == len(toks):
                if left < int(toks[0]):
                    continue
            else:
                t2, p2, d2 = gvars[toks[len(toks) - len(cur_ids) - 1]]
                if t2 == DELETION:
                    p2 = p2 + int(d2) - 1
                if left <= p2:
                    continue
            hit = True
            for alt in loc.alts_left[key]:
                atoks = alt.split("-")
                a_right = int(atoks[-1])
                assert a_right <= cur_right
                seq_pos, cur_pos = cur_right - a_right, a_right
                part = []
                for v in reversed(atoks[1:-1]):
                    vt, vp, vd = gvars[v]
                    dl = 0
                    if vt == DELETION:
                        dl = int(vd)
                        vp = vp + dl - 1
                    assert vp <= cur_pos
                    nxt = seq_pos + (cur_pos - vp)
                    if nxt >= cur_len:
                        break
                    if vt == "single":
                        nxt += 1
                        npos = vp - 1
                    else:
                        assert vt == DELETION
                        npos = vp - dl
                    part.insert(0, v)
                    if nxt >= cur_len:
                        break
                    seq_pos, cur_pos = nxt, npos
                if part:
                    seq_left = cur_len - seq_pos - 1
                    s = "%d-%s" % (cur_pos - seq_left, "-".join(part))
                    if found:
                        mid = hv_ids_between(i + 1, cmp_left)
                        if mid:
                            s += "-" + "-".join(mid)
                    left_alts.add(s)
        if hit:
            if not found:
                cmp_left = i + 1
                left_alts.add(("%d-%s" % (left, joined)) if cur_ids else str(left))
            found = True
    if not found:
        left_alts.add(str(left))

    # ---- right end ------------------------------------------------------------------------------
    found = False
    for i in range(n):
        typ, cur_left, length = cmp_list[i][:3]
        vid = cmp_list[i][3] if typ in (MISMATCH, DELETION) else ""
        if typ != MATCH and not vid.startswith("hv"):
            continue
        cur_right = cur_left + length - 1 if typ in (MATCH, DELETION) else cur_left
        cur_ids, cur_len = ids_and_len(cmp_list[i:])
        joined = "-".join(cur_ids)
        ar = loc.alts_right_list
        hit = False
        for j in range(lower_bound(ar, cur_left), len(ar)):
            hpos, key = ar[j]
            if hpos > cur_right:
                break
            if hpos < cur_left:
                continue
            if cur_ids and key.find(joined) == -1:
                continue
            toks = key.split("-")[1:]
            if len(cur_ids) + 1 == len(toks):
                if right > int(toks[-1]):
                    continue
            else:
                p2 = gvars[toks[len(cur_ids)]][1]
                if right >= p2:
                    continue
            hit = True
            for alt in loc.alts_right[key]:
                atoks = alt.split("-")
                a_left = int(atoks[0])
                assert cur_left <= a_left
                seq_pos, cur_pos = a_left - cur_left, a_left
                part = []
                for v in atoks[1:-1]:
                    vt, vp, vd = gvars[v]
                    assert vp >= cur_pos
                    nxt = seq_pos + (vp - cur_pos)
                    if nxt >= cur_len:
                        break
                    if vt == "single":
                        nxt += 1
                        npos = vp + 1
                    else:
                        assert vt == DELETION
                        npos = vp + int(vd)
                    part.append(v)
                    if nxt >= cur_len:
                        break
                    seq_pos, cur_pos = nxt, npos
                if part:
                    seq_left = cur_len - seq_pos - 1
                    assert seq_left >= 0
                    s = ""
                    if found:
                        mid = hv_ids_between(cmp_right + 1, i)
                        if mid:
                            s = "-".join(mid) + "-"
                    s += "%s-%d" % ("-".join(part), cur_pos + seq_left)
                    right_alts.add(s)
        if hit:
            if not found:
                cmp_right = i - 1
                right_alts.add(("%s-%d" % (joined, right)) if cur_ids else str(right))
            found = True
    if not found:
        right_alts.add(str(right))
    if cmp_right < cmp_left:
        cmp_left = 0
        left_alts = {str(left)}
    # check_amb_uniqueness (validation_check.py:313-341; SANITY_CHECK is effectively always on)
    seen = set()
    for s in left_alts:
        k = "-".join(s.split("-")[1:])
        if k == "":
            continue
        if k in seen:
            raise AmbiguityError(k)
        seen.add(k)
    for s in right_alts:
        k = "-".join(s.split("-")[:-1])
        if k == "":
            continue
        if k in seen:
            raise AmbiguityError(k)
        seen.add(k)
    return cmp_left, cmp_right, left_alts, right_alts


# ------------------------------------------------------------------------------------------------
# Haplotype -> allele set (core:626-677) and exon clipping (core:718-792)
# ------------------------------------------------------------------------------------------------
def _novel_fields(vid):
    # canonical novel id "nv<T><pos>_<len>" (see walk_record)
    t = INSERTION if vid[2] == "I" else DELETION
    p, ln = vid[3:].split("_")
    return t, int(p), int(ln)


def _var_fields(loc, vid):
    """(type, pos, length-as-int-for-deletions) of a known or novel variant id."""
    if vid.startswith("nv"):
        return _novel_fields(vid)
    t, p, d = loc.gene_vars[vid]
    return t, p, (int(d) if t == DELETION else d)


def compat_set(loc, ht, table_mask):
    """Alleles compatible with haplotype string ht, restricted to a table (add_count, core:626-677)."""
    toks = ht.split("-")
    left, right = int(toks[0]), int(toks[-1])
    assert left <= right
    ids = toks[1:-1]
    alleles = loc.all_mask
    for v in ids:
        if v.startswith("nv") or v not in loc.links:
            continue
        alleles &= loc.bits[loc.row_of[v]]
    idset = set(ids)
    neg = 0
    vl = loc.var_list
    j = min(lower_bound(vl, right + 1), len(vl) - 1)
    while j >= 0:
        vid = vl[j][1]
        if vid in idset or vid not in loc.links:
            j -= 1
            continue
        if loc.maxright[j] < left:
            break
        vleft = vl[j][0]
        vright = loc.var_right(vid)
        if left <= vleft <= right or left <= vright <= right:
            neg |= loc.bits[j]
        j -= 1
    return alleles & ~neg & table_mask


def exon_haplotypes(loc, ht, exons):
    """Clip a haplotype to each overlapping exon (core:718-792)."""
    toks = ht.split("-")
    h_left, h_right, ids = int(toks[0]), int(toks[-1]), toks[1:-1]
    res = []
    for e_left, e_right in exons:
        if e_left > h_right or e_right < h_left:
            continue
        left, right, cur = h_left, h_right, list(ids)
        if left < e_left:
            split = False
            for i, v in enumerate(cur):
                t, p, d = _var_fields(loc, v)
                if (t != DELETION and p >= e_left) or (t == DELETION and p - 1 >= e_left):
                    left, cur, split = e_left, cur[i:], True
                    break
                if t == DELETION and p + d >= e_left:
                    left, cur, split = p + d, cur[i + 1:], True
                    break
            if not split:
                left, cur = e_left, []
        if right > e_right:
            split = False
            for i in range(len(cur) - 1, -1, -1):
                t, p, d = _var_fields(loc, cur[i])
                r = p + d - 1 if t == DELETION else p
                if (t != DELETION and r <= e_right) or (t == DELETION and r + 1 <= e_right):
                    right, cur, split = e_right, cur[:i + 1], True
                    break
                if t == DELETION and r - d <= e_right:
                    right, cur, split = r - d, cur[:i], True
                    break
            if not split:
                right, cur = e_right, []
                # NB: the reference rebuilds [ht_left, ht_right] here using the *already clipped* left
        assert left <= right
        res.append("-".join([str(left)] + cur + [str(right)]))
    return res


# ------------------------------------------------------------------------------------------------
# CIGAR x MD x Zs walk of one alignment record (core:876-1095)
# ------------------------------------------------------------------------------------------------
class Record:
    __slots__ = ("read_id", "flag", "pos", "cigar", "seq", "NM", "NH", "MD", "Zs")


def parse_record(line, simulation, base_locus):
    cols = line.strip().split()
    r = Record()
    r.read_id = cols[0].split("|")[0] if simulation else cols[0]
    r.flag = int(cols[1])
    r.pos = int(cols[3]) - (base_locus + 1)
    r.cigar = cols[5]
    r.seq = cols[9]
    r.NM, r.NH, r.MD, r.Zs = "", "", "", ""
    for c in cols[11:]:
        if c.startswith("Zs"):
            r.Zs = c[5:]
        elif c.startswith("MD"):
            r.MD = c[5:]
        elif c.startswith("NM"):
            r.NM = int(c[5:])
        elif c.startswith("NH"):
            r.NH = int(c[5:])
    return r


def walk_record(loc, rec, counts, nt_sets, error_correction):
    """Return (cmp_list, right_pos, num_error_correction, likely_misalignment, read_seq)."""
    MD, seq = rec.MD, rec.seq
    assert MD != ""
    zs = []
    if rec.Zs:
        for item in rec.Zs.split(","):
            off, kind, vid = item.split("|")
            zs.append((int(off), kind, vid))
    md_i, md_len, zs_i = 0, 0, 0
    zs_pos = zs[0][0] if zs else 0
    read_pos, right_pos = 0, rec.pos
    cmp_list, ncorr, misaligned = [], 0, False
    cig = parse_cigar(rec.cigar)
    for ci, (op, length) in enumerate(cig):
        if op == "M":
            first, used, seg_start = True, 0, len(cmp_list)
            while True:
                if not first or md_len == 0:
                    if MD[md_i].isdigit():
                        num = 0
                        while md_i < len(MD) and MD[md_i].isdigit():
                            num = num * 10 + int(MD[md_i])
                            md_i += 1
                        md_len += num
                if md_len >= length:
                    md_len -= length
                    if length > used:
                        cmp_list.append([MATCH, right_pos + used, length - used])
                    break
                first = False
                base = seq[read_pos + md_len]
                assert MD[md_i] in "ACGT"
                md_i += 1
                if md_len > used:
                    cmp_list.append([MATCH, right_pos + used, md_len - used])
                if read_pos + md_len == zs_pos and zs_i < len(zs):
                    assert zs[zs_i][1] == "S"
                    vid = zs[zs_i][2]
                    zs_i += 1
                    zs_pos += 1
                    if zs_i < len(zs):
                        zs_pos += zs[zs_i][0]
                else:
                    vid = _known_single(loc, right_pos + md_len, base)
                cmp_list.append([MISMATCH, right_pos + md_len, 1, vid])
                used = md_len + 1
                md_len += 1
                if md_len == length:
                    md_len = 0
                    break
            if error_correction:
                seg, seq, k = error_correct(loc, seq, read_pos, nt_sets, cmp_list[seg_start:])
                cmp_list = cmp_list[:seg_start] + seg
                ncorr += k
        elif op == "I":
            vid = UNKNOWN
            if read_pos == zs_pos and zs_i < len(zs):
                assert zs[zs_i][1] == "I"
                vid = zs[zs_i][2]
                zs_i += 1
                if zs_i < len(zs):
                    zs_pos += zs[zs_i][0]
            else:
                vl = loc.var_list
                j = lower_bound(vl, right_pos)
                while j < len(vl) and vl[j][0] == right_pos:
                    t, _, d = loc.gene_vars[vl[j][1]]
                    if t == INSERTION and len(d) == length:
                        vid = vl[j][1]
                        break
                    j += 1
            cmp_list.append([INSERTION, right_pos, length, vid])
            if "N" in seq[read_pos:read_pos + length]:
                misaligned = True
        elif op == "D":
            if MD[md_i] == "0":
                md_i += 1
            assert MD[md_i] == "^"
            md_i += 1
            while md_i < len(MD) and MD[md_i] in "ACGT":
                md_i += 1
            vid = UNKNOWN
            if read_pos == zs_pos and zs_i < len(zs) and zs[zs_i][1] == "D":
                vid = zs[zs_i][2]
                zs_i += 1
                if zs_i < len(zs):
                    zs_pos += zs[zs_i][0]
            else:
                vl = loc.var_list
                j = lower_bound(vl, right_pos)
                while j < len(vl) and vl[j][0] == right_pos:
                    t, _, d = loc.gene_vars[vl[j][1]]
                    if t == DELETION and int(d) == length:
                        vid = vl[j][1]
                        break
                    j += 1
            cmp_list.append([DELETION, right_pos, length, vid])
            if right_pos < len(counts):
                dels = counts[right_pos].get("D", 0)
                nts = sum(k for nt, k in counts[right_pos].items() if nt != "D")
                if loc.base_fname == "hla" and dels * 6 < nts:
                    misaligned = True
        elif op == "S":
            if ci == 0:
                zs_pos += length
            else:
                assert ci + 1 == len(cig)
        else:
            raise AssertionError("unsupported CIGAR op %s (core:1086-1088)" % op)
        if op in "MND":
            right_pos += length
        if op in "MIS":
            read_pos += length
    # soft clips are cut from the read afterwards (core:1099-1107)
    if cig and cig[0][0] == "S":
        seq = seq[cig[0][1]:]
    if len(cig) > 1 and cig[-1][0] == "S":
        seq = seq[:-cig[-1][1]]
    return cmp_list, right_pos, ncorr, misaligned, seq


# ------------------------------------------------------------------------------------------------
# The per-read loop (core:598-1596)
# ------------------------------------------------------------------------------------------------
class Table:
    """One Gene_cmpt / Gene_counts pair.  classes: {bitset: [count, first_seen_pair]} in first-seen order."""

    def __init__(self, mask, active=True):
        self.mask = mask
        self.active = active
        self.classes = {}
        self.counts = {}  # allele index -> count, insertion order = reference dict order

    def cmpt_items(self, loc):
        return [["-".join(loc.names_of(b)), c[0]] for b, c in self.classes.items()]

    def count_items(self, loc):
        return [[loc.names[i], c] for i, c in self.counts.items()]


def type_locus(loc, sam_lines, simulation=False, num_editdist=2, error_correction=True, allow_discordant=False,
               base_locus=0, collect=None, collect_hts=None):
    """Stage (a) for one locus.  Returns dict with tables 'gene', 'exon', 'primary', num_reads, num_pairs."""
    hla = loc.base_fname == "hla"
    counts, nt_sets = get_mpileup(sam_lines, len(loc.ref_seq), base_locus, allow_discordant)
    tables = {
        "primary": Table(loc.primary_mask, hla),
        "exon": Table(loc.exon_mask, hla),
        "gene": Table(loc.all_mask, True),
    }
    # allele order of the per-read dicts = Gene_names order (core:1338-1347); add_stat walks it (core:1179)
    gn_order = [loc.index[n] for n in loc.table_names]
    num_reads = num_pairs = 0
    seen = {"L": set(), "R": set(), "U": set()}
    prev_id = None
    left_hts, right_hts = set(), set()

    def flush(pair_index):
        per = {k: {} for k in tables}  # allele bit -> count, stored sparsely as list of compat bitsets
        sets = {k: [] for k in tables}
        for ht in left_hts | right_hts:
            for eh in exon_haplotypes(loc, ht, loc.primary_exons):
                sets["primary"].append(compat_set(loc, eh, tables["primary"].mask))
            for eh in exon_haplotypes(loc, ht, loc.exons):
                sets["exon"].append(compat_set(loc, eh, tables["exon"].mask))
            sets["gene"].append(compat_set(loc, ht, tables["gene"].mask))
        for k, tb in tables.items():
            if not tb.active or tb.mask == 0:
                continue
            # max count per allele over the list of compat sets (add_stat, core:1171-1236)
            level = [tb.mask]  # level[c] = alleles with count >= c
            for s in sets[k]:
                level.append(0)
                for c in range(len(level) - 1, 0, -1):
                    level[c] |= level[c - 1] & s
            best = tb.mask
            for c in range(len(level) - 1, 0, -1):
                if level[c]:
                    best = level[c]
                    break
            for i in gn_order:
                if (best >> i) & 1:
                    tb.counts[i] = tb.counts.get(i, 0) + 1
            if best in tb.classes:
                tb.classes[best][0] += 1
            else:
                tb.classes[best] = [1, pair_index]
            if collect is not None:
                collect.append((pair_index, k, best))

    for line_i, line in enumerate(sam_lines):
        rec = parse_record(line, simulation, base_locus)
        if rec.pos < 0:
            continue
        if rec.flag & 0x4:
            continue
        if rec.NM > num_editdist:
            continue
        if rec.NH > 1:
            continue
        if not allow_discordant and not (rec.flag & 0x2):
            continue
        is_left = bool(rec.flag & 0x40)
        kind = "L" if is_left else ("R" if rec.flag & 0x80 else "U")
        if kind == "U":
            assert allow_discordant
        if rec.read_id in seen[kind]:
            continue
        seen[kind].add(rec.read_id)
        cmp_list, right_pos, ncorr, misaligned, seq = walk_record(loc, rec, counts, nt_sets, error_correction)
        if right_pos > len(loc.ref_seq):
            continue
        if ncorr > max(1, num_editdist):
            continue
        if misaligned:
            continue
        # novel variants (core:1126-1164): unknown indels get a canonical novel id; unknown mismatches
        # become matches below whether or not they were registered
        for e in cmp_list:
            if e[0] in (INSERTION, DELETION) and e[3] == UNKNOWN:
                e[3] = "nv%s%d_%d" % ("I" if e[0] == INSERTION else "D", e[1], e[2])
        num_reads += 1
        if rec.read_id != prev_id:
            if prev_id is not None:
                flush(num_pairs)
                num_pairs += 1
            left_hts, right_hts = set(), set()
        # cmp_list2 (core:1351-1368)
        c2 = []
        for e in cmp_list:
            if e[0] == MATCH or (e[0] == MISMATCH and (e[3] == UNKNOWN or e[3].startswith("nv"))):
                ln = e[2] if e[0] == MATCH else 1
                if c2 and c2[-1][0] == MATCH:
                    c2[-1][2] += ln
                else:
                    c2.append([MATCH, e[1], ln])
            else:
                c2.append(list(e))
        l, r, la, ra = identify_ambiguous_diffs(loc, c2)
        mid = [e[3] for e in c2[l:r + 1] if e[0] != MATCH]
        own = set()
        for a in la:
            for b in ra:
                ht = "-".join(a.split("-") + mid + b.split("-"))
                (left_hts if is_left else right_hts).add(ht)
                own.add(ht)
        if collect_hts is not None:  # the read's own haplotype strings (core:1386-1406)
            collect_hts.append((line_i, rec.flag, sorted(own)))
        prev_id = rec.read_id
    if prev_id is not None:
        flush(num_pairs)
        num_pairs += 1
    return {"tables": tables, "num_reads": num_reads, "num_pairs": num_pairs, "counts": counts, "nt_sets": nt_sets}


# ------------------------------------------------------------------------------------------------
# Stage (b): EM (common:1282-1410), literal dict form for small inputs.  The C restatement used for
# timing and larger cases is oracle/em_oracle.c.
# ------------------------------------------------------------------------------------------------
def single_abundance(cmpt, remove_low=False, lengths=None):
    import math
    lengths = lengths or {}

    def norm(p):
        if lengths:
            tot = 0
            for a, m in p.items():
                tot += m / lengths[a]
            for a, m in p.items():
                p[a] = m / lengths[a] / tot
        else:
            tot = sum(p.values())
            for a, m in p.items():
                p[a] = m / tot

    items = [(k.split("-"), c) for k, c in cmpt.items()]
    prob = {}
    for als, c in items:
        for a in als:
            prob[a] = prob.get(a, 0.0) + float(c) / len(als)
    norm(prob)

    def step(p):
        q = {}
        for als, c in items:
            s = 0.0
            for a in als:
                if a in p:
                    s += p[a]
            if s <= 0.0:
                continue
            for a in als:
                if a in p:
                    q[a] = q.get(a, 0.0) + float(c) * p[a] / s
        norm(q)
        return q

    def prune(p):
        if not p:
            return p
        mx = max(p.values())
        return {a: v for a, v in p.items() if v >= mx / 10.0}

    diff, it = 1.0, 0
    while diff > 0.0001 and it < 1000:
        p1 = step(prob)
        p2 = step(p1)
        sr = sv = 0.0
        r, v = {}, {}
        for a in prob:
            r[a] = p1[a] - prob[a]
            sr += r[a] * r[a]
            v[a] = p2[a] - p1[a] - r[a]
            sv += v[a] * v[a]
        if sv > 0.0:
            g = -math.sqrt(sr / sv)
            for a in prob:
                p2[a] = max(0.0, prob[a] - 2 * g * r[a] + g * g * v[a])
            p1 = step(p2)
        diff = 0.0
        for a in prob:
            diff += abs(prob[a] - p1[a]) if a in p1 else prob[a]
        prob = p1
        if it >= 10 and remove_low:
            prob = prune(prob)
        it += 1
    if remove_low:
        prob = prune(prob)
    norm(prob)
    return sorted(([a, p] for a, p in prob.items()), key=lambda x: x[1], reverse=True), it


# ------------------------------------------------------------------------------------------------
# EM driver of typing() (core:1679-1789) on the tables of type_locus(): Gene_prob as the report ranks it.
# `em` = a single_abundance implementation returning (ranked list, iterations); default the C restatement.
# ------------------------------------------------------------------------------------------------
def locus_abundance(loc, res, remove_low=True, em=None):
    if em is None:
        import em_oracle
        em = em_oracle.single_abundance
    gene_cmpt = dict(map(tuple, res["tables"]["gene"].cmpt_items(loc)))
    if loc.base_fname != "hla":
        # core:1783-1789 (a single class raises TypeError on Python 3: Gene_cmpt.keys()[0])
        if len(gene_cmpt) <= 1:
            if len(gene_cmpt) == 1:
                raise TypeError("'dict_keys' object is not subscriptable")
            return []
        return em(gene_cmpt, False, {})[0]
    exon_cmpt = dict(map(tuple, res["tables"]["exon"].cmpt_items(loc)))
    exon_prob = em(exon_cmpt, remove_low, {})[0]                                     # core:1732-1737
    exon_alleles, exon_prob_sum = set(), 0.0
    for i, (allele, prob) in enumerate(exon_prob):                                  # core:1739-1749
        if i >= 10 and prob < 0.03:
            break
        group = loc.allele_rep_groups[allele]
        if len(group) <= 1:
            continue
        exon_prob_sum += prob
        exon_alleles |= set(group)
    if not exon_alleles:
        return exon_prob
    cmpt2 = {}
    for key, cnt in gene_cmpt.items():                                              # core:1753-1766
        k2 = "-".join(a for a in key.split("-") if a in exon_alleles)
        if k2:
            cmpt2[k2] = cmpt2.get(k2, 0) + cnt
    full = em(cmpt2, True, loc.gene_lengths)[0]                                     # core:1767
    comb = {a: p for a, p in exon_prob if a not in exon_alleles}                    # core:1771-1782
    for a, p in full:
        comb[a] = p * exon_prob_sum
    return sorted(([a, p] for a, p in comb.items()), key=lambda x: x[1], reverse=True)
