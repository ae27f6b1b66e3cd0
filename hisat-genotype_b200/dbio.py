"""Readers for the typing database files, producing the same containers the reference builds in
genotyping_locus (reference hisatgenotype_modules/hisatgenotype_typing_core.py:2417-2485).

Each reader takes the *text* of a file so that tests can feed fixtures without touching disk;
`load_database(prefix)` reads the files of `<ix_dir>/<base>`.

  read_locus        <- hisatgenotype_typing_common.py:279-309   (.locus, 7 columns, non-genome form)
  read_variants     <- hisatgenotype_typing_common.py:339-368   (.snp)
  read_links        <- hisatgenotype_typing_common.py:388-403   (.link)
  read_backbone     <- hisatgenotype_typing_common.py:313-334   (_backbone.fa)
  build_genes       <- hisatgenotype_typing_core.py:2199-2237, 2462-2485
  read_haplotypes / read_index_variants   <- hisatgenotype_typing_process.py:1088-1102, 1215-1220 (.haplotype, .index.snp)
  write_database_text   the files of process.py:1055-1106, 1242-1244 from the containers (round trip of the readers)
"""
from __future__ import annotations

import os


def read_locus(text):
    refGenes, refGene_loci = {}, {}
    for line in text.strip("\n").split("\n"):
        if not line.strip():
            continue
        name, chrom, left, right, _, exon_str, _strand = line.split()
        gene = name.split("*")[0]
        assert gene not in refGenes
        refGenes[gene] = name
        exons, primary = [], []
        for ex in exon_str.split(","):
            is_primary = ex.endswith("p")
            if is_primary:
                ex = ex[:-1]
            a, b = ex.split("-")
            exons.append([int(a), int(b)])
            if is_primary:
                primary.append([int(a), int(b)])
        refGene_loci[gene] = [name, chrom, int(left), int(right), exons, primary]
    return refGenes, refGene_loci


def read_variants(text):
    Vars, Var_list = {}, {}
    for line in text.strip("\n").split("\n"):
        if not line:
            continue
        var_id, var_type, name, pos, data = line.split("\t")
        gene = name.split("*")[0]
        Vars.setdefault(gene, {})
        Var_list.setdefault(gene, [])
        assert var_id not in Vars[gene]
        Vars[gene][var_id] = [var_type, int(pos), data]
        Var_list[gene].append([int(pos), var_id])
    for gene in Var_list:
        Var_list[gene].sort(key=lambda x: x[0])  # stable: file order inside one position
    return Vars, Var_list


def read_links(text):
    links = {}
    for line in text.strip("\n").split("\n"):
        if not line:
            continue
        cols = line.replace(" ", "\t").split("\t")
        assert cols[0] not in links
        links[cols[0]] = cols[1:]
    return links


def read_backbone(text):
    Genes = {}
    for chunk in text.strip("\n").split(">")[1:]:
        nl = chunk.find("\n")
        name = chunk[:nl]
        gene = name.split("*")[0]
        Genes.setdefault(gene, {})
        if name in Genes[gene]:
            raise SystemExit("Error: Nonunique sequence name: %s" % name)
        Genes[gene][name] = chunk[nl:].replace("\n", "")
    return Genes


def build_genes(Genes, Vars, Var_list, Links, allele_names):
    """Allele sequences = backbone + linked variants; then alleles identical to the backbone."""
    for gene in Genes:
        assert len(Genes[gene]) == 1
        bb_name, bb = list(Genes[gene].items())[0]
        gene_vars = Vars.get(gene, {})
        allele_vars = {}
        for _, var_id in Var_list.get(gene, []):
            for allele in Links.get(var_id, []):
                allele_vars.setdefault(allele, []).append(var_id)
        for allele, ids in allele_vars.items():
            out, prev = [], 0
            for var_id in ids:
                t, pos, data = gene_vars[var_id]
                assert prev <= pos
                if pos > prev:
                    out.append(bb[prev:pos])
                if t == "single":
                    out.append(data)
                    prev = pos + 1
                elif t == "deletion":
                    prev = pos + int(data)
                else:
                    assert t == "insertion"
                    out.append(data)
                    prev = pos
            if prev < len(bb):
                out.append(bb[prev:])
            Genes[gene][allele] = "".join(out)
        if len(Genes[gene]) <= 1:
            Genes[gene]["%s*GRCh38" % gene] = bb
    # alleles only listed in .allele are identical to the backbone (core:2462-2467).  The reference
    # iterates a Python set here, so their relative order is unpinned; sorted order is used.
    for allele in sorted(allele_names):
        gene = allele.split("*")[0]
        assert gene in Genes
        if allele not in Genes[gene]:
            Genes[gene][allele] = Genes[gene]["%s*BACKBONE" % gene]
    Gene_names = {g: list(d.keys()) for g, d in Genes.items()}
    Gene_lengths = {g: {a: len(s) for a, s in d.items()} for g, d in Genes.items()}
    return Gene_names, Gene_lengths


def read_haplotypes(text):
    """.haplotype: `htN \\t backbone \\t left \\t right \\t id,id,...` (written by process.py:1215-1220, read by hisat2-build
    --haplotype): {gene: [[ht_id, left, right, [var ids]], ...]} in file order."""
    out = {}
    for line in text.strip("\n").split("\n"):
        if not line:
            continue
        ht_id, name, left, right, ids = line.split("\t")
        out.setdefault(name.split("*")[0], []).append([ht_id, int(left), int(right), [v for v in ids.split(",") if v]])
    return out


def read_index_variants(text):
    """.index.snp: the subset of .snp that went into the graph index (same five columns, process.py:1088-1102); HISAT2 names
    these ids in the Zs tag of its alignments.  Returns the set of ids per gene."""
    Vars, _ = read_variants(text)
    return {gene: set(v) for gene, v in Vars.items()}


def write_database_text(d):
    """The inverse of load_database_text for the files this package reads: {extension: text} from the containers (variants
    in Var_list order, one allele name per line).  Round trip: load_database_text(write_database_text(d)) == d."""
    locus, snp, link, backbone, alleles = [], [], [], [], []
    for gene, name in d["refGenes"].items():
        _, chrom, left, right, exons, primary = d["refGene_loci"][gene]
        ex = ",".join("%d-%d%s" % (a, b, "p" if [a, b] in primary else "") for a, b in exons)
        seq = d["Genes"][gene][name]
        locus.append("%s\t%s\t%d\t%d\t%d\t%s\t+" % (name, chrom, left, right, len(seq), ex))
        backbone.append(">%s" % name)
        backbone += [seq[i:i + 60] for i in range(0, len(seq), 60)]
        for pos, var_id in d["Var_list"].get(gene, []):
            t, p, data = d["Vars"][gene][var_id]
            snp.append("%s\t%s\t%s\t%d\t%s" % (var_id, t, name, p, data))
            if var_id in d["Links"]:
                link.append("%s\t%s" % (var_id, " ".join(d["Links"][var_id])))
        alleles += [a for a in d["Gene_names"][gene] if a != name and not a.endswith("*GRCh38")]
    return {".locus": "\n".join(locus) + "\n", ".snp": "\n".join(snp) + "\n", ".link": "\n".join(link) + "\n",
            "_backbone.fa": "\n".join(backbone) + "\n", ".allele": "\n".join(alleles) + "\n",
            ".partial": "".join(a + "\n" for a in sorted(d.get("partial_alleles", ())))}


def load_database_text(db):
    """db: {'.locus': text, '.snp': text, '.link': text, '_backbone.fa': text, '.allele': text, '.partial': text}"""
    refGenes, refGene_loci = read_locus(db[".locus"])
    Vars, Var_list = read_variants(db[".snp"])
    Links = read_links(db[".link"])
    Genes = read_backbone(db["_backbone.fa"])
    alleles = [l.strip() for l in db[".allele"].split("\n") if l.strip()]
    partial = set(l.strip() for l in db.get(".partial", "").split("\n") if l.strip())
    for gene in refGene_loci:
        if gene not in Vars:
            Vars[gene], Var_list[gene] = {}, []
    Gene_names, Gene_lengths = build_genes(Genes, Vars, Var_list, Links, alleles)
    return dict(refGenes=refGenes, refGene_loci=refGene_loci, Vars=Vars, Var_list=Var_list, Links=Links,
                Genes=Genes, Gene_names=Gene_names, Gene_lengths=Gene_lengths, partial_alleles=partial)


def load_database(prefix):
    db = {}
    for ext in ["_backbone.fa", ".locus", ".snp", ".link", ".allele", ".partial"]:
        path = prefix + ext
        if not os.path.exists(path):
            raise SystemExit("Error: index files missing (%s)" % path)
        db[ext] = open(path).read()
    return load_database_text(db)


# ---- native reader (libhgt hgt_db_*, csrc/dbio.cpp): the same containers from the same files ---------------------------------
_DB = dict(EXONS=0, VAR_TYPE=1, VAR_POS=2, VAR_IN_INDEX=3, HAP_RANGE=4, GENE=16, BACKBONE=17, VAR_ID=18, VAR_DATA=19, VAR_LINKS=20,
           ALLELES=21, PARTIAL=22, HAP_ID=23, HAP_VARS=24)
_TYPES = ["single", "deletion", "insertion"]


def load_database_native(prefix):
    """load_database(prefix) with the files parsed by libhgt (hgt_db_open ...): returns the same dict, plus "haplotypes"
    (read_haplotypes form) and "index_vars" (read_index_variants form) when those files exist."""
    import ctypes

    import numpy as np

    from . import _lib
    L = _lib.lib()
    L.hgt_db_open.restype = ctypes.c_int
    L.hgt_db_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    L.hgt_db_close.restype = None
    L.hgt_db_close.argtypes = [ctypes.c_void_p]
    L.hgt_db_n_genes.restype = ctypes.c_int32
    L.hgt_db_n_genes.argtypes = [ctypes.c_void_p]
    L.hgt_db_sizes.restype = ctypes.c_int
    L.hgt_db_sizes.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    L.hgt_db_ints.restype = ctypes.c_int
    L.hgt_db_ints.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]
    L.hgt_db_text.restype = ctypes.c_int
    L.hgt_db_text.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p),
                              ctypes.POINTER(ctypes.c_size_t)]
    h = ctypes.c_void_p()
    rc = L.hgt_db_open(prefix.encode(), ctypes.byref(h))
    if rc == _lib.HGT_ERR_ARG and "index files missing" in _lib.last_error():
        raise SystemExit(_lib.last_error())
    _lib.check(rc)

    def ints(g, what, n):
        out = np.zeros(max(n, 1), np.int64)
        _lib.check(L.hgt_db_ints(h, g, _DB[what], out.ctypes.data_as(ctypes.c_void_p), n))
        return out[:n].tolist()

    def text(g, what, strip=True):
        p, n = ctypes.c_char_p(), ctypes.c_size_t(0)
        _lib.check(L.hgt_db_text(h, g, _DB[what], ctypes.byref(p), ctypes.byref(n)))
        s = ctypes.string_at(p, n.value).decode()
        if what == "BACKBONE":
            return s
        return s.split("\n")[:-1] if s else []

    try:
        refGenes, refGene_loci, Vars, Var_list, Links, Genes = {}, {}, {}, {}, {}, {}
        alleles, partial, haplotypes, index_vars = [], set(), {}, {}
        for g in range(L.hgt_db_n_genes(h)):
            sz = (ctypes.c_int64 * 8)()
            _lib.check(L.hgt_db_sizes(h, g, sz))
            n_seq, n_ex, n_var, n_al, n_pa, n_hap, left, right = list(sz)
            gene, name, chrom, _strand = text(g, "GENE")
            ex = ints(g, "EXONS", 3 * n_ex)
            exons = [[ex[3 * k], ex[3 * k + 1]] for k in range(n_ex)]
            primary = [[ex[3 * k], ex[3 * k + 1]] for k in range(n_ex) if ex[3 * k + 2]]
            refGenes[gene] = name
            refGene_loci[gene] = [name, chrom, left, right, exons, primary]
            Genes[gene] = {name: text(g, "BACKBONE")}
            ids, data = text(g, "VAR_ID"), text(g, "VAR_DATA")
            types, pos, in_idx = ints(g, "VAR_TYPE", n_var), ints(g, "VAR_POS", n_var), ints(g, "VAR_IN_INDEX", n_var)
            Vars[gene] = {ids[k]: [_TYPES[types[k]], pos[k], data[k]] for k in range(n_var)}
            Var_list[gene] = [[pos[k], ids[k]] for k in range(n_var)]
            for k, line in enumerate(text(g, "VAR_LINKS")):
                if line:
                    Links[ids[k]] = line.split(" ")
            alleles += text(g, "ALLELES")
            partial |= set(text(g, "PARTIAL"))
            if n_hap:
                rng, hid, hv = ints(g, "HAP_RANGE", 2 * n_hap), text(g, "HAP_ID"), text(g, "HAP_VARS")
                haplotypes[gene] = [[hid[k], rng[2 * k], rng[2 * k + 1], [v for v in hv[k].split(",") if v]] for k in range(n_hap)]
            if any(in_idx):
                index_vars[gene] = {ids[k] for k in range(n_var) if in_idx[k]}
    finally:
        L.hgt_db_close(h)
    Gene_names, Gene_lengths = build_genes(Genes, Vars, Var_list, Links, alleles)
    out = dict(refGenes=refGenes, refGene_loci=refGene_loci, Vars=Vars, Var_list=Var_list, Links=Links, Genes=Genes,
               Gene_names=Gene_names, Gene_lengths=Gene_lengths, partial_alleles=partial)
    if haplotypes:
        out["haplotypes"] = haplotypes
    if index_vars:
        out["index_vars"] = index_vars
    return out


def copy_database_native(prefix_in, prefix_out):
    """Read the database files with the native reader and write them back with the native writer (hgt_db_open ->
    hgt_db_write): the canonical form of the same tables (variants in Var_list order, 60-column FASTA)."""
    import ctypes

    from . import _lib
    L = _lib.lib()
    L.hgt_db_open.restype = ctypes.c_int
    L.hgt_db_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    L.hgt_db_write.restype = ctypes.c_int
    L.hgt_db_write.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.hgt_db_close.restype = None
    L.hgt_db_close.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p()
    _lib.check(L.hgt_db_open(prefix_in.encode(), ctypes.byref(h)))
    try:
        _lib.check(L.hgt_db_write(h, prefix_out.encode()))
    finally:
        L.hgt_db_close(h)
