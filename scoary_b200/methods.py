"""Host-side mirror of the reference's interface for the accelerated path.

Same function names, argument meaning, return structures, error behaviour
(sys.exit(message)) and output files as scoary/methods.py, so the call sites in
`main` read the same:

    Setup_results(genedic, traitsdic, collapse)            scoary/methods.py:757
    StoreResults(...) / StoreTraitResult(...)              :987 / :1003
    PairWiseComparisons((domain, argdict))                 :1208
    ConvertUPGMAtoPhyloTree(tree, GTC)                     :1386
    Permute(tree, GTC, permutations, cutoffs)              :1314

but every per-gene number on the path -- 2x2 tables, Fisher p, pair counts,
permutation hit counts -- comes from libscoary_b200.so on the GPU.  What stays
on the host is what SURVEY.md 8(b) leaves there: CSV parsing and bit packing,
collapse grouping, Bonferroni / Benjamini-Hochberg, the two binomial tests per
reported gene (SciPy, as in the reference), sorting, filtering and CSV writing.
There is no CPU fallback for the GPU part.
"""
import argparse
import csv
import logging
import operator
import os
import sys
import time
from collections.abc import Mapping

import numpy as np

from . import __version__
from . import distributed as dist
from . import engine as eng
from . import tree as treemod

log = logging.getLogger("scoary_b200")
log.setLevel(logging.DEBUG)

MISSING = ("NA", "-", ".", " ", "")
_ENGINE = None
PERMUTATION_SEED = 0x5C0A27B200   # fixed default: runs are repeatable (the reference's are not)


def get_engine():
    """The process-wide engine (one GPU per process: LOCAL_RANK, else device 0)."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = eng.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _ENGINE


# ============================================================================ inputs
class GeneTable(Mapping):
    """Gene presence/absence table: the packed form of the reference's `genedic`
    (dict gene -> {isolate: 0/1, "Non-unique Gene name", "Annotation", "<col>_name"}).
    Behaves like that dict for readers; the data live as uint64 bitset rows [G][W] (what
    sb_set_genes takes) and are unpacked to uint8 [G][N] only on demand (`.matrix`)."""

    def __init__(self, names, nugn, annotation, strains, matrix=None, extra=None, bits=None):
        self.names = list(names)
        self.nugn = list(nugn)
        self.annotation = list(annotation)
        self.strains = list(strains)
        self.extra = extra or {}
        if bits is None:
            m = np.ascontiguousarray(matrix, dtype=np.uint8).reshape(len(self.names), len(self.strains))
            bits = eng.pack_rows(m) if len(self.names) else np.zeros((0, eng.words_for(len(self.strains))), np.uint64)
        self.bits = np.ascontiguousarray(bits, dtype=np.uint64)
        self.all_bits = self.bits        # every line of the file, duplicates included: what the tree is built from
        self._matrix = None
        # the reference's dict keeps the LAST row of a duplicated identifier at the position of
        # its FIRST occurrence (methods.py:450,462)
        last = dict(zip(self.names, range(len(self.names))))
        unique = len(last) == len(self.names)
        if not unique:
            first = dict(zip(reversed(self.names), range(len(self.names) - 1, -1, -1)))
            order = sorted(last, key=lambda n: first[n])
            rows = [last[n] for n in order]
            self.names = order
            self.nugn = [self.nugn[r] for r in rows]
            self.annotation = [self.annotation[r] for r in rows]
            self.bits = np.ascontiguousarray(self.bits[rows])
            self.extra = {k: [v[r] for r in rows] for k, v in self.extra.items()}
        self.index = last if unique else dict(zip(self.names, range(len(self.names))))
        self.col = {s: j for j, s in enumerate(self.strains)}

    @property
    def matrix(self):
        """uint8 [G][N], unpacked on first use"""
        if self._matrix is None:
            self._matrix = self.rows_matrix(np.arange(len(self.names)))
        return self._matrix

    def rows_matrix(self, rows):
        """uint8 [len(rows)][N] for a subset of gene rows"""
        b = np.ascontiguousarray(self.bits[np.asarray(rows, dtype=np.int64)])
        by = b.view(np.uint8).reshape(b.shape[0], -1)
        return np.unpackbits(by, axis=1, bitorder="little")[:, :len(self.strains)]

    @classmethod
    def from_dict(cls, genedic, strains=None):
        """Reference-style dict of dicts -> GeneTable."""
        if isinstance(genedic, GeneTable):
            return genedic
        names = list(genedic.keys())
        if strains is None:
            first = genedic[names[0]] if names else {}
            strains = [k for k, v in first.items() if isinstance(v, (int, np.integer)) and not isinstance(v, bool)]
        m = np.zeros((len(names), len(strains)), dtype=np.uint8)
        for i, g in enumerate(names):
            row = genedic[g]
            m[i] = [row[s] for s in strains]
        extra = {}
        if names:
            for k in genedic[names[0]]:
                if isinstance(k, str) and k.endswith("_name"):
                    extra[k] = [str(genedic[g].get(k, "")) for g in names]
        return cls(names, [genedic[g].get("Non-unique Gene name", "") for g in names],
                   [genedic[g].get("Annotation", "") for g in names], strains, m, extra)

    def __len__(self):
        return len(self.names)

    def __iter__(self):
        return iter(self.names)

    def __getitem__(self, gene):
        i = self.index[gene]
        row = {"Non-unique Gene name": self.nugn[i], "Annotation": self.annotation[i]}
        row.update(zip(self.strains, self.rows_matrix([i])[0].tolist()))
        for k, v in self.extra.items():
            row[k] = v[i]
        return row


ROARY_COLUMNS = ["Gene", "Non-unique Gene name", "Annotation", "No. isolates", "No. sequences",
                 "Avg sequences per isolate", "Genome Fragment", "Order within Fragment", "Accessory Fragment",
                 "Accessory Order with Fragment", "QC", "Min group size nuc", "Max group size nuc",
                 "Avg group size nuc", "Order within fragment", "Genome fragment", "Accessory fragment"]


def Csv_to_dic_Roary(genefile, delimiter, grabcols, startcol=14, allowed_isolates=None, writereducedset=False,
                     time="", outdir="./"):
    """Gene presence/absence CSV -> tables (scoary/methods.py:335-508).

    Returns the reference's keys: "Roarydic" (a GeneTable), "Zero_ones_matrix"
    (isolates x variable genes, here a uint8 array), "Strains", "Extracols",
    "Firstcolnames".  A cell counts as present unless it is "", "0" or "-"."""
    opened = None
    if writereducedset:
        opened = open(ReduceSet(genefile, delimiter, grabcols, startcol, allowed_isolates, time, outdir), "r")
        genefile = opened
    import io
    path = getattr(genefile, "name", None)
    raw = getattr(getattr(genefile, "buffer", None), "raw", None)
    native = (isinstance(path, str) and isinstance(raw, io.FileIO) and os.path.isfile(path)     # a plain file on disk
              and os.environ.get("SCOARY_B200_PY_CSV") != "1")
    rdr = csv.reader(genefile, skipinitialspace=True, delimiter=delimiter)
    header = next(rdr)
    if grabcols == [-999]:
        grabcols = list(range(3, len(header)))
    if startcol >= len(header):
        sys.exit("The startcol (-s) you have specified does not seem to correspond to any column in your gene "
                 "presence/absence file.")
    strains = header[startcol:]
    extracolstoprint = [header[c] for c in grabcols]
    roaryfile = header[0:3] == ROARY_COLUMNS[0:3]
    if roaryfile and strains[0] in ROARY_COLUMNS:
        guess = next((startcol + c for c in range(len(strains)) if strains[c] not in ROARY_COLUMNS), startcol)
        log.error("ERROR: Make sure you have set the -s parameter correctly. You are running with -s %s. This "
                  "correponds to the column %s. If this is not an isolate, Scoary might crash or produce strange "
                  "results. Scoary thinks you should have run with -s %s instead" % (startcol + 1, strains[0], guess + 1))
    if roaryfile:
        before = header[:startcol][::-1]
        censored = []
        for name in before:
            if name not in ROARY_COLUMNS:
                censored.append(name)
            else:
                if censored:
                    log.error("ERROR: Make sure you have set the -s parameter correctly. You are running with -s %s. "
                              "Scoary thinks you should have used %s. This excludes the following, which Scoary "
                              "thinks are isolates: %s" % (startcol + 1, startcol - len(censored) + 1, ", ".join(censored)))
                break
    keep_cols = [c for c, s in enumerate(strains) if allowed_isolates is None or s in allowed_isolates]
    strain_names_allowed = [strains[c] for c in keep_cols]
    if roaryfile:
        try:
            genecol, nugcol, anncol = (header.index("Gene"), header.index("Non-unique Gene name"),
                                       header.index("Annotation"))
        except ValueError:
            log.error("ERROR: Could not properly detect the correct names for all columns in the ROARY table.")
            genecol, nugcol, anncol = 0, 1, 2
        firstcolnames = ["Gene", "Non-unique Gene name", "Annotation"]
    else:
        genecol, nugcol, anncol = 0, 1, 2
        firstcolnames = header[0:3]
    src_cols = [startcol + c for c in keep_cols]
    table = None
    if native:
        table = _native_gene_table(path, delimiter, roaryfile, genecol, nugcol, anncol, grabcols, header, src_cols,
                                   strain_names_allowed)
    if table is not None:        # None: ragged rows -- the csv-module loop below is the semantic reference for those
        if opened:
            opened.close()
        s = _popcount_rows(table.bits)
        variable = (s > 0) & (s < len(strain_names_allowed))
        # small tables: the matrix itself; large ones: the same cells, unpacked only if somebody asks for them
        zero_ones = LazyZeroOnes(table, np.flatnonzero(variable))
        if int(variable.sum()) * len(strain_names_allowed) <= 200_000_000:
            zero_ones = np.asarray(zero_ones)
        return {"Roarydic": table, "Zero_ones_matrix": zero_ones, "Strains": strain_names_allowed,
                "Extracols": extracolstoprint, "Firstcolnames": firstcolnames}
    names, nugn, ann, rows = [], [], [], []
    extra = {header[c] + "_name": [] for c in grabcols}
    absent = ("", "0", "-")
    for q in rdr:
        try:
            ident = q[genecol] if roaryfile else q[genecol] + "_|_" + q[nugcol] + "_|_" + q[anncol]
            nug, an = q[nugcol], q[anncol]
            row = np.fromiter((q[c] not in absent for c in src_cols), dtype=np.uint8, count=len(src_cols))
        except IndexError:
            sys.exit("CRITICAL: Could not read gene presence absence file. Verify that this file is a proper Roary "
                     "file using the specified delimiter (default is ',').")
        names.append(ident)
        nugn.append(nug)
        ann.append(an)
        rows.append(row)
        for c in grabcols:
            extra[header[c] + "_name"].append(q[c])
    if opened:
        opened.close()
    matrix = np.vstack(rows) if rows else np.zeros((0, len(src_cols)), dtype=np.uint8)
    table = GeneTable(names, nugn, ann, strain_names_allowed, matrix, extra)
    s = matrix.sum(axis=1)                                            # every line counts, duplicates included (:496)
    variable = (s > 0) & (s < matrix.shape[1])
    zero_ones = np.ascontiguousarray(matrix[variable].T)              # isolates x variable genes
    return {"Roarydic": table, "Zero_ones_matrix": zero_ones, "Strains": strain_names_allowed,
            "Extracols": extracolstoprint, "Firstcolnames": firstcolnames}


class LazyZeroOnes:
    """The reference's Zero_ones_matrix (isolates x variable genes, uint8; scoary/methods.py:496-497, the input of
    CreateTriangularDistanceMatrix) for tables too large to unpack up front: shape and len are known at once, the
    cells are unpacked from the bitset rows on first use (np.asarray(m), m[i], iteration)."""

    def __init__(self, table, rows):
        self._table, self._rows, self._m = table, np.asarray(rows, dtype=np.int64), None
        self.shape = (len(table.strains), len(self._rows))
        self.dtype = np.dtype(np.uint8)

    def _cells(self):
        if self._m is None:
            self._m = np.ascontiguousarray(self._table.rows_matrix(self._rows).T)
        return self._m

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        m = self._cells()
        return m if dtype is None else m.astype(dtype)

    def __getitem__(self, key):
        return self._cells()[key]

    def __iter__(self):
        return iter(self._cells())


def _popcount_rows(bits):
    b = np.ascontiguousarray(bits, dtype=np.uint64)
    if not b.size:
        return np.zeros(b.shape[0], np.int64)
    if hasattr(np, "bitwise_count"):                       # NumPy >= 2.0
        return np.bitwise_count(b).sum(axis=1, dtype=np.int64)
    return np.unpackbits(b.view(np.uint8), axis=1).sum(axis=1, dtype=np.int64)


def _native_gene_table(path, delimiter, roaryfile, genecol, nugcol, anncol, grabcols, header, src_cols, strains):
    """The row loop of Csv_to_dic_Roary (methods.py:445-497) through libscoary_b200's native
    packer (sb_csv_row_starts / sb_csv_pack_rows): presence bits are packed straight into the
    uint64 rows the GPU takes; only the few text fields are sliced out here."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    import mmap
    with open(path, "rb") as fh:
        try:        # mapped, not read: the parallel scan below is then also what pulls the file in
            buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            view = np.frombuffer(buf, dtype=np.uint8)
            pbuf = ctypes.c_char_p(view.ctypes.data)
        except (ValueError, OSError):       # empty file, or a file system without mmap
            buf = fh.read()
            pbuf = buf
    delim = delimiter.encode()[:1]
    found = ctypes.c_void_p()
    n = lib.sb_csv_scan_rows(pbuf, len(buf), delim, ctypes.byref(found), None)      # one pass, in parallel pieces
    if n < 0:
        sys.exit("CRITICAL: Could not read gene presence absence file.")
    starts = np.empty(max(n, 1), dtype=np.int64)
    if n:
        ctypes.memmove(starts.ctypes.data, found.value, 8 * n)
    lib.sb_csv_free(found)
    W = eng.words_for(len(src_cols))
    bits = np.empty((n, W), dtype=np.uint64)
    lead = sorted(set([genecol, nugcol, anncol] + list(grabcols)))
    lead_arr = np.asarray(lead, dtype=np.int32)
    keep_arr = np.asarray(src_cols, dtype=np.int32)
    ranges = np.empty((n, len(lead), 2), dtype=np.int64)
    nfields = np.empty(max(n, 1), dtype=np.int32)
    rc = lib.sb_csv_pack_rows(pbuf, len(buf), delim, starts.ctypes.data_as(ctypes.c_void_p), n,
                              keep_arr.ctypes.data_as(ctypes.c_void_p), len(src_cols),
                              bits.ctypes.data_as(ctypes.c_void_p), W, lead_arr.ctypes.data_as(ctypes.c_void_p),
                              len(lead), ranges.ctypes.data_as(ctypes.c_void_p), nfields.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        sys.exit("CRITICAL: Could not read gene presence absence file. Verify that this file is a proper Roary "
                 "file using the specified delimiter (default is ',').")
    if n and np.any(nfields[:n] != len(header)):
        return None              # rows wider or narrower than the header: leave them to the csv module
    slot = {c: k for k, c in enumerate(lead)}

    def field(r, c):
        b, e = int(ranges[r, slot[c], 0]), int(ranges[r, slot[c], 1])
        if b < 0:      # escaped quotes or text after a closing quote: let the csv module unescape it
            raw = buf[-b - 1 - 1:e].decode("utf-8", "replace")
            return next(csv.reader([raw], skipinitialspace=True, delimiter=delimiter))[0]
        return buf[b:e].decode("utf-8", "replace")

    def column(c):
        """field c of every row: gathered natively into one blob, decoded once, sliced n times"""
        k = slot[c]
        offs = np.empty(n + 1, dtype=np.int64)
        rp, op = ranges.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p)
        total = lib.sb_csv_gather_fields(pbuf, rp, n, len(lead), k, None, 0, op)
        blob = np.empty(max(int(total), 1), dtype=np.uint8)
        if total < 0 or lib.sb_csv_gather_fields(pbuf, rp, n, len(lead), k, blob.ctypes.data_as(ctypes.c_void_p),
                                                 int(total), op) != total:
            return [field(r, c) for r in range(n)]
        raw, o = blob[:total].tobytes(), offs.tolist()
        if raw.isascii():
            text = raw.decode("ascii")
            col = [text[o[r]:o[r + 1]] for r in range(n)]
        else:               # byte offsets are not character offsets
            col = [raw[o[r]:o[r + 1]].decode("utf-8", "replace") for r in range(n)]
        for r in np.flatnonzero(ranges[:n, k, 0] < 0).tolist():      # escaped quotes: the csv module unescapes them
            col[r] = field(r, c)
        return col

    gene, nug, ann = column(genecol), column(nugcol), column(anncol)
    names = gene if roaryfile else [g + "_|_" + u + "_|_" + a for g, u, a in zip(gene, nug, ann)]
    extra = {header[c] + "_name": column(c) for c in grabcols}
    return GeneTable(names, nug, ann, strains, extra=extra, bits=bits)


def ReduceSet(genefile, delimiter, grabcols, startcol=14, allowed_isolates=None, time="", outdir="./"):
    """-w: write the gene table restricted to the allowed isolates (scoary/methods.py:510-544)."""
    rdr = csv.reader(genefile, skipinitialspace=True, delimiter=delimiter)
    header = next(rdr)
    keep = list(range(startcol)) + [c for c in range(len(header)) if header[c] in allowed_isolates]
    log.info("Writing gene presence absence file for the reduced set of isolates")
    fname = "%sgene_presence_absence_reduced%s.csv" % (outdir, time)
    with open(fname, "w") as out:
        w = csv.writer(out, delimiter=delimiter)
        w.writerow([header[a] for a in keep])
        for r in rdr:
            w.writerow(tuple(r[a] for a in keep))
    log.info("Finished writing reduced gene presence absence list to file %s" % fname)
    return fname


def Csv_to_dic(csvfile, delimiter, allowed_isolates, strains):
    """Traits CSV -> ({trait: {isolate: "0"/"1"}}, Prunedic) (scoary/methods.py:546-614).
    Missing values are dropped per trait and listed in Prunedic[trait] (+ a trailing None)."""
    tab = list(zip(*csv.reader(csvfile, delimiter=delimiter)))
    if len(tab) < 2:
        sys.exit("Please check that your traits file is formatted properly and contains at least one trait")
    r, prunedic = {}, {}
    for k in range(1, len(tab)):
        p = dict(zip(tab[0], tab[k]))
        if "" in p:
            name = p.pop("")
        elif "Name" in p:
            name = p.pop("Name")
        else:
            sys.exit("Make sure the top-left cell in the traits file is either empty or 'Name'. Do not include "
                     "empty rows")
        if allowed_isolates is not None:
            p = {s: v for s, v in p.items() if s in allowed_isolates}
        allowed_values = ["0", "1", "NA", ".", "-", " ", ""]
        if not all(v in allowed_values for v in p.values()):
            sys.exit("Unrecognized character found in trait file. Allowed values (no commas): %s"
                     % ",".join(allowed_values))
        missing = [s for s, v in p.items() if v in MISSING]
        if missing:
            log.warning("WARNING: Some isolates have missing values for trait %s. Missing-value isolates will not "
                        "be counted in association analysis towards this trait." % name)
        prunedic[name] = list(missing)
        if not all(s in p for s in strains):
            log.error("ERROR: Some isolates in your gene presence absence file were not represented in your traits "
                      "file. These will count as MISSING data and will not be included.")
            prunedic[name] += [s for s in strains if s not in p and s not in prunedic[name]]
        r[name] = {s: v for s, v in p.items() if v not in MISSING}
        prunedic[name] += [None]
    return r, prunedic


# ============================================================================ tree building (host; off the hot path)
def upgma(table):
    """UPGMA tree (nested lists of isolate names) of the isolates of a GeneTable, built on the
    GPU (sb_upgma): relative Hamming distances over the variable genes, then the reference's
    merge order -- CreateTriangularDistanceMatrix -> PopulateQuadTreeWithDistances -> upgma
    (scoary/methods.py:619-707) with the QuadTree argmin tie-break (classes.py:155-196)."""
    if len(table.strains) < 2:
        sys.exit("Need at least two isolates to build a tree")
    e = get_engine()
    # the reference's distance matrix counts every line of the file, also those a later line with the same
    # identifier replaces in the gene dictionary (methods.py:496 vs :458-463)
    e.set_genes(table.all_bits, len(table.strains))
    return treemod.from_merges(table.strains, e.upgma())


def PruneForMissing(tree, Prunedic):
    """scoary/methods.py:709-739"""
    return treemod.prune(tree, Prunedic)


def StoreUPGMAtreeToFile(upgmatree, outdir, time=""):
    """scoary/methods.py:741-752"""
    fname = str(outdir + ("Tree%s.nwk" % time))
    with open(fname, "w") as fh:
        fh.write(treemod.to_scoary_newick(upgmatree))
    log.info("Wrote the UPGMA tree to file: %s" % fname)


def ReadTreeFromFile(path):
    """Custom tree (-n), nwkhandler.ReadTreeFromFile (scoary/nwkhandler.py:10-22).  The reference uses ete3, which is
    not installable here; treemod.from_newick reads the same files (branch lengths, support values / internal labels,
    quoted or bare names) and resolves polytomies the way ete3's resolve_polytomy(recursive=True) does."""
    try:
        with open(path) as fh:
            nested = treemod.from_newick(fh.read())
    except (OSError, ValueError) as e:
        sys.exit("Corrupted or non-existing custom tree file? %s" % e)
    return nested, treemod.leaves(nested)


# ============================================================================ per-gene statistics (A1-A3)
class _TraitGTC(Mapping):
    """GTC[trait]: gene -> {isolate: "AB"|"Ab"|"aB"|"ab"}, materialised on demand
    (the reference builds all G x N strings up front, methods.py:939-965)."""

    def __init__(self, table, isolates, labels, row_of):
        self.table, self.isolates, self.labels, self.row_of = table, isolates, labels, row_of
        self.cols = np.asarray([table.col[s] for s in isolates], dtype=np.int64)

    def __len__(self):
        return len(self.row_of)

    def __iter__(self):
        return iter(self.row_of)

    def __getitem__(self, gene):
        g = self.table.rows_matrix([self.row_of[gene]])[0][self.cols]
        return {s: ("A" if gi else "a") + ("B" if ti else "b") for s, gi, ti in zip(self.isolates, g, self.labels)}

    def gene_bits(self, genes, isolates):
        cols = np.asarray([self.table.col[s] for s in isolates], dtype=np.int64)
        rows = np.asarray([self.row_of[g] for g in genes], dtype=np.int64)
        return self.table.rows_matrix(rows)[:, cols]

    def trait_bits(self, isolates):
        lab = dict(zip(self.isolates, self.labels))
        return np.asarray([lab[s] for s in isolates], dtype=np.uint8)


def _trait_vector(table, trait_values):
    """{isolate: "0"/"1"} -> int8 [N] over the table's columns (-1 = not in the trait)."""
    vec = np.full(len(table.strains), -1, dtype=np.int8)
    for s, v in trait_values.items():
        if s not in table.col:
            log.critical("CRITICAL: Could not find %s in the genes file." % str(s))
            sys.exit("Make sure strains are named the same in your traits file as in your gene presence/absence "
                     "file")
        if v in ("NA", "-", "."):
            continue
        try:
            iv = int(v)
        except ValueError:
            iv = -9
        if iv not in (0, 1):
            sys.exit("There was a problem with comparing your traits and gene presence/absence files. Make sure you "
                     "have formatted the traits file to specification and only use 1s and 0s, as well as NA, - or . "
                     "for missing data. Also make sure the Roary file contains empty cells for non-present genes and "
                     "non-empty text cells for present genes.")
        vec[table.col[s]] = iv
    return vec


def benjamini_hochberg(p_sorted, number_of_tests):
    """Step-up BH exactly as scoary/methods.py:903-919: the least significant entry keeps
    its p; an entry tied with its less significant neighbour inherits that neighbour's value."""
    p = np.asarray(p_sorted, dtype=np.float64)
    n = len(p)
    if n == 0:
        return p
    vals = (p * number_of_tests) / np.arange(1.0, n + 1.0)
    vals[-1] = p[-1]
    tie = np.zeros(n, dtype=bool)
    tie[:-1] = p[:-1] == p[1:]
    vals[tie] = np.inf
    return np.minimum.accumulate(vals[::-1])[::-1]


DEVICE_EPILOGUE_MIN = 16_384       # rows from which the adjusted p-values and the p-sort run on the GPU (50 000 rows:
                                   # 0.5 ms with the copies, against 5 ms in NumPy)
DEVICE_BINOMIAL_MIN = 4096         # genes from which PairWiseComparisons' binomial tests run on the GPU


def adjust_pvalues(e, p, keep, n_tests=0, device_min=None):
    """Bonferroni, Benjamini-Hochberg and the p-sort of one trait (scoary/methods.py:900-925, :1448-1454) for the genes
    flagged in `keep`: -> (order: tested genes by ascending p, ties in gene order; bonferroni [n]; bh [n]; NaN where
    not tested).  Large vectors go through the device epilogue (sb_adjust_pvalues, csrc/epilogue.cuh), small ones
    through NumPy; the two are the same operations in the same order and agree bit for bit (tests/test_gpu_parity.py)."""
    p = np.asarray(p, dtype=np.float64)
    keep = np.asarray(keep, dtype=bool)
    device_min = DEVICE_EPILOGUE_MIN if device_min is None else device_min
    if len(p) >= device_min and hasattr(e, "adjust_pvalues"):
        return e.adjust_pvalues(p, keep, n_tests)
    idx = np.flatnonzero(keep)
    m = n_tests or len(idx)
    order = idx[np.argsort(p[idx], kind="stable")]
    bonf = np.full(len(p), np.nan)
    bh = np.full(len(p), np.nan)
    bonf[idx] = np.minimum(p[idx] * m, 1.0)
    bh[order] = np.minimum(benjamini_hochberg(p[order], m), 1.0)
    return order, bonf, bh


def Setup_results(genedic, traitsdic, collapse):
    """Counting, Fisher's exact test and multiple-testing adjustment for every trait
    (scoary/methods.py:757-928); the counting and Fisher part runs on the GPU
    (sb_contingency_fisher).  Returns {"Results": ..., "Gene_trait_combinations": ...}."""
    table = GeneTable.from_dict(genedic)
    e = get_engine()
    world, rank = dist.world_rank()
    G = table.bits.shape[0]
    lo, hi = dist.shard_bounds(G, world)[rank]           # N > 1: this rank's contiguous block of genes
    if hi > lo:
        e.set_genes(table.bits[lo:hi], len(table.strains))
    all_traits, gtc = {}, {}
    trait_names = list(traitsdic)
    TRAITS_PER_PASS = 8       # sb_contingency_fisher_multi reads every gene row once for this many traits
    staged = {}               # trait -> (counts, p, hashes) of the pass it was computed in
    for t_idx, trait in enumerate(trait_names):
        if trait not in staged:
            chunk = trait_names[t_idx:t_idx + TRAITS_PER_PASS]
            if hi > lo:       # a rank without genes (fewer genes than GPUs) only takes part in the gather
                for slot, name in enumerate(chunk):
                    e.set_trait_vector(slot, _trait_vector(table, traitsdic[name]))
                c_all, p_all, h_all = e.contingency_fisher_multi(0, len(chunk), want_hash=bool(collapse))
            else:
                c_all = np.zeros((len(chunk), 0, 4), np.int32)
                p_all = np.zeros((len(chunk), 0), np.float64)
                h_all = np.zeros((len(chunk), 0, 2), np.uint64) if collapse else None
            staged = {name: (c_all[k], p_all[k], h_all[k] if collapse else None) for k, name in enumerate(chunk)}
        log.info("Gene-wise counting and Fisher's exact tests for trait: %s" % str(trait))
        counts, pvals, hashes = staged[trait]
        if world > 1:                                     # one all-gather; the rest is the same on every rank
            rec = np.zeros((hi - lo, 10), dtype=np.int32)
            rec[:, 0:4] = counts
            rec[:, 4:6] = np.ascontiguousarray(pvals, dtype=np.float64).view(np.int32).reshape(-1, 2)
            if collapse:
                rec[:, 6:10] = np.ascontiguousarray(hashes, dtype=np.uint64).view(np.int32).reshape(-1, 4)
            rec = dist.gather_blocks(rec, G)
            counts = rec[:, 0:4]
            pvals = np.ascontiguousarray(rec[:, 4:6]).view(np.float64).reshape(-1)
            hashes = np.ascontiguousarray(rec[:, 6:10]).view(np.uint64).reshape(-1, 2)
        tpgp, tngp, tpgn, tngn = (counts[:, k].astype(np.int64) for k in range(4))
        keep = ((tpgp + tngp) > 0) & ((tpgn + tngn) > 0)          # methods.py:804-814
        number_of_tests = int(keep.sum())
        num_pos, num_neg = tpgp + tpgn, tngp + tngn
        with np.errstate(divide="ignore", invalid="ignore"):
            sens = np.where(num_pos > 0, tpgp.astype(np.float64) / np.maximum(num_pos, 1) * 100, 0.0)
            spes = np.where(num_neg > 0, tngn.astype(np.float64) / np.maximum(num_neg, 1) * 100, 0.0)
            odds = np.where((tngp > 0) & (tpgn > 0), (tpgp * tngn) / np.maximum(tngp * tpgn, 1).astype(np.float64),
                            np.inf)
            # a table with an empty row or column has no odds ratio in SciPy (nan, p = 1): only the trait margins
            # can be empty here, genes present in all or no isolates were skipped above
            odds = np.where((num_pos == 0) | (num_neg == 0), np.nan, odds)
        res = {}
        row_of = {}
        idx = np.flatnonzero(keep)
        names = table.names
        cols = [x[idx].tolist() for x in (tpgp, tngp, tpgn, tngn, sens, spes, odds, np.asarray(pvals, np.float64))]
        if not collapse:
            # every tested gene is its own row: the adjusted p-values are array operations (methods.py:900-925)
            _, bonf_all, bh_all = adjust_pvalues(e, pvals, keep, number_of_tests)
            log.info("Adding p-values adjusted for testing multiple hypotheses")
            nugn, annotation = table.nugn, table.annotation
            for i, a, b, c, d, se, sp, od, p_v, b_p, bh_p in zip(idx.tolist(), *cols, bonf_all[idx].tolist(),
                                                                 bh_all[idx].tolist()):
                gene = names[i]
                res[gene] = {"NUGN": nugn[i], "Annotation": annotation[i], "tpgp": a, "tngp": b, "tpgn": c, "tngn": d,
                             "sens": se, "spes": sp, "OR": od, "p_v": p_v, "B_p": b_p, "BH_p": bh_p}
                row_of[gene] = i
        else:
            p_list_names, p_list_vals = [], []
            owner = {}      # pattern hash -> current (possibly merged) name
            for i, a, b, c, d, se, sp, od, p_v in zip(idx.tolist(), *cols):
                gene = names[i]
                entry = {"NUGN": table.nugn[i], "Annotation": table.annotation[i], "tpgp": a, "tngp": b, "tpgn": c,
                         "tngn": d, "sens": se, "spes": sp, "OR": od, "p_v": p_v}
                key = (int(hashes[i, 0]), int(hashes[i, 1]))
                prev = owner.get(key)
                if prev is not None:                                   # methods.py:823-835, :874-892
                    number_of_tests -= 1
                    newname = prev + "--" + gene
                    old = res.pop(prev)
                    entry["NUGN"] = old["NUGN"] + "--" + entry["NUGN"]
                    entry["Annotation"] = old["Annotation"] + "--" + entry["Annotation"]
                    row_of.pop(prev)
                    owner[key] = newname
                    gene = newname
                else:
                    owner[key] = gene
                res[gene] = entry
                row_of[gene] = i
                p_list_names.append(gene)
                p_list_vals.append(p_v)
            log.info("Adding p-values adjusted for testing multiple hypotheses")
            pv = np.asarray(p_list_vals, dtype=np.float64)
            order = np.argsort(pv, kind="stable")
            bh = benjamini_hochberg(pv[order], number_of_tests)
            bh_by_name = {}
            for k, o in enumerate(order.tolist()):
                bh_by_name[p_list_names[o]] = bh[k]
            for gene, entry in res.items():
                entry["B_p"] = min(entry["p_v"] * number_of_tests, 1.0)
                entry["BH_p"] = min(float(bh_by_name[gene]), 1.0)
        all_traits[trait] = res
        isolates = list(traitsdic[trait].keys())
        labels = np.asarray([1 if traitsdic[trait][s] == "1" else 0 for s in isolates], dtype=np.uint8)
        gtc[trait] = _TraitGTC(table, isolates, labels, row_of)
    return {"Results": all_traits, "Gene_trait_combinations": gtc}


# ============================================================================ pairwise comparisons + permutations (A4-A10)
_BINOM_CACHE = {}


def _binom_two_sided_many(k, n):
    """ss.binom_test(k, n, 0.5) (scoary/methods.py:1267-1275) for arrays of (k, n): host, memoised, SciPy as in
    the reference -- but ONE vectorised binomtest call over the (k, n) pairs not seen before instead of a
    1.5 ms scalar call each (elementwise the same arithmetic: tests/test_host_logic.py checks bit equality)."""
    k = np.asarray(k, dtype=np.int64).reshape(-1)
    n = np.asarray(n, dtype=np.int64).reshape(-1)
    key = k * (int(n.max(initial=0)) + 1) + n
    uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    uk, un = k[first], n[first]
    vals = np.full(len(uniq), np.nan)                   # n = 0 (no contrasting pair): binomtest has no answer -> nan
    todo = [i for i in range(len(uniq)) if un[i] > 0 and (int(uk[i]), int(un[i])) not in _BINOM_CACHE]
    if todo:
        from scipy import stats as ss
        tk, tn = uk[todo], un[todo]
        try:
            pv = np.asarray(ss.binomtest(tk, tn, 0.5).pvalue, dtype=np.float64).reshape(-1)
        except (TypeError, ValueError):                     # a SciPy whose binomtest takes scalars only
            pv = np.asarray([float(ss.binomtest(int(a), int(b), 0.5).pvalue) for a, b in zip(tk, tn)])
        for a, b, v in zip(tk.tolist(), tn.tolist(), pv.tolist()):
            _BINOM_CACHE[(a, b)] = v
    for i in range(len(uniq)):
        if un[i] > 0:
            vals[i] = _BINOM_CACHE[(int(uk[i]), int(un[i]))]
    return vals[inverse]


def _binom_two_sided(k, n):
    return float(_binom_two_sided_many([k], [n])[0])


_RMIN_CACHE = {}


def early_stop_table(P):
    """rmin[i] = least r for which the reference aborts after permutation i:
    1 - ss.binom.cdf(r, i, 0.1) < 0.05 (scoary/methods.py:1360-1361), for all i at once.

    The upper tail P[X > r], X ~ Binomial(i, 0.1), is summed directly from log-factorials (NumPy only: importing
    scipy.stats costs more than a whole C3 run on the GPU) in a window around the normal approximation of the
    threshold; an entry whose tail comes within 1e-9 of 0.05, or whose window does not bracket the threshold, is
    decided by the reference's own SciPy expression instead (none does for P <= 100 000; tests/test_host_logic.py
    compares the whole table with SciPy's)."""
    t = _RMIN_CACHE.get(P)
    if t is None:
        t = np.full(max(P, 1), np.iinfo(np.int32).max, dtype=np.int32)
        if P > 30:
            i = np.arange(30, P, dtype=np.int64)
            lf = np.concatenate([[0.0], np.cumsum(np.log(np.arange(1, P + 1, dtype=np.float64)))])
            sd = np.sqrt(0.09 * i)
            r0 = np.clip(np.floor(0.1 * i + 1.645 * sd).astype(np.int64) - 4, 0, None)       # window start
            K = int(8 + 12 * sd.max() + 8)
            k = r0[:, None] + 1 + np.arange(K, dtype=np.int64)[None, :]                       # terms k = r0 + 1 ..
            ok = k <= i[:, None]
            kk = np.where(ok, k, 0)
            logpmf = (lf[i][:, None] - lf[kk] - lf[i[:, None] - kk] + kk * np.log(0.1) + (i[:, None] - kk) * np.log(0.9))
            pmf = np.where(ok, np.exp(logpmf), 0.0)
            sf = np.cumsum(pmf[:, ::-1], axis=1)[:, ::-1]                                     # sf[:, j] = P[X > r0 + j]
            stops = sf[:, :9] < 0.05
            first = np.argmax(stops, axis=1)
            r = r0 + first
            doubt = (~stops.any(axis=1)) | (stops[:, 0] & (r0 > 0)) | (np.abs(sf[:, :9] - 0.05).min(axis=1) < 1e-9)
            if doubt.any():
                from scipy import stats as ss
                for j in np.flatnonzero(doubt):
                    rr = 0
                    while not (1 - ss.binom.cdf(rr, int(i[j]), 0.1)) < 0.05:
                        rr += 1
                    r[j] = rr
            t[30:P] = r
        _RMIN_CACHE[P] = t
    return t


def _gtc_arrays(GTC, genes, isolates):
    """gene bits [S][n] and trait bits [n] for `isolates`, from either a lazy GTC or the
    reference's dict-of-dicts of "AB"/"Ab"/"aB"/"ab" strings."""
    if isinstance(GTC, _TraitGTC):
        return GTC.gene_bits(genes, isolates), GTC.trait_bits(isolates)
    g = np.zeros((len(genes), len(isolates)), dtype=np.uint8)
    t = None
    for a, gene in enumerate(genes):
        row = GTC[gene]
        try:
            codes = [row[s] for s in isolates]
        except KeyError as ex:
            sys.exit("Isolate %s of the tree has no gene-trait combination" % ex)
        g[a] = [c[0] == "A" for c in codes]
        if t is None:
            t = np.asarray([c[-1] == "B" for c in codes], dtype=np.uint8)
    return g, t


def _walk_setup(tree, GTC, genes):
    """Load genes x tree-leaves into the engine (trait slot 0) and return the engine."""
    left, right, names = treemod.flatten(tree)
    e = get_engine()
    if isinstance(GTC, _TraitGTC):
        # the packed rows go up as they are (table column order); the tree finds its columns through leaf_to_col
        table = GTC.table
        rows = np.fromiter((GTC.row_of[g] for g in genes), dtype=np.int64, count=len(genes))
        vec = np.full(len(table.strains), -1, dtype=np.int8)
        vec[GTC.cols] = GTC.labels
        try:
            cols = np.asarray([table.col[n] for n in names], dtype=np.int32)
        except KeyError as ex:
            sys.exit("Isolate %s of the tree has no gene-trait combination" % ex)
        if np.any(vec[cols] < 0):
            sys.exit("Isolate '%s' of the tree has no gene-trait combination" % names[int(np.argmax(vec[cols] < 0))])
        e.set_genes(table.bits[rows], len(table.strains))
        e.set_trait_vector(0, vec)
        e.set_tree(0, left, right, cols)
        return e
    g, t = _gtc_arrays(GTC, genes, names)
    e.set_genes(eng.pack_rows(g), len(names))
    e.set_trait_vector(0, t.astype(np.int8))
    e.set_tree(0, left, right, np.arange(len(names), dtype=np.int32))
    return e


def ConvertUPGMAtoPhyloTree(tree, GTC):
    """Max contrasting / supporting / opposing pairs for one gene (scoary/methods.py:1386-1402,
    classes.PhyloTree); GTC = {isolate: "AB"|...}.  GPU walk (sb_pairwise)."""
    e = _walk_setup(tree, {"_": GTC}, ["_"])
    total, pro, anti = (int(x) for x in e.pairwise(0)[0])
    return {"Total": total, "Pro": pro, "Anti": anti}


def Permute(tree, GTC, permutations, cutoffs, seed=PERMUTATION_SEED):
    """Empirical p by label switching for one gene (scoary/methods.py:1314-1369), with the
    reference's sequential early stop; GPU (sb_permute)."""
    if permutations < 10:
        sys.stdout.write("Number of permutations too few. The absolute minimum is 10.")
        return None
    e = _walk_setup(tree, {"_": GTC}, ["_"])
    _, r, nd = e.permute(0, int(permutations), seed=seed, early_stop=True, rmin=early_stop_table(int(permutations)))
    return (int(r[0]) + 1.0) / (int(nd[0]) + 1.0)


def PairWiseComparisons(nestedlist):
    """(domain, {"si","tree","GTC","cutoffs","cp","perm","Trait","Threaded"}) -> {gene: {...}}
    (scoary/methods.py:1208-1312).  All genes of the domain are walked in one launch;
    the reference's "break at the first gene failing I/B/BH" is applied first, on the host,
    because those cut-offs do not depend on the walk."""
    domain, a = nestedlist[0], nestedlist[1]
    si, tree, GTC, cutoffs, perm, Trait = a["si"], a["tree"], a["GTC"], a["cutoffs"], a["perm"], a["Trait"]
    genes = []
    for genenumber in domain:
        g = si[genenumber]
        if decideifbreak(cutoffs, Trait[g]):
            break
        genes.append(g)
    out = {}
    if not genes:
        return out
    world, rank = dist.world_rank()
    mine = genes[rank::world]                            # N > 1: strided, as the reference deals out its domains
    pairs, r, nd = np.zeros((0, 3), np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32)
    if mine:
        e = _walk_setup(tree, GTC, mine)
        if perm >= 10:
            pairs, r, nd = e.permute(0, int(perm), seed=a.get("seed", PERMUTATION_SEED), early_stop=True,
                                     rmin=early_stop_table(int(perm)))
        else:
            pairs = e.pairwise(0)
    if world > 1:
        rec = np.zeros((len(mine), 5), dtype=np.int32)
        rec[:, 0:3] = pairs
        if perm >= 10:
            rec[:, 3], rec[:, 4] = r, nd
        rec = dist.gather_strided(rec, len(genes))
        pairs, r, nd = rec[:, 0:3], rec[:, 3], rec[:, 4]
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 3)
    e = get_engine()
    if len(pairs) >= DEVICE_BINOMIAL_MIN and hasattr(e, "binom_two_sided"):
        # many genes: the binomial tests run on the GPU (sb_binom_two_sided, <= 1e-13 relative to SciPy's value)
        both = e.binom_two_sided(np.concatenate([pairs[:, 1], pairs[:, 0] - pairs[:, 2]]), np.concatenate([pairs[:, 0]] * 2))
        p_pro, p_anti = both[:len(pairs)].tolist(), both[len(pairs):].tolist()
    else:   # few genes: SciPy itself, memoised -- the columns are then bit-identical to the reference's
        p_pro = _binom_two_sided_many(pairs[:, 1], pairs[:, 0]).tolist()
        p_anti = _binom_two_sided_many(pairs[:, 0] - pairs[:, 2], pairs[:, 0]).tolist()
    pairs = pairs.tolist()
    for k, g in enumerate(genes):
        total, pro, anti = pairs[k]
        best, worst = p_pro[k], p_anti[k]
        if pro < anti:                                   # names swap, methods.py:1259-1265
            best, worst = worst, best
        d = {"max_total_pairs": total, "max_propairs": pro, "max_antipairs": anti, "Pbest": best, "Pworst": worst,
             "Plowest": min(best, worst), "Pboth": max(best, worst),
             "p_v": Trait[g]["p_v"], "B_p": Trait[g]["B_p"], "BH_p": Trait[g]["BH_p"]}
        if perm >= 10:
            d["Empirical_p"] = (int(r[k]) + 1.0) / (int(nd[k]) + 1.0)
        out[g] = d
    return out


def decideifbreak(cutoffs, currentgene):
    """scoary/methods.py:1489-1508"""
    for key, field in (("I", "p_v"), ("B", "B_p"), ("BH", "BH_p")):
        if key in cutoffs and currentgene[field] > cutoffs[key]:
            return True
    return False


def SortResultsAndSetKey(genedic, key="p_v"):
    """rank -> gene, ascending by `key`, ties in dict order (scoary/methods.py:1448-1454)."""
    return dict(enumerate(sorted(genedic, key=lambda g: genedic[g][key])))


def SortResultsAndSetKeyPairwise(genedic, correctionmethod):
    """scoary/methods.py:1456-1470"""
    if "EPW" in correctionmethod:
        return SortResultsAndSetKey(genedic, "Pboth")
    if "PW" in correctionmethod:
        return SortResultsAndSetKey(genedic, "Plowest")
    sys.exit("Something went wrong when using this set of correction methods. Please report this bug")


CUT_FIELDS = {"I": "p_v", "B": "B_p", "BH": "BH_p", "PW": "Plowest", "EPW": "Pboth", "P": "Empirical_p"}


def StoreResults(Results, max_hits, cutoffs, upgmatree, GTC, Prunedic, outdir, permutations, num_threads,
                 no_pairwise, genedic, extracolstoprint, firstcolnames, time="", delimiter=","):
    """scoary/methods.py:987-1001"""
    for Trait in Results:
        sys.stdout.write("\n")
        log.info("Storing results: " + Trait)
        StoreTraitResult(Results[Trait], Trait, max_hits, cutoffs, upgmatree, GTC, Prunedic, outdir, permutations,
                         num_threads, no_pairwise, genedic, extracolstoprint, firstcolnames, time, delimiter)


def StoreTraitResult(Trait, Traitname, max_hits, cutoffs, upgmatree, GTC, Prunedic, outdir, permutations,
                     num_threads, no_pairwise, genedic, extracolstoprint, firstcolnames, time="", delimiter=","):
    """One <Trait>.results.csv (scoary/methods.py:1003-1197): same columns, quoting, ordering
    and final `<=` filter.  --threads is accepted and ignored: the walk runs on the GPU."""
    permutations = int(permutations)
    fname = outdir + Traitname + time + ".results.csv"
    with open(fname, "w") as outfile:
        sort_instructions = SortResultsAndSetKey(Trait)
        if max_hits is None:
            max_hits = len(Trait)
        num_results = min(max_hits, len(Trait))
        columns = list(firstcolnames) + ["Number_pos_present_in", "Number_neg_present_in", "Number_pos_not_present_in",
                                         "Number_neg_not_present_in", "Sensitivity", "Specificity", "Odds_ratio",
                                         "Naive_p", "Bonferroni_p", "Benjamini_H_p"]
        if not no_pairwise:
            columns += ["Max_Pairwise_comparisons", "Max_supporting_pairs", "Max_opposing_pairs",
                        "Best_pairwise_comp_p", "Worst_pairwise_comp_p"]
        if permutations >= 10:
            columns.append("Empirical_p")
        columns += list(extracolstoprint)
        outfile.write(delimiter.join('"' + c + '"' for c in columns) + "\n")
        if permutations >= 10:
            log.info("Calculating max number of contrasting pairs for each significant gene and performing %s "
                     "permutations" % str(permutations))
        elif no_pairwise:
            log.info("Skipping population structure-aware analyses.")
        else:
            log.info("Calculating max number of contrasting pairs for each nominally significant gene")
        if not no_pairwise:
            if len(Prunedic[Traitname]) > 0:
                upgmatree = PruneForMissing(upgmatree, Prunedic[Traitname])
            args = {"si": sort_instructions, "tree": upgmatree, "GTC": GTC[Traitname], "cutoffs": cutoffs,
                    "cp": CUT_FIELDS, "perm": permutations, "Trait": Trait, "Threaded": False}
            walked = PairWiseComparisons((list(range(num_results)), args))
            filtered = {}
            for g, d in walked.items():
                d.update(Trait[g])
                filtered[g] = d
            if ("I" in cutoffs) or ("B" in cutoffs) or ("BH" in cutoffs):
                sort_instructions = SortResultsAndSetKey(filtered)
            elif ("PW" in cutoffs) or ("EPW" in cutoffs):
                sort_instructions = SortResultsAndSetKeyPairwise(filtered, cutoffs)
            elif "P" in cutoffs:
                sort_instructions = SortResultsAndSetKey(filtered, key="Empirical_p")
            else:
                log.info("No filtration applied")
        else:
            filtered = Trait                             # same genes, same key: the ranking above stands
        log.info("Storing results to file")
        keys = ["tpgp", "tngp", "tpgn", "tngn", "sens", "spes", "OR", "p_v", "B_p", "BH_p"]
        if not no_pairwise:
            keys += ["max_total_pairs", "max_propairs", "max_antipairs", "Pbest", "Pworst"]
            if permutations >= 10:
                keys.append("Empirical_p")
        stats_of = operator.itemgetter(*keys)
        cuts = [(CUT_FIELDS[m], cutoffs[m]) for m in cutoffs]
        sep = '"' + delimiter + '"'                       # every cell double-quoted (methods.py:1159-1197)
        lines = []
        for x in range(min(num_results, len(filtered))):
            gene = sort_instructions[x]
            row = filtered[gene]
            if not all(row[f] <= c for f, c in cuts):
                continue
            out = gene.split("_|_") if "_|_" in gene else [gene, str(row["NUGN"]), str(row["Annotation"])]
            out += map(str, stats_of(row))
            for colname in extracolstoprint:
                parts = gene.split("--") if "--" in gene else [gene]
                out.append("--".join(str(genedic[g][colname + "_name"]) for g in parts))
            lines.append('"' + sep.join(out) + '"\n')
            if len(lines) >= 65536:
                outfile.writelines(lines)
                lines = []
        outfile.writelines(lines)


# ============================================================================ CLI
def filtrationoptions(cutoffs, collapse):
    """scoary/methods.py:1472-1487"""
    names = {"I": "Individual (Naive)", "B": "Bonferroni", "BH": "Benjamini-Hochberg",
             "PW": "Pairwise comparison (Best)", "EPW": "Pairwise comparison (Entire range)",
             "P": "Empirical p-value (permutation-based)"}
    lines = ["-- Filtration options --"] + [names[k] + ":    " + str(v) for k, v in cutoffs.items()]
    lines.append("Collapse genes:    " + str(collapse) + "\n\n")
    return lines


def grabcoltype(string):
    """--include_input_columns: '4,6-8' or ALL -> zero-based columns (scoary/methods.py:1510-1549).
    As in the reference a range a-b needs b > a and stops BEFORE b, the three identifier columns are
    dropped, and the result is list(set(...)) -- the column order of the output follows from that."""
    if string == "ALL":
        return [-999]
    if string == "":
        return []
    cols = []
    for part in string.split(","):
        if "-" in part:
            try:
                lo, hi = (int(x) for x in part.split("-")[:2])
                if not hi > lo:
                    raise ValueError(part)
            except (ValueError, IndexError):
                sys.exit("Could not understand --include_input_columns argument %s" % part)
            cols += list(range(lo, hi))
        else:
            cols.append(int(part))
    cols = [c - 1 for c in cols if c - 1 not in (0, 1, 2)]
    if not all(c > 1 for c in cols):
        sys.exit("Could not understand --include_input_columns argument. Make sure all numbers are positive and "
                 "real.")
    return list(set(cols))


def ScoaryArgumentParser(argv=None):
    """Same flags as scoary/methods.py:1551-1744."""
    p = argparse.ArgumentParser(description="scoary_b200 %s - Scoary's pan-GWAS screen with the statistics, "
                                            "pairwise comparisons and permutations on a B200" % __version__)
    i = p.add_argument_group("Input options")
    i.add_argument("-t", "--traits", help="trait table (CSV; isolates in rows, traits in columns, 1/0/NA)")
    i.add_argument("-g", "--genes", help="gene presence/absence table (Roary CSV)")
    i.add_argument("-n", "--newicktree", default=None, help="custom binary Newick tree instead of the internal UPGMA tree")
    i.add_argument("-s", "--start_col", default=15, type=int, help="1-based column where isolates start (default 15)")
    i.add_argument("--delimiter", default=",", type=str, help="cell delimiter of inputs and outputs")
    i.add_argument("-r", "--restrict_to", help="file with a comma-separated subset of isolates to analyse")
    o = p.add_argument_group("Output options")
    o.add_argument("-o", "--outdir", default="./", help="output directory")
    o.add_argument("-u", "--upgma_tree", default=False, action="store_true", help="write the UPGMA tree to Tree.nwk")
    o.add_argument("-p", "--p_value_cutoff", nargs="+", default=[0.05], type=float,
                   help="one cut-off for all correction methods, or one per method in order")
    choices = ["I", "B", "BH", "PW", "EPW", "P"]
    o.add_argument("-c", "--correction", choices=choices, nargs="*", default=["I"],
                   help="filters: I naive, B Bonferroni, BH Benjamini-Hochberg, PW best pairwise, EPW entire pairwise "
                        "range, P empirical (permutation)")
    o.add_argument("-m", "--max_hits", type=int, help="report at most this many genes per trait")
    o.add_argument("--include_input_columns", dest="grabcols", type=grabcoltype, default=[],
                   help="copy input columns to the output, e.g. 4,6,8,16-23 or ALL")
    o.add_argument("-w", "--write_reduced", default=False, action="store_true",
                   help="with -r: also write the reduced gene table")
    o.add_argument("--no-time", default=False, action="store_true", help="no timestamp in output file names")
    a = p.add_argument_group("Analysis options")
    a.add_argument("-e", "--permute", type=int, default=0, help="label-switching permutations per reported gene (>= 10)")
    a.add_argument("--no_pairwise", default=False, action="store_true", help="population-structure-naive analysis only")
    a.add_argument("--collapse", default=False, action="store_true", help="merge genes with identical patterns")
    m = p.add_argument_group("Misc options")
    m.add_argument("--threads", type=int, default=1, help="accepted for compatibility; the GPU does the work")
    m.add_argument("--seed", type=int, default=PERMUTATION_SEED, help="seed of the permutation RNG (Philox)")
    m.add_argument("--test", default=False, action="store_true", help="run on the reference's example data if present")
    m.add_argument("--citation", default=False, action="store_true", help="show citation information and exit")
    m.add_argument("--version", action="version", version=__version__)
    args = p.parse_args(argv)
    if len(args.p_value_cutoff) == 1:
        cutoffs = {c: args.p_value_cutoff[0] for c in args.correction}
    else:
        cutoffs = dict(zip(args.correction, args.p_value_cutoff))
    return args, cutoffs


CITATION = ("Scoary: Brynildsrud O, Bohlin J, Scheffer L, Eldholm V. Rapid scoring of genes in microbial pan-genome-wide "
            "association studies with Scoary. Genome Biol. 2016;17:238.  This is scoary_b200, a B200-native engine "
            "for that method.")


def _close_run(fileh, console, scratch, ok):
    log.removeHandler(fileh)
    log.removeHandler(console)
    fileh.close()
    if ok:
        dist.finish()              # barrier + leave the process group (no-op on one GPU)
    if scratch is not None:
        import shutil
        shutil.rmtree(scratch, ignore_errors=True)


def main(**kwargs):
    """`scoary -g genes.csv -t traits.csv` (scoary/methods.py:49-330) with the same files out."""
    global PERMUTATION_SEED
    if "args" in kwargs:
        args, cutoffs = kwargs["args"], kwargs["cutoffs"]
    else:
        args, cutoffs = ScoaryArgumentParser(kwargs.get("argv"))
    if args.citation:
        sys.exit(CITATION)
    if args.test:
        # --test (scoary/methods.py:69-92): the reference's example data.  This repository carries it as test fixtures
        # (tests/golden/inputs, the gene table gzipped); it is unpacked next to the results.
        ex = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
        packed = os.path.join(ex, "Gene_presence_absence.csv.gz")
        if not os.path.isfile(packed):
            sys.exit("--test needs the example data of the source checkout (tests/golden/inputs); not found at %s" % ex)
        import gzip
        import shutil
        import tempfile
        unpacked = os.path.join(tempfile.mkdtemp(prefix="scoary_b200_test_"), "Gene_presence_absence.csv")
        with gzip.open(packed, "rb") as fi, open(unpacked, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        args.correction, cutoffs = ["I", "EPW"], {"I": 0.05, "EPW": 0.05}
        args.delimiter, args.grabcols, args.max_hits, args.newicktree = ",", [], None, None
        args.genes, args.traits = unpacked, os.path.join(ex, "Tetracycline_resistance.csv")
        args.no_pairwise, args.outdir, args.permute, args.p_value_cutoff = False, "./", 0, [0.05, 0.05]
        args.restrict_to, args.start_col, args.upgma_tree, args.write_reduced = None, 15, True, False
        args.no_time, args.collapse = False, False
    PERMUTATION_SEED = getattr(args, "seed", PERMUTATION_SEED)
    starttime = time.time()
    currenttime = "" if args.no_time else time.strftime("_%d_%m_%Y_%H%M")
    scratch = None
    if dist.init_from_env() and dist.world_rank()[1] != 0:
        # one process per GPU (torchrun): every rank runs the same host code, rank 0 owns the output
        # directory; the others write their (identical) files to a scratch directory that is removed
        import tempfile
        scratch = tempfile.mkdtemp(prefix="scoary_b200_rank%d_" % dist.world_rank()[1])
        args.outdir = scratch
    if not args.outdir.endswith("/"):
        args.outdir += "/"
    os.makedirs(args.outdir, exist_ok=True)
    console = logging.StreamHandler(sys.stdout if scratch is None else open(os.devnull, "w"))
    console.setFormatter(logging.Formatter("%(message)s"))
    console.setLevel(logging.INFO)
    log.addHandler(console)
    fileh = logging.FileHandler(os.path.join(args.outdir, "scoary%s.log" % currenttime), mode="w")
    fileh.setFormatter(logging.Formatter(fmt="%(asctime)s    %(message)s", datefmt="%m/%d/%Y %I:%M:%S %p"))
    log.addHandler(fileh)
    log.info("==== Scoary started ====")
    log.info("Command: " + " ".join(sys.argv))
    try:
        if args.traits is None or args.genes is None:
            sys.exit("The following arguments are required: -t/--traits, -g/--genes")
        if args.threads <= 0:
            sys.exit("Number of threads must be positive")
        if not os.path.isfile(args.traits):
            sys.exit("Could not find the traits file: %s" % args.traits)
        if not os.path.isfile(args.genes):
            sys.exit("Could not find the gene presence absence file: %s" % args.genes)
        if args.newicktree is not None and not os.path.isfile(args.newicktree):
            sys.exit("Could not find the custom tree file: %s" % args.newicktree)
        if not all(0.0 < p <= 1.0 for p in args.p_value_cutoff):
            sys.exit("P must be between 0.0 and 1.0 or exactly 1.0")
        if len(args.delimiter) > 1:
            sys.exit("Delimiter must be a single character string. There is no support for tab.")
        if len(args.p_value_cutoff) != len(args.correction) and len(args.p_value_cutoff) != 1:
            sys.exit("You can not use more p-value cutoffs than correction methods. Either provide a single p-value "
                     "that will be applied to all correction methods, or provide exactly as many as the number of "
                     "correction methods and in corresponding sequence. e.g. -c I EPW -p 0.1 0.05 will apply an "
                     "individual p-value cutoff of 0.1 AND a pairwise comparisons p-value cutoff of 0.05.")
        if "P" in cutoffs and args.permute == 0:
            sys.exit("Cannot use empirical p-values in filtration without performing permutations. Use "
                     "'--permute X' where X is a number equal to or larger than 10")
        if args.permute < 10 and args.permute != 0:
            sys.exit("The absolute minimum number of permutations is 10 (or 0 to deactivate)")
        if "P" in cutoffs and cutoffs["P"] < (1.0 / args.permute):
            sys.exit("Permutation cutoff too low for this number of permutations")
        if args.no_pairwise:
            log.info("Performing no pairwise comparisons. Ignoring all tree related options (user tree, population "
                     "aware-correction, permutations).")
            args.permute, args.newicktree = 0, None
            for m in ("PW", "EPW", "P"):
                cutoffs.pop(m, None)
        with open(args.genes, "r") as genes, open(args.traits, "r") as traits:
            if args.restrict_to is not None:
                allowed = {iso: "all" for line in open(args.restrict_to, "r") for iso in line.rstrip().split(",")}
            else:
                allowed = None
                if args.write_reduced:
                    sys.exit("You cannot use the -w argument without specifying a subset (-r)")
            if args.genes.lower().endswith(".vcf"):
                # a VCF goes straight to the packed table: what vcf2scoary followed by
                # `-s 11` gives, without writing and re-reading the converted CSV
                if args.grabcols or args.write_reduced:
                    sys.exit("--include_input_columns and -w need a converted table: run vcf2scoary first")
                log.info("Reading variants from the VCF file")
                from . import vcf2scoary
                table = vcf2scoary.vcf_to_table(args.genes, "ALL", allowed)
                pc = _popcount_rows(table.bits)
                parsed = {"Roarydic": table,
                          "Zero_ones_matrix": LazyZeroOnes(table, np.flatnonzero((pc > 0) & (pc < len(table.strains)))),
                          "Strains": table.strains, "Extracols": [],
                          "Firstcolnames": ["#CHROM", "POS", "ID"]}
            else:
                log.info("Reading gene presence absence file")
                parsed = Csv_to_dic_Roary(genes, args.delimiter, [-999] if args.grabcols == "ALL" else args.grabcols,
                                          startcol=int(args.start_col) - 1, allowed_isolates=allowed,
                                          writereducedset=args.write_reduced, time=currenttime,
                                          outdir=args.outdir)
            genedic, strains = parsed["Roarydic"], parsed["Strains"]
            if args.newicktree is None and not args.no_pairwise:
                log.info("Creating Hamming distance matrix based on gene presence/absence")
                log.info("Building UPGMA tree from distance matrix")
                upgmatree = upgma(genedic)
            elif args.no_pairwise:
                log.info("Ignoring relatedness among input sample and performing only population structure-naive "
                         "analysis.")
                upgmatree = None
            else:
                log.info("Reading custom tree file")
                upgmatree, members = ReadTreeFromFile(args.newicktree)
                if sorted(strains) != sorted(members):
                    if args.restrict_to is None:
                        sys.exit("CRITICAL: Please make sure that isolates in your custom tree match those in your "
                                 "gene presence absence file.")
                    if all(s in members for s in strains):
                        log.info("Pruning phylogenetic tree to correspond to set of included isolates")
                        keep = set(strains)
                        upgmatree = PruneForMissing(upgmatree, [m for m in members if m not in keep])
                    else:
                        sys.exit("CRITICAL: Your provided tree file did not contain all the isolates in your gene "
                                 "presence absence file.")
            log.info("Reading traits file")
            traitsdic, Prunedic = Csv_to_dic(traits, args.delimiter, allowed, strains)
        log.info("Finished loading files into memory.\n\n")
        log.info("==== Performing statistics ====")
        for line in filtrationoptions(cutoffs, args.collapse):
            log.info(line)
        log.info("Tallying genes and performing statistical analyses")
        both = Setup_results(genedic, traitsdic, args.collapse)
        if args.upgma_tree:
            StoreUPGMAtreeToFile(upgmatree, args.outdir, time=currenttime)
        StoreResults(both["Results"], args.max_hits, cutoffs, upgmatree, both["Gene_trait_combinations"], Prunedic,
                     args.outdir, args.permute, args.threads, args.no_pairwise, genedic, parsed["Extracols"],
                     parsed["Firstcolnames"], time=currenttime, delimiter=args.delimiter)
        log.info("\n")
        log.info("==== Finished ====")
        log.info("Checked a total of %d genes for associations to %d trait(s). Total time used: %d seconds."
                 % (len(genedic), len(traitsdic), int(time.time() - starttime)))
    except SystemExit:
        log.exception("CRITICAL:")
        _close_run(fileh, console, scratch, ok=False)
        _abort_peers()
        raise
    except Exception:
        # an engine / CUDA error on ONE rank of an N > 1 run would leave the others waiting in their next
        # collective for ever: log it, close the files and take the whole job down
        log.exception("CRITICAL:")
        _close_run(fileh, console, scratch, ok=False)
        _abort_peers()
        raise
    _close_run(fileh, console, scratch, ok=True)
    sys.exit(0)


def _abort_peers():
    """N > 1 only: a rank that fails leaves the process group the hard way, so that torchrun tears the other
    ranks down instead of letting them block in a collective (exit status 1)."""
    if dist.world_rank()[0] > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
