"""VCF 4.x -> Scoary mutation presence/absence table.

Same command line and output file as the reference's converter (scoary/vcf2scoary.py:50-218):
one output row per ALT allele (:178-192), a DUMMY column after FORMAT that says whether the row
was split from a multi-allelic site (:162,:190,:197), genotype = first colon field (:200-202),
and for split rows "." or another allele -> "0", this allele -> "1" (fixdummy, :204-218).

vcf_to_table() applies the same rules but builds the packed GeneTable directly (the form
sb_set_genes takes), so a variant table with a million rows never goes through a CSV file:
identifiers are CHROM_|_POS_|_ID as Csv_to_dic_Roary forms them for non-Roary input
(scoary/methods.py:457-463) and a cell is present unless it is "", "0" or "-" (:476-487).
"""
import argparse
import csv
import os
import re
import sys

import numpy as np

__version__ = "0.1b"


def _read_meta_and_header(rows):
    """Consume the ## lines; returns (meta dict of simple keys + FORMAT ids, header row)."""
    meta, formats = {}, {}
    for line in rows:
        if not line:
            continue
        if line[0][:2] == "##":
            key, _, val = line[0].partition("=")
            if key == "##FORMAT":
                m = re.search(r"ID=(\w+)", val)
                n = re.search(r"Number=([^,>]+)", val)
                if m:
                    formats[m.group(1)] = n.group(1) if n else None
            else:
                meta.setdefault(key, val)
        else:
            return meta, formats, line
    sys.exit("ERROR: There appears to be only metainformation (lines starting with ##) in your VCF file.")


def _check(meta, formats):
    fmt = meta.get("##fileformat")
    try:
        version = fmt.split("v")[1]
        if int(version[0]) != 4:
            print("WARNING: A VCF format other than 4.x detected. File parsing may proceed with errors.")
        else:
            print("VCF version %s detected" % version)
    except Exception:
        print("WARNING: Could not detect VCF format. Expected v4.x. File parsing may proceed with errors.")
    if formats.get("GT") != "1":
        sys.exit("ERROR: Expected a single allele per genotype. Scoary only works for haploid organisms.")


def _variant_rows(rows, types):
    """Yields (first nine fields, dummy flag, genotype cells) per output row."""
    for line in rows:
        if not line:
            continue
        if types != "ALL":
            m = re.search(r"TYPE=(\w+)", line[7])
            if m is None or m.group(1) not in types:
                continue
        gts = [cell.split(":")[0] for cell in line[9:]]
        if "," in line[4]:
            for c, alt in enumerate(line[4].split(","), start=1):
                head = line[:9]
                head[4] = alt
                cells = []
                for g in gts:
                    if g == ".":
                        cells.append("0")
                    else:
                        try:
                            cells.append("1" if int(g) == c else "0")
                        except ValueError:
                            print(gts, c)
                            sys.exit(-1)
                yield head, "True", cells
        else:
            yield line[:9], "False", gts


def convert(vcf_path, out_path, types="ALL"):
    with open(vcf_path, "r") as fh, open(out_path, "w") as out:
        rows = csv.reader(fh, delimiter="\t", quotechar='"')
        meta, formats, header = _read_meta_and_header(rows)
        _check(meta, formats)
        header = header[:9] + ["DUMMY"] + header[9:]
        out.write(",".join('"' + c + '"' for c in header) + "\n")
        for head, dummy, cells in _variant_rows(rows, types):
            out.write(",".join('"' + c + '"' for c in head + [dummy] + cells) + "\n")
    print("Reached the end of the file")


def _native_table(vcf_path, types, keep):
    """The packed table through sb_vcf_* (csrc/vcf_pack.cpp), or None when the file needs the Python
    parser (quotes, non-ASCII bytes, ragged lines, unusual genotypes) -- then the caller runs it."""
    import ctypes
    import io
    from . import _lib
    from . import engine as eng
    from .methods import GeneTable
    lib = _lib.load()
    with open(vcf_path, "rb") as fh:
        buf = fh.read()
    odd = ctypes.c_int32(0)
    n = lib.sb_vcf_line_starts(buf, len(buf), None, 0, ctypes.byref(odd))
    if odd.value or n < 1:
        return None
    starts = np.empty(n, dtype=np.int64)
    lib.sb_vcf_line_starts(buf, len(buf), starts.ctypes.data_as(ctypes.c_void_p), n, None)
    head_end = int(starts[1]) if n > 1 else len(buf)
    rows = csv.reader(io.StringIO(buf[:head_end].decode("utf-8"), newline=None), delimiter="\t", quotechar='"')
    meta, formats, header = _read_meta_and_header(rows)
    _check(meta, formats)
    strains = header[9:]
    keep_idx = np.full(len(strains), -1, dtype=np.int32)
    kept = [j for j, s in enumerate(strains) if keep is None or s in keep]
    keep_idx[kept] = np.arange(len(kept), dtype=np.int32)
    kept_names = [strains[j] for j in kept]
    tl = None if types == "ALL" else "\n".join(types).encode("ascii", "replace")
    n_lines = n - 1
    vstarts = np.ascontiguousarray(starts[1:])
    per_line = np.zeros(max(n_lines, 1), dtype=np.int32)
    total = lib.sb_vcf_count_rows(buf, len(buf), vstarts.ctypes.data_as(ctypes.c_void_p), n_lines, tl,
                                  len(tl) if tl is not None else 0, per_line.ctypes.data_as(ctypes.c_void_p))
    if total < 0:
        return None
    W = eng.words_for(len(kept))
    offs = np.zeros(max(n_lines, 1), dtype=np.int64)
    np.cumsum(per_line[:n_lines - 1], out=offs[1:n_lines]) if n_lines > 1 else None
    bits = np.zeros((total, W), dtype=np.uint64)
    ranges = np.zeros((max(total, 1), 3, 2), dtype=np.int64)
    rc = lib.sb_vcf_pack_rows(buf, len(buf), vstarts.ctypes.data_as(ctypes.c_void_p), n_lines,
                              per_line.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), tl,
                              len(tl) if tl is not None else 0, keep_idx.ctypes.data_as(ctypes.c_void_p), len(strains),
                              bits.ctypes.data_as(ctypes.c_void_p), W, ranges.ctypes.data_as(ctypes.c_void_p), None)
    if rc != 0:
        return None
    def column(k):
        """CHROM / POS / ID of every output row: gathered natively into one blob, decoded once, sliced `total` times"""
        offs = np.empty(total + 1, dtype=np.int64)
        rp, op = ranges.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p)
        size = lib.sb_csv_gather_fields(buf, rp, total, 3, k, None, 0, op)
        blob = np.empty(max(int(size), 1), dtype=np.uint8)
        if size < 0 or lib.sb_csv_gather_fields(buf, rp, total, 3, k, blob.ctypes.data_as(ctypes.c_void_p), int(size),
                                                op) != size:
            return [buf[b:e].decode("ascii") for b, e in ranges[:total, k].tolist()]
        text, o = blob[:size].tobytes().decode("ascii"), offs.tolist()      # the native parser only accepts ASCII files
        return [text[o[r]:o[r + 1]] for r in range(total)]

    chrom, pos, vid = column(0), column(1), column(2)
    names = [c + "_|_" + p + "_|_" + i for c, p, i in zip(chrom, pos, vid)]
    return GeneTable(names, pos, vid, kept_names, bits=bits)


def vcf_to_table(vcf_path, types="ALL", allowed_isolates=None):
    """VCF -> GeneTable, equal to Csv_to_dic_Roary(converted CSV, startcol=10)["Roarydic"]
    (restricted to `allowed_isolates` like its allowed_isolates argument).  Parsed natively
    (sb_vcf_*); files the native parser is not sure about go through the Python csv module."""
    from . import engine as eng
    from .methods import GeneTable
    if os.environ.get("SCOARY_B200_PY_CSV") != "1":
        table = _native_table(vcf_path, types, allowed_isolates)
        if table is not None:
            return table
    with open(vcf_path, "r") as fh:
        rows = csv.reader(fh, delimiter="\t", quotechar='"')
        meta, formats, header = _read_meta_and_header(rows)
        _check(meta, formats)
        strains = header[9:]
        kept = [j for j, s in enumerate(strains) if allowed_isolates is None or s in allowed_isolates]
        names, nug, ann, packed = [], [], [], []
        absent = ("", "0", "-")
        chunk = []
        for head, _, cells in _variant_rows(rows, types):
            names.append(head[0] + "_|_" + head[1] + "_|_" + head[2])
            nug.append(head[1])
            ann.append(head[2])
            chunk.append([cells[j] not in absent for j in kept])
            if len(chunk) == 4096:
                packed.append(eng.pack_rows(np.asarray(chunk, dtype=np.uint8).reshape(len(chunk), len(kept))))
                chunk = []
        if chunk:
            packed.append(eng.pack_rows(np.asarray(chunk, dtype=np.uint8).reshape(len(chunk), len(kept))))
    W = eng.words_for(len(kept))
    bits = np.concatenate(packed, axis=0) if packed else np.zeros((0, W), dtype=np.uint64)
    return GeneTable(names, nug, ann, [strains[j] for j in kept], bits=bits)


def main(argv=None):
    p = argparse.ArgumentParser(description="VCF 4.x -> mutation presence/absence table in the Roary/Scoary format")
    p.add_argument("--out", default="./mutations_presence_absence.csv", help="output file")
    p.add_argument("--types", default="ALL", help="comma-separated TYPE= values of the INFO column to keep, or ALL")
    p.add_argument("--version", action="version", version=__version__)
    p.add_argument("--force", action="store_true", default=False, help="overwrite the output file")
    p.add_argument("vcf", metavar="<VCF_file>", help="the VCF file to convert")
    args = p.parse_args(argv)
    types = args.types if args.types == "ALL" else args.types.split(",")
    if os.path.isfile(args.out) and not args.force:
        sys.exit("Outfile already exists. Change name of outfile or run with --force")
    if not os.path.isfile(args.vcf):
        sys.exit("Unable to locate input file %s" % args.vcf)
    convert(args.vcf, args.out, types)
    sys.exit(0)


if __name__ == "__main__":
    main()
