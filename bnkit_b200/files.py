"""Input parsers and output writers as a library (SURVEY.md section 8 f-4 and the writers of f-2): the file
formats asr.GRASP reads and writes around the hot path.  Host-side text work, mirroring the reference's readers'
acceptance rules and error messages so that a file GRASP rejects is rejected here for the same reason.

    dat.file.Utils.SneakPeek / loadAlignment / checkData   src/dat/file/Utils.java:30-189
    dat.file.FastaReader.loadGappy / extract               src/dat/file/FastaReader.java:255-306, :362-405
    dat.file.AlnReader.load (CLUSTAL)                      src/dat/file/AlnReader.java:60-196
    dat.EnumSeq.Alignment (equal lengths)                  src/dat/EnumSeq.java:387-401
    dat.file.TSVFile.loadEmpiricalFreqFile + GRASP's renormalisation  src/dat/file/TSVFile.java:920-981, src/asr/GRASP.java:385-401
    dat.file.AlnWriter.save                                src/dat/file/AlnWriter.java:64-120
    dat.file.Newick.sprint / save (MODE_DEFAULT / ANCESTOR / STRIPPED)   src/dat/file/Newick.java:275-378
"""
import re

from .pog import ASRRuntimeException, java_double


class ASRException(Exception):
    """asr.ASRException (checked in the reference): problems with the input files"""


ALPHABETS = {
    "LG": "ARNDCQEGHILKMFPSTWYV", "JTT": "ARNDCQEGHILKMFPSTWYV", "WAG": "ARNDCQEGHILKMFPSTWYV",
    "Dayhoff": "ARNDCQEGHILKMFPSTWYV", "JC": "ACGT", "Yang": "ACGT",
}


def sniff_format(text):
    """Utils.SneakPeek: the format of a file from its first line with content"""
    for line in text.splitlines():
        if not line.strip():
            continue
        if line.startswith("CLUSTAL"):
            return "CLUSTAL"
        if line.startswith(">"):
            return "FASTA"
        if line.startswith("#"):
            rest = line[1:].strip()
            if rest.startswith("STOCKHOLM"):
                return "STOCKHOLM"
            if rest.startswith("NEXUS"):
                return "NEXUS"
            return None
        if line.startswith("("):
            return "NEWICK"
        return None
    return None


def parse_fasta_gappy(text, alphabet, gap="-", to_upper=True):
    """FastaReader.loadGappy: [(name, [symbol or None, ...])].  The name is the first token of the defline
    (separators: space, tab, comma, semicolon); symbols are upper-cased; anything that is neither the gap symbol nor
    in the alphabet is an error -- X, B, Z, '*', '.' are NOT ambiguity codes here (FastaReader.java:277-285)."""
    out = []
    cur = None
    for line in text.splitlines():
        if ">" in line:                           # Pattern "\\>" .find(): a '>' anywhere in the line starts an entry
            first = re.split(r"[ \t,;]+", line.strip())[0] if line.strip() else ""
            if not first:
                raise RuntimeError("Invalid Sequence format (not proper FASTA)")
            cur = (first[1:], [], line)
            out.append(cur)
            continue
        if cur is None:
            continue
        trimmed = line.strip()
        for s, ch in enumerate(trimmed):
            if to_upper:
                ch = ch.upper()
            if ch == gap:
                cur[1].append(None)
            elif ch in alphabet:
                cur[1].append(ch)
            else:
                raise RuntimeError("Sequence \"%s\" is using an invalid symbol. Unrecognised symbol: %s at index %d"
                                   % (cur[2][1:], ch, s))
    return [(name, syms) for name, syms, _ in out]


def parse_clustal(text, alphabet):
    """AlnReader.load: the first line with content must start with CLUSTAL; an alignment line starts with a letter or a
    digit; its longest token after the first is the sequence, the tokens before it form the identifier; blocks are
    matched by row position when identifiers repeat; '/start-end' suffixes are dropped from names"""
    seq, order = {}, {}
    collect = True
    row_count = 0
    prec_empty = 0
    title = None
    count = 0
    row_in_block = 0
    for raw in text.splitlines():
        row_count += 1
        line = raw.strip()
        if len(line) < 1:
            prec_empty += 1
            continue
        if prec_empty > 0:
            row_in_block = 0
        else:
            row_in_block += 1
        if title is None:
            title = line.upper()
            if not title.startswith("CLUSTAL"):
                raise RuntimeError("Not a Clustal file. First row should contain \"CLUSTAL\".")
        elif line[0].isalpha() or line[0].isdigit():
            toks = line.split()
            if len(toks) > 1:
                longest = 1
                for i in range(1, len(toks)):
                    if len(toks[i]) >= len(toks[longest]):
                        longest = i
                strid = " ".join(toks[:longest]) + "\\" + str(row_in_block)
                seqstr = toks[longest]
                if strid in seq:
                    collect = False
                    seq[strid].append(seqstr)
                else:
                    orig = order.get(row_in_block + 1)
                    if orig is None:
                        if collect:
                            count += 1
                            seq[strid] = [seqstr]
                            order[count] = strid
                        else:
                            raise RuntimeError("Invalid identifier \"%s\" on row %d" % (strid, row_count))
                    else:
                        collect = False
                        seq[orig].append(seqstr)
        prec_empty = 0
    out = []
    for k in sorted(order):
        key = order[k]
        name = key
        i = name.find("\\")
        if i > 0:
            name = name[:i]
        i = name.find("/")
        if i > 0:
            name = name[:i]
        syms = []
        for index, ch in enumerate("".join(seq[key])):
            if ch == "-":
                syms.append(None)
            elif ch in alphabet:
                syms.append(ch)
            else:
                raise RuntimeError("Sequence \"%s\" is using an invalid symbol. Unrecognised symbol: %s at index %d" % (name, ch, index))
        out.append((name, syms))
    return out


def loadAlignment(path_or_text, alphabet, is_text=False):
    """Utils.loadAlignment: (names, rows) from a CLUSTAL or FASTA file; ASRException for anything else, for unknown
    symbols and for rows of different lengths"""
    text = path_or_text if is_text else open(path_or_text).read()
    fmt = sniff_format(text)
    if fmt is None:
        raise ASRException("Format of alignment is unknown")
    try:
        if fmt == "CLUSTAL":
            seqs = parse_clustal(text, alphabet)
        elif fmt == "FASTA":
            seqs = parse_fasta_gappy(text, alphabet)
        else:
            raise ASRException("Format is not an alignment. Detected format " + fmt)
    except RuntimeError as e:
        raise ASRException(str(e))
    w = -1
    for _, s in seqs:
        if w < 0:
            w = len(s)
        elif w != len(s):
            raise ASRException("Invalid alignment with sequences of different lengths.")
    return [n for n, _ in seqs], [s for _, s in seqs]


def checkData(names, tree):
    """Utils.checkData: duplicate identifiers in the tree's extants or in the alignment, identifiers present in one
    but missing from the other -- all collected, then raised as one ASRException"""
    issues = []
    in_tree, dup_tree = set(), set()
    for idx in tree.getLeaves():
        name = str(tree.getID(idx))
        if name in in_tree:
            dup_tree.add(name)
        in_tree.add(name)
    if dup_tree:
        issues.append(("Duplicate identifiers in tree: ", dup_tree))
    in_aln, dup_aln = set(), set()
    for name in names:
        if name in in_aln:
            dup_aln.add(name)
        in_aln.add(name)
    if dup_aln:
        issues.append(("Duplicate identifiers in alignment: ", dup_aln))
    missing_in_tree = {n for n in in_aln if n not in in_tree}
    if missing_in_tree:
        issues.append(("Identifiers found in alignment but missing in tree", missing_in_tree))
    missing_in_aln = {n for n in in_tree if n not in in_aln}
    if missing_in_aln:
        issues.append(("Identifiers found in tree but missing in alignment", missing_in_aln))
    if issues:
        e = ASRException("; ".join("%s %s" % (msg, ", ".join(sorted(s))) for msg, s in issues))
        e.issues = issues
        raise e


def load_empirical_freqs(path_or_text, model_name, is_text=False):
    """TSVFile.loadEmpiricalFreqFile + the renormalisation asr.GRASP applies: a tab-separated table with the columns
    'Character' and 'Proportion', one row per state of the model's alphabet (any order); frequencies that do not sum
    to 1 within 1e-5 are divided by their sum (GRASP.java:392-399).  Returns the frequencies in model order."""
    text = path_or_text if is_text else open(path_or_text).read()
    lines = [ln for ln in text.splitlines() if ln.strip()]
    if not lines:
        raise RuntimeError("Empirical frequency file must contain 'Character' and 'Proportion' columns.")
    header = lines[0].split("\t")
    if "Character" not in header or "Proportion" not in header:
        raise RuntimeError("Empirical frequency file must contain 'Character' and 'Proportion' columns.")
    ci, pi = header.index("Character"), header.index("Proportion")
    if model_name in ("LG", "JTT", "WAG", "Dayhoff"):
        order = "ARNDCQEGHILKMFPSTWYV"
    elif model_name == "Yang":
        order = "ACGT"
    else:
        raise ValueError("Model " + model_name + " is not supported for empirical frequencies.")
    rows = [ln.split("\t") for ln in lines[1:]]
    if len(rows) != len(order):
        raise RuntimeError("Empirical frequency file has incorrect number of entries for model " + model_name + ".")
    freqs = [0.0] * len(order)
    for r in rows:
        ch = r[ci] if ci < len(r) else ""
        if ch not in order or len(ch) != 1:
            raise TypeError("Empirical frequency file has an invalid character: " + ch)      # ClassCastException
        try:
            freqs[order.index(ch)] = float(r[pi])
        except (ValueError, IndexError):
            raise ValueError("Empirical frequency file has an invalid number format: " + (r[pi] if pi < len(r) else ""))
    total = 0.0
    for f in freqs:
        total += f
    if abs(total - 1.0) >= 1e-5:
        freqs = [f / total for f in freqs]
    return freqs


# ---- writers ----
def clustal_text(names, seqs, line_width=60):
    """AlnWriter.save(String[], Object[][]): "CLUSTAL", a blank line, then blocks of 60 columns: name TAB symbols, each
    block followed by two blank lines; '-' for null"""
    out = ["CLUSTAL\n", "\n"]
    width = len(seqs[0]) if seqs else 0
    line = 0
    while line < width:
        for name, seq in zip(names, seqs):
            out.append(name + "\t" + "".join("-" if x is None else str(x) for x in seq[line:line + line_width]) + "\n")
        line += line_width
        out.append("\n\n")
    return "".join(out)


NEWICK_DEFAULT, NEWICK_ANCESTOR, NEWICK_STRIPPED = 0, 1, 2
_EMBED = ":(), @#[]&${}=!*+"
_FORBIDDEN = ":(),"


def newick_text(tree, mode=NEWICK_DEFAULT):
    """Newick.sprint(root, MODE) + ";" (what Newick.save writes): MODE_DEFAULT labels as stored (quoted when they hold
    a character Newick reserves), MODE_ANCESTOR ancestors as N<number>, MODE_STRIPPED ancestors unlabelled; distances
    as Double.toString prints them; the root carries its distance like every other node"""
    def label_of(v):
        leaf = tree.isLeaf(v)
        if not leaf and mode in (NEWICK_ANCESTOR, NEWICK_STRIPPED):
            lab = "" if mode == NEWICK_STRIPPED else "N" + str(tree.getID(v))
        else:
            raw = str(tree.getLabel(v))
            lab = raw if mode == NEWICK_STRIPPED else ("'" + raw + "'" if any(ch in raw for ch in _EMBED) else raw)
        if mode == NEWICK_STRIPPED and any(ch in lab for ch in _FORBIDDEN):
            raise RuntimeError("Invalid label: \"" + lab + "\". Please use alternative format or DEFAULT mode.")
        return lab
    # iterative post-order (a 10,000-level caterpillar must not hit the recursion limit)
    text = {}
    stack = [(0, False)]
    while stack:
        v, done = stack.pop()
        kids = tree.getChildren(v)
        if not done and kids:
            stack.append((v, True))
            for c in reversed(kids):
                stack.append((c, False))
            continue
        d = ":" + java_double(tree.getDistance(v)) if tree.has_distance[v] else ""
        if kids:
            text[v] = "(" + ",".join(text.pop(c) for c in kids) + ")" + label_of(v) + d
        else:
            text[v] = label_of(v) + d
    return text[0] + ";\n"


__all__ = ["ASRException", "ASRRuntimeException", "sniff_format", "parse_fasta_gappy", "parse_clustal", "loadAlignment",
           "checkData", "load_empirical_freqs", "clustal_text", "newick_text", "NEWICK_DEFAULT", "NEWICK_ANCESTOR",
           "NEWICK_STRIPPED"]
